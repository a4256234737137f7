// Micro-benchmark: the "stacked taps" formulation of the 64 -> 64 channel 3x3 convolution.
// Shipping trunk: D[128 pixels x 64 ch] += A[128 pixels x 16] (smem, shifted halo view) * B[16 x 64 ch] (smem):
// 4 KB + 2 KB of shared-memory operand reads per 32-cycle MMA = 48 cycles at 128 B/clk, 36 MMAs per 128 pixels.
// Alternative measured here: D[128 = 2 taps x 64 ch, N pixels] += A[2 taps' weights, 128 x 16] * B[16 x N pixels]
// with the stacked weights held in TENSOR MEMORY (TS form: only B comes from shared memory) or in shared
// memory (SS form).  Nine taps need 5 such MMAs per 16-channel slice (3 pairs + 1 pair + 1 half-empty),
// i.e. 20 MMAs per N pixels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I joshupscale_b200/csrc/kernels \
//        bench_tools/mma_rate3.cu -o gpurun_out/mma_rate3 && gpurun_out/mma_rate3
#include <cstdio>
#include <cuda_runtime.h>

#include "tc_common.cuh"

using namespace ju::tc;

__device__ __forceinline__ void spin(uint32_t bar, uint32_t parity) {
	while (!mbar_try_wait(bar, parity)) {
	}
}

__device__ void fill_smem(unsigned char *smem_raw, int bytes) {
	uint32_t x = 0x9E3779B9u * (threadIdx.x + 1) + blockIdx.x;
	uint32_t *w = reinterpret_cast<uint32_t *>(smem_raw);
	for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) {
		x = x * 1664525u + 1013904223u;
		const uint32_t lo = 0x3800u | ((x >> 3) & 0x83FFu), hi = 0x3800u | ((x >> 17) & 0x83FFu);
		w[i] = lo | (hi << 16);
	}
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void umma_ts_f16(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
	asm volatile(
	    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
	    "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
	    "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
	    : "memory");
}

// N pixels per tile (8 pixels per tile row -> N/8 rows, halo 10 wide), 20 MMAs per tile
template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, long long *cycles) {
	extern __shared__ __align__(1024) unsigned char smem_raw[];
	__shared__ uint32_t tmem_slot;
	__shared__ __align__(8) unsigned long long bar;
	const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
	constexpr int kHaloBytes = (N / 8 + 2) * 10 * 128;
	const uint32_t x_base = base;                                   // pixel halo tile
	const uint32_t w_base = base + ((kHaloBytes + 1023) & ~1023);    // SS form: 5 groups x 128 rows x 128 B
	const int warp = threadIdx.x / 32;
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
	}
	if (threadIdx.x == 0) {
		mbar_init(smem_u32(&bar), 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	fill_smem(smem_raw, 1024 + ((kHaloBytes + 1023) & ~1023) + 5 * 128 * 128);
	tcgen05_fence_before();
	__syncthreads();
	tcgen05_fence_after();
	const uint32_t tmem = tmem_slot;
	constexpr int kAccCols = N == 256 ? 256 : 2 * N;       // two accumulator sets per tile (one at N = 256: TMEM is full)
	constexpr int kBuffers = (2 * kAccCols + 160 <= 512) ? 2 : 1;
	constexpr int kWeightCol = kBuffers * kAccCols;       // 5 groups x 4 slices x 8 columns
	const uint32_t idesc = make_idesc(N);
	if (warp == 0) {
		long long t0 = clock64();
		for (int it = 0; it < iters; ++it) {
			const uint32_t dbase = tmem + (it % kBuffers) * kAccCols;
#pragma unroll
			for (int g = 0; g < 5; ++g) {
				// pixel shift of the B window for this group of stacked taps (conv-like addressing)
				const int dy = g < 3 ? g : (g == 3 ? 0 : 2), dx = g < 3 ? 0 : 2;
				const uint32_t d = dbase + ((g < 3 || N == 256) ? 0 : N);
#pragma unroll
				for (int ks = 0; ks < 4; ++ks) {
					if (elect_one_sync()) {
						const uint64_t bd = make_smem_desc(x_base + (dy * 10 + dx) * 128 + ks * 32, 1280u, 0);
						const uint32_t acc = (g == 0 || g == 3) ? (ks != 0) : 1u;
						if (TS) {
							umma_ts_f16(d, tmem + kWeightCol + (g * 4 + ks) * 8, bd, idesc, acc);
						} else {
							const uint64_t ad = make_smem_desc(w_base + g * 128 * 128 + ks * 32, 1024u, 0);
							umma_f16(d, ad, bd, idesc, acc);
						}
					}
					__syncwarp();
				}
			}
		}
		if (elect_one_sync()) umma_commit(smem_u32(&bar));
		__syncwarp();
		spin(smem_u32(&bar), 0);
		long long t1 = clock64();
		if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
	}
	tcgen05_fence_before();
	__syncthreads();
	if (warp == 0) {
		tcgen05_fence_after();
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
	}
}


// The exact issue pattern of trunk_ts_tc.cu: halo pitch 66 pixels, SBO 1024, three accumulator stages of 64
// columns, weights from column WCOL, optionally the output-lane masks on half of the MMAs
__device__ __forceinline__ void umma_ts_f16_masked(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc,
    uint32_t m01, uint32_t m23) {
	asm volatile(
	    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
	    "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %6, %6}, p;\n\t}" ::"r"(d_tmem),
	    "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc), "r"(m01), "r"(m23)
	    : "memory");
}

template <int PITCH, int WCOL, bool MASKED, int NACC, int COMMIT = 0>
__global__ void __launch_bounds__(128, 1) ts_pattern_kernel(int iters, long long *cycles) {
	extern __shared__ __align__(1024) unsigned char smem_raw[];
	__shared__ uint32_t tmem_slot;
	__shared__ __align__(8) unsigned long long bar;
	__shared__ __align__(8) unsigned long long bar2[2];
	const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
	const int warp = threadIdx.x / 32;
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
	}
	if (threadIdx.x == 0) {
		mbar_init(smem_u32(&bar), 1);
		mbar_init(smem_u32(&bar2[0]), 1);
		mbar_init(smem_u32(&bar2[1]), 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	fill_smem(smem_raw, 1024 + 3 * 4 * PITCH * 128);
	tcgen05_fence_before();
	__syncthreads();
	tcgen05_fence_after();
	const uint32_t tmem = tmem_slot;
	const uint32_t idesc = make_idesc(64);
	if (warp == 0) {
		long long t0 = clock64();
		for (int it = 0; it < iters; ++it) {
			const uint32_t d = tmem + (it % NACC) * 64;
			const uint32_t x_base = base + (it % 3) * 4 * PITCH * 128;
#pragma unroll
			for (int st = 0; st < 12; ++st) {
				const int kx = st >> 2, q4 = st & 3;
				const int r = q4 == 0 ? 1 : (q4 == 1 ? 2 : (q4 == 2 ? 0 : 3));
				const int g = kx * 3 + (q4 < 2 ? q4 + 1 : 0);
#pragma unroll
				for (int j = 0; j < 4; ++j) {
					if (elect_one_sync()) {
						const uint64_t bd = make_smem_desc(x_base + (r * PITCH + kx) * 128 + j * 32, 1024u, 0);
						const uint32_t a = tmem + WCOL + (g * 4 + j) * 8;
						const uint32_t acc = (st | j) != 0;
						if (MASKED && r == 0) umma_ts_f16_masked(d, a, bd, idesc, acc, 0u, 0xffffffffu);
						else if (MASKED && r == 3) umma_ts_f16_masked(d, a, bd, idesc, acc, 0xffffffffu, 0u);
						else umma_ts_f16(d, a, bd, idesc, acc);
					}
					__syncwarp();
				}
			}
			if (COMMIT >= 1) {
				if (elect_one_sync()) {
					umma_commit(smem_u32(&bar2[0]));
					if (COMMIT >= 2) umma_commit(smem_u32(&bar2[1]));
				}
				__syncwarp();
			}
		}
		if (elect_one_sync()) umma_commit(smem_u32(&bar));
		__syncwarp();
		spin(smem_u32(&bar), 0);
		long long t1 = clock64();
		if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
	}
	tcgen05_fence_before();
	__syncthreads();
	if (warp == 0) {
		tcgen05_fence_after();
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
	}
}

template <int PITCH, int WCOL, bool MASKED, int NACC, int COMMIT = 0>
void run_pattern(const char *what) {
	const int iters = 2000;
	long long *d;
	cudaMalloc(&d, 148 * sizeof(long long));
	cudaMemset(d, 0, 148 * sizeof(long long));
	const int smem = 2048 + 3 * 4 * PITCH * 128;
	fprintf(stderr, "pattern %s: smem %d\n", what, smem);
	cudaFuncSetAttribute(ts_pattern_kernel<PITCH, WCOL, MASKED, NACC, COMMIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	ts_pattern_kernel<PITCH, WCOL, MASKED, NACC, COMMIT><<<148, 128, smem>>>(iters, d);
	ts_pattern_kernel<PITCH, WCOL, MASKED, NACC, COMMIT><<<148, 128, smem>>>(iters, d);
	cudaError_t e = cudaDeviceSynchronize();
	fprintf(stderr, "sync: %s\n", cudaGetErrorString(e));
	long long h[148];
	cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
	double s = 0;
	for (int i = 0; i < 148; ++i) s += h[i];
	printf("trunk_ts pattern (%s): %.1f cycles per MMA, %.0f per unit %s\n", what, s / 148 / (iters * 48.0), s / 148 / iters,
	    e == cudaSuccess ? "" : cudaGetErrorString(e));
	cudaFree(d);
}

template <int N, bool TS>
void run() {
	const int iters = 2000;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	long long *d;
	cudaMalloc(&d, 148 * sizeof(long long));
	cudaMemset(d, 0, 148 * sizeof(long long));
	const int smem = 2048 + (N / 8 + 2) * 10 * 128 + 1024 + 5 * 128 * 128;
	cudaFuncSetAttribute(rate_kernel<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	rate_kernel<N, TS><<<148, 128, smem>>>(iters, d);
	cudaEventRecord(e0);
	rate_kernel<N, TS><<<148, 128, smem>>>(iters, d);
	cudaEventRecord(e1);
	cudaError_t e = cudaDeviceSynchronize();
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	long long h[148];
	cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
	double s = 0;
	for (int i = 0; i < 148; ++i) s += h[i];
	const double per = s / 148 / (iters * 20.0);
	// useful work: 9 taps x 64 x 64 x N pixels x 2 per tile
	const double useful = 148.0 * iters * 9.0 * 64 * 64 * N * 2;
	printf("stacked taps, weights in %s, N=%3d pixels: %.1f cycles per MMA, %.2f cycles per pixel (shipping form: 13.5 at its 48-cycle bound), "
	       "%.3f ms, %.0f useful TFLOP/s %s\n",
	    TS ? "TMEM" : "smem", N, per, per * 20 / N, ms, useful / (ms * 1e-3) / 1e12, e == cudaSuccess ? "" : cudaGetErrorString(e));
	cudaFree(d);
}

int main() {
	setvbuf(stdout, nullptr, _IONBF, 0);
	run_pattern<66, 192, true, 3>("pitch 66, weights at col 192, masks, 3 acc");
	run_pattern<66, 192, false, 3>("pitch 66, col 192, no masks, 3 acc");
	run_pattern<66, 192, false, 3, 1>("same + 1 commit per unit");
	run_pattern<66, 192, false, 3, 2>("same + 2 commits per unit");
	run_pattern<66, 224, false, 2>("pitch 66, col 224, no masks, 2 acc");
	run_pattern<64, 192, true, 3>("pitch 64, col 192, masks, 3 acc");
	run<64, true>();
	run<128, true>();
	run<256, true>();
	run<64, false>();
	run<128, false>();
	run<256, false>();
	return 0;
}
