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
	run<64, true>();
	run<128, true>();
	run<256, true>();
	run<64, false>();
	run<128, false>();
	run<256, false>();
	return 0;
}
