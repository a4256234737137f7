// Micro-benchmark: issue rate of tcgen05.mma (kind::f16, M=128, K=16, SS operands) for the
// operand shapes the conv kernels use.  Shared memory holds garbage; only timing matters.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I joshupscale_b200/csrc/kernels \
//        bench_tools/mma_rate.cu -o gpurun_out/mma_rate && gpurun_out/mma_rate
#include <cstdio>
#include <cuda_runtime.h>

#include "tc_common.cuh"

using namespace ju::tc;

// mode 0: conv-like (A = shifted views of one 18x10 halo, SBO 1280; B = 9 resident tap slices)
// mode 1: plain GEMM-like (A tile 128 rows x 128 B, SBO 1024, same tile every time)
template <int N>
__global__ void __launch_bounds__(384, 1) mma_rate_kernel(int iters, int mode, int lsu_noise, int fill, long long *cycles) {
	extern __shared__ __align__(1024) unsigned char smem_raw[];
	__shared__ uint32_t tmem_slot;
	__shared__ __align__(8) unsigned long long bar;
	__shared__ __align__(8) unsigned long long dummy_bar[2];
	const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
	const uint32_t a_base = base;                 // 18*10*128 = 23040 B -> 24 KB
	const uint32_t w_base = base + 24 * 1024;     // 9 * N * 128 B
	const int warp = threadIdx.x / 32;
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
	}
	if (threadIdx.x == 0) {
		mbar_init(smem_u32(&bar), 1);
		mbar_init(smem_u32(&dummy_bar[0]), 1);
		mbar_init(smem_u32(&dummy_bar[1]), 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	tcgen05_fence_before();
	__syncthreads();
	tcgen05_fence_after();
	if (fill) {
		// pseudo-random fp16 data in [-1, 1): data-dependent power / throttling shows up as slower MMAs
		uint32_t x = 0x9E3779B9u * (threadIdx.x + 1) + blockIdx.x;
		uint32_t *w = reinterpret_cast<uint32_t *>(smem_raw);
		const int words = (1024 + 24 * 1024 + 9 * N * 128) / 4;
		for (int i = threadIdx.x; i < words; i += blockDim.x) {
			x = x * 1664525u + 1013904223u;
			const uint32_t lo = 0x3800u | ((x >> 3) & 0x83FFu), hi = 0x3800u | ((x >> 17) & 0x83FFu);
			w[i] = lo | (hi << 16);
		}
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		__syncthreads();
	}
	const uint32_t tmem = tmem_slot;
	const uint32_t idesc = make_idesc(N);
	if (warp == 0) {
		long long t0 = clock64();
		for (int it = 0; it < iters; ++it) {
			const uint32_t d = tmem + ((mode & 4) ? (it & 3) : (it & 1)) * N;
#pragma unroll
			for (int tap = 0; tap < 9; ++tap) {
				const int dy = tap / 3, dx = tap % 3;
				const uint32_t a0 = !(mode & 1) ? a_base + (dy * 10 + dx) * 128 : a_base;
				const uint32_t sbo = !(mode & 1) ? 1280u : 1024u;
				const uint32_t b0 = w_base + tap * N * 128;
#pragma unroll
				for (int ks = 0; ks < 4; ++ks) {
					if (elect_one_sync()) {
						umma_f16(d, make_smem_desc(a0 + ks * 32, sbo, 0), make_smem_desc(b0 + ks * 32, 1024, 0), idesc,
						    (tap | ks) != 0);
					}
					__syncwarp();
				}
			}
			if (mode & 2) {
				if (elect_one_sync()) {
					umma_commit(smem_u32(&dummy_bar[0]));
					umma_commit(smem_u32(&dummy_bar[1]));
				}
				__syncwarp();
			}
		}
		if (elect_one_sync()) umma_commit(smem_u32(&bar));
		__syncwarp();
		mbar_wait(smem_u32(&bar), 0, nullptr, 0);
		long long t1 = clock64();
		if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
	} else if (lsu_noise == -2 || lsu_noise == -3) {
		// polling warps: spin on an mbarrier phase that never completes, like idle pipeline stages do
		long long t0 = clock64();
		uint32_t fails = 0;
		const int n = iters * 36;
		for (int i = 0; i < n; ++i) {
			if (!mbar_try_wait(smem_u32(&dummy_bar[1]), 0)) ++fails;
			if (lsu_noise == -3) __nanosleep(64);
		}
		long long t1 = clock64();
		if (threadIdx.x == 32 && blockIdx.x == 0) cycles[148] = (t1 - t0) / n;
		if (fails == 0x12345678) cycles[0] = 0;
	} else if (lsu_noise < 0) {
		uint32_t r[32];
		uint32_t acc = 0;
		for (int i = 0; i < iters * 8; ++i) {
			tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (i & 3) * 64, r);
			tmem_ld_wait();
			acc += r[0] + r[31];
		}
		if (acc == 0x12345678) cycles[0] = 0;
	} else if (lsu_noise) {
		// a second warp hammering shared memory with conflict-free 128-bit loads/stores
		uint32_t addr = base + 24 * 1024 + 9 * N * 128 + (threadIdx.x & 31) * 16;
		uint32_t acc = 0;
		for (int i = 0; i < lsu_noise; ++i) {
			uint32_t x, y, z, w;
			asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(addr + (i & 7) * 512));
			acc += x + y + z + w;
			asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr + 4096 + (i & 7) * 512), "r"(acc), "r"(y), "r"(z), "r"(w));
		}
		if (acc == 0x12345678) cycles[0] = 0;
	}
	tcgen05_fence_before();
	__syncthreads();
	if (warp == 0) {
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
	}
}

#define N_TMEM_NOISE 1
template <int N>
void run(int mode, int noise, int fill = 0) {
	const int iters = 2000;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	long long *d;
	cudaMalloc(&d, 149 * sizeof(long long));
	cudaMemset(d, 0, 149 * sizeof(long long));
	const int smem = 1024 + 24 * 1024 + 9 * N * 128 + 16 * 1024;
	cudaFuncSetAttribute(mma_rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	mma_rate_kernel<N><<<148, 384, smem>>>(iters, mode, noise, fill, d);
	cudaEventRecord(e0);
	for (int rep = 0; rep < 1; ++rep) mma_rate_kernel<N><<<148, 384, smem>>>(iters, mode, noise, fill, d);
	cudaEventRecord(e1);
	cudaError_t e = cudaDeviceSynchronize();
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	long long h[149];
	cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
	double s = 0;
	for (int i = 0; i < 148; ++i) s += h[i];
	const double per = s / 148 / (iters * 36.0);
	printf("N=%3d mode=%d noise=%d fill=%d: %.1f cycles/MMA  (%.0f cycles per 36-MMA tile, %.0f MAC/clk/SM), %.1f ns/MMA => %.2f GHz, %.0f TFLOP/s %s\n",
	    N, mode, noise, fill, per, per * 36, 128.0 * N * 16 / per, ms * 1e6 / (iters * 36.0), per / (ms * 1e6 / (iters * 36.0)),
	    148.0 * 2 * 128 * N * 16 * iters * 36 / (ms * 1e-3) / 1e12, e == cudaSuccess ? "" : cudaGetErrorString(e));
	if (noise <= -2) printf("   (failed try_wait: %lld cycles per poll, %d polling warps)\n", h[148], 384 / 32 - 1);
	cudaFree(d);
}

int main() {
	run<64>(0, 0, 1);
	run<64>(2, 0, 1);
	run<64>(4, 0, 1);
	run<64>(6, 0, 1);
	run<64>(4, -2, 1);
	run<64>(4, -3, 1);
	return 0;
}
