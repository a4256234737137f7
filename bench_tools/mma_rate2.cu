// Micro-benchmark: issue rate of tcgen05.mma.cta_group::2 (M = 256 over a CTA pair, N = 64, K = 16,
// fp16, SS operands) with the trunk's conv-like operand addressing, next to the cta_group::1
// M128 N64 figure.  Question it answers: how far does halving the B operand fetch (each CTA supplies
// its own A tile and HALF of B) lift the shared-memory operand bound of the N = 64 layers?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I joshupscale_b200/csrc/kernels \
//        bench_tools/mma_rate2.cu -o gpurun_out/mma_rate2 && gpurun_out/mma_rate2
#include <cstdio>
#include <cuda_runtime.h>

#include "tc_common.cuh"

using namespace ju::tc;

__device__ __forceinline__ uint32_t cluster_ctarank() {
	uint32_t r;
	asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
	return r;
}
__device__ __forceinline__ void cluster_sync_all() {
	asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
	asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
	asm volatile(
	    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
	    "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
	    "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
	    : "memory");
}
__device__ __forceinline__ void umma2_commit_both(uint32_t bar) {
	asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
	    "h"(static_cast<uint16_t>(3))
	    : "memory");
}

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

// pair = 1: cta_group::2, M = 256, each CTA holds a full A halo and 32 of the 64 B rows per tap
// pair = 0: cta_group::1, M = 128, N = 64 (the shipping trunk)
// nstage accumulator stages rotate like the trunk's (64 columns each)
template <int PAIR>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, int nacc, long long *cycles) {
	extern __shared__ __align__(1024) unsigned char smem_raw[];
	__shared__ uint32_t tmem_slot;
	__shared__ __align__(8) unsigned long long bar;
	const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
	const uint32_t a_base = base;              // 18*10*128 = 23040 B -> 24 KB
	const uint32_t w_base = base + 24 * 1024;  // 9 taps x (PAIR ? 32 : 64) rows x 128 B
	constexpr uint32_t kBSlice = (PAIR ? 32u : 64u) * 128u;
	const int warp = threadIdx.x / 32;
	const bool leader = !PAIR || cluster_ctarank() == 0;
	if (warp == 0) {
		if (PAIR) {
			asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256));
			asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
		} else {
			asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256));
			asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
		}
	}
	if (threadIdx.x == 0) {
		mbar_init(smem_u32(&bar), 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	fill_smem(smem_raw, 1024 + 24 * 1024 + 9 * kBSlice);
	tcgen05_fence_before();
	if (PAIR) cluster_sync_all(); else __syncthreads();
	tcgen05_fence_after();
	const uint32_t tmem = tmem_slot;
	const uint32_t idesc = (1u << 4) | (static_cast<uint32_t>(64 >> 3) << 17) | (static_cast<uint32_t>((PAIR ? 256 : 128) >> 4) << 24);
	if (warp == 0 && leader) {
		long long t0 = clock64();
		for (int it = 0; it < iters; ++it) {
			const uint32_t d = tmem + (it % nacc) * 64;
#pragma unroll
			for (int tap = 0; tap < 9; ++tap) {
				const int dy = tap / 3, dx = tap % 3;
				const uint32_t a0 = a_base + (dy * 10 + dx) * 128;
				const uint32_t b0 = w_base + tap * kBSlice;
#pragma unroll
				for (int ks = 0; ks < 4; ++ks) {
					if (elect_one_sync()) {
						const uint64_t ad = make_smem_desc(a0 + ks * 32, 1280u, 0), bd = make_smem_desc(b0 + ks * 32, 1024u, 0);
						if (PAIR) umma2_f16(d, ad, bd, idesc, (tap | ks) != 0);
						else umma_f16(d, ad, bd, idesc, (tap | ks) != 0);
					}
					__syncwarp();
				}
			}
		}
		if (elect_one_sync()) {
			if (PAIR) umma2_commit_both(smem_u32(&bar)); else umma_commit(smem_u32(&bar));
		}
		__syncwarp();
		spin(smem_u32(&bar), 0);
		long long t1 = clock64();
		if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
	} else if (warp == 0 && PAIR) {
		spin(smem_u32(&bar), 0);  // the follower's copy of the multicast commit
	}
	tcgen05_fence_before();
	if (PAIR) cluster_sync_all(); else __syncthreads();
	if (warp == 0) {
		tcgen05_fence_after();
		if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
		else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
	}
}

// Weight-stationary form: tcgen05.mma.ws keeps the B operand (one tap's 64 x 16 weight slice) in a
// collector buffer, so a group of G tiles (own halo stage and accumulator each) fetches it once:
// per MMA the shared-memory port carries 4 KB of A + 2/G KB of B instead of 4 + 2 KB.
template <int G>
__global__ void __launch_bounds__(128, 1) ws_rate_kernel(int iters, long long *cycles) {
	extern __shared__ __align__(1024) unsigned char smem_raw[];
	__shared__ uint32_t tmem_slot;
	__shared__ __align__(8) unsigned long long bar;
	const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
	const uint32_t a_base = base;                  // G halo tiles of 24 KB
	const uint32_t w_base = base + G * 24 * 1024;  // 9 taps x 64 rows x 128 B
	const int warp = threadIdx.x / 32;
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
	}
	if (threadIdx.x == 0) {
		mbar_init(smem_u32(&bar), 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	fill_smem(smem_raw, 1024 + G * 24 * 1024 + 9 * 64 * 128);
	tcgen05_fence_before();
	__syncthreads();
	tcgen05_fence_after();
	const uint32_t tmem = tmem_slot;
	const uint32_t idesc = make_idesc(64);
	if (warp == 0) {
		long long t0 = clock64();
		for (int it = 0; it < iters; ++it) {
			const uint32_t dbase = tmem + ((it & 1) * G) * 64;  // two groups of accumulators alternate
#pragma unroll
			for (int tap = 0; tap < 9; ++tap) {
				const int dy = tap / 3, dx = tap % 3;
#pragma unroll
				for (int ks = 0; ks < 4; ++ks) {
					if (elect_one_sync()) {
						const uint64_t bd = make_smem_desc(w_base + tap * 64 * 128 + ks * 32, 1024u, 0);
#pragma unroll
						for (int g = 0; g < G; ++g) {
							const uint64_t ad = make_smem_desc(a_base + g * 24 * 1024 + (dy * 10 + dx) * 128 + ks * 32, 1280u, 0);
							const uint32_t d = dbase + g * 64;
							const uint32_t acc = (tap | ks) != 0;
							if (G == 1) {
								asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
								             "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::discard [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
							} else if (g == 0) {
								asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
								             "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::fill [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
							} else if (g == G - 1) {
								asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
								             "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::lastuse [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
							} else {
								asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
								             "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::use [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
							}
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

template <int G>
void run_ws() {
	const int iters = 1000;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	long long *d;
	cudaMalloc(&d, 148 * sizeof(long long));
	cudaMemset(d, 0, 148 * sizeof(long long));
	const int smem = 1024 + G * 24 * 1024 + 9 * 64 * 128 + 1024;
	cudaFuncSetAttribute(ws_rate_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	ws_rate_kernel<G><<<148, 128, smem>>>(iters, d);
	cudaEventRecord(e0);
	ws_rate_kernel<G><<<148, 128, smem>>>(iters, d);
	cudaEventRecord(e1);
	cudaError_t e = cudaDeviceSynchronize();
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	long long h[148];
	cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
	double s = 0;
	for (int i = 0; i < 148; ++i) s += h[i];
	const double per = s / 148 / (iters * 36.0 * G);
	printf("cta_group::1 .ws M=128 N=64, B slice shared by %d tile(s): %.1f cycles per MMA, %.3f ms, %.0f TFLOP/s %s\n", G, per, ms,
	    148.0 * 2 * 128 * 64 * 16 * iters * 36 * G / (ms * 1e-3) / 1e12, e == cudaSuccess ? "" : cudaGetErrorString(e));
	cudaFree(d);
}

template <int PAIR>
void run(int nacc) {
	const int iters = 2000;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	long long *d;
	cudaMalloc(&d, 148 * sizeof(long long));
	cudaMemset(d, 0, 148 * sizeof(long long));
	const int smem = 1024 + 24 * 1024 + 9 * 64 * 128 + 1024;
	cudaFuncSetAttribute(rate_kernel<PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	cudaLaunchConfig_t cfg{};
	cfg.gridDim = dim3(148);
	cfg.blockDim = dim3(128);
	cfg.dynamicSmemBytes = smem;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = PAIR ? 2 : 1;
	attr[0].val.clusterDim.y = 1;
	attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr;
	cfg.numAttrs = 1;
	cudaLaunchKernelEx(&cfg, rate_kernel<PAIR>, iters, nacc, d);
	cudaEventRecord(e0);
	cudaLaunchKernelEx(&cfg, rate_kernel<PAIR>, iters, nacc, d);
	cudaEventRecord(e1);
	cudaError_t e = cudaDeviceSynchronize();
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	long long h[148];
	cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
	double s = 0;
	int n = 0;
	for (int i = 0; i < 148; ++i) {
		if (h[i] > 0) {
			s += h[i];
			++n;
		}
	}
	const double per = s / n / (iters * 36.0);  // cycles per MMA instruction (per issuing CTA)
	const double flops = 148.0 * 2 * 128 * 64 * 16 * iters * 36;  // every SM computes 128 x 64 x 16 per instruction
	printf("%s N=64 acc-stages=%d: %.1f cycles per MMA (%.0f per 36-MMA tile%s), %.3f ms, %.0f TFLOP/s %s\n",
	    PAIR ? "cta_group::2 M=256" : "cta_group::1 M=128", nacc, per, per * 36, PAIR ? " pair" : "", ms,
	    flops / (ms * 1e-3) / 1e12, e == cudaSuccess ? "" : cudaGetErrorString(e));
	cudaFree(d);
}

int main() {
	run<0>(2);
	run<0>(4);
	run<1>(2);
	run<1>(4);
	run_ws<1>();
	run_ws<2>();
	run_ws<4>();
	return 0;
}
