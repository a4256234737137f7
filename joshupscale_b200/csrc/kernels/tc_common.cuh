// Device-side PTX helpers shared by the tcgen05 kernels (conv_tc.cu, tail_tc.cu):
// mbarrier, TMA bulk tensor copies, tcgen05 MMA / TMEM access, descriptors.
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdint>

#include "kernels.h"

namespace ju {
namespace tc {

// ---- PTX helpers ----------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
	return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
	uint32_t ok;
	asm volatile(
	    "{\n\t.reg .pred p;\n\t"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
	    "selp.u32 %0, 1, 0, p;\n\t}"
	    : "=r"(ok)
	    : "r"(bar), "r"(parity)
	    : "memory");
	return ok;
}

// non-blocking phase test
__device__ __forceinline__ uint32_t mbar_test_wait(uint32_t bar, uint32_t parity) {
	uint32_t ok;
	asm volatile(
	    "{\n\t.reg .pred p;\n\t"
	    "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
	    "selp.u32 %0, 1, 0, p;\n\t}"
	    : "=r"(ok)
	    : "r"(bar), "r"(parity)
	    : "memory");
	return ok;
}

// ---- bounded, abortable waits ----------------------------------------------
// A protocol bug or a lost dependency must surface as a recoverable error, never as a hung GPU
// and never as a trap (a trap is a sticky context error that kills every runtime of the host
// process; the reference reports failures as exceptions, core/src/tensorrt_backend.cc:266).
//
// TcStatus lives in device memory, one per engine.  A wait that exceeds `timeout_ms` records its
// code (also in the host-mapped word the engine reads after the stream synchronize) and the
// waiting thread turns DEAD: every later wait returns at once and every side effect (TMA, MMA,
// arrive, publish) is skipped - the control flow itself is unchanged, so named barriers and the
// final __syncthreads are still reached by everybody and the kernel drains in microseconds.
// Other threads / CTAs learn about the abort from `code` in the slow path of their own waits.
constexpr int kTimeoutMsDefault = 4000;

__device__ __forceinline__ unsigned long long globaltimer_ns() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}

struct Waiter {
	TcStatus *status;
	int kernel_id;
	int layer;  // context reported with a time-out (persistent kernels keep it up to date)
	bool dead;
	__device__ __forceinline__ Waiter(TcStatus *s, int id) : status(s), kernel_id(id), layer(0), dead(false) {}

	// converged-warp roles: one lane's abort is everybody's
	__device__ __forceinline__ void sync_warp() { dead = __any_sync(0xffffffffu, dead) != 0; }

	__device__ __forceinline__ void fail(int code) {
		dead = true;
		if (!status) return;
		const int v = (kernel_id << 8) | code;
		atomicOr(&status->pending, 1u << (code & 31));
		if (atomicCAS(&status->code, 0, v) == 0) {
			const int where = (layer << 16) | static_cast<int>(blockIdx.x & 0xffffu);
			status->where = where;
			int *h = status->host_code;
			if (h) {
				*reinterpret_cast<volatile int *>(h + 1) = where;
				*reinterpret_cast<volatile int *>(h) = v;
				__threadfence_system();
			}
		}
	}

	// true when the barrier phase completed; false once the frame is aborted
	__device__ __forceinline__ bool wait(uint32_t bar, uint32_t parity, int code) {
		if (dead) return false;
		unsigned long long t0 = 0;
		for (uint32_t i = 0;; ++i) {
			if (mbar_try_wait(bar, parity)) return true;
			if (i < 128u) continue;  // tight polls first: most waits complete within a few hundred ns
			__nanosleep(128);
			if (i == 128u) t0 = globaltimer_ns();
			if ((i & 255u) == 0u && poll_expired(t0, code)) return false;
		}
	}

	// a global-memory poll loop (dataflow counters) calls this every few hundred spins
	__device__ __forceinline__ bool poll_expired(unsigned long long t0, int code) {
		if (status && *reinterpret_cast<volatile int *>(&status->code) != 0) {
			atomicOr(&status->pending, 1u << (code & 31));
			dead = true;
			return true;
		}
		unsigned long long limit = static_cast<unsigned long long>(kTimeoutMsDefault) * 1000000ull;
		if (status) {
			const int ms = *reinterpret_cast<volatile int *>(&status->timeout_ms);
			if (ms > 0) limit = static_cast<unsigned long long>(ms) * 1000000ull;
		}
		if (globaltimer_ns() - t0 > limit) {
			fail(code);
			return true;
		}
		return false;
	}
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0,
    int c1, int c2, int c3) {
	asm volatile(
	    "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
	    " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
	    "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
	    : "memory");
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0,
    int c1) {
	asm volatile(
	    "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
	    " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
	    "l"(map), "r"(bar), "r"(c0), "r"(c1)
	    : "memory");
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap *map, uint32_t src, int c0, int c1, int c2,
    int c3) {
	asm volatile(
	    "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map),
	    "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
	    : "memory");
	asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// the epilogue warps only (threads 64..)
template <int kCount = 128>
__device__ __forceinline__ void epilogue_barrier() {
	asm volatile("bar.sync 1, %0;" ::"n"(kCount) : "memory");
}

__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_launch_dependents() {
	asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// One lane of a CONVERGED warp is elected; keeping the warp converged lets the
// compiler hold descriptors in uniform registers and issue UTCHMMA / UTMALDG
// without the per-lane R2UR + ELECT retry loop it emits under `if (lane == 0)`.
__device__ __forceinline__ bool elect_one_sync() {
	uint32_t pred = 0;
	asm volatile(
	    "{\n\t.reg .pred p;\n\t"
	    "elect.sync _|p, 0xffffffff;\n\t"
	    "selp.u32 %0, 1, 0, p;\n\t}"
	    : "=r"(pred)::"memory");
	return pred != 0;
}

__device__ __forceinline__ void tcgen05_fence_before() {
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t bar) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
	             : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (fp16 in, fp32 accumulate)
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
    uint32_t accumulate) {
	asm volatile(
	    "{\n\t.reg .pred p;\n\t"
	    "setp.ne.b32 p, %4, 0;\n\t"
	    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
	    "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
	    : "memory");
}

// UMMA shared-memory matrix descriptor, K-major, SWIZZLE_128B (cute
// UMMA::SmemDescriptor bit layout): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), base_offset [49,52), layout [61,64).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_off) {
	uint64_t d = 0;
	d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
	d |= static_cast<uint64_t>(1) << 16;
	d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
	d |= static_cast<uint64_t>(1) << 46;
	d |= static_cast<uint64_t>(base_off & 7u) << 49;
	d |= static_cast<uint64_t>(2) << 61;
	return d;
}

// Instruction descriptor, kind::f16: D fp32, A/B fp16, both K-major, M=128.
__device__ __forceinline__ uint32_t make_idesc(int n) {
	return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
	asm volatile(
	    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
	    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
	    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
	    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
	      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
	      "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
	      "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
	      "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
	    : "r"(taddr)
	    : "memory");
}

// GPU-scope loads of the dataflow counters: relaxed while spinning (no L1 invalidation per poll),
// one acquire once the expected value has been seen
__device__ __forceinline__ unsigned int ld_relaxed_gpu(const unsigned int *p) {
	unsigned int v;
	asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int *p) {
	unsigned int v;
	asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}

__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }


}  // namespace tc
}  // namespace ju
