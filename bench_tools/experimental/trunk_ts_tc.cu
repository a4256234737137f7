// Persistent ResBlock trunk, STACKED-ROWS version with the weights in tensor memory.
//
// Same job and same dataflow protocol as trunk_df_tc.cu (all 3x3 64->64 layers of the generator's
// residual stack - scripts/training/models.py:193-254, 544-550 - in one launch, per-(layer, wave)
// release/acquire counters instead of grid barriers), but the GEMM is transposed:
//
//   D[128 = (output row y | output row y+1) x 64 channels, 64 pixels]
//        += A[tensor memory: two stacked 64 x 16 weight slices] * B[shared memory: 64 pixels x 16 channels]
//
// For input row offset r = 0..3 and column offset kx = 0..2 the upper half of A is W(ky = r, kx)
// (what output row y takes from input row y - 1 + r) and the lower half W(ky = r - 1, kx) (what
// output row y + 1 takes from the same input row); halves without a tap are zero.  12 groups x 4
// slices of 16 input channels = 48 MMAs (M128 N64 K16) per unit of 2 x 64 output pixels.
//
// Why: with pixels as M (trunk_df_tc.cu) every MMA fetches 4 KB of A + 2 KB of B from shared
// memory for 32 tensor cycles, i.e. 48 cycles of the 128 B/clk port: the layer cannot exceed 0.67
// of the tensor rate and the port is ~92 % busy.  Here A never touches shared memory (it is
// written once per layer with tcgen05.st), B is 2 KB per MMA, and the instruction issues at the
// tensor rate (32.0 cycles, bench_tools/mma_rate3.cu); 6 of the 12 groups are half empty, so the
// useful rate is 0.75 of the tensor peak = 12 cycles per pixel against 13.5 (bound) / 18.9
// (achieved) before.  Numerics of every building block: bench_tools/ts_unit_test.cu.
//
// Unit = 2 output rows x 64 pixels; halo = 4 x 66 pixels (TMA box, SWIZZLE_128B, one 128-byte row
// per pixel); the B window of group (r, kx) is the halo shifted by r rows and kx pixels, i.e. the
// same shared-memory bytes addressed with a shifted descriptor start.
//
// Warp roles (18 warps, 1 CTA / SM):
//   warp 0        halo producer (stage-empty wait, dataflow polls, TMA)
//   warp 1, 16    MMA issuers, alternate units (accumulator stage = unit parity)
//   warps 2..9    epilogue, two groups of four on alternate units.  A warp owns one TMEM lane
//                 quadrant = 32 channels of one output row: tcgen05.ld.16x256b hands it the
//                 accumulator in mma-fragment layout, the shortcut tile is copied into the staging
//                 tile with cp.async and read back with ldmatrix.trans, the result goes back in
//                 place with stmatrix.trans - which is the channel-major -> NHWC transposition
//   warps 10..13  weight writers: shared-memory tap ring -> registers -> tcgen05.st, group by
//                 group behind the last units of a layer (as the tap-by-tap swap of trunk_df_tc.cu)
//   warps 14, 15  store + publish, one per epilogue group
//   warp 17       weight TMA: streams the 9 taps of the next layer into a 7-slot ring
#include <cstring>

#include "kernels.h"
#include "tc_common.cuh"

namespace ju {

namespace {

using namespace tc;

constexpr int kUnitW = 64, kUnitH = 2;
constexpr int kHaloW = kUnitW + 2, kHaloH = kUnitH + 2;
constexpr uint32_t kHaloBytes = static_cast<uint32_t>(kHaloH * kHaloW) * 128u;  // 33792 = 33 KB (a multiple of 1024)
constexpr int kStages = 3;  // halo stages; every (stage, issuer) pair has its own `full` barrier
constexpr uint32_t kStageTile = static_cast<uint32_t>(kUnitH * kUnitW) * 128u;  // 16 KB staging tile per epilogue group
constexpr int kTapSlots = 7;
constexpr uint32_t kTapBytes = 64u * 128u;
// Stored weight groups: per kx, t = 1, 2 hold [W(t, kx); W(t - 1, kx)] (input rows 1 and 2 of the unit's
// halo feed both output rows); t = 0 holds [W(0, kx); W(2, kx)] and is used twice with an output-lane
// mask: for input row 0 only the upper half of D is written, for input row 3 only the lower half.
// 9 groups x 32 columns = 288 TMEM columns, which leaves room for THREE accumulator stages - the
// issuer <-> epilogue hand-over takes ~2000 cycles and does not fit behind one unit's MMAs.
constexpr int kGroups = 9;
constexpr int kAccStages = 3;
constexpr uint32_t kWCol = 64u * kAccStages;  // TMEM: accumulators first, weight slices (g * 4 + j) * 8 behind them
static_assert(kWCol + kGroups * 32u <= 512u, "tensor memory budget");
constexpr int kThreadsTs = 32 * 18;
constexpr int kIssuer2 = 16, kTapWarp = 17;
constexpr uint32_t kSmemLimitTs = 227 * 1024;
constexpr uint32_t kBarBytesTs = 512u;
constexpr uint32_t kResBytes = 64u * 64u;  // per epilogue warp: 64 pixels x 32 channels of the shortcut
constexpr uint32_t kSmemTs = 1024u + kStages * kHaloBytes + 2u * kStageTile + 8u * kResBytes + kTapSlots * kTapBytes + kBarBytesTs;
static_assert(kHaloBytes % 1024u == 0, "halo stages must keep the 1024-byte swizzle phase");
static_assert(kSmemTs <= kSmemLimitTs, "shared memory budget");

struct TsParams {
	int batch, h, w;
	int units_x, units_y, total_units;
	int n_layers;
	int act;
	float slope;
	int pdl;
	const float *bias;            // [n_layers][64]
	unsigned int *sync_counter;   // [0] finished-warp counter, [1] launch epoch
	unsigned int *flags;          // [n_layers][n_waves] stored-unit counters, never reset
	TcStatus *status;
	const __half *buffers[3];     // T0, T1, T2
	int cstride;
	int lead;
	int dbg;  // what-if timing switches (results are garbage when set)
};

struct TsMaps {
	CUtensorMap in[4];    // halo boxes (64 ch, 66, 4, 1) over T0, T1, T2 and the lead layer's input
	CUtensorMap tile[3];  // unit tiles (64 ch, 64, 2, 1) over T0, T1, T2: output stores
	CUtensorMap w;        // weights of all layers: rows [layer][tap][cout], 64 ch each; box = one tap
};

__device__ __forceinline__ void umma_ts_f16(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
	asm volatile(
	    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
	    "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
	    "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
	    : "memory");
}
// same with an output-lane mask: bit i of word w set = lane 32 * w + i of D is NOT written
__device__ __forceinline__ void umma_ts_f16_masked(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc,
    uint32_t m01, uint32_t m23) {
	asm volatile(
	    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
	    "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %6, %6}, p;\n\t}" ::"r"(d_tmem),
	    "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc), "r"(m01), "r"(m23)
	    : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
	asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
	    "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
	    : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t (&r)[32]) {
	asm volatile(
	    "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
	    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
	    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
	    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
	      "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
	      "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
	      "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
	    : "r"(taddr)
	    : "memory");
}
__device__ __forceinline__ void stmatrix_x4_trans(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
	asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d)
	             : "memory");
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d) {
	asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
	             : "=r"(a), "=r"(b), "=r"(c), "=r"(d)
	             : "r"(addr)
	             : "memory");
}
// 16 bytes global -> shared; src_bytes = 0 zero-fills (pixels outside the image)
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void *src, uint32_t src_bytes) {
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// byte offset of (128-byte line `line`, channel `ch`) in a SWIZZLE_128B tile with a 1024-byte aligned base
__device__ __forceinline__ uint32_t swz(uint32_t line, uint32_t ch) {
	return line * 128u + ((((ch >> 3) ^ (line & 7u)) << 4) | ((ch & 7u) << 1));
}

__global__ void __launch_bounds__(kThreadsTs, 1) trunk_ts_tc_kernel(const __grid_constant__ TsMaps maps, const TsParams p) {
	extern __shared__ uint8_t smem_raw[];
	const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
	const uint32_t stage_base = smem_base + kStages * kHaloBytes;  // two staging tiles
	const uint32_t res_base = stage_base + 2u * kStageTile;        // private shortcut buffers of the 8 epilogue warps
	const uint32_t tap_base = res_base + 8u * kResBytes;           // weight tap ring
	const uint32_t bar_base = tap_base + kTapSlots * kTapBytes;
	// Two issuers and two epilogue groups take units alternately while halo and accumulator stages
	// rotate with period 3, so consecutive phases of one stage belong to different waiters.  A parity
	// wait is only unambiguous for a waiter that has seen every earlier phase of its barrier, hence
	// one barrier per (stage, waiter): full[s][issuer], tfull[a][group], tempty[a][issuer].
	auto full_bar = [&](int s, int issuer) { return bar_base + 8u * (2 * s + issuer); };
	auto empty_bar = [&](int s) { return bar_base + 8u * (6 + s); };
	auto tfull_bar = [&](int a, int group) { return bar_base + 8u * (9 + 2 * a + group); };
	auto tempty_bar = [&](int a, int issuer) { return bar_base + 8u * (15 + 2 * a + issuer); };
	auto sready_bar = [&](int s) { return bar_base + 8u * (21 + s); };
	auto sfree_bar = [&](int s) { return bar_base + 8u * (23 + s); };
	auto wfull_bar = [&](int g) { return bar_base + 8u * (25 + g); };
	auto wempty_bar = [&](int g) { return bar_base + 8u * (34 + g); };
	auto tapfull_bar = [&](int s) { return bar_base + 8u * (43 + s); };
	auto tapempty_bar = [&](int s) { return bar_base + 8u * (50 + s); };
	const uint32_t tmem_slot = bar_base + 8u * 57;
	static_assert(8u * 58 <= kBarBytesTs, "barrier block");
	static_assert(kStages == 3 && kAccStages == 3 && kTapSlots == 7, "barrier map");

	const int warp = threadIdx.x >> 5;
	const int lane = threadIdx.x & 31;

	if (warp == 0 && lane == 0) {
		for (int s = 0; s < kStages; ++s) {
			mbar_init(full_bar(s, 0), 1);
			mbar_init(full_bar(s, 1), 1);
			mbar_init(empty_bar(s), 1);
		}
		for (int a = 0; a < kAccStages; ++a) {
			for (int w2 = 0; w2 < 2; ++w2) {
				mbar_init(tfull_bar(a, w2), 1);
				mbar_init(tempty_bar(a, w2), 4);
			}
		}
		for (int g2 = 0; g2 < 2; ++g2) {
			mbar_init(sready_bar(g2), 4);
			mbar_init(sfree_bar(g2), 1);
		}
		for (int g = 0; g < kGroups; ++g) {
			mbar_init(wfull_bar(g), 4);
			mbar_init(wempty_bar(g), 2);
		}
		for (int s = 0; s < kTapSlots; ++s) {
			mbar_init(tapfull_bar(s), 1);
			mbar_init(tapempty_bar(s), 4);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 1) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	tcgen05_fence_before();
	__syncthreads();
	tcgen05_fence_after();
	uint32_t tmem_base;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

	if (p.pdl) grid_launch_dependents();

	auto decode = [&](int unit, int &b, int &y0, int &x0) {
		const int ux = unit % p.units_x;
		const int rest = unit / p.units_x;
		y0 = (rest % p.units_y) * kUnitH;
		x0 = ux * kUnitW;
		b = rest / p.units_y;
	};
	// buffer rotation: identical to trunk_df_tc.cu
	const int lead = p.lead;
	auto layer_in = [lead](int l) {
		if (lead && l == 0) return 3;
		l -= lead;
		return (l & 1) ? 1 : (((l >> 1) & 1) ? 2 : 0);
	};
	auto layer_res = [lead](int l) {
		if (lead && l == 0) return -1;
		l -= lead;
		return (l & 1) ? (((l >> 1) & 1) ? 2 : 0) : -1;
	};
	auto layer_out = [lead](int l) {
		if (lead && l == 0) return 0;
		l -= lead;
		return (l & 1) ? (((l >> 1) & 1) ? 0 : 2) : 1;
	};

	const unsigned int epoch = *reinterpret_cast<volatile unsigned int *>(p.sync_counter + 1);
	const int grid = static_cast<int>(gridDim.x);
	const int first = static_cast<int>(blockIdx.x);
	const int n_waves = (p.total_units + grid - 1) / grid;
	// units of this CTA per layer
	const int cnt = first < p.total_units ? (p.total_units - first + grid - 1) / grid : 0;
	// the 3x3 unit neighbourhood spans unit indices u +- (units_x + 1)
	const int wave_reach = (first + p.units_x + 1) / grid;

	if (warp == 0) {
		// ===================== halo producer =====================
		Waiter W(p.status, TC_KERNEL_TRUNK);
		if (p.pdl) grid_dependency_wait();
		int it = 0;
		for (int l = 0; l < p.n_layers; ++l) {
			W.layer = l;
			const CUtensorMap *min = &maps.in[layer_in(l)];
			int known = -1;
			for (int k = 0; k < cnt; ++k, ++it) {
				const int unit = first + k * grid;
				int b, y0, x0;
				decode(unit, b, y0, x0);
				const int s = it % kStages;
				const uint32_t ph = (it / kStages) & 1;
				W.wait(empty_bar(s), ph ^ 1u, 1);
				W.sync_warp();
				bool polled = false;
				if (l > 0 && !W.dead && !(p.dbg & 1)) {
					const int needw = k + wave_reach < n_waves ? k + wave_reach : n_waves - 1;
					while (known < needw && !W.dead) {
						const int wv = known + 1 + lane;
						const bool mine = wv <= needw;
						const int cntw = wv + 1 < n_waves ? grid : p.total_units - wv * grid;
						const unsigned int target = (epoch + 1u) * static_cast<unsigned int>(cntw);
						const unsigned int *ctr = p.flags + (l - 1) * n_waves + (mine ? wv : 0);
						unsigned int spins = 0;
						unsigned long long t0 = 0;
						while (true) {
							const bool ok = !mine || static_cast<int>(ld_relaxed_gpu(ctr) - target) >= 0;
							if (__all_sync(0xffffffffu, ok)) break;
							if (spins > 64) __nanosleep(32);
							++spins;
							if (spins == 64u) t0 = globaltimer_ns();
							const bool expired = (spins & 1023u) == 0u && W.poll_expired(t0, 8);
							if (__any_sync(0xffffffffu, expired)) {
								W.dead = true;
								break;
							}
						}
						if (W.dead) break;
						if (mine) (void)ld_acquire_gpu(ctr);
						__syncwarp();
						known = known + 32 < needw ? known + 32 : needw;
						polled = true;
					}
				}
				if (lane == 0 && !W.dead) {
					if (polled) asm volatile("fence.proxy.async;" ::: "memory");
					const uint32_t fb = full_bar(s, it & 1);
					if (p.dbg & 4) {
						mbar_arrive(fb);
					} else {
						mbar_arrive_expect_tx(fb, kHaloBytes);
						tma_load_4d(smem_base + s * kHaloBytes, min, fb, 0, x0 - 1, y0 - 1, b);
					}
				}
				__syncwarp();
			}
		}
	} else if (warp == kTapWarp) {
		// ===================== weight TMA: tap sequence (layer, kx, ky) through the ring ==========
		if (lane == 0) {
			Waiter W(p.status, TC_KERNEL_TRUNK);
			const int total = p.n_layers * 9;
			for (int seq = 0; seq < total; ++seq) {
				const int slot = seq % kTapSlots;
				const uint32_t n = static_cast<uint32_t>(seq / kTapSlots);
				const int l = seq / 9, t = seq % 9;
				const int kx = t / 3, pos = t % 3;
				const int ky = pos == 0 ? 1 : (pos == 1 ? 0 : 2);  // the order the weight writers need them
				W.layer = l;
				W.wait(tapempty_bar(slot), (n & 1u) ^ 1u, 9);
				if (W.dead) continue;
				mbar_arrive_expect_tx(tapfull_bar(slot), kTapBytes);
				tma_load_2d(tap_base + slot * kTapBytes, &maps.w, tapfull_bar(slot), 0, (l * 9 + ky * 3 + kx) * 64);
			}
		}
	} else if (warp >= 10 && warp < 14) {
		// ===================== weight writers: tap ring -> tensor memory =====================
		// thread m = quadrant * 32 + lane owns row m of every stacked slice: half h = m / 64 is
		// output row y + h, channel c = m % 64; group (kx, r) takes tap ky = r - h from it
		Waiter W(p.status, TC_KERNEL_TRUNK);
		const int quad = warp & 3;
		const int h = quad >> 1;
		const int c = (quad & 1) * 32 + lane;
		const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
		for (int l = 0; l < p.n_layers; ++l) {
			W.layer = l;
			for (int step = 0; step < kGroups; ++step) {
				// in the order the issuers release and need them: per kx the groups t = 1, 2, 0
				const int kx = step / 3, t = step % 3 == 2 ? 0 : step % 3 + 1;
				const int g = kx * 3 + t;
				const int ky = t == 0 ? (h ? 2 : 0) : t - h;
				const int pos = ky == 1 ? 0 : (ky == 0 ? 1 : 2);
				const int seq = l * 9 + kx * 3 + pos;
				const int slot = seq % kTapSlots;
				uint32_t v[4][8];
				W.wait(tapfull_bar(slot), static_cast<uint32_t>(seq / kTapSlots) & 1u, 12);
				{
					const uint32_t row = tap_base + slot * kTapBytes + static_cast<uint32_t>(c) * 128u;
#pragma unroll
					for (int j = 0; j < 4; ++j) {
#pragma unroll
						for (int e = 0; e < 2; ++e) {
							const uint32_t a = row + (((static_cast<uint32_t>(2 * j + e)) ^ (static_cast<uint32_t>(c) & 7u)) << 4);
							asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
							             : "=r"(v[j][4 * e]), "=r"(v[j][4 * e + 1]), "=r"(v[j][4 * e + 2]), "=r"(v[j][4 * e + 3])
							             : "r"(a)
							             : "memory");
						}
					}
				}
				if (l > 0) W.wait(wempty_bar(g), static_cast<uint32_t>((l - 1) & 1), 13);
				W.sync_warp();
				if (!W.dead && !(p.dbg & 32)) {
					tcgen05_fence_after();
#pragma unroll
					for (int j = 0; j < 4; ++j) tmem_st8(tmem_base + lane_base + kWCol + (g * 4 + j) * 8, v[j]);
					tmem_st_wait();
				}
				tcgen05_fence_before();
				__syncwarp();
				if (lane == 0 && !W.dead) {
					mbar_arrive(tapempty_bar(slot));
					mbar_arrive(wfull_bar(g));
				}
			}
		}
	} else if (warp == 14 || warp == 15) {
		// ===================== store + publish, one warp per epilogue group =====================
		if (lane == 0) {
			Waiter W(p.status, TC_KERNEL_TRUNK);
			const bool stall = p.status && *reinterpret_cast<volatile int *>(&p.status->inject) == TC_KERNEL_TRUNK;
			if (p.pdl) grid_dependency_wait();
			const int gi = warp - 14;
			const bool skip_publish = stall || (p.dbg & 2);
			unsigned int *pending = nullptr;  // counter of a stored but not yet published unit
			int it = 0;
			uint32_t n = 0;  // units of this group so far
			for (int l = 0; l < p.n_layers; ++l) {
				W.layer = l;
				const CUtensorMap *mout = &maps.tile[layer_out(l)];
				for (int k = 0; k < cnt; ++k, ++it) {
					if ((it & 1) != gi) continue;
					const int unit = first + k * grid;
					int b, y0, x0;
					decode(unit, b, y0, x0);
					W.wait(sready_bar(gi), n & 1u, 10);
					++n;
					if (W.dead) continue;
					if (!(p.dbg & 2)) tma_store_4d(mout, stage_base + gi * kStageTile, 0, x0, y0, b);
					asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
					mbar_arrive(sfree_bar(gi));
					// publish: the previous store has certainly completed once at most this one is pending; this
					// one is published right away only if the group's next unit is not already waiting
					// (no unit is ever left unpublished while this warp blocks on a barrier)
					if (pending) {
						asm volatile("cp.async.bulk.wait_group 1;" ::: "memory");
						asm volatile("fence.proxy.async;" ::: "memory");
						if (!skip_publish) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(pending) : "memory");
						pending = nullptr;
					}
					unsigned int *flag = p.flags + l * n_waves + k;
					if (mbar_test_wait(sready_bar(gi), n & 1u)) {
						pending = flag;  // next unit is ready: overlap this store's completion with the next store
					} else {
						asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
						asm volatile("fence.proxy.async;" ::: "memory");
						if (!skip_publish) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(flag) : "memory");
					}
				}
			}
			if (pending) {
				asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
				asm volatile("fence.proxy.async;" ::: "memory");
				if (!skip_publish) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(pending) : "memory");
			}
			__threadfence();
			const unsigned int old = atomicAdd(p.sync_counter, 1u);
			if (old == gridDim.x * 2u - 1u) {
				atomicExch(p.sync_counter, 0u);
				atomicAdd(p.sync_counter + 1, 1u);
			}
		}
	} else if (warp == 1 || warp == kIssuer2) {
		// ===================== MMA issuers (alternate units) =====================
		const int mi = warp == 1 ? 0 : 1;
		Waiter W(p.status, TC_KERNEL_TRUNK);
		const uint32_t idesc = make_idesc(kUnitW);
		const uint32_t b_hi = static_cast<uint32_t>(make_smem_desc(0, 1024u, 0) >> 32);
		const uint32_t lo_flags = 1u << 16;
		int it = 0;
		for (int l = 0; l < p.n_layers; ++l) {
			W.layer = l;
			for (int pos = 0; pos < cnt; ++pos, ++it) {
				const bool solo = (p.dbg & 1024) != 0;  // experiment: warp 1 issues every unit
				if (solo ? mi != 0 : (it & 1) != mi) continue;
				const int par = it & 1;
				const int as = it % kAccStages;
				const int s = it % kStages;
				// phases of the (stage, waiter) barriers: every 6th unit
				const uint32_t ph = static_cast<uint32_t>(it / 6) & 1u;
				// the accumulator stage was drained by the epilogue of unit it - 3 (the other group)
				if (it >= kAccStages) W.wait(tempty_bar(as, par), static_cast<uint32_t>((it - kAccStages) / 6) & 1u, 3);
				W.wait(full_bar(s, par), ph, 4);
				W.sync_warp();
				tcgen05_fence_after();
				const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * 64);
				const uint32_t b_lo = lo_flags | ((smem_base + s * kHaloBytes) >> 4);
				// the first two units of a layer (one per issuer) wait group by group for the new weights;
				// the last two release the groups one by one (wempty counts 2 arrivals)
				const bool fresh = pos < 2;
				const int releases = pos + 2 >= cnt ? (cnt == 1 ? 2 : 1) : 0;
				if (elect_one_sync()) {
#pragma unroll
					for (int st = 0; st < 12; ++st) {
						// per kx: input rows 1, 2 (full groups t = 1, 2), then rows 0 and 3 through the shared group t = 0
						const int kx = st >> 2, q4 = st & 3;
						const int r = q4 == 0 ? 1 : (q4 == 1 ? 2 : (q4 == 2 ? 0 : 3));
						const int g = kx * 3 + (q4 < 2 ? q4 + 1 : 0);
						if (fresh && q4 != 3) {
							W.wait(wfull_bar(g), static_cast<uint32_t>(l & 1), 2);
							tcgen05_fence_after();
						}
						if (W.dead) break;
						const uint32_t b_g = b_lo + static_cast<uint32_t>(r * kHaloW + kx) * 8u;
						const uint32_t a_g = tmem_base + kWCol + static_cast<uint32_t>(g * 4) * 8u;
#pragma unroll
						for (int j = 0; j < 4; ++j) {
							const uint64_t b_desc = (static_cast<uint64_t>(b_hi) << 32) | (b_g + j * 2u);
							const uint32_t acc = (st | j) != 0 ? 1u : 0u;
							if (p.dbg & 16) {
							} else if (r == 0) {
								umma_ts_f16_masked(d_tmem, a_g + j * 8u, b_desc, idesc, acc, 0u, 0xffffffffu);
							} else if (r == 3) {
								umma_ts_f16_masked(d_tmem, a_g + j * 8u, b_desc, idesc, acc, 0xffffffffu, 0u);
							} else {
								umma_ts_f16(d_tmem, a_g + j * 8u, b_desc, idesc, acc);
							}
						}
						if (q4 != 2) {  // the shared group is released after its second use
							if (releases >= 1) umma_commit(wempty_bar(g));
							if (releases == 2) umma_commit(wempty_bar(g));
						}
					}
					if (!W.dead) {
						umma_commit(empty_bar(s));
						umma_commit(tfull_bar(as, par));
					}
				}
				__syncwarp();
				W.sync_warp();
			}
		}
	} else {
		// ===================== epilogue: warps 2..9, group = (warp - 2) / 4 =====================
		const int q = warp & 3;  // TMEM lane quadrant
		const int gi = (warp - 2) >> 2;
		const int orow = q >> 1;                  // output row of the unit
		const int cq = (q & 1) * 32;              // first channel of this warp
		const uint32_t stile = stage_base + gi * kStageTile;
		const uint32_t rbuf = res_base + static_cast<uint32_t>(warp - 2) * kResBytes;
		Waiter W(p.status, TC_KERNEL_TRUNK);
		if (p.pdl) grid_dependency_wait();
		int it = 0;
		uint32_t n = 0;  // units of this group so far
		// stmatrix / ldmatrix row of this lane: matrix mi = lane / 8 -> (channel block mi & 1, pixel block pb + (mi >> 1))
		const int mi = lane >> 3, mr = lane & 7;
		// private shortcut buffer: 64 pixels x 64 bytes, 16-byte chunk index XORed with (pixel / 2) % 4 so that
		// the 8 rows of an ldmatrix tile (8 consecutive pixels) fall into different banks
		auto roff = [](uint32_t px, uint32_t sub) { return px * 64u + ((sub ^ ((px >> 1) & 3u)) << 4); };
		for (int l = 0; l < p.n_layers; ++l) {
			W.layer = l;
			const bool has_res = layer_res(l) >= 0;
			const bool use_res = has_res && !(p.dbg & 8);
			const __half *res_buf = has_res ? p.buffers[layer_res(l)] : nullptr;
			// this thread's channels: (cq + 16 * hb + lane / 4) and + 8
			float bias_r[4];
#pragma unroll
			for (int i = 0; i < 4; ++i) bias_r[i] = __ldg(p.bias + l * 64 + cq + 16 * (i >> 1) + (lane >> 2) + 8 * (i & 1));
			for (int k = 0; k < cnt; ++k, ++it) {
				if ((it & 1) != gi) continue;
				const int unit = first + k * grid;
				const int as = it % kAccStages;
				const uint32_t aph = static_cast<uint32_t>(it / 6) & 1u;
				if (use_res && !W.dead) {
					// shortcut: this warp's 64 pixels x 32 channels of output row `orow`, requested before
					// the accumulator wait (the buffer is private: free since this warp's previous unit)
					int b, y0, x0;
					decode(unit, b, y0, x0);
					const int y = y0 + orow;
					const __half *src_row = res_buf + ((static_cast<size_t>(b) * p.h + y) * p.w + x0) * static_cast<size_t>(p.cstride) + cq;
#pragma unroll
					for (int i = 0; i < 8; ++i) {
						const int chunk = i * 32 + lane;
						const int px = chunk >> 2, sub = chunk & 3;
						const bool inside = y < p.h && x0 + px < p.w;
						const __half *src = inside ? src_row + static_cast<size_t>(px) * p.cstride + sub * 8 : res_buf;
						cp_async_16(rbuf + roff(static_cast<uint32_t>(px), static_cast<uint32_t>(sub)), src, inside ? 16u : 0u);
					}
					asm volatile("cp.async.commit_group;" ::: "memory");
				}
				W.wait(tfull_bar(as, gi), aph, 5);
				tcgen05_fence_after();
				// drain the whole accumulator stage first and hand it back: the issuer <-> epilogue loop over
				// only two stages is the tightest loop of the kernel
				uint32_t r[2][32];
				__syncwarp();
				if (!(p.dbg & 128)) {
					tmem_ld_16x256b_x8(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * 64), r[0]);
					tmem_ld_16x256b_x8(tmem_base + (static_cast<uint32_t>(q * 32 + 16) << 16) + static_cast<uint32_t>(as * 64), r[1]);
					tmem_ld_wait();
				}
				tcgen05_fence_before();
				__syncwarp();
				W.sync_warp();
				if (lane == 0 && !W.dead) mbar_arrive(tempty_bar(as, gi ^ 1));  // its next user is the other issuer (unit it + 3)
				// the staging tile is free once the store of this group's previous unit has read it
				W.wait(sfree_bar(gi), (n & 1u) ^ 1u, 11);
				++n;
				if (use_res) {
					asm volatile("cp.async.wait_group 0;" ::: "memory");
					__syncwarp();
				}
				if (!(p.dbg & 64)) {
#pragma unroll
					for (int hb = 0; hb < 2; ++hb) {
						const float b_lo = bias_r[2 * hb], b_hi2 = bias_r[2 * hb + 1];
#pragma unroll
						for (int pb = 0; pb < 8; pb += 2) {
							const uint32_t px = static_cast<uint32_t>((pb + (mi >> 1)) * 8 + mr);
							const uint32_t sub = static_cast<uint32_t>(2 * hb + (mi & 1));
							uint32_t qq[4] = {0u, 0u, 0u, 0u};
							if (use_res) ldmatrix_x4_trans(rbuf + roff(px, sub), qq[0], qq[1], qq[2], qq[3]);
							uint32_t o[4];
#pragma unroll
							for (int i = 0; i < 4; ++i) {
								const int blk = pb + (i >> 1);
								const bool hi = (i & 1) != 0;
								float v0 = __uint_as_float(r[hb][blk * 4 + (hi ? 2 : 0)]) + (hi ? b_hi2 : b_lo);
								float v1 = __uint_as_float(r[hb][blk * 4 + (hi ? 3 : 1)]) + (hi ? b_hi2 : b_lo);
								if (use_res) {
									const float2 rr = __half22float2(*reinterpret_cast<const __half2 *>(&qq[i]));
									v0 += rr.x;
									v1 += rr.y;
								}
								if (p.act == ACT_RELU) {
									v0 = fmaxf(v0, 0.f);
									v1 = fmaxf(v1, 0.f);
								} else if (p.act == ACT_LRELU) {
									v0 = v0 >= 0.f ? v0 : v0 * p.slope;
									v1 = v1 >= 0.f ? v1 : v1 * p.slope;
								}
								const __half2 hh = __floats2half2_rn(v0, v1);
								o[i] = *reinterpret_cast<const uint32_t *>(&hh);
							}
							stmatrix_x4_trans(stile + swz(static_cast<uint32_t>(orow * kUnitW) + px, static_cast<uint32_t>(cq) + sub * 8u), o[0], o[1],
							    o[2], o[3]);
						}
					}
				}
				// generic-proxy smem writes -> visible to the TMA store (async proxy), then hand the tile over
				asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
				__syncwarp();
				W.sync_warp();
				if (lane == 0 && !W.dead) mbar_arrive(sready_bar(gi));
			}
		}
	}

	tcgen05_fence_before();
	__syncthreads();
	if (warp == 1) {
		tcgen05_fence_after();
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
	}
}

using EncodeTiledFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encodeTiledTS() {
	static EncodeTiledFn fn = nullptr;
	if (!fn) {
		void *p = nullptr;
		cudaDriverEntryPointQueryResult q;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
		    q != cudaDriverEntryPointSuccess) {
			return nullptr;
		}
		fn = reinterpret_cast<EncodeTiledFn>(p);
	}
	return fn;
}

}  // namespace

int trunk_ts_units(int h, int w) { return ((h + kUnitH - 1) / kUnitH) * ((w + kUnitW - 1) / kUnitW); }

cudaError_t trunk_ts_tc_prepare(const TrunkArgs &a, TrunkTcLaunch *out) {
	EncodeTiledFn encode = encodeTiledTS();
	if (!encode) return cudaErrorNotSupported;
	const int lead = a.lead_in ? 1 : 0;
	if (a.cstride != 64 || a.n_layers - lead < 2 || ((a.n_layers - lead) & 1)) return cudaErrorInvalidValue;
	static_assert(sizeof(TsParams) <= sizeof(out->params), "TrunkTcLaunch::params too small");
	static_assert(sizeof(TsMaps) <= sizeof(out->maps), "TrunkTcLaunch::maps too small");
	TsParams p{};
	p.batch = a.batch;
	p.h = a.h;
	p.w = a.w;
	p.units_x = (a.w + kUnitW - 1) / kUnitW;
	p.units_y = (a.h + kUnitH - 1) / kUnitH;
	p.total_units = a.batch * p.units_x * p.units_y;
	p.n_layers = a.n_layers;
	p.act = a.act;
	p.slope = a.slope;
	p.pdl = a.cooperative ? 0 : 1;
	p.bias = a.bias;
	p.sync_counter = a.sync_counter;
	p.flags = a.flags;
	if (!a.flags) return cudaErrorInvalidValue;
	for (int i = 0; i < 3; ++i) p.buffers[i] = static_cast<const __half *>(a.buffers[i]);
	p.cstride = a.cstride;
	p.lead = lead;
	p.dbg = a.debug_skip;
	TsMaps maps;
	std::memset(&maps, 0, sizeof(maps));
	cuuint32_t estr[4] = {1, 1, 1, 1};
	cuuint64_t dims[4] = {static_cast<cuuint64_t>(a.cstride), static_cast<cuuint64_t>(a.w), static_cast<cuuint64_t>(a.h),
	    static_cast<cuuint64_t>(a.batch)};
	cuuint64_t strides[3] = {static_cast<cuuint64_t>(a.cstride) * 2, static_cast<cuuint64_t>(a.w) * a.cstride * 2,
	    static_cast<cuuint64_t>(a.h) * a.w * a.cstride * 2};
	cuuint32_t hbox[4] = {64, kHaloW, kHaloH, 1};
	cuuint32_t tbox[4] = {64, kUnitW, kUnitH, 1};
	for (int i = 0; i < 3; ++i) {
		if (encode(&maps.in[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, a.buffers[i], dims, strides, hbox, estr,
		        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
		        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
		    encode(&maps.tile[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, a.buffers[i], dims, strides, tbox, estr,
		        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
		        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
			return cudaErrorInvalidValue;
		}
	}
	if (lead) {
		if (encode(&maps.in[3], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(a.lead_in), dims, strides, hbox, estr,
		        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
		        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
			return cudaErrorInvalidValue;
		}
	}
	{
		cuuint64_t wd[2] = {64, static_cast<cuuint64_t>(a.n_layers) * 9 * 64};
		cuuint64_t ws[1] = {128};
		cuuint32_t wb[2] = {64, 64};
		if (encode(&maps.w, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(a.weights), wd, ws, wb, estr,
		        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
		        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
			return cudaErrorInvalidValue;
		}
	}
	std::memcpy(out->maps, &maps, sizeof(maps));
	std::memcpy(out->params, &p, sizeof(p));
	cudaError_t attrErr =
	    cudaFuncSetAttribute(trunk_ts_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemTs));
	if (attrErr != cudaSuccess) return attrErr;
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	out->grid = p.total_units < sms ? p.total_units : sms;
	out->pair = 0;
	out->ts = 1;
	out->smem_bytes = kSmemTs;
	out->sync_counter = a.sync_counter;
	out->cooperative = a.cooperative;
	return cudaSuccess;
}

cudaError_t trunk_ts_tc_launch(const TrunkTcLaunch &l, TcStatus *status, cudaStream_t s) {
	TsMaps maps;
	TsParams p;
	std::memcpy(&maps, l.maps, sizeof(maps));
	std::memcpy(&p, l.params, sizeof(p));
	p.status = status;
	cudaLaunchConfig_t cfg{};
	cfg.gridDim = dim3(l.grid);
	cfg.blockDim = dim3(kThreadsTs);
	cfg.dynamicSmemBytes = l.smem_bytes;
	cfg.stream = s;
	cudaLaunchAttribute attr[2];
	int n = 0;
	if (p.pdl) {
		attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
		attr[n].val.programmaticStreamSerializationAllowed = 1;
		++n;
	}
	if (l.cooperative) {
		attr[n].id = cudaLaunchAttributeCooperative;
		attr[n].val.cooperative = 1;
		++n;
	}
	cfg.attrs = attr;
	cfg.numAttrs = n;
	return cudaLaunchKernelEx(&cfg, trunk_ts_tc_kernel, maps, p);
}

}  // namespace ju
