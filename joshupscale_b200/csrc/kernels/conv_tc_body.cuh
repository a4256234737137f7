// Device side of the tcgen05 implicit-GEMM convolution (see conv_tc.cu for the design notes):
// the per-layer pipeline as a device function shared by conv_tc_kernel (one launch per layer)
// and the persistent flow-net kernel (flow_df_tc.cu, all layers of the flow net in one launch).
#pragma once

#include <cuda.h>

#include "kernels.h"
#include "tc_common.cuh"

namespace ju {
namespace tcconv {

constexpr int kTileH = 16;  // pixel rows per M tile (= UMMA core-matrix groups)
constexpr int kTileW = 8;   // pixel columns per M tile (= rows per core-matrix group)
// warps: 0 TMA producer, 1 MMA issuer, then the epilogue warps, then a second producer and a
// second issuer that are only active in `dual` mode (two independent tile pipelines)
constexpr int kThreads = 256;       // 4 epilogue warps
constexpr int kThreadsWide = 384;   // 8 epilogue warps (every shared-memory epilogue mode)
constexpr int kMaxStages = 8;
constexpr uint32_t kSmemLimit = 227 * 1024;

struct TcParams {
	int batch, h, w;
	int tiles_x, tiles_y, n_tiles, total_tiles;
	int kb;      // Cin / 64
	int ks;      // 1 or 3
	int nt;      // N tile (32 or 64)
	int cout, cout_stride;
	int pitch;   // halo pitch in pixels
	int nbox;    // 1, or 3 = one 8-wide box per horizontal tap (canonical 1024B-aligned starts)
	int base_off_mode;
	int b_resident;
	int stages;
	int tma_epi;  // shared-memory epilogue mode: 0 direct, 1 fp16 N=64, 2 fp16 N=32, 3 fp32 N=32
	int pool;     // fuse MaxPool2D(2) into the epilogue (tma_epi 1/2, no residual)
	int pdl;      // launched with programmatic stream serialization
	int dual;     // even / odd tiles run through two independent producer + issuer pipelines
	int last_k16; // K steps of 16 channels in the last 64-channel block that are not all zero (1..4)
	uint32_t a_box_bytes, a_region_bytes, stage_bytes, b_slice_bytes;
	const float *bias;
	const __half *residual;
	void *out;
	int act;
	float slope;
	int out_f32;
	int shuffle2;
	TcStatus *status;
};

using namespace tc;

struct TileCoord {
	int b, y0, x0, n0;
};

__device__ __forceinline__ TileCoord decode_tile(const TcParams &p, int idx) {
	TileCoord t;
	int nt_idx = idx % p.n_tiles;
	int rest = idx / p.n_tiles;
	int tx = rest % p.tiles_x;
	rest /= p.tiles_x;
	int ty = rest % p.tiles_y;
	t.b = rest / p.tiles_y;
	t.y0 = ty * kTileH;
	t.x0 = tx * kTileW;
	t.n0 = nt_idx * p.nt;
	return t;
}

// Per-layer context of the persistent flow-net kernel (PERSIST = true), see conv_tc_body.
struct PersistLayer {
	uint32_t tmem_base;        // TMEM allocation of the whole kernel (128 columns)
	int item_begin, item_end;  // work items of this layer handled in this pass (a chunk of streams)
	// Completion counter of the layer this one reads (null: the input comes from an earlier kernel):
	// every CTA adds 1 (release) when its part is stored; complete at dep_want.
	const unsigned int *dep_counter;
	unsigned int dep_want;
	uint32_t bar_base;         // shared-memory address of the kernel's mbarrier block (kBarBytes)
	// The barriers are initialised ONCE per kernel and never again: re-initialising an mbarrier
	// shortly after an asynchronous arrive (tcgen05.commit, TMA complete_tx) leaves it in an
	// unpredictable phase (observed on B200: the raw words of freshly re-initialised barriers
	// differ in the phase bit).  Instead every thread carries the parity of the phases each
	// barrier has completed so far: bit s of stage_par for full/empty[s], bit a of acc_par for
	// tfull/tempty[a], w_par for the weight barrier; conv_tc_body XORs them into its waits and
	// advances them (arithmetically, from this CTA's item count) when the layer is done.
	uint32_t stage_par, acc_par, w_par;
};

// Wait until the previous layer of the persistent kernel is complete (all CTAs have published),
// then order the caller's later async-proxy (TMA) and generic reads after it.
__device__ __forceinline__ void persist_wait_layer(const PersistLayer &pl, Waiter &W, int code) {
	if (!pl.dep_counter || W.dead) return;
	unsigned int spins = 0;
	unsigned long long t0 = 0;
	while (static_cast<int>(ld_relaxed_gpu(pl.dep_counter) - pl.dep_want) < 0) {
		if (spins > 16) __nanosleep(64);
		++spins;
		if (spins == 16u) t0 = globaltimer_ns();
		if ((spins & 1023u) == 0u && W.poll_expired(t0, code)) return;
	}
	asm volatile("fence.acq_rel.gpu;" ::: "memory");
	asm volatile("fence.proxy.async;" ::: "memory");
}

constexpr uint32_t kBarBytes = 8u * (2 * kMaxStages + 10);
constexpr uint32_t kPersistTemptyCount = 8;  // arrivals per accumulator hand-back in the persistent kernel

// One convolution layer on the calling CTA: every role of the pipeline, for the work items
// [item_begin + blockIdx.x, item_end) strided by the grid.
//
// PERSIST = false: the whole kernel body of conv_tc_kernel (one launch per layer).
// PERSIST = true : one layer inside the persistent flow-net kernel (flow_df_tc.cu).  Differences:
//   * TMEM and the mbarriers are set up once by the caller (phase parities carry over, see
//     PersistLayer); the caller separates layers by a CTA-wide barrier: every accumulator has been
//     read by then (all MMAs and their operand fetches are complete) and every store leader has
//     waited for its last bulk store;
//   * instead of griddepcontrol the producers wait for the completion counter of the PREVIOUS
//     layer (persist_wait_layer) before their first activation load; the caller publishes this
//     layer's completion after its CTA-wide barrier;
//   * work items are a sub-range of the layer (a chunk of streams).
template <int KS, int EPI, bool PERSIST>
__device__ __forceinline__ void conv_tc_body(const CUtensorMap *map_a_p, const CUtensorMap *map_b_p,
    const CUtensorMap *map_c_p, const CUtensorMap *map_r_p, const TcParams &p, uint8_t *smem_raw,
    PersistLayer &pl, Waiter &W) {
	const CUtensorMap &map_a = *map_a_p;
	const CUtensorMap &map_b = *map_b_p;
	const CUtensorMap &map_c = *map_c_p;
	const CUtensorMap &map_r = *map_r_p;
	// SWIZZLE_128B operands need 1024-byte aligned tiles
	const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
	constexpr int taps = KS * KS;
	const uint32_t stages_bytes = static_cast<uint32_t>(p.stages) * p.stage_bytes;
	const uint32_t resb_base = smem_base + stages_bytes;
	const uint32_t resb_bytes = p.b_resident ? static_cast<uint32_t>(taps * p.kb) * p.b_slice_bytes : 0u;
	// epilogue staging (tma_epi): 2 output tiles + 2 residual tiles of 128 px x 128 B, bias
	constexpr int kNT = EPI == 1 ? 64 : 32;                       // channels per staged row
	constexpr uint32_t kRowB = EPI == 2 ? 64u : 128u;             // bytes per staged row
	constexpr int kChunks = static_cast<int>(kRowB / 16u);        // 16-byte chunks per row
	constexpr uint32_t kEpiTile = 128u * kRowB;
	const uint32_t epi_out_base = resb_base + resb_bytes;
	const uint32_t epi_res_base = epi_out_base + (p.tma_epi ? 2u * kEpiTile : 0u);
	const uint32_t bias_base = epi_res_base + ((p.tma_epi && p.residual) ? 2u * kEpiTile : 0u);
	// 8-byte aligned (all sizes are multiples of 256); the persistent kernel keeps its barriers
	// outside the per-layer layout, alternating between two blocks from layer to layer
	const uint32_t bar_base = PERSIST ? pl.bar_base : bias_base + 256u;
	auto full_bar = [&](int s) { return bar_base + 8u * s; };
	auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
	auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + s); };
	auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 2 + s); };
	const uint32_t w_bar = bar_base + 8u * (2 * kMaxStages + 4);
	const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages + 5);
	auto rfull_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 6 + s); };
	auto rempty_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 8 + s); };

	const int warp = threadIdx.x >> 5;
	const int lane = threadIdx.x & 31;
	const uint32_t tmem_cols = p.nt * 2 <= 64 ? 64u : 128u;
	const int item_begin = PERSIST ? pl.item_begin : 0;
	const int item_end = PERSIST ? pl.item_end : p.total_tiles;

	// phases completed before this layer (all zero in the one-launch-per-layer kernel)
	const uint32_t stage_par = PERSIST ? pl.stage_par : 0u;
	const uint32_t acc_par = PERSIST ? pl.acc_par : 0u;
	const uint32_t w_par = PERSIST ? pl.w_par : 0u;
	auto bs = [&](int s) { return (stage_par >> s) & 1u; };
	auto ba = [&](int a) { return (acc_par >> a) & 1u; };

	if (!PERSIST && warp == 0 && lane == 0) {
		for (int s = 0; s < p.stages; ++s) {
			mbar_init(full_bar(s), 1);
			mbar_init(empty_bar(s), 1);
		}
		for (int s = 0; s < 2; ++s) {
			mbar_init(tfull_bar(s), 1);
			mbar_init(tempty_bar(s), EPI == 1 ? 8 : 4);  // one arrival per epilogue warp
			mbar_init(rfull_bar(s), 1);
			mbar_init(rempty_bar(s), EPI == 1 ? 8 : 4);
		}
		mbar_init(w_bar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (!PERSIST && warp == 1) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
		             "r"(tmem_cols)
		             : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	if (!PERSIST) {
		tcgen05_fence_before();
		__syncthreads();
		tcgen05_fence_after();
	}
	uint32_t tmem_base;
	if (PERSIST) {
		tmem_base = pl.tmem_base;
	} else {
		asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
	}

	constexpr int pad = (KS - 1) / 2;

	// Programmatic dependent launch: everything above (barrier init, TMEM
	// allocation) and the weight loads below do not depend on the previous
	// kernel in the stream; activations are only touched after the wait.
	if (!PERSIST && p.pdl) grid_launch_dependents();

	constexpr int kEpiWarps = EPI != 0 ? 8 : 4;
	// Dual mode (kb == 1, resident weights, no residual, even stage count): tiles alternate
	// between two producer / issuer pairs.  Stage s = tile % stages and TMEM stage = tile & 1, so
	// each pair owns a disjoint half of the halo ring and one accumulator: two independent
	// single-producer / single-consumer pipelines.  One issuer's barrier round trips (~300 cycles
	// each through the shared-memory pipe the operand fetch saturates) then overlap the other
	// issuer's MMAs instead of draining the tensor pipe between tiles.
	const bool second = warp >= 2 + kEpiWarps;
	const int pipe = second ? 1 : 0;
	if (warp == 0 || (p.dual && warp == 2 + kEpiWarps)) {
		// ===================== TMA producer =====================
		if (lane == 0) {
			if (p.b_resident && !second) {
				mbar_arrive_expect_tx(w_bar, resb_bytes);
				for (int s = 0; s < taps * p.kb; ++s) {
					// slice s = tap * kb + kbi ; rows [s*cout, s*cout + nt)
					tma_load_2d(resb_base + s * p.b_slice_bytes, &map_b, w_bar, 0, s * p.cout);
				}
			}
			if (!PERSIST && p.pdl) grid_dependency_wait();
			bool layer_ready = !PERSIST;  // PERSIST: the previous layer's completion has been acquired
			// The residual tile of tile i is consumed by the epilogue one tile after
			// the MMAs of tile i start, so it is requested after the halo of tile
			// i+1: the halo ring never waits behind the (shallower) residual ring.
			auto load_residual = [&](int tc, int tile) {
				const TileCoord t = decode_tile(p, tile);
				const int rb = tc & 1;
				const uint32_t rph = (tc >> 1) & 1;
				if (!W.wait(rempty_bar(rb), rph ^ 1u, 6)) return;
				mbar_arrive_expect_tx(rfull_bar(rb), kEpiTile);
				tma_load_4d(epi_res_base + rb * kEpiTile, &map_r, rfull_bar(rb), t.n0, t.x0, t.y0, t.b);
			};
			const bool with_res = p.tma_epi && p.residual;
			int tcount = 0, prev_tile = -1;
			for (int tile = item_begin + blockIdx.x; tile < item_end; tile += gridDim.x, ++tcount) {
				if (p.dual && (tcount & 1) != pipe) continue;
				const TileCoord t = decode_tile(p, tile);
				if (PERSIST && !layer_ready) {
					persist_wait_layer(pl, W, 8);
					layer_ready = true;
				}
				for (int kbi = 0; kbi < p.kb; ++kbi) {
					const int it = tcount * p.kb + kbi;
					const int s = it % p.stages;
					const uint32_t ph = (it / p.stages) & 1;
					if (!W.wait(empty_bar(s), (ph ^ bs(s)) ^ 1u, 1)) continue;
					const uint32_t stage = smem_base + s * p.stage_bytes;
					const uint32_t bytes = p.a_box_bytes + (p.b_resident ? 0u : taps * p.b_slice_bytes);
					mbar_arrive_expect_tx(full_bar(s), bytes);
					tma_load_4d(stage, &map_a, full_bar(s), kbi * 64, t.x0 - pad, t.y0 - pad, t.b);
					if (!p.b_resident) {
						for (int tap = 0; tap < taps; ++tap) {
							tma_load_2d(stage + p.a_region_bytes + tap * p.b_slice_bytes, &map_b, full_bar(s), 0,
							    (tap * p.kb + kbi) * p.cout + t.n0);
						}
					}
				}
				if (with_res && prev_tile >= 0) load_residual(tcount - 1, prev_tile);
				prev_tile = tile;
			}
			if (with_res && prev_tile >= 0) load_residual(tcount - 1, prev_tile);
		}
	} else if (warp == 1 || (p.dual && warp == 3 + kEpiWarps)) {
		// ===================== MMA issuer =====================
		// One thread issues every tcgen05.mma of the CTA, so its instruction
		// stream is the critical path: descriptors are split into a constant high
		// word and a 32-bit low word (start address >> 4) that only needs one add
		// per MMA; tap / k offsets are compile-time constants after unrolling.
		{
			// the whole warp stays converged; one elected lane issues the MMAs and commits
			const uint32_t idesc = make_idesc(p.nt);
			const uint32_t a_sbo = static_cast<uint32_t>(p.pitch) * 128u;
			const uint32_t a_hi = static_cast<uint32_t>(make_smem_desc(0, a_sbo, 0) >> 32);
			const uint32_t b_hi = static_cast<uint32_t>(make_smem_desc(0, 1024u, 0) >> 32);
			const uint32_t lo_flags = 1u << 16;            // LBO field (unused for swizzled K-major)
			const uint32_t row_off = static_cast<uint32_t>(p.pitch) * 8u;  // one halo row, in 16-byte units
			const uint32_t b_slice16 = p.b_slice_bytes >> 4;
			if (p.b_resident) {
				W.wait(w_bar, w_par, 2);
			}
			int tcount = 0;
			for (int tile = item_begin + blockIdx.x; tile < item_end; tile += gridDim.x, ++tcount) {
				if (p.dual && (tcount & 1) != pipe) continue;
				const int as = tcount & 1;
				const uint32_t aph = (tcount >> 1) & 1;
				W.wait(tempty_bar(as), (aph ^ ba(as)) ^ 1u, 3);
				tcgen05_fence_after();
				const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * p.nt);
				for (int kbi = 0; kbi < p.kb; ++kbi) {
					const int it = tcount * p.kb + kbi;
					const int s = it % p.stages;
					const uint32_t ph = (it / p.stages) & 1;
					W.wait(full_bar(s), ph ^ bs(s), 4);
					W.sync_warp();
					if (W.dead) continue;  // aborted frame: nothing is issued or committed any more
					tcgen05_fence_after();
					const uint32_t stage = smem_base + s * p.stage_bytes;
					const uint32_t a_lo = lo_flags | (stage >> 4);
					const uint32_t b_lo = lo_flags | ((p.b_resident ? resb_base + kbi * p.b_slice_bytes
					                                               : stage + p.a_region_bytes) >> 4);
					const uint32_t b_tap16 = p.b_resident ? b_slice16 * p.kb : b_slice16;
					uint32_t first = kbi == 0 ? 0u : 1u;
					// channels beyond the layer's real Cin are zero in both operands: skip those K steps
					const int nk16 = kbi == p.kb - 1 ? p.last_k16 : 4;
					if (elect_one_sync()) {
#pragma unroll
					for (int tap = 0; tap < taps; ++tap) {
						const uint32_t a_tap = a_lo + (tap / KS) * row_off + (tap % KS) * 8u;
						const uint32_t b_tap = b_lo + tap * b_tap16;
#pragma unroll
						for (int k16 = 0; k16 < 4; ++k16) {
							if (k16 < nk16) {
								const uint64_t a_desc = (static_cast<uint64_t>(a_hi) << 32) | (a_tap + k16 * 2u);
								const uint64_t b_desc = (static_cast<uint64_t>(b_hi) << 32) | (b_tap + k16 * 2u);
								umma_f16(d_tmem, a_desc, b_desc, idesc, (tap | k16) != 0 ? 1u : first);
							}
						}
					}
					umma_commit(empty_bar(s));  // frees the stage when these MMAs have read it
					if (kbi == p.kb - 1) umma_commit(tfull_bar(as));  // accumulator complete -> epilogue
					}
					__syncwarp();
				}
			}
		}
	} else if (warp >= 2 && warp < 2 + kEpiWarps) {
		// ===================== epilogue (warps 2..5 or 2..9) =====================
		const int q = warp & 3;  // TMEM lane quarter this warp may access
		const int row = q * 32 + lane;
		const int cpp = p.shuffle2 ? p.cout / 4 : p.cout;  // channels per output pixel
		// N = 64 (EPI 1): the two warp quartets split the channels of every tile.  N = 32 (EPI 2, 3):
		// they form two independent groups that take even / odd tiles, each with its own TMEM
		// stage, staging tile, named barrier and bulk-store group - two tiles drain at once.
		constexpr bool kGrouped = EPI == 2 || EPI == 3;
		const int group = kGrouped ? ((warp - 2) >> 2) : 0;
		const int etid = threadIdx.x - 64 - group * 128;  // index within this warp's epilogue group
		const int half = EPI == 1 ? ((warp - 2) >> 2) : 0;  // 32-channel half handled by this warp
		bool first_tile = true;
		auto group_barrier = [&]() {
			if constexpr (EPI == 1) {
				epilogue_barrier<256>();
			} else {
				asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");
			}
		};
		uint8_t *smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));  // generic pointer to the aligned base
		float bias_reg[32];
		if (!PERSIST && p.pdl) grid_dependency_wait();
		int tcount = 0;
		for (int tile = item_begin + blockIdx.x; tile < item_end; tile += gridDim.x, ++tcount) {
			const TileCoord t = decode_tile(p, tile);
			if (kGrouped && (tcount & 1) != group) continue;
			const int as = tcount & 1;
			const uint32_t aph = (tcount >> 1) & 1;
			if constexpr (EPI != 0) {
				// ---- shared-memory epilogue: every global access is a TMA bulk copy ----
				// A thread owns 32 channels of one pixel: the fp16 N=64 mode runs 8 epilogue
				// warps (two per TMEM lane quarter, one per 32-channel half), the N=32 modes 4.
				// Pixel `row` is one row (kRowB bytes) of the swizzled staging tiles; 16-byte
				// chunk c of row r lives at r*kRowB + ((c ^ sw(r)) << 4) with sw = r&7 (128B
				// swizzle) or (r>>1)&3 (64B swizzle), so row-per-thread LDS.128 / STS.128 are
				// bank-conflict free.  Latency is hidden by ILP inside the thread: bias stays
				// in registers across tiles, residual and accumulator are fetched up-front,
				// TMEM is released before the math.
				constexpr int kTC = EPI == 3 ? 8 : 4;  // 16-byte chunks produced per thread
				const int coff = EPI == 1 ? half * 4 : 0;  // first chunk of this thread's channels
				if (first_tile || p.n_tiles > 1) {
#pragma unroll
					for (int c = 0; c < 32; ++c) bias_reg[c] = p.bias ? __ldg(p.bias + t.n0 + half * 32 + c) : 0.f;
					first_tile = false;
				}
				if (etid == 0 && tcount >= 2) {
					// the bulk store that read staging[as] two tiles ago must have drained (a grouped
					// leader only has its own group's stores outstanding: that one is its latest)
					if (kGrouped) {
						asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
					} else {
						asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
					}
				}
				const uint32_t sw = EPI == 2 ? static_cast<uint32_t>((row >> 1) & 3) : static_cast<uint32_t>(row & 7);
				uint4 res[4];
				if (EPI != 3 && p.residual) {
					W.wait(rfull_bar(as), aph, 7);
					const uint4 *res_row = reinterpret_cast<const uint4 *>(
					    smem_gen + (epi_res_base - smem_base) + as * kEpiTile + row * kRowB);
#pragma unroll
					for (int c = 0; c < 4; ++c) res[c] = res_row[(coff + c) ^ sw];
				}
				W.wait(tfull_bar(as), aph ^ ba(as), 5);
				tcgen05_fence_after();
				uint32_t acc[32];
				const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
				                       static_cast<uint32_t>(as * p.nt + half * 32);
				__syncwarp();
				tmem_ld32(taddr, acc);
				tmem_ld_wait();
				// TMEM and residual tile are in registers -> hand both back early
				tcgen05_fence_before();
				__syncwarp();
				W.sync_warp();
				if (lane == 0 && !W.dead) {
					if (PERSIST && kGrouped) {
						// the persistent kernel's barrier always expects 8 arrivals: 4 warps x 2
						asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], 2;" ::"r"(tempty_bar(as)) : "memory");
					} else {
						mbar_arrive(tempty_bar(as));
					}
					if (EPI != 3 && p.residual) mbar_arrive(rempty_bar(as));
				}
				group_barrier();  // staging[as] free (wait_group.read above)
				float v[32];
#pragma unroll
				for (int c = 0; c < 32; ++c) v[c] = __uint_as_float(acc[c]) + bias_reg[c];
				if (EPI != 3 && p.residual) {
#pragma unroll
					for (int c = 0; c < 4; ++c) {
						const __half2 *h2 = reinterpret_cast<const __half2 *>(&res[c]);
#pragma unroll
						for (int e = 0; e < 4; ++e) {
							const float2 f = __half22float2(h2[e]);
							v[c * 8 + e * 2] += f.x;
							v[c * 8 + e * 2 + 1] += f.y;
						}
					}
				}
				if (p.act == ACT_RELU) {
#pragma unroll
					for (int c = 0; c < 32; ++c) v[c] = fmaxf(v[c], 0.f);
				} else if (p.act == ACT_LRELU) {
#pragma unroll
					for (int c = 0; c < 32; ++c) v[c] = v[c] >= 0.f ? v[c] : v[c] * p.slope;
				}
				uint8_t *out_tile = smem_gen + (epi_out_base - smem_base) + as * kEpiTile;
				if constexpr (EPI == 3) {
					// fp32 rows: 32 floats = 8 chunks of 4
					uint4 *out_row = reinterpret_cast<uint4 *>(out_tile + row * kRowB);
#pragma unroll
					for (int c = 0; c < kTC; ++c) {
						uint4 o;
						o.x = __float_as_uint(v[c * 4 + 0]);
						o.y = __float_as_uint(v[c * 4 + 1]);
						o.z = __float_as_uint(v[c * 4 + 2]);
						o.w = __float_as_uint(v[c * 4 + 3]);
						out_row[c ^ sw] = o;
					}
				} else {
					uint32_t packed[16];
#pragma unroll
					for (int c = 0; c < 16; ++c) {
						__half2 h = __floats2half2_rn(v[2 * c], v[2 * c + 1]);
						packed[c] = *reinterpret_cast<uint32_t *>(&h);
					}
					if (p.pool) {
						// MaxPool2D(2): lane = (ty%4)*8 + tx, so the 2x2 window partners are
						// lane^1 (x) and lane^8 (y); max on packed fp16 pairs is exact
#pragma unroll
						for (int c = 0; c < 16; ++c) {
							__half2 h = *reinterpret_cast<__half2 *>(&packed[c]);
							uint32_t o1 = __shfl_xor_sync(0xffffffffu, packed[c], 1);
							h = __hmax2(h, *reinterpret_cast<__half2 *>(&o1));
							uint32_t hv = *reinterpret_cast<uint32_t *>(&h);
							uint32_t o8 = __shfl_xor_sync(0xffffffffu, hv, 8);
							h = __hmax2(h, *reinterpret_cast<__half2 *>(&o8));
							packed[c] = *reinterpret_cast<uint32_t *>(&h);
						}
						if (((row & 1) | ((row >> 3) & 1)) == 0) {
							// pooled pixel (ty/2, tx/2) of the 8x4 pooled tile
							const int pr = (row >> 4) * 4 + ((row & 7) >> 1);
							const uint32_t psw = EPI == 2 ? static_cast<uint32_t>((pr >> 1) & 3) : static_cast<uint32_t>(pr & 7);
							uint4 *out_row = reinterpret_cast<uint4 *>(out_tile + pr * kRowB);
#pragma unroll
							for (int c = 0; c < kTC; ++c) {
								out_row[(coff + c) ^ psw] = make_uint4(packed[c * 4], packed[c * 4 + 1], packed[c * 4 + 2], packed[c * 4 + 3]);
							}
						}
					} else {
						uint4 *out_row = reinterpret_cast<uint4 *>(out_tile + row * kRowB);
#pragma unroll
						for (int c = 0; c < kTC; ++c) {
							out_row[(coff + c) ^ sw] = make_uint4(packed[c * 4], packed[c * 4 + 1], packed[c * 4 + 2], packed[c * 4 + 3]);
						}
					}
				}
				// make the generic-proxy smem writes visible to the TMA (async proxy)
				asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
				group_barrier();
				if (etid == 0 && !W.dead) {
					// out-of-range rows/columns of ragged tiles are clipped by the TMA store
					if (p.pool) {
						tma_store_4d(&map_c, epi_out_base + as * kEpiTile, t.n0, t.x0 >> 1, t.y0 >> 1, t.b);
					} else {
						tma_store_4d(&map_c, epi_out_base + as * kEpiTile, t.n0, t.x0, t.y0, t.b);
					}
				}
				continue;
			}
			if constexpr (EPI == 0) {
			W.wait(tfull_bar(as), aph ^ ba(as), 5);
			tcgen05_fence_after();
			const int y = t.y0 + (row >> 3), x = t.x0 + (row & 7);
			const bool valid = y < p.h && x < p.w;
			for (int half = 0; half < p.nt / 32; ++half) {
				uint32_t acc[32];
				const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
				                       static_cast<uint32_t>(as * p.nt + half * 32);
				__syncwarp();  // tcgen05.ld is .sync.aligned: the warp must be converged
				tmem_ld32(taddr, acc);
				tmem_ld_wait();
				if (valid) {
				const int co = t.n0 + half * 32;  // first of 32 consecutive output channels
				size_t opix;
				int oc;
				if (p.shuffle2) {
					const int sub = co / cpp;
					oc = co % cpp;
					opix = (static_cast<size_t>(t.b) * 2 * p.h + 2 * y + (sub >> 1)) * 2 * p.w + 2 * x + (sub & 1);
				} else {
					oc = co;
					opix = (static_cast<size_t>(t.b) * p.h + y) * p.w + x;
				}
				float v[32];
#pragma unroll
				for (int c = 0; c < 32; ++c) v[c] = __uint_as_float(acc[c]);
				if (p.bias) {
					const float4 *bp = reinterpret_cast<const float4 *>(p.bias + co);
#pragma unroll
					for (int c4 = 0; c4 < 8; ++c4) {
						const float4 bv = __ldg(bp + c4);
						v[c4 * 4 + 0] += bv.x;
						v[c4 * 4 + 1] += bv.y;
						v[c4 * 4 + 2] += bv.z;
						v[c4 * 4 + 3] += bv.w;
					}
				}
				if (p.residual) {
					const uint4 *rp = reinterpret_cast<const uint4 *>(p.residual + opix * p.cout_stride + oc);
#pragma unroll
					for (int c8 = 0; c8 < 4; ++c8) {
						const uint4 rv = __ldg(rp + c8);
						const __half2 *h2 = reinterpret_cast<const __half2 *>(&rv);
#pragma unroll
						for (int e = 0; e < 4; ++e) {
							const float2 f = __half22float2(h2[e]);
							v[c8 * 8 + e * 2] += f.x;
							v[c8 * 8 + e * 2 + 1] += f.y;
						}
					}
				}
				if (p.act == ACT_RELU) {
#pragma unroll
					for (int c = 0; c < 32; ++c) v[c] = fmaxf(v[c], 0.f);
				} else if (p.act == ACT_LRELU) {
#pragma unroll
					for (int c = 0; c < 32; ++c) v[c] = v[c] >= 0.f ? v[c] : v[c] * p.slope;
				}
				if (p.out_f32) {
					float4 *op = reinterpret_cast<float4 *>(static_cast<float *>(p.out) + opix * p.cout_stride + oc);
#pragma unroll
					for (int c4 = 0; c4 < 8; ++c4) {
						op[c4] = make_float4(v[c4 * 4], v[c4 * 4 + 1], v[c4 * 4 + 2], v[c4 * 4 + 3]);
					}
				} else {
					uint4 *op = reinterpret_cast<uint4 *>(static_cast<__half *>(p.out) + opix * p.cout_stride + oc);
#pragma unroll
					for (int c8 = 0; c8 < 4; ++c8) {
						uint4 o;
						__half2 h0 = __floats2half2_rn(v[c8 * 8 + 0], v[c8 * 8 + 1]);
						__half2 h1 = __floats2half2_rn(v[c8 * 8 + 2], v[c8 * 8 + 3]);
						__half2 h2 = __floats2half2_rn(v[c8 * 8 + 4], v[c8 * 8 + 5]);
						__half2 h3 = __floats2half2_rn(v[c8 * 8 + 6], v[c8 * 8 + 7]);
						o.x = *reinterpret_cast<uint32_t *>(&h0);
						o.y = *reinterpret_cast<uint32_t *>(&h1);
						o.z = *reinterpret_cast<uint32_t *>(&h2);
						o.w = *reinterpret_cast<uint32_t *>(&h3);
						op[c8] = o;
					}
				}
				}  // valid
			}
			__syncwarp();
			// all of this warp's TMEM reads are complete (wait::ld) -> release the stage
			tcgen05_fence_before();
			__syncwarp();
			W.sync_warp();
			if (lane == 0 && !W.dead) mbar_arrive(tempty_bar(as));
			}  // EPI == 0
		}
		if (PERSIST && EPI != 0 && etid == 0) {
			// this leader's last bulk stores must be complete before the caller publishes the layer;
			// the proxy fence orders them before that (generic-proxy) release
			asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
			asm volatile("fence.proxy.async;" ::: "memory");
		}
	}

	if (PERSIST) {
		// advance the carried phase parities by what this CTA did in this layer (same for every thread)
		const int first = item_begin + static_cast<int>(blockIdx.x);
		const int n_items = first < item_end ? (item_end - first + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x) : 0;
		const int n_it = n_items * p.kb;
		for (int s = 0; s < p.stages; ++s) {
			const int uses = n_it > s ? (n_it - s + p.stages - 1) / p.stages : 0;
			if (uses & 1) pl.stage_par ^= 1u << s;
		}
		if (((n_items + 1) / 2) & 1) pl.acc_par ^= 1u;
		if ((n_items / 2) & 1) pl.acc_par ^= 2u;
		if (p.b_resident) pl.w_par ^= 1u;
		return;  // the caller's next layer (or its teardown) starts with a CTA-wide barrier
	}
	if (EPI != 0 && (threadIdx.x == 64 || ((EPI == 2 || EPI == 3) && threadIdx.x == 64 + 128))) {
		asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
	}
	tcgen05_fence_before();
	__syncthreads();
	if (warp == 1) {
		tcgen05_fence_after();
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols)
		             : "memory");
	}
}

}  // namespace tcconv
}  // namespace ju
