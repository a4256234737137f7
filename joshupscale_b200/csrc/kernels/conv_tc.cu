// Implicit-GEMM 3x3 / 1x1 convolution on the 5th-generation tensor cores
// (tcgen05.mma, accumulators in TMEM, operands staged by TMA) for sm_100a.
//
// Replaces, inside the reference's TensorRT engine, layers.Conv2D(3x3/1x1,
// SAME) + BatchNormalization + activation (+ Add) - the ResBlock stack that is
// 91% of the frame's MACs (scripts/training/models.py:193-254, 531-550), the
// flow net's convs (models.py:257-331, 377-479) and Conv2DTranspose(k2,s2) as
// a 1x1 GEMM with a pixel-shuffle store (models.py:559-572).
//
// GEMM view: D[128 pixels, N=Cout tile] += A[128 pixels, 64 ch] * B[64 ch, N],
// looped over the 9 taps and Cin/64 channel blocks, fp16 operands, fp32
// accumulation in TMEM.
//
// Data movement (the B200-specific part):
//  * An M tile is a 16-row x 8-column pixel patch.  One TMA box load brings its
//    (16+2) x (8+2) halo of 64-channel pixels (128 B each, 128B-swizzled) into
//    shared memory ONCE; all 9 taps are then expressed as UMMA shared-memory
//    descriptors into that same halo tile: an 8-row core-matrix group is 8
//    horizontally adjacent pixels (8 x 128 B contiguous), consecutive groups
//    are consecutive image rows (stride = halo pitch), and a tap (dy, dx) only
//    shifts the descriptor start address by (dy*pitch + dx) * 128 B.  L2->SMEM
//    traffic is 1.4x the activation tensor instead of 9x (im2col-style loads).
//  * SAME zero padding and ragged right/bottom edges come from TMA out-of-bounds
//    zero fill (negative / past-the-end box coordinates); masked in the store.
//  * Weights (B operand) stay resident in shared memory for the whole
//    persistent CTA when they fit (64->64: 72 KB), else stream with the halo.
//  * Warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (one elected
//    thread) + TMEM allocator, warps 2-5 = epilogue (tcgen05.ld -> +bias
//    [folded BN] -> +residual -> ReLU/LeakyReLU -> fp16/fp32 store); two TMEM
//    accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <cuda.h>

#include <cstring>

#include "kernels.h"
#include "tc_common.cuh"

namespace ju {

namespace {

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

template <int KS, int EPI>
__global__ void __launch_bounds__(EPI != 0 ? kThreadsWide : kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
    const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_r,
    const TcParams p) {
	extern __shared__ uint8_t smem_raw[];
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
	const uint32_t bar_base = bias_base + 256u;  // 8-byte aligned (all sizes are multiples of 256)
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

	if (warp == 0 && lane == 0) {
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
	if (warp == 1) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
		             "r"(tmem_cols)
		             : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	tcgen05_fence_before();
	__syncthreads();
	tcgen05_fence_after();
	uint32_t tmem_base;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

	constexpr int pad = (KS - 1) / 2;

	// Programmatic dependent launch: everything above (barrier init, TMEM
	// allocation) and the weight loads below do not depend on the previous
	// kernel in the stream; activations are only touched after the wait.
	if (p.pdl) grid_launch_dependents();

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
			Waiter W(p.status, TC_KERNEL_CONV);
			if (p.b_resident && !second) {
				mbar_arrive_expect_tx(w_bar, resb_bytes);
				for (int s = 0; s < taps * p.kb; ++s) {
					// slice s = tap * kb + kbi ; rows [s*cout, s*cout + nt)
					tma_load_2d(resb_base + s * p.b_slice_bytes, &map_b, w_bar, 0, s * p.cout);
				}
			}
			if (p.pdl) grid_dependency_wait();
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
			for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcount) {
				if (p.dual && (tcount & 1) != pipe) continue;
				const TileCoord t = decode_tile(p, tile);
				for (int kbi = 0; kbi < p.kb; ++kbi) {
					const int it = tcount * p.kb + kbi;
					const int s = it % p.stages;
					const uint32_t ph = (it / p.stages) & 1;
					if (!W.wait(empty_bar(s), ph ^ 1u, 1)) continue;
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
			Waiter W(p.status, TC_KERNEL_CONV);
			if (p.b_resident) {
				W.wait(w_bar, 0, 2);
			}
			int tcount = 0;
			for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcount) {
				if (p.dual && (tcount & 1) != pipe) continue;
				const int as = tcount & 1;
				const uint32_t aph = (tcount >> 1) & 1;
				W.wait(tempty_bar(as), aph ^ 1u, 3);
				tcgen05_fence_after();
				const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * p.nt);
				for (int kbi = 0; kbi < p.kb; ++kbi) {
					const int it = tcount * p.kb + kbi;
					const int s = it % p.stages;
					const uint32_t ph = (it / p.stages) & 1;
					W.wait(full_bar(s), ph, 4);
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
		Waiter W(p.status, TC_KERNEL_CONV);
		if (p.pdl) grid_dependency_wait();
		int tcount = 0;
		for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcount) {
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
				W.wait(tfull_bar(as), aph, 5);
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
					mbar_arrive(tempty_bar(as));
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
			W.wait(tfull_bar(as), aph, 5);
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

// ---- host side --------------------------------------------------------------

using EncodeTiledFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encodeTiled() {
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

// process defaults (ju_set_option); every engine takes a copy when it is created
ConvTcOptions g_TcDefaults{0, 1, 1, 1};

// the kernel instance a prepared launch runs
const void *conv_tc_function(int ks, int tma_epi) {
	if (ks == 3) {
		switch (tma_epi) {
		case 1: return reinterpret_cast<const void *>(conv_tc_kernel<3, 1>);
		case 2: return reinterpret_cast<const void *>(conv_tc_kernel<3, 2>);
		case 3: return reinterpret_cast<const void *>(conv_tc_kernel<3, 3>);
		default: return reinterpret_cast<const void *>(conv_tc_kernel<3, 0>);
		}
	}
	switch (tma_epi) {
	case 1: return reinterpret_cast<const void *>(conv_tc_kernel<1, 1>);
	case 2: return reinterpret_cast<const void *>(conv_tc_kernel<1, 2>);
	case 3: return reinterpret_cast<const void *>(conv_tc_kernel<1, 3>);
	default: return reinterpret_cast<const void *>(conv_tc_kernel<1, 0>);
	}
}

}  // namespace

ConvTcOptions &conv_tc_default_options() { return g_TcDefaults; }

bool conv_tc_supported(const ConvArgs &a) {
	if (a.ksize != 1 && a.ksize != 3) return false;
	if (a.cin_stride % 64 || a.cin % 64 || a.cin > a.cin_stride) return false;
	if (a.cout % 32) return false;
	if (a.cout > 64 && a.cout % 64) return false;
	if (a.cout_stride % 8) return false;
	if (a.shuffle2 && (a.cout / 4) % 32) return false;
	return true;
}

size_t conv_tc_weight_bytes(int ksize, int cin_padded, int cout) {
	return static_cast<size_t>(ksize) * ksize * cin_padded * cout * sizeof(__half);
}

// kernel: Keras (kh, kw, Cin, Cout) fp32 -> [tap][kb][cout][64] fp16 (K-major B slices)
void conv_tc_pack_weights(const float *kernel, const float *scale, int ksize, int cin, int cin_padded,
    int cout, __half *dst) {
	std::memset(dst, 0, conv_tc_weight_bytes(ksize, cin_padded, cout));
	const int kb = cin_padded / 64;
	for (int tap = 0; tap < ksize * ksize; ++tap) {
		for (int c = 0; c < cin; ++c) {
			for (int o = 0; o < cout; ++o) {
				float v = kernel[(static_cast<size_t>(tap) * cin + c) * cout + o];
				if (scale) v = v * scale[o];
				const size_t idx = ((static_cast<size_t>(tap) * kb + c / 64) * cout + o) * 64 + (c % 64);
				dst[idx] = __float2half_rn(v);
			}
		}
	}
}

cudaError_t conv_tc_prepare(const ConvArgs &a, const ConvTcOptions &opt, ConvTcLaunch *out) {
	const int variant = opt.variant;
	if (!conv_tc_supported(a)) return cudaErrorInvalidValue;
	EncodeTiledFn encode = encodeTiled();
	if (!encode) return cudaErrorNotSupported;
	static_assert(sizeof(TcParams) <= sizeof(out->params), "ConvTcLaunch::params too small");
	static_assert(sizeof(CUtensorMap) == 128, "unexpected CUtensorMap size");
	TcParams p{};
	p.batch = a.batch;
	p.h = a.h;
	p.w = a.w;
	p.tiles_x = (a.w + kTileW - 1) / kTileW;
	p.tiles_y = (a.h + kTileH - 1) / kTileH;
	p.nt = a.cout >= 64 ? 64 : a.cout;
	p.n_tiles = a.cout / p.nt;
	p.total_tiles = a.batch * p.tiles_x * p.tiles_y * p.n_tiles;
	p.kb = a.cin / 64;
	p.last_k16 = 4;
	if (a.cin_live > 0 && a.cin_live <= a.cin) {
		const int live = a.cin_live - 64 * (p.kb - 1);
		if (live >= 1) p.last_k16 = (live + 15) / 16;  // a fully dead last block is left alone (not expected)
	}
	p.ks = a.ksize;
	p.cout = a.cout;
	p.cout_stride = a.cout_stride;
	const int taps = a.ksize * a.ksize;
	const int halo_h = kTileH + a.ksize - 1;
	// Halo layout.  Probed on B200 (profiles/r01_tc_probe_variants.json): the
	// 128B swizzle XOR is a function of the absolute shared-memory address
	// bits for both TMA and UMMA, so descriptor start addresses may be any
	// multiple of 128 B inside a 1024B-aligned TMA tile, SBO need not be a
	// multiple of 1024, and the descriptor base_offset field must stay 0
	// (setting it to (addr>>7)&7 produces wrong results).  variant 1 keeps a
	// 16-pixel pitch (SBO = 2048) as a cross-check layout.
	p.base_off_mode = 0;
	p.nbox = 1;
	if (a.ksize == 1) {
		p.pitch = kTileW;
	} else if ((variant & 3) == 1) {
		p.pitch = 16;
	} else {
		p.pitch = kTileW + 2;
	}
	p.a_box_bytes = static_cast<uint32_t>(halo_h * p.pitch * 128);
	p.a_region_bytes = (p.nbox * p.a_box_bytes + 1023u) & ~1023u;
	p.b_slice_bytes = static_cast<uint32_t>(p.nt * 128);
	const uint32_t all_b = static_cast<uint32_t>(taps * p.kb) * p.b_slice_bytes;
	p.b_resident = (p.n_tiles == 1 && all_b <= 96 * 1024) ? 1 : 0;
	p.stage_bytes = p.a_region_bytes + (p.b_resident ? 0u : taps * p.b_slice_bytes);
	// shared-memory epilogue (TMA residual load + TMA store) for the common
	// fp16, 64-channel-tile, non-shuffled case
	p.tma_epi = 0;
	if (opt.tma_epilogue && !a.shuffle2 && a.cout_stride % 8 == 0) {
		if (!a.out_f32 && p.nt == 64) p.tma_epi = 1;
		else if (!a.out_f32 && p.nt == 32) p.tma_epi = 2;
		else if (a.out_f32 && p.nt == 32 && !a.residual) p.tma_epi = 3;
	}
	p.pool = 0;
	if (a.pool) {
		if ((p.tma_epi != 1 && p.tma_epi != 2) || a.residual || (a.h & 1) || (a.w & 1)) return cudaErrorInvalidValue;
		p.pool = 1;
	}
	auto epiBytes = [&]() -> uint32_t {
		if (!p.tma_epi) return 0u;
		const uint32_t tile = 128u * (p.tma_epi == 2 ? 64u : 128u);
		return (a.residual ? 4u : 2u) * tile;
	};
	uint32_t epi_bytes = epiBytes();
	uint32_t fixed = 1024u + 512u + (p.b_resident ? all_b : 0u) + epi_bytes;
	if (fixed + 2 * p.stage_bytes > kSmemLimit && !p.pool) {
		// streamed-weight layers with big stages: fall back to the register epilogue
		p.tma_epi = 0;
		epi_bytes = 0;
		fixed = 1024u + 512u + (p.b_resident ? all_b : 0u);
	}
	if (fixed + 2 * p.stage_bytes > kSmemLimit) return cudaErrorInvalidValue;
	int stages = static_cast<int>((kSmemLimit - fixed) / p.stage_bytes);
	if (stages > kMaxStages) stages = kMaxStages;
	// dual pipelines need an even stage count (disjoint halves of the ring, see the kernel)
	p.dual = 0;
	if (opt.dual && p.kb == 1 && p.b_resident && !a.residual && stages >= 4) {
		p.dual = 1;
		stages &= ~1;
	}
	p.stages = stages;
	p.pdl = opt.pdl ? 1 : 0;
	p.bias = a.bias;
	p.residual = a.residual;
	p.out = a.out;
	p.act = a.act;
	p.slope = a.slope;
	p.out_f32 = a.out_f32;
	p.shuffle2 = a.shuffle2;
	p.status = nullptr;

	CUtensorMap mapA, mapB, mapC, mapR;
	std::memset(&mapC, 0, sizeof(mapC));
	std::memset(&mapR, 0, sizeof(mapR));
	if (p.tma_epi) {
		// output tensor [batch, oh, ow, cout_stride]; one (pooled) pixel tile x N-tile per copy
		const bool f32 = p.tma_epi == 3;
		const cuuint64_t esz = f32 ? 4 : 2;
		const int oh = p.pool ? a.h / 2 : a.h, ow = p.pool ? a.w / 2 : a.w;
		cuuint64_t dims[4] = {static_cast<cuuint64_t>(a.cout_stride), static_cast<cuuint64_t>(ow),
		    static_cast<cuuint64_t>(oh), static_cast<cuuint64_t>(a.batch)};
		cuuint64_t strides[3] = {static_cast<cuuint64_t>(a.cout_stride) * esz,
		    static_cast<cuuint64_t>(ow) * a.cout_stride * esz, static_cast<cuuint64_t>(oh) * ow * a.cout_stride * esz};
		cuuint32_t box[4] = {static_cast<cuuint32_t>(p.nt), static_cast<cuuint32_t>(p.pool ? kTileW / 2 : kTileW),
		    static_cast<cuuint32_t>(p.pool ? kTileH / 2 : kTileH), 1};
		cuuint32_t estr[4] = {1, 1, 1, 1};
		const CUtensorMapSwizzle swz = p.tma_epi == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
		CUresult r = encode(&mapC, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, a.out,
		    dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
		    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
		if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
		if (a.residual) {
			cuuint64_t rdims[4] = {static_cast<cuuint64_t>(a.cout_stride), static_cast<cuuint64_t>(a.w),
			    static_cast<cuuint64_t>(a.h), static_cast<cuuint64_t>(a.batch)};
			cuuint64_t rstrides[3] = {static_cast<cuuint64_t>(a.cout_stride) * 2,
			    static_cast<cuuint64_t>(a.w) * a.cout_stride * 2, static_cast<cuuint64_t>(a.h) * a.w * a.cout_stride * 2};
			cuuint32_t rbox[4] = {static_cast<cuuint32_t>(p.nt), kTileW, kTileH, 1};
			r = encode(&mapR, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half *>(a.residual), rdims, rstrides,
			    rbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
			    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
			if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
		}
	}
	{
		cuuint64_t dims[4] = {static_cast<cuuint64_t>(a.cin_stride), static_cast<cuuint64_t>(a.w),
		    static_cast<cuuint64_t>(a.h), static_cast<cuuint64_t>(a.batch)};
		cuuint64_t strides[3] = {static_cast<cuuint64_t>(a.cin_stride) * 2,
		    static_cast<cuuint64_t>(a.w) * a.cin_stride * 2,
		    static_cast<cuuint64_t>(a.h) * a.w * a.cin_stride * 2};
		cuuint32_t box[4] = {64, static_cast<cuuint32_t>(p.pitch), static_cast<cuuint32_t>(halo_h), 1};
		cuuint32_t estr[4] = {1, 1, 1, 1};
		CUresult r = encode(&mapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half *>(a.in), dims, strides,
		    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
		    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
		if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
	}
	{
		cuuint64_t dims[2] = {64, static_cast<cuuint64_t>(taps) * p.kb * a.cout};
		cuuint64_t strides[1] = {128};
		cuuint32_t box[2] = {64, static_cast<cuuint32_t>(p.nt)};
		cuuint32_t estr[2] = {1, 1};
		CUresult r = encode(&mapB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(a.weights), dims, strides,
		    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
		    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
		if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
	}
	std::memcpy(out->map_a, &mapA, 128);
	std::memcpy(out->map_b, &mapB, 128);
	std::memcpy(out->map_c, &mapC, 128);
	std::memcpy(out->map_r, &mapR, 128);
	std::memcpy(out->params, &p, sizeof(p));
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	out->grid = p.total_tiles < sms ? p.total_tiles : sms;
	out->smem_bytes = fixed + static_cast<uint32_t>(p.stages) * p.stage_bytes;
	out->pdl = p.pdl;
	// per device and cheap: set at every prepare (plan time), never on the launch path
	return cudaFuncSetAttribute(conv_tc_function(p.ks, p.tma_epi), cudaFuncAttributeMaxDynamicSharedMemorySize,
	    static_cast<int>(kSmemLimit));
}

cudaError_t conv_tc_launch(const ConvTcLaunch &l, TcStatus *status, cudaStream_t s) {
	CUtensorMap mapA, mapB, mapC, mapR;
	TcParams p;
	std::memcpy(&mapA, l.map_a, 128);
	std::memcpy(&mapB, l.map_b, 128);
	std::memcpy(&mapC, l.map_c, 128);
	std::memcpy(&mapR, l.map_r, 128);
	std::memcpy(&p, l.params, sizeof(p));
	p.status = status;
	cudaLaunchConfig_t cfg{};
	cfg.gridDim = dim3(l.grid);
	cfg.blockDim = dim3(p.tma_epi != 0 ? kThreadsWide : kThreads);
	cfg.dynamicSmemBytes = l.smem_bytes;
	cfg.stream = s;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr;
	cfg.numAttrs = l.pdl ? 1 : 0;
	void *args[5] = {&mapA, &mapB, &mapC, &mapR, &p};
	return cudaLaunchKernelExC(&cfg, conv_tc_function(p.ks, p.tma_epi), args);
}

}  // namespace ju
