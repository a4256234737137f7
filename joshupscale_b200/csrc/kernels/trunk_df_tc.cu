// Persistent ResBlock trunk, DATAFLOW version: like trunk_tc.cu (all 3x3 64->64
// layers of the generator's residual stack - scripts/training/models.py:193-254,
// 544-550 - in one launch), but without any grid-wide barrier.
//
// Dependencies.  A tile of layer l only needs the 3x3 tile neighbourhood of layer
// l-1.  With the static tile->CTA striding (tile = cta + k*grid) all CTAs advance
// in waves k, and the neighbours of a wave-k tile lie in waves k-1..k+1, so the
// dependency is kept per (layer, wave): every completed TMA store bumps a counter
// (release), and a TMA producer needs counter[l-1][<= k+1] to be full (relaxed
// spin, then one acquire) before it requests a halo of layer l, wave k.  The
// needed wave was stored ~6 tile periods earlier, so in steady state nobody waits
// at a layer boundary and halos of layer l+1 are prefetched while layer l runs.
// (Exact per-tile flags - 9 acquire polls per tile - were measured slower.)
//
// Warp roles (16 warps, 1 CTA/SM).  Every mbarrier operation has to get through
// the shared-memory pipe that the UMMA operand fetch saturates (~300 cycles per
// round trip, measured with an in-kernel timeline), and the tensor pipe only
// queues ~4 MMAs, so every serial per-tile chain is split over warps that take
// tiles round-robin:
//   2 TMA producers   (wait stage-empty, poll wave counters, request the halo)
//   2 MMA issuers     (their barrier waits overlap the other warp's MMAs)
//   8 epilogue warps  (TMEM -> regs, +bias, +shortcut, act, fp16 -> staging tile)
//   3 store warps     (TMA store, wait for completion, GPU-scope publish ~1 us)
//   1 weight loader
// The shortcut (ResBlock input) is read straight from L2 with 256-bit loads by
// the epilogue threads, requested before the accumulator wait.
//
// Weights.  The resident 72 KB are swapped tap by tap: the last two tiles of a
// layer (one per issuer) commit one barrier per tap, the loader warp refills that
// tap's 8 KB slice with layer l+1 immediately, and the first two tiles of layer
// l+1 wait per tap - the reload hides behind the last tiles' own MMAs.
//
// Buffer reuse is safe without extra dependencies: a buffer is rewritten two
// layers after it was read, and the read-after-write chain of the writer
// (neighbours of neighbours) covers every reader of the old contents.
//
// Counters are never reset: launch number `epoch` (kept in global memory, advanced
// by the last CTA to finish) expects (epoch+1) * tiles_in_wave, compared wrap-safe,
// so the captured CUDA graph replays without any memset node.
#include <cstring>

#include "kernels.h"
#include "tc_common.cuh"

namespace ju {

namespace {

using namespace tc;

constexpr int kTileH = 16, kTileW = 8;
constexpr int kStoreWarps = 3;  // one staging tile each
constexpr int kSecondMmaWarp = 11 + kStoreWarps;
constexpr int kSecondProducerWarp = kSecondMmaWarp + 1;
constexpr int kThreadsT = 32 * (kSecondProducerWarp + 1);  // producer, MMA, 8 epilogue warps, weight loader, store/publish warps, 2nd MMA warp
constexpr int kMaxStages = 8;
constexpr uint32_t kSmemLimit = 227 * 1024;
constexpr uint32_t kABox = 18u * 10u * 128u;
constexpr uint32_t kARegion = (kABox + 1023u) & ~1023u;
constexpr uint32_t kBSlice = 64u * 128u;
constexpr uint32_t kBBytes = 9u * kBSlice;  // one layer's weights
// CTA-pair version (PAIR): a cluster of two CTAs issues one tcgen05.mma.cta_group::2 (M = 256) for
// both pixel tiles; each CTA supplies its own halo and HALF of the weights (32 of the 64 output
// channels' rows).  Measured issue rate of the bare MMA stream: 43 instead of 48 cycles
// (profiles/r02_mma_rate2.txt) - the weight fetch is what the pair shares.
constexpr uint32_t kBSliceP = 32u * 128u;
constexpr uint32_t kBBytesP = 9u * kBSliceP;
constexpr uint32_t kEpiTile = 128u * 128u;
constexpr int kAccStages = 4;  // TMEM accumulator stages of 64 columns

struct TrunkParams {
	int batch, h, w;
	int tiles_x, tiles_y, total_tiles;
	int stages;
	int n_layers;
	int act;
	float slope;
	int pdl;
	const float *bias;            // [n_layers][64]
	unsigned int *sync_counter;   // [0] finished-CTA counter, [1] launch epoch
	unsigned int *flags;          // [n_layers][n_waves] stored-tile counters, never reset
	TcStatus *status;
	const __half *buffers[3];     // T0, T1, T2 (residual rows are read straight from global memory)
	int cstride;
	int lead;  // layer 0 is a plain conv reading maps.in[3] and writing T0; ResBlock layers follow
};

struct TrunkMaps {
	CUtensorMap in[4];    // halo boxes (64 ch, 10, 18, 1) over T0, T1, T2 and the lead layer's input
	CUtensorMap tile[3];  // pixel tiles (64 ch, 8, 16, 1) over T0, T1, T2: output stores
	CUtensorMap w;        // weights of all layers: rows [layer][tap][cout], 64 ch each
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
	uint32_t r;
	asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
	return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
	uint32_t r;
	asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
	return r;
}
__device__ __forceinline__ void cluster_sync_all() {
	asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
	asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
	asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads of a CTA pair: the data lands in the issuing CTA, the bytes are counted on the barrier
// at `bar_cluster` (the leader's)
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const CUtensorMap *map, uint32_t bar_cluster, int c0, int c1, int c2,
    int c3) {
	asm volatile(
	    "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
	    " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
	    "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
	    : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar_cluster, int c0, int c1) {
	asm volatile(
	    "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
	    " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
	    "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
	    : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
	asm volatile(
	    "{\n\t.reg .pred p;\n\t"
	    "setp.ne.b32 p, %4, 0;\n\t"
	    "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
	    "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
	    : "memory");
}
// arrive on the barrier at this CTA-relative offset in both CTAs of the pair
__device__ __forceinline__ void umma2_commit_both(uint32_t bar) {
	asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
	    "h"(static_cast<uint16_t>(3))
	    : "memory");
}

// 32 bytes (16 channels) from L2, bypassing L1
__device__ __forceinline__ void ld_global_256(const __half *p, uint4 &a, uint4 &b) {
	asm volatile("ld.global.cg.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
	             : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
	             : "l"(p)
	             : "memory");
}

template <bool PAIR>
__global__ void __launch_bounds__(kThreadsT, 1)
trunk_df_tc_kernel(const __grid_constant__ TrunkMaps maps, const TrunkParams p) {
	extern __shared__ uint8_t smem_raw[];
	const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
	uint8_t *smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
	const uint32_t resb_base = smem_base + static_cast<uint32_t>(p.stages) * kARegion;
	constexpr uint32_t kSlice = PAIR ? kBSliceP : kBSlice;  // this CTA's rows of one tap's weights
	const uint32_t epi_out_base = resb_base + 9u * kSlice;
	// PAIR: the two CTAs of a cluster work on tiles 2i, 2i+1 of every round; rank 0 issues the MMAs
	const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
	const bool leader = rank == 0;
	// first tile of this CTA's (pair's) first round: tile = first + round * grid (+ rank)
	const int first = PAIR ? static_cast<int>(blockIdx.x & ~1u) : static_cast<int>(blockIdx.x);
	const uint32_t bar_base = epi_out_base + kStoreWarps * kEpiTile;
	auto full_bar = [&](int s) { return bar_base + 8u * s; };
	auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
	auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + s); };
	auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + kAccStages + s); };
	const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages + 8);
	auto wfull_tap = [&](int t) { return bar_base + 8u * (2 * kMaxStages + 10 + t); };
	auto wempty_tap = [&](int t) { return bar_base + 8u * (2 * kMaxStages + 19 + t); };
	auto sready_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 28 + s); };  // staging tile written (s < 4)
	auto sfree_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 32 + s); };   // staging tile read by the TMA store

	const int warp = threadIdx.x >> 5;
	const int lane = threadIdx.x & 31;
	constexpr uint32_t kTmemCols = 64u * kAccStages;

	if (warp == 0 && lane == 0) {
		for (int s = 0; s < p.stages; ++s) {
			mbar_init(full_bar(s), 1);
			mbar_init(empty_bar(s), 1);
		}
		for (int s = 0; s < kAccStages; ++s) {
			mbar_init(tfull_bar(s), 1);
			mbar_init(tempty_bar(s), PAIR ? 16 : 8);  // the leader's barrier also counts the partner's epilogue warps
		}
		for (int s = 0; s < kStoreWarps; ++s) {
			mbar_init(sready_bar(s), 8);
			mbar_init(sfree_bar(s), 1);
		}
		for (int t = 0; t < 9; ++t) {
			mbar_init(wfull_tap(t), 1);
			mbar_init(wempty_tap(t), 2);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 1) {
		if (PAIR) {
			asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols)
			             : "memory");
			asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
		} else {
			asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols)
			             : "memory");
			asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
		}
	}
	tcgen05_fence_before();
	if (PAIR) {
		cluster_sync_all();  // the partner's barriers are initialised, its TMEM allocated
	} else {
		__syncthreads();
	}
	tcgen05_fence_after();
	uint32_t tmem_base;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

	if (p.pdl) grid_launch_dependents();

	auto decode = [&](int tile, int &b, int &y0, int &x0) {
		const int tx = tile % p.tiles_x;
		const int rest = tile / p.tiles_x;
		y0 = (rest % p.tiles_y) * kTileH;
		x0 = tx * kTileW;
		b = rest / p.tiles_y;
	};
	// layer l = 2*block + conv: buffers (see header)
	// with a lead layer everything shifts by one: lead reads buffer 3 and writes T0
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

	// launch epoch: identical for every CTA of this launch (advanced by the last CTA to finish)
	const unsigned int epoch = *reinterpret_cast<volatile unsigned int *>(p.sync_counter + 1);
	const int n_waves = (p.total_tiles + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
	// the 3x3 tile neighbourhood spans tile indices t +- (tiles_x + 1): that many waves ahead must be complete
	// ... for THIS CTA: its furthest neighbour tile, index + tiles_x + 1, belongs to CTA
	// (blockIdx.x + tiles_x + 1) % grid in wave k + (blockIdx.x + tiles_x + 1) / grid - so the first
	// grid - tiles_x - 1 CTAs only need wave k itself, which was stored a full tile period earlier
	const int wave_reach = (static_cast<int>(blockIdx.x) + p.tiles_x + 1) / static_cast<int>(gridDim.x);
	// PAIR: the last round may hold a tile without a partner; the partner then runs a dummy tile
	// (coordinates beyond the batch: TMA loads zero-fill it, the TMA store clips it) that neither
	// waits for nor publishes anything
	auto my_tile = [&](int base) { return PAIR ? base + static_cast<int>(rank) : base; };

	if (warp == 0 || warp == kSecondProducerWarp) {
		const int pme = warp == 0 ? 0 : 1;  // two producers, alternating tiles
		// ===================== TMA producer (warp converged; lanes 0..8 poll neighbour flags) =====
		Waiter W(p.status, TC_KERNEL_TRUNK);
		if (p.pdl) grid_dependency_wait();
		int it = 0;
		for (int l = 0; l < p.n_layers; ++l) {
			const CUtensorMap *min = &maps.in[layer_in(l)];
			int known = -1;  // highest wave of layer l-1 known to be completely stored
			for (int base = first; base < p.total_tiles; base += gridDim.x, ++it) {
				const int tile = my_tile(base);
				int b, y0, x0;
				decode(tile, b, y0, x0);
				if ((it & 1) != pme) continue;
				const int s = it % p.stages;
				const uint32_t ph = (it / p.stages) & 1;
				W.wait(empty_bar(s), ph ^ 1u, 1);
				W.sync_warp();
				bool polled = false;
				if (l > 0 && !W.dead && tile < p.total_tiles) {
					// waves <= k+1 of layer l-1 must be completely stored (covers the 3x3 neighbourhood)
					const int k = (tile - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x);
					const int needw = k + wave_reach < n_waves ? k + wave_reach : n_waves - 1;
					while (known < needw && !W.dead) {
						// lane i polls wave known+1+i: relaxed spins (no L1 invalidation per poll), then one
						// acquire load per counter once all of them are complete
						const int wv = known + 1 + lane;
						const bool mine = wv <= needw;
						const int cntw = wv + 1 < n_waves ? static_cast<int>(gridDim.x)
						                                  : p.total_tiles - wv * static_cast<int>(gridDim.x);
						const unsigned int target = (epoch + 1u) * static_cast<unsigned int>(cntw);
						const unsigned int *ctr = p.flags + (l - 1) * n_waves + (mine ? wv : 0);
						unsigned int spins = 0;
						unsigned long long t0 = 0;
						while (true) {
							const bool ok = !mine || static_cast<int>(ld_relaxed_gpu(ctr) - target) >= 0;
							if (__all_sync(0xffffffffu, ok)) break;
							if (spins > 64) __nanosleep(32);  // tight relaxed polls first: the dependency is usually a few hundred ns away
							++spins;
							if (spins == 64u) t0 = globaltimer_ns();
							// a dependency that never arrives aborts the frame (recoverable), it does not trap
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
					if (polled) {
						// order the async-proxy (TMA) reads below after the acquire loads above
						asm volatile("fence.proxy.async;" ::: "memory");
					}
					if (PAIR) {
						// both halos of the pair are counted on the leader's barrier
						if (leader) mbar_arrive_expect_tx(full_bar(s), 2u * kABox);
						tma2_load_4d(smem_base + s * kARegion, min, mapa(full_bar(s), 0), 0, x0 - 1, y0 - 1, b);
					} else {
						mbar_arrive_expect_tx(full_bar(s), kABox);
						tma_load_4d(smem_base + s * kARegion, min, full_bar(s), 0, x0 - 1, y0 - 1, b);
					}
				}
				__syncwarp();
			}
		}
	} else if (warp == 10) {
		// ===================== weight loader: tap slices follow the MMA warp layer by layer ======
		if (lane == 0) {
			Waiter W(p.status, TC_KERNEL_TRUNK);
			for (int l = 0; l < p.n_layers; ++l) {
				for (int t = 0; t < 9; ++t) {
					if (l > 0) W.wait(wempty_tap(t), static_cast<uint32_t>((l - 1) & 1), 9);
					if (W.dead) continue;
					if (PAIR) {
						// this CTA's 32 rows of the tap; both halves are counted on the leader's barrier
						if (leader) mbar_arrive_expect_tx(wfull_tap(t), 2u * kBSliceP);
						tma2_load_2d(resb_base + t * kBSliceP, &maps.w, mapa(wfull_tap(t), 0), 0,
						    (l * 9 + t) * 64 + static_cast<int>(rank) * 32);
					} else {
						mbar_arrive_expect_tx(wfull_tap(t), kBSlice);
						tma_load_2d(resb_base + t * kBSlice, &maps.w, wfull_tap(t), 0, (l * 9 + t) * 64);
					}
				}
			}
		}
	} else if (warp >= 11 && warp < 11 + kStoreWarps) {
		// ===================== store + publish warps =====================
		// kStoreWarps warps take the finished tiles round-robin: TMA store, wait for its completion,
		// publish (GPU-scope release, ~1 us) - so no epilogue warp ever waits on a store or a fence,
		// and the publish latency of one tile overlaps the stores of the next ones.
		if (lane == 0) {
			Waiter W(p.status, TC_KERNEL_TRUNK);
			// test hook: a launch whose status block names this kernel never publishes a tile, i.e.
			// every consumer of layer 0 stalls until its wait expires
			const bool stall = p.status && *reinterpret_cast<volatile int *>(&p.status->inject) == TC_KERNEL_TRUNK;
			if (p.pdl) grid_dependency_wait();
			const int me = warp - 11;
			int it = 0;
			for (int l = 0; l < p.n_layers; ++l) {
				const CUtensorMap *mout = &maps.tile[layer_out(l)];
				for (int base = first; base < p.total_tiles; base += gridDim.x, ++it) {
					const int tile = my_tile(base);
					if (it % kStoreWarps != me) continue;
					int b, y0, x0;
					decode(tile, b, y0, x0);
					const int as = me;  // this warp's staging tile
					const uint32_t aph = (it / kStoreWarps) & 1;
					W.wait(sready_bar(as), aph, 10);
					if (W.dead) continue;
					tma_store_4d(mout, epi_out_base + as * kEpiTile, 0, x0, y0, b);
					asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // smem tile consumed
					mbar_arrive(sfree_bar(as));
					asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // global writes complete
					// the bulk store has completed (async proxy): order it before the generic-proxy
					// release below, which makes it visible to every acquiring producer warp
					asm volatile("fence.proxy.async;" ::: "memory");
					if (stall || tile >= p.total_tiles) continue;
					asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.flags + l * n_waves + (base - first) / static_cast<int>(gridDim.x))
					             : "memory");
				}
			}
			__threadfence();
			// the last store warp of the last CTA to finish advances the epoch for the next launch
			const unsigned int old = atomicAdd(p.sync_counter, 1u);
			if (old == gridDim.x * kStoreWarps - 1u) {
				atomicExch(p.sync_counter, 0u);
				atomicAdd(p.sync_counter + 1, 1u);
			}
		}
	} else if ((warp == 1 || warp == kSecondMmaWarp) && !leader) {
		// PAIR: the partner's issuer warps have nothing to do - the leader issues for both CTAs
	} else if (warp == 1 || warp == kSecondMmaWarp) {
		// ===================== MMA issuers (two warps, alternating tiles) =====================
		// A barrier check has to get through the shared-memory pipe that the UMMA operand fetch
		// saturates (~300 cycles each) and the tensor pipe only queues ~4 instructions, so a single
		// issuer lets the pipe drain between tiles.  With two issuers one warp does its waits while
		// the other warp's MMAs execute.  Tiles are independent (own TMEM stage, own halo stage);
		// only the resident weights couple them, see `pos` below.
		const int mi = warp == 1 ? 0 : 1;
		Waiter W(p.status, TC_KERNEL_TRUNK);
		// PAIR: the commit arrives on the same barrier of BOTH CTAs (stage free, accumulator ready,
		// weight tap free)
		auto commit = [&](uint32_t bar) {
			if (PAIR) {
				umma2_commit_both(bar);
			} else {
				umma_commit(bar);
			}
		};
		// PAIR: M = 256 (128 rows per CTA), N = 64
		const uint32_t idesc = PAIR ? ((1u << 4) | (static_cast<uint32_t>(64 >> 3) << 17) | (static_cast<uint32_t>(256 >> 4) << 24))
		                            : make_idesc(64);
		const uint32_t a_hi = static_cast<uint32_t>(make_smem_desc(0, 1280u, 0) >> 32);
		const uint32_t b_hi = static_cast<uint32_t>(make_smem_desc(0, 1024u, 0) >> 32);
		const uint32_t lo_flags = 1u << 16;
		// tiles of this CTA per layer
		const int cnt = first < p.total_tiles ? (p.total_tiles - first + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x) : 0;
		int it = 0;
		for (int l = 0; l < p.n_layers; ++l) {
			for (int pos = 0; pos < cnt; ++pos, ++it) {
				if ((it & 1) != mi) continue;
				const int as = it % kAccStages;
				const uint32_t aph = (it / kAccStages) & 1;
				const int s = it % p.stages;
				const uint32_t ph = (it / p.stages) & 1;
				W.wait(tempty_bar(as), aph ^ 1u, 3);
				W.wait(full_bar(s), ph, 4);
				W.sync_warp();
				tcgen05_fence_after();
				const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * 64);
				const uint32_t a_lo = lo_flags | ((smem_base + s * kARegion) >> 4);
				const uint32_t b_lo = lo_flags | (resb_base >> 4);
				// the first two tiles of a layer (one per issuer) wait tap by tap for the new weights;
				// the last two release the slices tap by tap (wempty_tap counts 2 arrivals)
				const bool fresh = pos < 2;
				const int releases = pos + 2 >= cnt ? (cnt == 1 ? 2 : 1) : 0;
				if (elect_one_sync()) {
#pragma unroll
					for (int tap = 0; tap < 9; ++tap) {
						if (fresh) W.wait(wfull_tap(tap), static_cast<uint32_t>(l & 1), 2);
						if (W.dead) break;  // aborted frame: nothing is issued or committed any more
						const uint32_t a_tap = a_lo + (tap / 3) * 80u + (tap % 3) * 8u;
						const uint32_t b_tap = b_lo + tap * (kSlice >> 4);
#pragma unroll
						for (int k16 = 0; k16 < 4; ++k16) {
							const uint64_t a_desc = (static_cast<uint64_t>(a_hi) << 32) | (a_tap + k16 * 2u);
							const uint64_t b_desc = (static_cast<uint64_t>(b_hi) << 32) | (b_tap + k16 * 2u);
							if (PAIR) {
								umma2_f16(d_tmem, a_desc, b_desc, idesc, (tap | k16) != 0 ? 1u : 0u);
							} else {
								umma_f16(d_tmem, a_desc, b_desc, idesc, (tap | k16) != 0 ? 1u : 0u);
							}
						}
						// this tap's slice may be overwritten once these MMAs (of both issuers) retire
						if (releases >= 1) commit(wempty_tap(tap));
						if (releases == 2) commit(wempty_tap(tap));
					}
					if (!W.dead) {
						commit(empty_bar(s));
						commit(tfull_bar(as));
					}
				}
				__syncwarp();
				W.sync_warp();
			}
		}
	} else {
		// ===================== epilogue (8 warps) =====================
		const int q = warp & 3;
		const int half = (warp - 2) >> 2;
		const int row = q * 32 + lane;
		const int etid = threadIdx.x - 64;
		const uint32_t sw = static_cast<uint32_t>(row & 7);
		const int coff = half * 4;
		Waiter W(p.status, TC_KERNEL_TRUNK);
		if (p.pdl) grid_dependency_wait();
		int it = 0;
		for (int l = 0; l < p.n_layers; ++l) {
			float bias_reg[32];
#pragma unroll
			for (int c = 0; c < 32; ++c) bias_reg[c] = __ldg(p.bias + l * 64 + half * 32 + c);
			const bool has_res = layer_res(l) >= 0;
			const __half *res_buf = has_res ? p.buffers[layer_res(l)] : nullptr;
			for (int base = first; base < p.total_tiles; base += gridDim.x, ++it) {
				const int tile = my_tile(base);
				const int as = it % kAccStages;
				const uint32_t aph = (it / kAccStages) & 1;
				const int ss = it % kStoreWarps;  // staging tile
				const uint32_t sph = (it / kStoreWarps) & 1;
				// The shortcut row comes straight from global memory (L2) and never touches shared
				// memory, whose port the MMA operand fetch saturates.  It was stored two layers ago
				// by this CTA's own TMA stores; requested before the accumulator wait, so the
				// latency hides behind the MMAs of this tile.
				uint4 res[4];
				if (has_res) {
					int b, y0, x0;
					decode(tile, b, y0, x0);
					const int y = y0 + (row >> 3), x = x0 + (row & 7);
					if (tile < p.total_tiles && y < p.h && x < p.w) {
						const __half *src = res_buf +
						    ((static_cast<size_t>(b) * p.h + y) * p.w + x) * static_cast<size_t>(p.cstride) + half * 32;
						ld_global_256(src, res[0], res[1]);
						ld_global_256(src + 16, res[2], res[3]);
					} else {
						res[0] = res[1] = res[2] = res[3] = make_uint4(0u, 0u, 0u, 0u);
					}
				}
				W.wait(tfull_bar(as), aph, 5);
				tcgen05_fence_after();
				uint32_t acc[32];
				const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
				                       static_cast<uint32_t>(as * 64 + half * 32);
				__syncwarp();
				tmem_ld32(taddr, acc);
				tmem_ld_wait();
				tcgen05_fence_before();
				__syncwarp();
				W.sync_warp();
				if (lane == 0 && !W.dead) {
					if (PAIR) {
						mbar_arrive_cluster(mapa(tempty_bar(as), 0));  // the leader's MMA warps own the accumulators of both CTAs
					} else {
						mbar_arrive(tempty_bar(as));
					}
				}
				W.wait(sfree_bar(ss), sph ^ 1u, 11);  // staging[ss] consumed by the store of tile it-2
				float v[32];
#pragma unroll
				for (int c = 0; c < 32; ++c) v[c] = __uint_as_float(acc[c]) + bias_reg[c];
				if (has_res) {
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
				uint4 *out_row = reinterpret_cast<uint4 *>(
				    smem_gen + (epi_out_base - smem_base) + ss * kEpiTile + row * 128u);
#pragma unroll
				for (int c = 0; c < 4; ++c) {
					__half2 h0 = __floats2half2_rn(v[c * 8 + 0], v[c * 8 + 1]);
					__half2 h1 = __floats2half2_rn(v[c * 8 + 2], v[c * 8 + 3]);
					__half2 h2 = __floats2half2_rn(v[c * 8 + 4], v[c * 8 + 5]);
					__half2 h3 = __floats2half2_rn(v[c * 8 + 6], v[c * 8 + 7]);
					out_row[(coff + c) ^ sw] = make_uint4(*reinterpret_cast<uint32_t *>(&h0),
					    *reinterpret_cast<uint32_t *>(&h1), *reinterpret_cast<uint32_t *>(&h2),
					    *reinterpret_cast<uint32_t *>(&h3));
				}
				// generic-proxy smem writes -> visible to the TMA (async proxy), then hand the tile over
				asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
				__syncwarp();
				W.sync_warp();
				if (lane == 0 && !W.dead) mbar_arrive(sready_bar(ss));
			}
		}
	}

	// PAIR: the leader's MMAs read the partner's shared memory and the partner arrives on the
	// leader's barriers: neither CTA may exit (or free TMEM) before both are done
	tcgen05_fence_before();
	if (PAIR) {
		cluster_sync_all();
	} else {
		__syncthreads();
	}
	if (warp == 1) {
		tcgen05_fence_after();
		if (PAIR) {
			asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
		} else {
			asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
		}
	}
}

using EncodeTiledFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encodeTiledDF() {
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

constexpr uint32_t kFixed = 1024u + 768u + kBBytes + kStoreWarps * kEpiTile;  // alignment slack, barriers, weights, staging
constexpr uint32_t kFixedP = 1024u + 768u + kBBytesP + kStoreWarps * kEpiTile;  // CTA pair: half of the weights per CTA

}  // namespace

cudaError_t trunk_df_tc_prepare(const TrunkArgs &a, TrunkTcLaunch *out) {
	EncodeTiledFn encode = encodeTiledDF();
	if (!encode) return cudaErrorNotSupported;
	const int lead = a.lead_in ? 1 : 0;
	if (a.cstride % 64 || a.n_layers - lead < 2 || ((a.n_layers - lead) & 1)) return cudaErrorInvalidValue;
	static_assert(sizeof(TrunkParams) <= sizeof(out->params), "TrunkTcLaunch::params too small");
	static_assert(sizeof(TrunkMaps) <= sizeof(out->maps), "TrunkTcLaunch::maps too small");
	TrunkParams p{};
	p.batch = a.batch;
	p.h = a.h;
	p.w = a.w;
	p.tiles_x = (a.w + kTileW - 1) / kTileW;
	p.tiles_y = (a.h + kTileH - 1) / kTileH;
	p.total_tiles = a.batch * p.tiles_x * p.tiles_y;
	p.n_layers = a.n_layers;
	p.act = a.act;
	p.slope = a.slope;
	p.pdl = a.cooperative ? 0 : 1;  // a cooperative grid is gang-scheduled: no early (programmatic) start
	p.bias = a.bias;
	p.sync_counter = a.sync_counter;
	p.flags = a.flags;
	if (!a.flags) return cudaErrorInvalidValue;
	for (int i = 0; i < 3; ++i) p.buffers[i] = static_cast<const __half *>(a.buffers[i]);
	p.cstride = a.cstride;
	p.lead = lead;
	const bool pair = a.pair != 0;
	int stages = static_cast<int>((kSmemLimit - (pair ? kFixedP : kFixed)) / kARegion);
	if (stages > kMaxStages) stages = kMaxStages;
	// The two producer / issuer pairs take tiles alternately.  With an EVEN stage count each pair
	// owns a disjoint set of halo stages, i.e. two independent single-producer / single-consumer
	// rings; an odd count would let one pair wait on a barrier whose previous phase belongs to the
	// other pair and may not even have started (mbarrier parity waits alias two phases apart).
	stages &= ~1;
	static_assert(kAccStages % 2 == 0, "TMEM stages must split evenly between the two issuers");
	if (stages < 2) return cudaErrorInvalidValue;
	p.stages = stages;
	TrunkMaps maps;
	std::memset(&maps, 0, sizeof(maps));
	cuuint32_t estr[4] = {1, 1, 1, 1};
	cuuint64_t dims[4] = {static_cast<cuuint64_t>(a.cstride), static_cast<cuuint64_t>(a.w),
	    static_cast<cuuint64_t>(a.h), static_cast<cuuint64_t>(a.batch)};
	cuuint64_t strides[3] = {static_cast<cuuint64_t>(a.cstride) * 2, static_cast<cuuint64_t>(a.w) * a.cstride * 2,
	    static_cast<cuuint64_t>(a.h) * a.w * a.cstride * 2};
	for (int i = 0; i < 3; ++i) {
		cuuint32_t hbox[4] = {64, 10, 18, 1};
		cuuint32_t tbox[4] = {64, kTileW, kTileH, 1};
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
		cuuint32_t hbox[4] = {64, 10, 18, 1};
		if (encode(&maps.in[3], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(a.lead_in), dims, strides, hbox,
		        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
		        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
			return cudaErrorInvalidValue;
		}
	}
	{
		cuuint64_t wd[2] = {64, static_cast<cuuint64_t>(a.n_layers) * 9 * 64};
		cuuint64_t ws[1] = {128};
		cuuint32_t wb[2] = {64, pair ? 32u : 64u};  // CTA pair: each CTA loads its 32 rows of a tap
		if (encode(&maps.w, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(a.weights), wd, ws, wb, estr,
		        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
		        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
			return cudaErrorInvalidValue;
		}
	}
	std::memcpy(out->maps, &maps, sizeof(maps));
	std::memcpy(out->params, &p, sizeof(p));
	// per device and cheap: set at every prepare (plan time), never on the launch path
	cudaError_t attrErr = cudaFuncSetAttribute(pair ? trunk_df_tc_kernel<true> : trunk_df_tc_kernel<false>,
	    cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemLimit));
	if (attrErr != cudaSuccess) return attrErr;
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	// every CTA must be co-resident (inter-CTA dependencies): at most one per SM; CTA pairs need an even grid
	out->grid = p.total_tiles < sms ? p.total_tiles : sms;
	if (pair) {
		out->grid = (out->grid + 1) & ~1;
		if (out->grid > sms) out->grid = sms & ~1;
	}
	out->pair = pair ? 1 : 0;
	out->smem_bytes = (pair ? kFixedP : kFixed) + static_cast<uint32_t>(p.stages) * kARegion;
	out->sync_counter = a.sync_counter;
	out->cooperative = a.cooperative;
	return cudaSuccess;
}

cudaError_t trunk_df_tc_launch(const TrunkTcLaunch &l, TcStatus *status, cudaStream_t s) {
	TrunkMaps maps;
	TrunkParams p;
	std::memcpy(&maps, l.maps, sizeof(maps));
	std::memcpy(&p, l.params, sizeof(p));
	p.status = status;
	cudaLaunchConfig_t cfg{};
	cfg.gridDim = dim3(l.grid);
	cfg.blockDim = dim3(kThreadsT);
	cfg.dynamicSmemBytes = l.smem_bytes;
	cfg.stream = s;
	cudaLaunchAttribute attr[3];
	int n = 0;
	if (l.pair) {
		attr[n].id = cudaLaunchAttributeClusterDimension;
		attr[n].val.clusterDim.x = 2;
		attr[n].val.clusterDim.y = 1;
		attr[n].val.clusterDim.z = 1;
		++n;
	}
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
	return l.pair ? cudaLaunchKernelEx(&cfg, trunk_df_tc_kernel<true>, maps, p) : cudaLaunchKernelEx(&cfg, trunk_df_tc_kernel<false>, maps, p);
}

}  // namespace ju
