// CTA-pair (cta_group::2) variant of the tcgen05 3x3 convolution for the
// generator's 64 -> 64 layers (the ResBlock stack, 91 % of the frame's MACs:
// scripts/training/models.py:193-254, 544-550).
//
// Same dataflow as conv_tc.cu (halo tile + shifted UMMA descriptors, weights
// resident, TMA residual load / TMA store epilogue), but two CTAs on the two
// SMs of a TPC form a cluster and the leader issues ONE tcgen05.mma
// .cta_group::2 with M = 256 for both pixel tiles: each CTA supplies its own
// 128-pixel A tile and only HALF of B (32 of the 64 output channels' weight
// rows), the tensor cores of the pair exchange the halves.  The shared-memory
// port - the bound of the single-CTA kernel - then carries
// 4 KB (A) + 1 KB (B) instead of 4 + 2 KB per MMA, and the resident weights
// shrink to 36 KB per CTA.
//
// Protocol differences to the single-CTA kernel:
//   * both CTAs' TMA loads (halo, weights) complete on the LEADER's full / weight
//     barriers (cp.async.bulk.tensor .cta_group::2, barrier address mapped with mapa);
//   * the leader's tcgen05.commit multicasts to the empty / tmem-full barriers of
//     both CTAs; the follower's epilogue warps arrive remotely on the leader's
//     tmem-empty barrier;
//   * TMEM is allocated / freed with .cta_group::2 by both CTAs, the pair is
//     fenced with cluster barriers at start-up and teardown.
#include <cstring>

#include "kernels.h"
#include "tc_common.cuh"

namespace ju {

namespace {

using namespace tc;

constexpr int kTileH = 16, kTileW = 8;
constexpr int kThreads2 = 320;  // producer warp, MMA warp, 8 epilogue warps
constexpr int kMaxStages = 8;
constexpr uint32_t kSmemLimit = 227 * 1024;
constexpr uint32_t kABox = 18u * 10u * 128u;          // halo tile (18 x 10 pixels x 64 ch fp16)
constexpr uint32_t kARegion = (kABox + 1023u) & ~1023u;
constexpr uint32_t kBSlice = 32u * 128u;              // this CTA's half of one tap's weights
constexpr uint32_t kBBytes = 9u * kBSlice;
constexpr uint32_t kEpiTile = 128u * 128u;
constexpr int kResDepth = 3;  // residual tiles in flight

struct Tc2Params {
	int batch, h, w;
	int tiles_x, tiles_y, total_tiles;
	int stages;
	int act;
	float slope;
	int pdl;
	int has_residual;
	const float *bias;
	int *error_flag;
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
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const CUtensorMap *map, uint32_t bar_cluster, int c0,
    int c1, int c2, int c3) {
	asm volatile(
	    "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
	    " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
	    "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
	    : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar_cluster, int c0,
    int c1) {
	asm volatile(
	    "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
	    " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
	    "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
	    : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
    uint32_t accumulate) {
	asm volatile(
	    "{\n\t.reg .pred p;\n\t"
	    "setp.ne.b32 p, %4, 0;\n\t"
	    "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
	    "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
	    : "memory");
}
// arrive on the barrier at this CTA-relative offset in both CTAs of the pair
__device__ __forceinline__ void umma2_commit_both(uint32_t bar) {
	asm volatile(
	    "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
	    "h"(static_cast<uint16_t>(3))
	    : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads2, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
    const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_r, const Tc2Params p) {
	extern __shared__ uint8_t smem_raw[];
	const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
	uint8_t *smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
	const uint32_t resb_base = smem_base + static_cast<uint32_t>(p.stages) * kARegion;
	const uint32_t epi_out_base = resb_base + kBBytes;
	const uint32_t epi_res_base = epi_out_base + 2u * kEpiTile;
	const uint32_t bar_base = epi_res_base + (p.has_residual ? kResDepth * kEpiTile : 0u);
	auto full_bar = [&](int s) { return bar_base + 8u * s; };
	auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
	auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + s); };
	auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 2 + s); };
	const uint32_t w_bar = bar_base + 8u * (2 * kMaxStages + 4);
	const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages + 5);
	auto rfull_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 6 + s); };
	auto rempty_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 6 + kResDepth + s); };

	const int warp = threadIdx.x >> 5;
	const int lane = threadIdx.x & 31;
	const uint32_t rank = cluster_ctarank();
	const bool leader = rank == 0;
	const int cluster_id = blockIdx.x >> 1;
	const int n_clusters = gridDim.x >> 1;
	const int n_pairs = (p.total_tiles + 1) >> 1;
	constexpr uint32_t kTmemCols = 128;  // 2 accumulator stages x 64 columns

	if (warp == 0 && lane == 0) {
		for (int s = 0; s < p.stages; ++s) {
			mbar_init(full_bar(s), 1);   // leader: its producer's arrive.expect_tx (bytes of both CTAs)
			mbar_init(empty_bar(s), 1);  // each CTA: the leader's multicast commit
		}
		for (int s = 0; s < 2; ++s) {
			mbar_init(tfull_bar(s), 1);    // each CTA: the leader's multicast commit
			mbar_init(tempty_bar(s), 16);  // leader: 8 epilogue warps of each CTA
		}
		for (int s = 0; s < kResDepth; ++s) {
			mbar_init(rfull_bar(s), 1);
			mbar_init(rempty_bar(s), 8);
		}
		mbar_init(w_bar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 1) {
		asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
		             "r"(kTmemCols)
		             : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
	}
	tcgen05_fence_before();
	cluster_sync_all();  // barriers of both CTAs initialised, TMEM allocated
	tcgen05_fence_after();
	uint32_t tmem_base;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

	if (p.pdl) grid_launch_dependents();

	// tile of this CTA in pair `pr`; an odd tile count leaves one dummy tile whose
	// coordinates are out of range: TMA loads zero-fill it, the TMA store clips it
	auto tile_coords = [&](int pr, int &b, int &y0, int &x0) {
		const int tile = 2 * pr + static_cast<int>(rank);
		if (tile >= p.total_tiles) {
			b = 0;
			y0 = p.tiles_y * kTileH;
			x0 = 0;
			return;
		}
		const int tx = tile % p.tiles_x;
		const int rest = tile / p.tiles_x;
		y0 = (rest % p.tiles_y) * kTileH;
		x0 = tx * kTileW;
		b = rest / p.tiles_y;
	};

	if (warp == 0) {
		// ===================== TMA producer (each CTA loads its own tile / weight half) ==========
		if (lane == 0) {
			const uint32_t w_bar_leader = mapa(w_bar, 0);
			if (leader) mbar_arrive_expect_tx(w_bar, 2u * kBBytes);
			for (int s = 0; s < 9; ++s) {
				// tap s, this CTA's 32 output channels: rows [s*64 + rank*32, +32)
				tma2_load_2d(resb_base + s * kBSlice, &map_b, w_bar_leader, 0, s * 64 + static_cast<int>(rank) * 32);
			}
			if (p.pdl) grid_dependency_wait();
			auto load_residual = [&](int tc, int pr) {
				int b, y0, x0;
				tile_coords(pr, b, y0, x0);
				const int rb = tc % kResDepth;
				const uint32_t rph = (tc / kResDepth) & 1;
				mbar_wait(rempty_bar(rb), rph ^ 1u, p.error_flag, 6);
				mbar_arrive_expect_tx(rfull_bar(rb), kEpiTile);
				tma_load_4d(epi_res_base + rb * kEpiTile, &map_r, rfull_bar(rb), 0, x0, y0, b);
			};
			int it = 0, prev = -1;
			for (int pr = cluster_id; pr < n_pairs; pr += n_clusters, ++it) {
				int b, y0, x0;
				tile_coords(pr, b, y0, x0);
				const int s = it % p.stages;
				const uint32_t ph = (it / p.stages) & 1;
				mbar_wait(empty_bar(s), ph ^ 1u, p.error_flag, 1);
				if (leader) mbar_arrive_expect_tx(full_bar(s), 2u * kABox);
				tma2_load_4d(smem_base + s * kARegion, &map_a, mapa(full_bar(s), 0), 0, x0 - 1, y0 - 1, b);
				if (p.has_residual && prev >= 0) load_residual(it - 1, prev);
				prev = pr;
			}
			if (p.has_residual && prev >= 0) load_residual(it - 1, prev);
		}
	} else if (warp == 1) {
		// ===================== MMA issuer (leader CTA only) =====================
		if (leader) {
			// M = 256 (128 rows per CTA), N = 64, K = 16
			const uint32_t idesc = (1u << 4) | (static_cast<uint32_t>(64 >> 3) << 17) | (static_cast<uint32_t>(256 >> 4) << 24);
			const uint32_t a_hi = static_cast<uint32_t>(make_smem_desc(0, 1280u, 0) >> 32);
			const uint32_t b_hi = static_cast<uint32_t>(make_smem_desc(0, 1024u, 0) >> 32);
			const uint32_t lo_flags = 1u << 16;
			mbar_wait(w_bar, 0, p.error_flag, 2);
			int it = 0;
			for (int pr = cluster_id; pr < n_pairs; pr += n_clusters, ++it) {
				const int as = it & 1;
				const uint32_t aph = (it >> 1) & 1;
				mbar_wait(tempty_bar(as), aph ^ 1u, p.error_flag, 3);
				const int s = it % p.stages;
				const uint32_t ph = (it / p.stages) & 1;
				mbar_wait(full_bar(s), ph, p.error_flag, 4);
				tcgen05_fence_after();
				const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * 64);
				const uint32_t a_lo = lo_flags | ((smem_base + s * kARegion) >> 4);
				const uint32_t b_lo = lo_flags | (resb_base >> 4);
				if (elect_one_sync()) {
#pragma unroll
					for (int tap = 0; tap < 9; ++tap) {
						const uint32_t a_tap = a_lo + (tap / 3) * 80u + (tap % 3) * 8u;  // halo pitch 10 px = 80 x 16 B
						const uint32_t b_tap = b_lo + tap * (kBSlice >> 4);
#pragma unroll
						for (int k16 = 0; k16 < 4; ++k16) {
							const uint64_t a_desc = (static_cast<uint64_t>(a_hi) << 32) | (a_tap + k16 * 2u);
							const uint64_t b_desc = (static_cast<uint64_t>(b_hi) << 32) | (b_tap + k16 * 2u);
							umma2_f16(d_tmem, a_desc, b_desc, idesc, (tap | k16) != 0 ? 1u : 0u);
						}
					}
					umma2_commit_both(empty_bar(s));
					umma2_commit_both(tfull_bar(as));
				}
				__syncwarp();
			}
		}
	} else {
		// ===================== epilogue (8 warps per CTA, identical to conv_tc EPI=1) ============
		const int q = warp & 3;
		const int half = (warp - 2) >> 2;
		const int row = q * 32 + lane;
		const int etid = threadIdx.x - 64;
		const uint32_t tempty_leader0 = mapa(tempty_bar(0), 0), tempty_leader1 = mapa(tempty_bar(1), 0);
		float bias_reg[32];
#pragma unroll
		for (int c = 0; c < 32; ++c) bias_reg[c] = p.bias ? __ldg(p.bias + half * 32 + c) : 0.f;
		if (p.pdl) grid_dependency_wait();
		const uint32_t sw = static_cast<uint32_t>(row & 7);
		const int coff = half * 4;
		int it = 0;
		for (int pr = cluster_id; pr < n_pairs; pr += n_clusters, ++it) {
			int b, y0, x0;
			tile_coords(pr, b, y0, x0);
			const int as = it & 1;
			const uint32_t aph = (it >> 1) & 1;
			if (etid == 0 && it >= 2) {
				asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
			}
			mbar_wait(tfull_bar(as), aph, p.error_flag, 5);
			tcgen05_fence_after();
			uint32_t acc[32];
			const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
			                       static_cast<uint32_t>(as * 64 + half * 32);
			__syncwarp();
			tmem_ld32(taddr, acc);
			tmem_ld_wait();
			// the accumulator is in registers: release TMEM to the leader's MMA warp first
			tcgen05_fence_before();
			__syncwarp();
			if (lane == 0) mbar_arrive_cluster(as ? tempty_leader1 : tempty_leader0);
			uint4 res[4];
			if (p.has_residual) {
				const int rb = it % kResDepth;
				mbar_wait(rfull_bar(rb), (it / kResDepth) & 1, p.error_flag, 7);
				const uint4 *res_row = reinterpret_cast<const uint4 *>(
				    smem_gen + (epi_res_base - smem_base) + rb * kEpiTile + row * 128u);
#pragma unroll
				for (int c = 0; c < 4; ++c) res[c] = res_row[(coff + c) ^ sw];
				__syncwarp();
				if (lane == 0) mbar_arrive(rempty_bar(rb));
			}
			epilogue_barrier<256>();
			float v[32];
#pragma unroll
			for (int c = 0; c < 32; ++c) v[c] = __uint_as_float(acc[c]) + bias_reg[c];
			if (p.has_residual) {
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
			    smem_gen + (epi_out_base - smem_base) + as * kEpiTile + row * 128u);
#pragma unroll
			for (int c = 0; c < 4; ++c) {
				__half2 h0 = __floats2half2_rn(v[c * 8 + 0], v[c * 8 + 1]);
				__half2 h1 = __floats2half2_rn(v[c * 8 + 2], v[c * 8 + 3]);
				__half2 h2 = __floats2half2_rn(v[c * 8 + 4], v[c * 8 + 5]);
				__half2 h3 = __floats2half2_rn(v[c * 8 + 6], v[c * 8 + 7]);
				out_row[(coff + c) ^ sw] = make_uint4(*reinterpret_cast<uint32_t *>(&h0), *reinterpret_cast<uint32_t *>(&h1),
				    *reinterpret_cast<uint32_t *>(&h2), *reinterpret_cast<uint32_t *>(&h3));
			}
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
			epilogue_barrier<256>();
			if (etid == 0) tma_store_4d(&map_c, epi_out_base + as * kEpiTile, 0, x0, y0, b);
		}
		if (etid == 0) {
			asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
		}
	}

	// the leader's MMAs read the follower's shared memory and the follower arrives on
	// the leader's barriers: neither CTA may exit (or free TMEM) before both are done
	tcgen05_fence_before();
	cluster_sync_all();
	if (warp == 1) {
		tcgen05_fence_after();
		asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols)
		             : "memory");
	}
}

using EncodeTiledFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encodeTiled2() {
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

bool conv_tc2_supported(const ConvArgs &a) {
	return a.ksize == 3 && a.cin == 64 && a.cin_stride % 64 == 0 && a.cout == 64 && a.cout_stride % 8 == 0 &&
	       !a.out_f32 && !a.shuffle2 && !a.pool;
}

cudaError_t conv_tc2_prepare(const ConvArgs &a, ConvTcLaunch *out) {
	if (!conv_tc2_supported(a)) return cudaErrorInvalidValue;
	EncodeTiledFn encode = encodeTiled2();
	if (!encode) return cudaErrorNotSupported;
	static_assert(sizeof(Tc2Params) <= sizeof(out->params), "ConvTcLaunch::params too small");
	Tc2Params p{};
	p.batch = a.batch;
	p.h = a.h;
	p.w = a.w;
	p.tiles_x = (a.w + kTileW - 1) / kTileW;
	p.tiles_y = (a.h + kTileH - 1) / kTileH;
	p.total_tiles = a.batch * p.tiles_x * p.tiles_y;
	p.act = a.act;
	p.slope = a.slope;
	p.pdl = 1;
	p.has_residual = a.residual ? 1 : 0;
	p.bias = a.bias;
	const uint32_t fixed = 1024u + 512u + kBBytes + (a.residual ? 2u + kResDepth : 2u) * kEpiTile;
	int stages = static_cast<int>((kSmemLimit - fixed) / kARegion);
	if (stages > kMaxStages) stages = kMaxStages;
	if (stages < 2) return cudaErrorInvalidValue;
	p.stages = stages;

	CUtensorMap mapA, mapB, mapC, mapR;
	std::memset(&mapR, 0, sizeof(mapR));
	cuuint32_t estr[4] = {1, 1, 1, 1};
	{
		cuuint64_t dims[4] = {static_cast<cuuint64_t>(a.cin_stride), static_cast<cuuint64_t>(a.w),
		    static_cast<cuuint64_t>(a.h), static_cast<cuuint64_t>(a.batch)};
		cuuint64_t strides[3] = {static_cast<cuuint64_t>(a.cin_stride) * 2,
		    static_cast<cuuint64_t>(a.w) * a.cin_stride * 2, static_cast<cuuint64_t>(a.h) * a.w * a.cin_stride * 2};
		cuuint32_t box[4] = {64, 10, 18, 1};
		if (encode(&mapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half *>(a.in), dims, strides, box, estr,
		        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
		        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
			return cudaErrorInvalidValue;
		}
	}
	{
		// weights packed by conv_tc_pack_weights: [tap][kb=1][cout=64][64]; half-N boxes
		cuuint64_t dims[2] = {64, 9 * 64};
		cuuint64_t strides[1] = {128};
		cuuint32_t box[2] = {64, 32};
		if (encode(&mapB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(a.weights), dims, strides, box, estr,
		        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
		        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
			return cudaErrorInvalidValue;
		}
	}
	{
		cuuint64_t dims[4] = {static_cast<cuuint64_t>(a.cout_stride), static_cast<cuuint64_t>(a.w),
		    static_cast<cuuint64_t>(a.h), static_cast<cuuint64_t>(a.batch)};
		cuuint64_t strides[3] = {static_cast<cuuint64_t>(a.cout_stride) * 2,
		    static_cast<cuuint64_t>(a.w) * a.cout_stride * 2, static_cast<cuuint64_t>(a.h) * a.w * a.cout_stride * 2};
		cuuint32_t box[4] = {64, kTileW, kTileH, 1};
		if (encode(&mapC, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, a.out, dims, strides, box, estr,
		        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
		        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
			return cudaErrorInvalidValue;
		}
		if (a.residual && encode(&mapR, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half *>(a.residual), dims,
		                      strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
		                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
			return cudaErrorInvalidValue;
		}
	}
	std::memcpy(out->map_a, &mapA, 128);
	std::memcpy(out->map_b, &mapB, 128);
	std::memcpy(out->map_c, &mapC, 128);
	std::memcpy(out->map_r, &mapR, 128);
	std::memcpy(out->params, &p, sizeof(p));
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	const int pairs = (p.total_tiles + 1) / 2;
	const int clusters = pairs < sms / 2 ? pairs : sms / 2;
	out->grid = 2 * clusters;
	out->smem_bytes = fixed + static_cast<uint32_t>(p.stages) * kARegion;
	out->pdl = 1;
	return cudaSuccess;
}

cudaError_t conv_tc2_launch(const ConvTcLaunch &l, int *error_flag, cudaStream_t s) {
	static bool attr_set[16] = {false};
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev >= 0 && dev < 16 && !attr_set[dev]) {
		cudaError_t e = cudaFuncSetAttribute(conv_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
		    static_cast<int>(kSmemLimit));
		if (e != cudaSuccess) return e;
		attr_set[dev] = true;
	}
	CUtensorMap mapA, mapB, mapC, mapR;
	Tc2Params p;
	std::memcpy(&mapA, l.map_a, 128);
	std::memcpy(&mapB, l.map_b, 128);
	std::memcpy(&mapC, l.map_c, 128);
	std::memcpy(&mapR, l.map_r, 128);
	std::memcpy(&p, l.params, sizeof(p));
	p.error_flag = error_flag;
	cudaLaunchConfig_t cfg{};
	cfg.gridDim = dim3(l.grid);
	cfg.blockDim = dim3(kThreads2);
	cfg.dynamicSmemBytes = l.smem_bytes;
	cfg.stream = s;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr;
	cfg.numAttrs = l.pdl ? 1 : 0;
	return cudaLaunchKernelEx(&cfg, conv_tc2_kernel, mapA, mapB, mapC, mapR, p);
}

}  // namespace ju
