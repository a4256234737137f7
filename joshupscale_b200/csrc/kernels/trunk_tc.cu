// Persistent ResBlock trunk: ALL 3x3 64->64 convolutions of the generator's
// residual stack (scripts/training/models.py:193-254, 544-550 - 91 % of the
// frame's MACs) in ONE kernel launch.
//
// Per layer the work is exactly conv_tc.cu's EPI=1 path (halo tile + shifted
// UMMA descriptors, resident weights, TMA residual load / TMA store epilogue,
// 8 epilogue warps).  What the persistent form removes is the per-layer launch
// boundary, which at batch 1 costs ~30 % of a layer (CTA launch, barrier init,
// TMEM allocation, weight fetch behind a cold pipeline, drain of the previous
// grid): the 148 CTAs stay resident, keep their TMEM and mbarrier state, fetch
// the next layer's 72 KB of weights while their own epilogue drains, and meet at
// a grid-wide barrier (one release/acquire counter in global memory) before the
// halo loads of the next layer may read what other CTAs stored.
//
// Buffer rotation (identical to the engine's per-layer plan): three NHWC fp16
// trunk buffers T0,T1,T2; block i reads cur, writes T1, then reads T1 + residual
// cur and writes nxt; cur and nxt swap.  cur = T0 for even i, T2 for odd i.
#include <cstring>

#include "kernels.h"
#include "tc_common.cuh"

namespace ju {

namespace {

using namespace tc;

constexpr int kTileH = 16, kTileW = 8;
constexpr int kThreadsT = 320;  // producer warp, MMA warp, 8 epilogue warps
constexpr int kMaxStages = 8;
constexpr uint32_t kSmemLimit = 227 * 1024;
constexpr uint32_t kABox = 18u * 10u * 128u;
constexpr uint32_t kARegion = (kABox + 1023u) & ~1023u;
constexpr uint32_t kBSlice = 64u * 128u;
constexpr uint32_t kBBytes = 9u * kBSlice;  // one layer's weights
constexpr uint32_t kEpiTile = 128u * 128u;

struct TrunkParams {
	int batch, h, w;
	int tiles_x, tiles_y, total_tiles;
	int stages;
	int n_layers;
	int act;
	float slope;
	int pdl;
	const float *bias;            // [n_layers][64]
	unsigned int *sync_counter;   // zero at launch; re-armed to zero by the last arrival
	int *error_flag;
};

struct TrunkMaps {
	CUtensorMap in[3];    // halo boxes (64 ch, 10, 18, 1) over T0, T1, T2
	CUtensorMap tile[3];  // pixel tiles (64 ch, 8, 16, 1) over T0, T1, T2: residual loads and output stores
	CUtensorMap w;        // weights of all layers: rows [layer][tap][cout], 64 ch each
};

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int *p) {
	unsigned int v;
	asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}

__global__ void __launch_bounds__(kThreadsT, 1)
trunk_tc_kernel(const __grid_constant__ TrunkMaps maps, const TrunkParams p) {
	extern __shared__ uint8_t smem_raw[];
	const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
	uint8_t *smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
	const uint32_t resb_base = smem_base + static_cast<uint32_t>(p.stages) * kARegion;
	const uint32_t epi_out_base = resb_base + kBBytes;
	const uint32_t epi_res_base = epi_out_base + 2u * kEpiTile;
	const uint32_t bar_base = epi_res_base + 2u * kEpiTile;
	auto full_bar = [&](int s) { return bar_base + 8u * s; };
	auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
	auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + s); };
	auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 2 + s); };
	const uint32_t wfull_bar = bar_base + 8u * (2 * kMaxStages + 4);
	const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages + 5);
	auto rfull_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 6 + s); };
	auto rempty_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 8 + s); };
	const uint32_t wempty_bar = bar_base + 8u * (2 * kMaxStages + 10);

	const int warp = threadIdx.x >> 5;
	const int lane = threadIdx.x & 31;
	constexpr uint32_t kTmemCols = 128;

	if (warp == 0 && lane == 0) {
		for (int s = 0; s < p.stages; ++s) {
			mbar_init(full_bar(s), 1);
			mbar_init(empty_bar(s), 1);
		}
		for (int s = 0; s < 2; ++s) {
			mbar_init(tfull_bar(s), 1);
			mbar_init(tempty_bar(s), 8);
			mbar_init(rfull_bar(s), 1);
			mbar_init(rempty_bar(s), 8);
		}
		mbar_init(wfull_bar, 1);
		mbar_init(wempty_bar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 1) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
		             "r"(kTmemCols)
		             : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	tcgen05_fence_before();
	__syncthreads();
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
	auto layer_in = [](int l) { return (l & 1) ? 1 : (((l >> 1) & 1) ? 2 : 0); };
	auto layer_res = [](int l) { return (l & 1) ? (((l >> 1) & 1) ? 2 : 0) : -1; };
	auto layer_out = [](int l) { return (l & 1) ? (((l >> 1) & 1) ? 0 : 2) : 1; };

	if (warp == 0) {
		// ===================== TMA producer =====================
		if (lane == 0) {
			auto load_weights = [&](int l) {
				mbar_arrive_expect_tx(wfull_bar, kBBytes);
				for (int s = 0; s < 9; ++s) tma_load_2d(resb_base + s * kBSlice, &maps.w, wfull_bar, 0, (l * 9 + s) * 64);
			};
			load_weights(0);
			if (p.pdl) grid_dependency_wait();
			int it = 0, tcount = 0;
			for (int l = 0; l < p.n_layers; ++l) {
				if (l > 0) {
					// grid barrier: every CTA has stored (and published) its tiles of layer l-1
					const unsigned int target = static_cast<unsigned int>(l) * gridDim.x;
					unsigned int spins = 0;
					while (ld_acquire_gpu(p.sync_counter) < target) {
						__nanosleep(64);
						if (++spins > (1u << 24)) {
							if (p.error_flag) atomicExch(p.error_flag, 8);
							__trap();
						}
					}
					// order the async-proxy (TMA) reads below after the acquire above
					asm volatile("fence.proxy.async;" ::: "memory");
				}
				const CUtensorMap *min = &maps.in[layer_in(l)];
				const int r = layer_res(l);
				auto load_residual = [&](int tc, int tile) {
					int b, y0, x0;
					decode(tile, b, y0, x0);
					const int rb = tc & 1;
					const uint32_t rph = (tc >> 1) & 1;
					mbar_wait(rempty_bar(rb), rph ^ 1u, p.error_flag, 6);
					mbar_arrive_expect_tx(rfull_bar(rb), kEpiTile);
					tma_load_4d(epi_res_base + rb * kEpiTile, &maps.tile[r], rfull_bar(rb), 0, x0, y0, b);
				};
				int prev_tile = -1, rcount_base = tcount;
				(void) rcount_base;
				for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
					int b, y0, x0;
					decode(tile, b, y0, x0);
					const int s = it % p.stages;
					const uint32_t ph = (it / p.stages) & 1;
					mbar_wait(empty_bar(s), ph ^ 1u, p.error_flag, 1);
					mbar_arrive_expect_tx(full_bar(s), kABox);
					tma_load_4d(smem_base + s * kARegion, min, full_bar(s), 0, x0 - 1, y0 - 1, b);
					if (r >= 0 && prev_tile >= 0) load_residual(tcount++, prev_tile);
					prev_tile = tile;
				}
				if (r >= 0 && prev_tile >= 0) load_residual(tcount++, prev_tile);
				if (l + 1 < p.n_layers) {
					// the MMAs of this layer have consumed the resident weights -> fetch the next layer's
					mbar_wait(wempty_bar, static_cast<uint32_t>(l & 1), p.error_flag, 9);
					load_weights(l + 1);
				}
			}
		}
	} else if (warp == 1) {
		// ===================== MMA issuer =====================
		const uint32_t idesc = make_idesc(64);
		const uint32_t a_hi = static_cast<uint32_t>(make_smem_desc(0, 1280u, 0) >> 32);
		const uint32_t b_hi = static_cast<uint32_t>(make_smem_desc(0, 1024u, 0) >> 32);
		const uint32_t lo_flags = 1u << 16;
		int it = 0;
		for (int l = 0; l < p.n_layers; ++l) {
			mbar_wait(wfull_bar, static_cast<uint32_t>(l & 1), p.error_flag, 2);
			for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
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
				const bool last = tile + static_cast<int>(gridDim.x) >= p.total_tiles;
				if (elect_one_sync()) {
#pragma unroll
					for (int tap = 0; tap < 9; ++tap) {
						const uint32_t a_tap = a_lo + (tap / 3) * 80u + (tap % 3) * 8u;
						const uint32_t b_tap = b_lo + tap * (kBSlice >> 4);
#pragma unroll
						for (int k16 = 0; k16 < 4; ++k16) {
							const uint64_t a_desc = (static_cast<uint64_t>(a_hi) << 32) | (a_tap + k16 * 2u);
							const uint64_t b_desc = (static_cast<uint64_t>(b_hi) << 32) | (b_tap + k16 * 2u);
							umma_f16(d_tmem, a_desc, b_desc, idesc, (tap | k16) != 0 ? 1u : 0u);
						}
					}
					umma_commit(empty_bar(s));
					umma_commit(tfull_bar(as));
					if (last) umma_commit(wempty_bar);  // weights of layer l no longer needed
				}
				__syncwarp();
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
		if (p.pdl) grid_dependency_wait();
		int it = 0, rcount = 0;
		for (int l = 0; l < p.n_layers; ++l) {
			float bias_reg[32];
#pragma unroll
			for (int c = 0; c < 32; ++c) bias_reg[c] = __ldg(p.bias + l * 64 + half * 32 + c);
			const bool has_res = (l & 1) != 0;
			const CUtensorMap *mout = &maps.tile[layer_out(l)];
			for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
				int b, y0, x0;
				decode(tile, b, y0, x0);
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
				tcgen05_fence_before();
				__syncwarp();
				if (lane == 0) mbar_arrive(tempty_bar(as));
				uint4 res[4];
				if (has_res) {
					const int rb = rcount & 1;
					mbar_wait(rfull_bar(rb), static_cast<uint32_t>((rcount >> 1) & 1), p.error_flag, 7);
					const uint4 *res_row = reinterpret_cast<const uint4 *>(
					    smem_gen + (epi_res_base - smem_base) + rb * kEpiTile + row * 128u);
#pragma unroll
					for (int c = 0; c < 4; ++c) res[c] = res_row[(coff + c) ^ sw];
					__syncwarp();
					if (lane == 0) mbar_arrive(rempty_bar(rb));
					++rcount;
				}
				epilogue_barrier<256>();  // staging[as] free (wait_group.read above)
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
				    smem_gen + (epi_out_base - smem_base) + as * kEpiTile + row * 128u);
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
				asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
				epilogue_barrier<256>();
				if (etid == 0) tma_store_4d(mout, epi_out_base + as * kEpiTile, 0, x0, y0, b);
			}
			// publish this CTA's tiles of layer l to the grid
			if (etid == 0) {
				asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores complete
				asm volatile("fence.proxy.async;" ::: "memory");           // async-proxy writes -> generic proxy
				__threadfence();
				const unsigned int old = atomicAdd(p.sync_counter, 1u);
				// the last arrival of the last layer re-arms the counter for the next launch
				if (l == p.n_layers - 1 && old == static_cast<unsigned int>(p.n_layers) * gridDim.x - 1u) {
					atomicExch(p.sync_counter, 0u);
				}
			}
		}
	}

	tcgen05_fence_before();
	__syncthreads();
	if (warp == 1) {
		tcgen05_fence_after();
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols)
		             : "memory");
	}
}

using EncodeTiledFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encodeTiledT() {
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

constexpr uint32_t kFixed = 1024u + 512u + kBBytes + 4u * kEpiTile;

}  // namespace

cudaError_t trunk_tc_prepare(const TrunkArgs &a, TrunkTcLaunch *out) {
	EncodeTiledFn encode = encodeTiledT();
	if (!encode) return cudaErrorNotSupported;
	if (a.cstride % 64 || a.n_layers < 1 || (a.n_layers & 1) || a.lead_in) return cudaErrorInvalidValue;
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
	p.pdl = 1;
	p.bias = a.bias;
	p.sync_counter = a.sync_counter;
	int stages = static_cast<int>((kSmemLimit - kFixed) / kARegion);
	if (stages > kMaxStages) stages = kMaxStages;
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
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	// every CTA must be co-resident (grid barrier): at most one per SM
	out->grid = p.total_tiles < sms ? p.total_tiles : sms;
	out->smem_bytes = kFixed + static_cast<uint32_t>(p.stages) * kARegion;
	out->sync_counter = a.sync_counter;
	return cudaSuccess;
}

cudaError_t trunk_tc_launch(const TrunkTcLaunch &l, int *error_flag, cudaStream_t s) {
	static bool attr_set[16] = {false};
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev >= 0 && dev < 16 && !attr_set[dev]) {
		cudaError_t e = cudaFuncSetAttribute(trunk_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
		    static_cast<int>(kSmemLimit));
		if (e != cudaSuccess) return e;
		attr_set[dev] = true;
	}
	TrunkMaps maps;
	TrunkParams p;
	std::memcpy(&maps, l.maps, sizeof(maps));
	std::memcpy(&p, l.params, sizeof(p));
	p.error_flag = error_flag;
	cudaLaunchConfig_t cfg{};
	cfg.gridDim = dim3(l.grid);
	cfg.blockDim = dim3(kThreadsT);
	cfg.dynamicSmemBytes = l.smem_bytes;
	cfg.stream = s;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr;
	cfg.numAttrs = p.pdl ? 1 : 0;
	return cudaLaunchKernelEx(&cfg, trunk_tc_kernel, maps, p);
}

}  // namespace ju
