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

#include "conv_tc_body.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace ju {

namespace {

using namespace tc;
using namespace tcconv;

template <int KS, int EPI>
__global__ void __launch_bounds__(EPI != 0 ? kThreadsWide : kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
    const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_r,
    const TcParams p) {
	extern __shared__ uint8_t smem_raw[];
	Waiter W(p.status, TC_KERNEL_CONV);
	PersistLayer none{};
	conv_tc_body<KS, EPI, false>(&map_a, &map_b, &map_c, &map_r, p, smem_raw, none, W);
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
ConvTcOptions g_TcDefaults{0, 1, 1, 1, 0};

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
	const uint32_t smem_limit = opt.smem_limit ? opt.smem_limit : kSmemLimit;
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
	if (fixed + 2 * p.stage_bytes > smem_limit && !p.pool) {
		// streamed-weight layers with big stages: fall back to the register epilogue
		p.tma_epi = 0;
		epi_bytes = 0;
		fixed = 1024u + 512u + (p.b_resident ? all_b : 0u);
	}
	if (fixed + 2 * p.stage_bytes > smem_limit) return cudaErrorInvalidValue;
	int stages = static_cast<int>((smem_limit - fixed) / p.stage_bytes);
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
