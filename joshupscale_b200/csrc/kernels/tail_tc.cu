// Fused generator tail on tcgen05: conv_trans_1 (+BN+act) as a 1x1 GEMM to
// 4 x 32 channels, and - entirely inside its epilogue - conv_trans_2 (a second,
// tiny GEMM whose A operand the epilogue threads write into tensor memory) + bias +
// tanh + legacy-bilinear x4 of the input frame + add + clip + uint8 BGRX pack +
// fp16 recurrent-state write.  The [2H,2W,32] intermediate never exists in
// memory: 16 epilogue warps, each thread owns one mid-resolution pixel (one of
// the 4 sub-pixels of an LR pixel) = a 2x2 block of HR pixels.
//
// Replaces, inside the reference's TensorRT engine and C++ glue:
//   Conv2DTranspose(32,k2,s2)+BN+act            scripts/training/models.py:559-572
//   Conv2DTranspose(3,k2,s2)+bias, tanh          models.py:573-583
//   UpscaleLayer(scale=4) + Add + ClipLayer      models.py:584-593; keras_layers.py:46-52, 253-266
//   PostprocessLayer (truncating uint8 cast)     keras_layers.py:227-230
//   out_raw -> pre_gen' recurrent output         models.py:808-823
//   castKernel fp->u8 BGRX, X = 0                core/src/cuda_convert.cc.cu:39-45, 95-108
//
// Exactness: everything after the tanh that feeds a byte is fp32 in the
// reference's operation order with round-to-nearest intrinsics (no FMA
// contraction), identical to final_kernel in pixel_io.cu.
#include <cstring>

#include "kernels.h"
#include "pixel_common.cuh"
#include "tc_common.cuh"

namespace ju {

namespace {

using namespace tc;

constexpr int kTileH = 16, kTileW = 8;
constexpr int kEpiWarps = 16;  // 4 TMEM lane quarters x 4 sub-pixels
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kStages = 6;
constexpr uint32_t kATile = 128u * 128u;  // 128 pixels x 64 ch fp16
constexpr uint32_t kBBytes = 128u * 128u; // 128 output channels x 64 ch fp16

struct TailParams {
	int batch, h, w;
	int tiles_x, tiles_y, total_tiles;
	int tile_begin;  // first tile of this launch (a launch may cover a band of tile rows only)
	int act;
	float slope;
	int pdl;
	// conv_trans_2 (s = i2*2+j2, o, c; fp16-representable values: every CTA turns them into the 2 KB
	// B tile of the second GEMM), its bias and the folded BN bias of conv_trans_1 (the same 32 values
	// for each of its 4 sub-pixels) travel IN the kernel parameters: constant-bank operands
	float w2[4 * 3 * 32];
	float bias2[4];
	float bias1[32];
	const FrameIO *io;
	__half *pre_gen_next;  // [batch,4H,4W,4]
	float *out_raw;        // optional [batch,4H,4W,3]
	const float *brightness;  // optional [batch]
	TcStatus *status;
};

// tanh(x) = 1 - 2 / (exp(2x) + 1) with ex2.approx / fast division: absolute error
// below 1e-6 (vs 1 - 2 ulp of tanhf), ~6 instructions instead of ~40.
__device__ __forceinline__ float tanh_fast(float x) {
	const float e = __expf(2.f * x);
	return 1.f - __fdividef(2.f, e + 1.f);
}

// conv_trans_2 on the tensor core: the activation of conv_trans_1 is rounded to fp16 anyway (the
// engine's storage contract) and w2 holds fp16-representable values, so z = m * w2^T is one tiny
// TS-form GEMM per sub-pixel group: every thread writes its pixel's 32 channels as one row of the A
// operand into tensor memory (tcgen05.st, lane = its own TMEM lane), one elected thread of the
// group's four warps issues two M128 N16 K16 MMAs against the 2 KB w2 tile in shared memory, and
// every thread reads its 12 sums back with tcgen05.ld.  This replaces 192 packed fp32 FMAs and as
// many constant loads per pixel in a kernel that is bound by instruction issue.
constexpr uint32_t kA2Col = 256u;  // 4 groups x 16 columns: [128 lanes x 32 ch fp16]
constexpr uint32_t kD2Col = 320u;  // 4 groups x 16 columns: [128 lanes x 16 fp32] (12 used)
constexpr uint32_t kW2Tile = 16u * 128u;

__device__ __forceinline__ void umma_ts_f16(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
	asm volatile(
	    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
	    "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
	    "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
	    : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
	asm volatile(
	    "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
	    "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
	    "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
	    : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
	asm volatile(
	    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
	    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
	      "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
	    : "r"(taddr)
	    : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
tail_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
    const TailParams p) {
	extern __shared__ uint8_t smem_raw[];
	const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
	uint8_t *smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
	const uint32_t b_base = smem_base + kStages * kATile;
	const uint32_t bar_base = smem_base + kStages * kATile + kBBytes;
	auto full_bar = [&](int s) { return bar_base + 8u * s; };
	auto empty_bar = [&](int s) { return bar_base + 8u * (8 + s); };
	auto tfull_bar = [&](int s) { return bar_base + 8u * (16 + s); };
	auto tempty_bar = [&](int s) { return bar_base + 8u * (18 + s); };
	const uint32_t w_bar = bar_base + 8u * 20;
	const uint32_t tmem_slot = bar_base + 8u * 21;
	auto d2_bar = [&](int g) { return bar_base + 8u * (24 + g); };
	const uint32_t w2_base = bar_base + 1024u;  // SWIZZLE_128B tile: 16 rows (s2 * 3 + o) x 64 ch fp16, zero padded

	const int warp = threadIdx.x >> 5;
	const int lane = threadIdx.x & 31;
	constexpr uint32_t kTmemCols = 512;  // 2 accumulator stages x 128 columns, conv_trans_2 operand and result

	if (warp == 0 && lane == 0) {
		for (int s = 0; s < kStages; ++s) {
			mbar_init(full_bar(s), 1);
			mbar_init(empty_bar(s), 1);
		}
		for (int s = 0; s < 2; ++s) {
			mbar_init(tfull_bar(s), 1);
			mbar_init(tempty_bar(s), kEpiWarps);
		}
		mbar_init(w_bar, 1);
		for (int g = 0; g < 4; ++g) mbar_init(d2_bar(g), 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	for (int idx = threadIdx.x; idx < 16 * 64; idx += kThreads) {
		const uint32_t n = static_cast<uint32_t>(idx) >> 6, c = static_cast<uint32_t>(idx) & 63u;
		const float v = (n < 12u && c < 32u) ? p.w2[n * 32u + c] : 0.f;
		*reinterpret_cast<__half *>(smem_gen + (w2_base - smem_base) + n * 128u + ((((c >> 3) ^ (n & 7u)) << 4) | ((c & 7u) << 1))) =
		    __float2half_rn(v);
	}
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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
		int tx = tile % p.tiles_x;
		int rest = tile / p.tiles_x;
		int ty = rest % p.tiles_y;
		b = rest / p.tiles_y;
		y0 = ty * kTileH;
		x0 = tx * kTileW;
	};

	if (warp == 0) {
		if (lane == 0) {
			Waiter W(p.status, TC_KERNEL_TAIL);
			mbar_arrive_expect_tx(w_bar, kBBytes);
			tma_load_2d(b_base, &map_b, w_bar, 0, 0);
			if (p.pdl) grid_dependency_wait();
			int it = 0;
			for (int tile = p.tile_begin + blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
				int b, y0, x0;
				decode(tile, b, y0, x0);
				const int s = it % kStages;
				const uint32_t ph = (it / kStages) & 1;
				if (!W.wait(empty_bar(s), ph ^ 1u, 1)) continue;
				mbar_arrive_expect_tx(full_bar(s), kATile);
				tma_load_4d(smem_base + s * kATile, &map_a, full_bar(s), 0, x0, y0, b);
			}
		}
	} else if (warp == 1) {
		if (lane == 0) {
			const uint32_t idesc = make_idesc(128);
			const uint32_t hi = static_cast<uint32_t>(make_smem_desc(0, 1024u, 0) >> 32);
			const uint32_t lo_flags = 1u << 16;
			Waiter W(p.status, TC_KERNEL_TAIL);
			W.wait(w_bar, 0, 2);
			int it = 0;
			for (int tile = p.tile_begin + blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
				const int as = it & 1;
				const uint32_t aph = (it >> 1) & 1;
				W.wait(tempty_bar(as), aph ^ 1u, 3);
				const int s = it % kStages;
				const uint32_t ph = (it / kStages) & 1;
				if (!W.wait(full_bar(s), ph, 4)) continue;  // aborted frame: nothing is issued any more
				tcgen05_fence_after();
				const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * 128);
				const uint32_t a_lo = lo_flags | ((smem_base + s * kATile) >> 4);
				const uint32_t b_lo = lo_flags | (b_base >> 4);
#pragma unroll
				for (int k16 = 0; k16 < 4; ++k16) {
					const uint64_t a_desc = (static_cast<uint64_t>(hi) << 32) | (a_lo + k16 * 2u);
					const uint64_t b_desc = (static_cast<uint64_t>(hi) << 32) | (b_lo + k16 * 2u);
					umma_f16(d_tmem, a_desc, b_desc, idesc, k16 != 0 ? 1u : 0u);
				}
				umma_commit(empty_bar(s));
				umma_commit(tfull_bar(as));
			}
		}
	} else {
		// ===================== epilogue =====================
		// 16 warps: warp (quarter, q) owns TMEM lanes quarter*32.. (its 32 LR pixels) and
		// accumulator columns q*32.. (sub-pixel q = i*2+j of conv_trans_1), i.e. one thread
		// = one mid-resolution pixel (2y+i, 2x+j) = a 2x2 block of HR pixels.  Four warps
		// per SM sub-partition hide each other's latency.
		const int q4 = warp & 3;           // TMEM lane quarter this warp may access
		const int q = (warp - 2) >> 2;     // sub-pixel handled by this warp
		const int i = q >> 1, j = q & 1;
		const int row = q4 * 32 + lane;
		const float b2[3] = {p.bias2[0], p.bias2[1], p.bias2[2]};
		Waiter W(p.status, TC_KERNEL_TAIL);
		if (p.pdl) grid_dependency_wait();
		const int H4 = 4 * p.h, W4 = 4 * p.w;
		int it = 0;
		for (int tile = p.tile_begin + blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
			int b, y0, x0;
			decode(tile, b, y0, x0);
			const int as = it & 1;
			const uint32_t aph = (it >> 1) & 1;
			const int y = y0 + (row >> 3), x = x0 + (row & 7);
			const bool valid = y < p.h && x < p.w;
			// bilinear corners of the LR frame (src = dst/4, clamp at the far edge)
			float cr[4][3] = {};
			const FrameIO f = p.io[b];
			const float bright = p.brightness ? p.brightness[b] : 0.f;
			if (valid) {
				const int x1 = min(x + 1, p.w - 1), y1 = min(y + 1, p.h - 1);
				const uint8_t *r0 = f.in + y * f.in_stride, *r1 = f.in + y1 * f.in_stride;
				const uchar4 ptl = *reinterpret_cast<const uchar4 *>(r0 + x * 4ll);
				const uchar4 ptr_ = *reinterpret_cast<const uchar4 *>(r0 + x1 * 4ll);
				const uchar4 pbl = *reinterpret_cast<const uchar4 *>(r1 + x * 4ll);
				const uchar4 pbr = *reinterpret_cast<const uchar4 *>(r1 + x1 * 4ll);
				cr[0][0] = preprocess_px(ptl.x); cr[0][1] = preprocess_px(ptl.y); cr[0][2] = preprocess_px(ptl.z);
				cr[1][0] = preprocess_px(ptr_.x); cr[1][1] = preprocess_px(ptr_.y); cr[1][2] = preprocess_px(ptr_.z);
				cr[2][0] = preprocess_px(pbl.x); cr[2][1] = preprocess_px(pbl.y); cr[2][2] = preprocess_px(pbl.z);
				cr[3][0] = preprocess_px(pbr.x); cr[3][1] = preprocess_px(pbr.y); cr[3][2] = preprocess_px(pbr.z);
			}
			W.wait(tfull_bar(as), aph, 5);
			tcgen05_fence_after();
			const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) +
			                       static_cast<uint32_t>(as * 128 + q * 32);
			uint32_t acc[32];
			__syncwarp();
			tmem_ld32(taddr, acc);
			tmem_ld_wait();
			tcgen05_fence_before();
			__syncwarp();
			W.sync_warp();
			if (lane == 0 && !W.dead) mbar_arrive(tempty_bar(as));
			uint32_t mh[16];
#pragma unroll
			for (int c2 = 0; c2 < 16; ++c2) {
				float v0 = __uint_as_float(acc[2 * c2]) + p.bias1[2 * c2];
				float v1 = __uint_as_float(acc[2 * c2 + 1]) + p.bias1[2 * c2 + 1];
				if (p.act == ACT_RELU) {
					v0 = fmaxf(v0, 0.f);
					v1 = fmaxf(v1, 0.f);
				} else if (p.act == ACT_LRELU) {
					v0 = v0 >= 0.f ? v0 : v0 * p.slope;
					v1 = v1 >= 0.f ? v1 : v1 * p.slope;
				}
				// the engine's storage contract rounds this activation to fp16
				const __half2 hh = __floats2half2_rn(v0, v1);
				mh[c2] = *reinterpret_cast<const uint32_t *>(&hh);
			}
			// conv_trans_2: z[s][o] = sum_c m[c] * w2[s][o][c] as D2[128 pixels x 16] = A2[128 x 32] * w2^T
			tmem_st16(tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + kA2Col + static_cast<uint32_t>(q * 16), mh);
			asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
			tcgen05_fence_before();
			asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");  // the four warps of this sub-pixel group
			if (q4 == 0) {
				tcgen05_fence_after();
				if (elect_one_sync() && !W.dead) {
					const uint32_t hi2 = static_cast<uint32_t>(make_smem_desc(0, 1024u, 0) >> 32);
					const uint32_t lo2 = (1u << 16) | (w2_base >> 4);
					const uint32_t idesc2 = make_idesc(16);
#pragma unroll
					for (int k16 = 0; k16 < 2; ++k16) {
						umma_ts_f16(tmem_base + kD2Col + static_cast<uint32_t>(q * 16),
						    tmem_base + kA2Col + static_cast<uint32_t>(q * 16 + k16 * 8),
						    (static_cast<uint64_t>(hi2) << 32) | (lo2 + k16 * 2u), idesc2, k16 != 0 ? 1u : 0u);
					}
					umma_commit(d2_bar(q));
				}
				__syncwarp();
			}
			// while the second GEMM is in flight: the bilinear x4 of the input frame for this thread's
			// 2x2 HR pixels (independent of z; same operation order as the reference)
			float up[4][3];
#pragma unroll
			for (int i2 = 0; i2 < 2; ++i2) {
				const float ty = static_cast<float>(2 * i + i2) * 0.25f;
#pragma unroll
				for (int j2 = 0; j2 < 2; ++j2) {
					const float tx = static_cast<float>(2 * j + j2) * 0.25f;
#pragma unroll
					for (int o = 0; o < 3; ++o) {
						const float topv = __fadd_rn(cr[0][o], __fmul_rn(__fsub_rn(cr[1][o], cr[0][o]), tx));
						const float botv = __fadd_rn(cr[2][o], __fmul_rn(__fsub_rn(cr[3][o], cr[2][o]), tx));
						up[i2 * 2 + j2][o] = __fadd_rn(topv, __fmul_rn(__fsub_rn(botv, topv), ty));
					}
				}
			}
			W.wait(d2_bar(q), static_cast<uint32_t>(it & 1), 6);
			tcgen05_fence_after();
			uint32_t zr[16];
			__syncwarp();
			tmem_ld16(tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + kD2Col + static_cast<uint32_t>(q * 16), zr);
			tmem_ld_wait();
			tcgen05_fence_before();
			float z[4][3];
#pragma unroll
			for (int s2 = 0; s2 < 4; ++s2)
#pragma unroll
				for (int o = 0; o < 3; ++o) z[s2][o] = __uint_as_float(zr[s2 * 3 + o]);
			if (valid) {
#pragma unroll
				for (int i2 = 0; i2 < 2; ++i2) {
					const int R = 2 * i + i2;  // HR row inside the LR pixel's 4x4 block
					const int Y = 4 * y + R;
					uchar4 px[2];
					__align__(16) __half st[2][4];
#pragma unroll
					for (int j2 = 0; j2 < 2; ++j2) {
						unsigned char o8[3];
#pragma unroll
						for (int o = 0; o < 3; ++o) {
							const float zz = tanh_fast(z[i2 * 2 + j2][o] + b2[o]);
							const float r = fminf(fmaxf(__fadd_rn(up[i2 * 2 + j2][o], zz), -0.5f), 0.5f);
							o8[o] = static_cast<unsigned char>(static_cast<int>(__fmul_rn(__fadd_rn(r, 0.5f), 255.0f)));
							st[j2][o] = __float2half_rn(r - bright);
							if (p.out_raw) {
								p.out_raw[((static_cast<size_t>(b) * H4 + Y) * W4 + 4 * x + 2 * j + j2) * 3 + o] = r;
							}
						}
						st[j2][3] = __half(0.f);
						px[j2] = make_uchar4(o8[0], o8[1], o8[2], 0);
					}
					uchar4 *orow = reinterpret_cast<uchar4 *>(f.out + Y * f.out_stride + (4 * x + 2 * j) * 4ll);
					orow[0] = px[0];
					orow[1] = px[1];
					*reinterpret_cast<uint4 *>(
					    p.pre_gen_next + ((static_cast<size_t>(b) * H4 + Y) * W4 + 4 * x + 2 * j) * 4) =
					    *reinterpret_cast<const uint4 *>(&st[0][0]);
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

constexpr uint32_t kSmemBytes = 1024u + kStages * kATile + kBBytes + 1024u + kW2Tile;

}  // namespace

cudaError_t tail_tc_prepare(const TailArgs &a, TailTcLaunch *out) {
	EncodeTiledFn encode = encodeTiled();
	if (!encode) return cudaErrorNotSupported;
	if (a.cin_stride % 64 || a.cin_stride < 64) return cudaErrorInvalidValue;
	static_assert(sizeof(TailParams) <= sizeof(out->params), "TailTcLaunch::params too small");
	TailParams p{};
	p.batch = a.batch;
	p.h = a.h;
	p.w = a.w;
	p.tiles_x = (a.w + kTileW - 1) / kTileW;
	p.tiles_y = (a.h + kTileH - 1) / kTileH;
	p.total_tiles = a.batch * p.tiles_x * p.tiles_y;
	p.tile_begin = 0;
	if (a.tile_row_end > a.tile_row_begin) {
		// a band of tile rows (single stream): tiles are numbered row-major
		if (a.batch != 1 || a.tile_row_end > p.tiles_y || a.tile_row_begin < 0) return cudaErrorInvalidValue;
		p.tile_begin = a.tile_row_begin * p.tiles_x;
		p.total_tiles = a.tile_row_end * p.tiles_x;
	}
	p.act = a.act;
	p.slope = a.slope;
	p.pdl = a.pdl;
	if (!a.w2_host || !a.bias2_host || !a.bias1_host) return cudaErrorInvalidValue;
	std::memcpy(p.w2, a.w2_host, sizeof(p.w2));
	std::memcpy(p.bias2, a.bias2_host, 3 * sizeof(float));
	std::memcpy(p.bias1, a.bias1_host, sizeof(p.bias1));
	p.io = a.io;
	p.pre_gen_next = a.pre_gen_next;
	p.out_raw = a.out_raw;
	p.brightness = a.brightness;
	CUtensorMap mapA, mapB;
	{
		cuuint64_t dims[4] = {static_cast<cuuint64_t>(a.cin_stride), static_cast<cuuint64_t>(a.w),
		    static_cast<cuuint64_t>(a.h), static_cast<cuuint64_t>(a.batch)};
		cuuint64_t strides[3] = {static_cast<cuuint64_t>(a.cin_stride) * 2,
		    static_cast<cuuint64_t>(a.w) * a.cin_stride * 2, static_cast<cuuint64_t>(a.h) * a.w * a.cin_stride * 2};
		cuuint32_t box[4] = {64, kTileW, kTileH, 1};
		cuuint32_t estr[4] = {1, 1, 1, 1};
		if (encode(&mapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half *>(a.in), dims, strides, box, estr,
		        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
		        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
			return cudaErrorInvalidValue;
		}
	}
	{
		// conv_trans_1 weights packed by conv_tc_pack_weights(ksize 1, cin 64, cout 128): [128][64] fp16
		cuuint64_t dims[2] = {64, 128};
		cuuint64_t strides[1] = {128};
		cuuint32_t box[2] = {64, 128};
		cuuint32_t estr[2] = {1, 1};
		if (encode(&mapB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(a.weights1), dims, strides, box, estr,
		        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
		        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
			return cudaErrorInvalidValue;
		}
	}
	std::memcpy(out->map_a, &mapA, 128);
	std::memcpy(out->map_b, &mapB, 128);
	std::memcpy(out->params, &p, sizeof(p));
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	const int n_tiles = p.total_tiles - p.tile_begin;
	out->grid = n_tiles < sms ? n_tiles : sms;
	out->smem_bytes = kSmemBytes;
	out->pdl = a.pdl;
	// per device and cheap: set at every prepare (plan time), never on the launch path
	return cudaFuncSetAttribute(tail_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
	    static_cast<int>(kSmemBytes));
}

cudaError_t tail_tc_launch(const TailTcLaunch &l, TcStatus *status, cudaStream_t s) {
	CUtensorMap mapA, mapB;
	TailParams p;
	std::memcpy(&mapA, l.map_a, 128);
	std::memcpy(&mapB, l.map_b, 128);
	std::memcpy(&p, l.params, sizeof(p));
	p.status = status;
	cudaLaunchConfig_t cfg{};
	cfg.gridDim = dim3(l.grid);
	cfg.blockDim = dim3(kThreads);
	cfg.dynamicSmemBytes = l.smem_bytes;
	cfg.stream = s;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr;
	cfg.numAttrs = l.pdl ? 1 : 0;
	return cudaLaunchKernelEx(&cfg, tail_tc_kernel, mapA, mapB, p);
}

}  // namespace ju
