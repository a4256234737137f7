// Bilinear backward warp of the previous HR output, fused with
// SpaceToDepth(4) + Concatenate: writes the generator's 64-channel input tile
// directly (HBM/L2-bound gather).
//
// Replaces DenseWarpLayer -> tfa dense_image_warp
// (scripts/training/keras_layers.py:79-97; scripts/training/tfa/dense_image_warp.py:87-245),
// the flow `unpad` slice (scripts/training/models.py:791-798), DepthToSpace(4)
// of the flow head (models.py:476-479; keras_layers.py:175), SpaceToDepth(4)
// + Concatenate in the generator (models.py:523-530; keras_layers.py:129).
//
// Exactness contract ("warp/indexing exact"): query points, clamped floors and
// clamped alphas are computed in fp32 with the reference's operation order and
// without FMA contraction, so (fy, fx, ay, ax) are bit-identical to the fp32
// oracle given the same flow; the three lerps use the reference's order
// top = ax*(TR-TL)+TL ; bot = ax*(BR-BL)+BL ; out = ay*(bot-top)+top.
#include "kernels.h"
#include "pixel_common.cuh"

namespace ju {

namespace {

constexpr int kLrTile = 16;  // LR pixels per block (one LR row segment) -> 64x4 HR pixels
constexpr int kTilePitch = 72;  // halfs per staged pixel row (64 + 8 padding)

template <bool kTaps, bool kBright>
__global__ void __launch_bounds__(256) warp_s2d_kernel(const __half *__restrict__ pre_gen,
    const float *__restrict__ flow_head, const FrameIO *__restrict__ io, __half *__restrict__ gen_in,
    float *__restrict__ taps, const float *__restrict__ brightness, int h, int w, int ph, int pw,
    int cstride) {
	// staging tile: kLrTile LR pixels x 64 channels fp16, written out as full
	// 128-byte pixel rows (coalesced) after the gather.  Rows are padded to 144 bytes: a warp
	// spans 8 LR pixels (one HR row of 32 pixels keeps the gathers coalesced), and with a
	// 128-byte pitch its 2-byte stores to 8 rows hit the same banks (8-way conflict); the
	// 16-byte skew spreads them.
	__shared__ __align__(16) __half tile[kLrTile][kTilePitch];

	const int b = blockIdx.z;
	const int ly = blockIdx.y;
	const int lx0 = blockIdx.x * kLrTile;
	const int t = threadIdx.x;
	const int i = t >> 6;        // HR row within the 4x4 block
	const int xx = t & 63;       // HR column within the tile
	const int lxl = xx >> 2;     // LR pixel within the tile
	const int j = xx & 3;
	const int lx = lx0 + lxl;
	const int H = 4 * h, W = 4 * w;

	// channels 0..2 (current LR frame) and the zero tail 51..63
	if (t < kLrTile) {
		int x = lx0 + t;
		float c0 = 0.f, c1 = 0.f, c2 = 0.f;
		if (x < w) {
			const FrameIO f = io[b];
			uchar4 p = *reinterpret_cast<const uchar4 *>(f.in + ly * f.in_stride + x * 4ll);
			c0 = preprocess_px(p.x);
			c1 = preprocess_px(p.y);
			c2 = preprocess_px(p.z);
		}
		tile[t][0] = __float2half_rn(c0);
		tile[t][1] = __float2half_rn(c1);
		tile[t][2] = __float2half_rn(c2);
#pragma unroll
		for (int c = 51; c < 64; ++c) tile[t][c] = __half(0.f);
	}

	if (lx < w) {
		const int Y = 4 * ly + i, X = 4 * lx + j;
		// flow(Y, X) = depth_to_space(head)[Y + 4*top, X + 4*left]
		const int top = (ph - h) / 2, left = (pw - w) / 2;
		const float2 fl = __ldg(reinterpret_cast<const float2 *>(
		    flow_head + ((static_cast<size_t>(b) * ph + (ly + top)) * pw + (lx + left)) * 32 +
		    (i * 4 + j) * 2));
		// query = grid - flow (dense_image_warp.py:232-237), (dy, dx) order
		const float qy = __fsub_rn(static_cast<float>(Y), fl.x);
		const float qx = __fsub_rn(static_cast<float>(X), fl.y);
		// floor clamped to [0, size-2], alpha clamped to [0, 1] (113-139)
		const float fy = fminf(fmaxf(0.f, floorf(qy)), static_cast<float>(H - 2));
		const float fx = fminf(fmaxf(0.f, floorf(qx)), static_cast<float>(W - 2));
		const float ay = fminf(fmaxf(0.f, __fsub_rn(qy, fy)), 1.f);
		const float ax = fminf(fmaxf(0.f, __fsub_rn(qx, fx)), 1.f);
		const int iy = static_cast<int>(fy), ix = static_cast<int>(fx);
		if (kTaps) {
			*reinterpret_cast<float4 *>(taps + ((static_cast<size_t>(b) * H + Y) * W + X) * 4) =
			    make_float4(fy, fx, ay, ax);
		}
		// 32-bit element offsets inside one stream's frame (< 2^31 for any supported size)
		const __half *base = pre_gen + static_cast<size_t>(b) * H * W * 4 +
		                     (static_cast<unsigned int>(iy) * W + ix) * 4u;
		// 4 taps x 8 bytes (B,G,R,pad fp16); TL/TR are adjacent in memory
		const uint2 utl = __ldg(reinterpret_cast<const uint2 *>(base));
		const uint2 utr = __ldg(reinterpret_cast<const uint2 *>(base + 4));
		const uint2 ubl = __ldg(reinterpret_cast<const uint2 *>(base + W * 4));
		const uint2 ubr = __ldg(reinterpret_cast<const uint2 *>(base + W * 4 + 4));
		const __half *tl = reinterpret_cast<const __half *>(&utl);
		const __half *tr = reinterpret_cast<const __half *>(&utr);
		const __half *bl = reinterpret_cast<const __half *>(&ubl);
		const __half *br = reinterpret_cast<const __half *>(&ubr);
		const float bright = kBright ? brightness[b] : 0.f;
#pragma unroll
		for (int c = 0; c < 3; ++c) {
			const float vtl = __half2float(tl[c]), vtr = __half2float(tr[c]);
			const float vbl = __half2float(bl[c]), vbr = __half2float(br[c]);
			const float topv = __fadd_rn(__fmul_rn(ax, __fsub_rn(vtr, vtl)), vtl);
			const float botv = __fadd_rn(__fmul_rn(ax, __fsub_rn(vbr, vbl)), vbl);
			float v = __fadd_rn(__fmul_rn(ay, __fsub_rn(botv, topv)), topv);
			if (kBright) v = __fadd_rn(v, bright);
			// space_to_depth: channel 3 + (i*4+j)*3 + c  (keras_layers.py:129)
			tile[lxl][3 + (i * 4 + j) * 3 + c] = __float2half_rn(v);
		}
	}
	__syncthreads();
	// 16 pixels x 128 B = 2 KB: 128 threads x 16 B, fully coalesced
	if (t < kLrTile * 8) {
		int p = t >> 3, q = t & 7;
		if (lx0 + p < w) {
			uint4 v = *reinterpret_cast<const uint4 *>(&tile[p][q * 8]);
			*reinterpret_cast<uint4 *>(
			    gen_in + ((static_cast<size_t>(b) * h + ly) * w + lx0 + p) * cstride + q * 8) = v;
		}
	}
}

}  // namespace

cudaError_t launch_warp_s2d(const __half *pre_gen, const float *flow_head, const FrameIO *io,
    __half *gen_in, float *taps, const float *brightness, int batch, int h, int w, int ph, int pw,
    int cstride, cudaStream_t s) {
	if (cstride < 64 || cstride % 8) return cudaErrorInvalidValue;
	dim3 grid((w + kLrTile - 1) / kLrTile, h, batch);
	if (taps) {
		if (brightness) {
			warp_s2d_kernel<true, true><<<grid, 256, 0, s>>>(pre_gen, flow_head, io, gen_in, taps, brightness, h, w, ph, pw, cstride);
		} else {
			warp_s2d_kernel<true, false><<<grid, 256, 0, s>>>(pre_gen, flow_head, io, gen_in, taps, brightness, h, w, ph, pw, cstride);
		}
	} else if (brightness) {
		warp_s2d_kernel<false, true><<<grid, 256, 0, s>>>(pre_gen, flow_head, io, gen_in, taps, brightness, h, w, ph, pw, cstride);
	} else {
		warp_s2d_kernel<false, false><<<grid, 256, 0, s>>>(pre_gen, flow_head, io, gen_in, taps, brightness, h, w, ph, pw, cstride);
	}
	return cudaGetLastError();
}

}  // namespace ju
