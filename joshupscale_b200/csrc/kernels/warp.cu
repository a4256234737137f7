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

constexpr int kLrTile = 64;     // LR pixels per block (one LR row segment) -> 256 x 4 HR pixels
constexpr int kTilePitch = 72;  // halfs per staged pixel row (64 + 8 padding)

// One thread = one HR row of one LR pixel's 4x4 block = 4 horizontally adjacent HR pixels: the
// address arithmetic, the flow fetch (2 x LDG.128 for the 4 (dy, dx) pairs) and the staging
// stores are shared by the four pixels, which halves the instruction count per pixel against one
// thread per HR pixel (the kernel was issue-bound, not bandwidth-bound: 230 instructions per pixel).
template <bool kTaps, bool kBright>
__global__ void __launch_bounds__(256) warp_s2d_kernel(const __half *__restrict__ pre_gen,
    const float *__restrict__ flow_head, const FrameIO *__restrict__ io, __half *__restrict__ gen_in,
    float *__restrict__ taps, const float *__restrict__ brightness, int h, int w, int ph, int pw,
    int cstride) {
	// staging tile: kLrTile LR pixels x 64 channels fp16, written out as full 128-byte pixel rows
	// (coalesced) after the gather; rows are padded to 144 bytes to spread the banks
	__shared__ __align__(16) __half tile[kLrTile][kTilePitch];

	const int b = blockIdx.z;
	const int ly = blockIdx.y;
	const int lx0 = blockIdx.x * kLrTile;
	const int t = threadIdx.x;
	const int i = t & 3;     // HR row within the 4x4 block
	const int lxl = t >> 2;  // LR pixel within the tile
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
		const int Y = 4 * ly + i;
		// flow(Y, X) = depth_to_space(head)[Y + 4*top, X + 4*left]: the 4 pixels of this HR row are
		// channels (i*4+j)*2 + {0: dy, 1: dx}, j = 0..3, i.e. 8 consecutive floats
		const int top = (ph - h) / 2, left = (pw - w) / 2;
		const float4 *fp = reinterpret_cast<const float4 *>(
		    flow_head + ((static_cast<size_t>(b) * ph + (ly + top)) * pw + (lx + left)) * 32 + i * 8);
		const float4 f01 = __ldg(fp), f23 = __ldg(fp + 1);
		const float dy[4] = {f01.x, f01.z, f23.x, f23.z};
		const float dx[4] = {f01.y, f01.w, f23.y, f23.w};
		const __half *frame = pre_gen + static_cast<size_t>(b) * H * W * 4;
		const float bright = kBright ? brightness[b] : 0.f;
		const float Yf = static_cast<float>(Y);
		uint2 utl[4], utr[4], ubl[4], ubr[4];
		float ay[4], ax[4];
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			const int X = 4 * lx + j;
			// query = grid - flow (dense_image_warp.py:232-237), (dy, dx) order
			const float qy = __fsub_rn(Yf, dy[j]);
			const float qx = __fsub_rn(static_cast<float>(X), dx[j]);
			// floor clamped to [0, size-2], alpha clamped to [0, 1] (113-139)
			const float fy = fminf(fmaxf(0.f, floorf(qy)), static_cast<float>(H - 2));
			const float fx = fminf(fmaxf(0.f, floorf(qx)), static_cast<float>(W - 2));
			ay[j] = fminf(fmaxf(0.f, __fsub_rn(qy, fy)), 1.f);
			ax[j] = fminf(fmaxf(0.f, __fsub_rn(qx, fx)), 1.f);
			if (kTaps) {
				*reinterpret_cast<float4 *>(taps + ((static_cast<size_t>(b) * H + Y) * W + X) * 4) =
				    make_float4(fy, fx, ay[j], ax[j]);
			}
			// 32-bit element offsets inside one stream's frame (< 2^31 for any supported size);
			// 4 taps x 8 bytes (B,G,R,pad fp16); TL/TR are adjacent in memory
			const __half *base = frame + (static_cast<unsigned int>(static_cast<int>(fy)) * W + static_cast<int>(fx)) * 4u;
			utl[j] = __ldg(reinterpret_cast<const uint2 *>(base));
			utr[j] = __ldg(reinterpret_cast<const uint2 *>(base + 4));
			ubl[j] = __ldg(reinterpret_cast<const uint2 *>(base + W * 4));
			ubr[j] = __ldg(reinterpret_cast<const uint2 *>(base + W * 4 + 4));
		}
		// space_to_depth: channel 3 + (i*4+j)*3 + c (keras_layers.py:129): 12 consecutive halfs
		__half out[12];
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			const __half *tl = reinterpret_cast<const __half *>(&utl[j]);
			const __half *tr = reinterpret_cast<const __half *>(&utr[j]);
			const __half *bl = reinterpret_cast<const __half *>(&ubl[j]);
			const __half *br = reinterpret_cast<const __half *>(&ubr[j]);
#pragma unroll
			for (int c = 0; c < 3; ++c) {
				const float vtl = __half2float(tl[c]), vtr = __half2float(tr[c]);
				const float vbl = __half2float(bl[c]), vbr = __half2float(br[c]);
				const float topv = __fadd_rn(__fmul_rn(ax[j], __fsub_rn(vtr, vtl)), vtl);
				const float botv = __fadd_rn(__fmul_rn(ax[j], __fsub_rn(vbr, vbl)), vbl);
				float v = __fadd_rn(__fmul_rn(ay[j], __fsub_rn(botv, topv)), topv);
				if (kBright) v = __fadd_rn(v, bright);
				out[j * 3 + c] = __float2half_rn(v);
			}
		}
		// halfs [3 + 12 i, 15 + 12 i) of the pixel row: the first and the last one alone, the ten in
		// between as five aligned 32-bit words
		__half *dst = &tile[lxl][3 + 12 * i];
		dst[0] = out[0];
#pragma unroll
		for (int k = 0; k < 5; ++k) {
			*reinterpret_cast<__half2 *>(dst + 1 + 2 * k) = __halves2half2(out[1 + 2 * k], out[2 + 2 * k]);
		}
		dst[11] = out[11];
	}
	__syncthreads();
	// 64 pixels x 128 B = 8 KB: 256 threads x 2 x 16 B, fully coalesced
#pragma unroll
	for (int r = 0; r < 2; ++r) {
		const int idx = t + r * 256;
		const int p = idx >> 3, q = idx & 7;
		if (lx0 + p < w) {
			uint4 v = *reinterpret_cast<const uint4 *>(&tile[p][q * 8]);
			*reinterpret_cast<uint4 *>(gen_in + ((static_cast<size_t>(b) * h + ly) * w + lx0 + p) * cstride + q * 8) = v;
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
