// Reference-grade SIMT convolution (CUDA cores, fp16 in / fp32 accumulate).
//
// This is NOT the performance path: it is the always-correct device
// implementation used (a) to bring the whole graph up, (b) as the on-device
// cross-check for the tcgen05 implicit-GEMM kernel (conv_tc.cu), and (c) for
// layer shapes the tensor-core kernel does not cover.  Same fused epilogue as
// the tensor-core kernel: +bias (folded BatchNorm), +residual, ReLU/LeakyReLU,
// fp16 or fp32 store, optional 2x2 pixel-shuffle store for ConvTranspose k2s2.
//
// Replaces, inside the reference's TensorRT engine, layers.Conv2D(3x3 / 1x1,
// SAME) + BatchNormalization + activation (+ Add) (scripts/training/models.py:
// 193-254, 377-447, 531-550) and Conv2DTranspose(k2, s2) (models.py:559-572).
#include "kernels.h"

#include <cstring>

namespace ju {

namespace {

constexpr int kTilePx = 64;   // output pixels per block (one image-row segment)
constexpr int kTileCo = 64;   // output channels per block
constexpr int kCk = 16;       // input channels per smem chunk
constexpr int kXStride = 20;  // halves per staged pixel (16 + pad: conflict-free, 8B aligned)

template <int KS>
__global__ void __launch_bounds__(128) conv_simt_kernel(const ConvArgs a) {
	constexpr int PAD = (KS - 1) / 2;
	constexpr int XW = kTilePx + KS - 1;
	__shared__ __align__(16) __half Xs[KS][XW][kXStride];
	__shared__ __align__(16) __half Ws[KS * KS][kCk][kTileCo];

	const int tid = threadIdx.x;
	const int cg = tid & 15;   // 4 output channels
	const int pg = tid >> 4;   // 8 pixels
	const int x0 = blockIdx.x * kTilePx;
	const int y = blockIdx.y;
	const int co_tiles = (a.cout + kTileCo - 1) / kTileCo;
	const int b = blockIdx.z / co_tiles;
	const int co0 = (blockIdx.z % co_tiles) * kTileCo;

	float acc[8][4];
#pragma unroll
	for (int p = 0; p < 8; ++p)
#pragma unroll
		for (int q = 0; q < 4; ++q) acc[p][q] = 0.f;

	const __half *wsrc = static_cast<const __half *>(a.weights);

	for (int c0 = 0; c0 < a.cin; c0 += kCk) {
		// stage input rows y-PAD..y+PAD, pixels x0-PAD..x0+63+PAD, channels c0..c0+15
		for (int v = tid; v < KS * XW * 2; v += 128) {
			int half_idx = v & 1;
			int px = (v >> 1) % XW;
			int r = (v >> 1) / XW;
			int yy = y + r - PAD, xx = x0 + px - PAD;
			uint4 val = make_uint4(0, 0, 0, 0);
			if (yy >= 0 && yy < a.h && xx >= 0 && xx < a.w) {
				val = *reinterpret_cast<const uint4 *>(
				    a.in + ((static_cast<size_t>(b) * a.h + yy) * a.w + xx) * a.cin_stride + c0 +
				    half_idx * 8);
			}
			uint2 *dst = reinterpret_cast<uint2 *>(&Xs[r][px][half_idx * 8]);
			dst[0] = make_uint2(val.x, val.y);
			dst[1] = make_uint2(val.z, val.w);
		}
		// stage weights [tap][c0..c0+15][co0..co0+63]
		for (int v = tid; v < KS * KS * kCk * (kTileCo / 8); v += 128) {
			int cv = v % (kTileCo / 8);
			int c = (v / (kTileCo / 8)) % kCk;
			int tap = v / (kTileCo / 8 * kCk);
			uint4 val = make_uint4(0, 0, 0, 0);
			int co = co0 + cv * 8;
			if (co < a.cout) {
				// cout is a multiple of 8 (checked by the launcher)
				val = *reinterpret_cast<const uint4 *>(
				    wsrc + (static_cast<size_t>(tap) * a.cin + c0 + c) * a.cout + co);
			}
			*reinterpret_cast<uint4 *>(&Ws[tap][c][cv * 8]) = val;
		}
		__syncthreads();
#pragma unroll
		for (int ky = 0; ky < KS; ++ky) {
#pragma unroll 2
			for (int c = 0; c < kCk; c += 2) {
				float2 xv[8 + KS - 1];
#pragma unroll
				for (int p = 0; p < 8 + KS - 1; ++p) {
					xv[p] = __half22float2(
					    *reinterpret_cast<const __half2 *>(&Xs[ky][pg * 8 + p][c]));
				}
#pragma unroll
				for (int kx = 0; kx < KS; ++kx) {
					const int tap = ky * KS + kx;
					uint2 w0 = *reinterpret_cast<const uint2 *>(&Ws[tap][c][cg * 4]);
					uint2 w1 = *reinterpret_cast<const uint2 *>(&Ws[tap][c + 1][cg * 4]);
					float2 w0a = __half22float2(*reinterpret_cast<const __half2 *>(&w0.x));
					float2 w0b = __half22float2(*reinterpret_cast<const __half2 *>(&w0.y));
					float2 w1a = __half22float2(*reinterpret_cast<const __half2 *>(&w1.x));
					float2 w1b = __half22float2(*reinterpret_cast<const __half2 *>(&w1.y));
#pragma unroll
					for (int p = 0; p < 8; ++p) {
						const float2 x = xv[p + kx];
						acc[p][0] = fmaf(x.x, w0a.x, acc[p][0]);
						acc[p][1] = fmaf(x.x, w0a.y, acc[p][1]);
						acc[p][2] = fmaf(x.x, w0b.x, acc[p][2]);
						acc[p][3] = fmaf(x.x, w0b.y, acc[p][3]);
						acc[p][0] = fmaf(x.y, w1a.x, acc[p][0]);
						acc[p][1] = fmaf(x.y, w1a.y, acc[p][1]);
						acc[p][2] = fmaf(x.y, w1b.x, acc[p][2]);
						acc[p][3] = fmaf(x.y, w1b.y, acc[p][3]);
					}
				}
			}
		}
		__syncthreads();
	}

	// ---- fused epilogue ------------------------------------------------
	const int co = co0 + cg * 4;
	if (co >= a.cout) return;
	float bias[4] = {0.f, 0.f, 0.f, 0.f};
	if (a.bias) {
#pragma unroll
		for (int q = 0; q < 4; ++q) bias[q] = a.bias[co + q];
	}
	const int cpp = a.shuffle2 ? a.cout / 4 : a.cout;  // channels per output pixel
	const int sub = a.shuffle2 ? co / cpp : 0;
	const int oc = a.shuffle2 ? co % cpp : co;
#pragma unroll
	for (int p = 0; p < 8; ++p) {
		const int x = x0 + pg * 8 + p;
		if (x >= a.w) continue;
		size_t opix;
		if (a.shuffle2) {
			opix = (static_cast<size_t>(b) * 2 * a.h + 2 * y + (sub >> 1)) * 2 * a.w + 2 * x + (sub & 1);
		} else {
			opix = (static_cast<size_t>(b) * a.h + y) * a.w + x;
		}
		float v[4];
#pragma unroll
		for (int q = 0; q < 4; ++q) v[q] = acc[p][q] + bias[q];
		if (a.residual) {
			uint2 r = *reinterpret_cast<const uint2 *>(a.residual + opix * a.cout_stride + oc);
			float2 r0 = __half22float2(*reinterpret_cast<const __half2 *>(&r.x));
			float2 r1 = __half22float2(*reinterpret_cast<const __half2 *>(&r.y));
			v[0] += r0.x;
			v[1] += r0.y;
			v[2] += r1.x;
			v[3] += r1.y;
		}
		if (a.act == ACT_RELU) {
#pragma unroll
			for (int q = 0; q < 4; ++q) v[q] = fmaxf(v[q], 0.f);
		} else if (a.act == ACT_LRELU) {
#pragma unroll
			for (int q = 0; q < 4; ++q) v[q] = v[q] >= 0.f ? v[q] : v[q] * a.slope;
		}
		if (a.out_f32) {
			*reinterpret_cast<float4 *>(static_cast<float *>(a.out) + opix * a.cout_stride + oc) =
			    make_float4(v[0], v[1], v[2], v[3]);
		} else {
			__half2 h0 = __floats2half2_rn(v[0], v[1]);
			__half2 h1 = __floats2half2_rn(v[2], v[3]);
			uint2 o;
			o.x = *reinterpret_cast<unsigned int *>(&h0);
			o.y = *reinterpret_cast<unsigned int *>(&h1);
			*reinterpret_cast<uint2 *>(static_cast<__half *>(a.out) + opix * a.cout_stride + oc) = o;
		}
	}
}

}  // namespace

cudaError_t launch_conv_simt(const ConvArgs &a, cudaStream_t s) {
	if ((a.ksize != 1 && a.ksize != 3) || a.cin % kCk || a.cin_stride % 8 || a.cout % 8 ||
	    a.cout_stride % 4 || (a.shuffle2 && (a.cout % 16))) {
		return cudaErrorInvalidValue;
	}
	const int co_tiles = (a.cout + kTileCo - 1) / kTileCo;
	dim3 grid((a.w + kTilePx - 1) / kTilePx, a.h, a.batch * co_tiles);
	if (a.ksize == 3) {
		conv_simt_kernel<3><<<grid, 128, 0, s>>>(a);
	} else {
		conv_simt_kernel<1><<<grid, 128, 0, s>>>(a);
	}
	return cudaGetLastError();
}

size_t conv_simt_weight_bytes(int ksize, int cin_padded, int cout) {
	return static_cast<size_t>(ksize) * ksize * cin_padded * cout * sizeof(__half);
}

// kernel: Keras (kh, kw, Cin, Cout) fp32; scale: per-Cout fp32 or nullptr.
// dst: [tap][cin_padded][cout] fp16, zero for cin..cin_padded.
void conv_simt_pack_weights(const float *kernel, const float *scale, int ksize, int cin,
    int cin_padded, int cout, __half *dst) {
	std::memset(dst, 0, conv_simt_weight_bytes(ksize, cin_padded, cout));
	for (int tap = 0; tap < ksize * ksize; ++tap) {
		for (int c = 0; c < cin; ++c) {
			for (int o = 0; o < cout; ++o) {
				float v = kernel[(static_cast<size_t>(tap) * cin + c) * cout + o];
				if (scale) v = v * scale[o];
				dst[(static_cast<size_t>(tap) * cin_padded + c) * cout + o] = __float2half_rn(v);
			}
		}
	}
}

}  // namespace ju
