// Pixel I/O kernels (HBM-bound): u8 BGRX -> fp16 flow-net input with padding
// and history shift, and the fused output epilogue.
//
// Replaces the reference's castKernel + cudaMemcpy2D paths
// (core/src/cuda_convert.cc.cu:95-108, 360-436) and, inside the TensorRT
// engine, PreprocessLayer / ZeroPadding2D / conv_trans_2 / tanh / UpscaleLayer
// / Add / ClipLayer / PostprocessLayer (scripts/training/models.py:573-593,
// 768-789; keras_layers.py:208, 227-230).
#include "kernels.h"
#include "pixel_common.cuh"

namespace ju {

namespace {

// float32(u8) / 255 - 0.5 with the exact rounding of the fp32 reference
// (keras_layers.py:208); no FMA contraction.
// ---------------------------------------------------------------------
// preprocess: one thread per padded LR pixel.
//   next[0:3]   = cur (0 inside the zero padding)   models.py:780-789
//   next[3:3K]  = prev[0:3K-3]                       models.py:823
// Channels >= 3K of the buffers stay zero from allocation.
// ---------------------------------------------------------------------
__global__ void preprocess_kernel(const FrameIO *__restrict__ io, const __half *__restrict__ prev,
    __half *__restrict__ next, const float *__restrict__ brightness, int h, int w, int ph, int pw,
    int k, int cstride) {
	int x = blockIdx.x * blockDim.x + threadIdx.x;
	int y = blockIdx.y;
	int b = blockIdx.z;
	if (x >= pw) return;
	int top = (ph - h) / 2, left = (pw - w) / 2;
	size_t px = (static_cast<size_t>(b) * ph + y) * pw + x;
	const __half *src = prev + px * cstride;
	__half *dst = next + px * cstride;
	int sy = y - top, sx = x - left;
	float c0 = 0.f, c1 = 0.f, c2 = 0.f;
	if (sy >= 0 && sy < h && sx >= 0 && sx < w) {
		const FrameIO f = io[b];
		uchar4 p = *reinterpret_cast<const uchar4 *>(f.in + sy * f.in_stride + sx * 4ll);
		float br = brightness ? brightness[b] : 0.f;
		c0 = preprocess_px(p.x) - br;
		c1 = preprocess_px(p.y) - br;
		c2 = preprocess_px(p.z) - br;
	}
	int nch = 3 * k;
	if (nch <= 16 && cstride % 8 == 0) {
		// common case (K <= 5): two 16-byte vectors cover all live channels
		__align__(16) __half in16[16];
		__align__(16) __half out16[16];
		*reinterpret_cast<uint4 *>(in16) = *reinterpret_cast<const uint4 *>(src);
		*reinterpret_cast<uint4 *>(in16 + 8) = *reinterpret_cast<const uint4 *>(src + 8);
		out16[0] = __float2half_rn(c0);
		out16[1] = __float2half_rn(c1);
		out16[2] = __float2half_rn(c2);
#pragma unroll
		for (int c = 3; c < 16; ++c) out16[c] = c < nch ? in16[c - 3] : __half(0.f);
		*reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(out16);
		*reinterpret_cast<uint4 *>(dst + 8) = *reinterpret_cast<const uint4 *>(out16 + 8);
	} else {
		for (int c = nch - 1; c >= 3; --c) dst[c] = src[c - 3];
		dst[0] = __float2half_rn(c0);
		dst[1] = __float2half_rn(c1);
		dst[2] = __float2half_rn(c2);
	}
}

// ---------------------------------------------------------------------
// final epilogue: one thread per mid-resolution pixel (2H x 2W grid) ->
// a 2x2 block of HR pixels.
//   z   = tanh(conv_trans_2(mid) + bias)            models.py:573-583
//   up  = legacy bilinear x4 of cur (src = dst/4)   keras_layers.py:46-52
//   out = clip(up + z, -0.5, 0.5)                   models.py:588-593
//   u8  = trunc((out + 0.5) * 255), BGRX with X=0   keras_layers.py:227-230
//   pre_gen' = fp16(out)                            models.py:808-823
// ---------------------------------------------------------------------
__global__ void final_kernel(const __half *__restrict__ mid, const float *__restrict__ w2,
    const float *__restrict__ bias2, const FrameIO *__restrict__ io,
    __half *__restrict__ pre_gen_next, float *__restrict__ out_raw,
    const float *__restrict__ brightness, int h, int w) {
	// conv_trans_2 weights [q=i*2+j][o][c] and bias: 387 floats, broadcast reads
	__shared__ float c_w2[4 * 3 * 32];
	__shared__ float c_b2[3];
	for (int t = threadIdx.x; t < 4 * 3 * 32; t += blockDim.x) c_w2[t] = w2[t];
	if (threadIdx.x < 3) c_b2[threadIdx.x] = bias2[threadIdx.x];
	__syncthreads();
	int mx = blockIdx.x * blockDim.x + threadIdx.x;
	int my = blockIdx.y;
	int b = blockIdx.z;
	int mw = 2 * w, mh = 2 * h;
	if (mx >= mw) return;
	const FrameIO f = io[b];
	float br = brightness ? brightness[b] : 0.f;

	// 32 fp16 channels of this mid pixel
	float v[32];
	{
		const uint4 *p = reinterpret_cast<const uint4 *>(
		    mid + ((static_cast<size_t>(b) * mh + my) * mw + mx) * 32);
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			uint4 u = p[q];
			const __half2 *h2 = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
			for (int e = 0; e < 4; ++e) {
				float2 t = __half22float2(h2[e]);
				v[q * 8 + e * 2] = t.x;
				v[q * 8 + e * 2 + 1] = t.y;
			}
		}
	}
	int oh = 4 * h, ow = 4 * w;
#pragma unroll
	for (int i = 0; i < 2; ++i) {
		int Y = 2 * my + i;
		int y0 = Y >> 2;
		int y1 = min(y0 + 1, h - 1);
		float ty = static_cast<float>(Y & 3) * 0.25f;
		const uint8_t *row0 = f.in + y0 * f.in_stride;
		const uint8_t *row1 = f.in + y1 * f.in_stride;
		uchar4 px_out[2];
		__align__(8) __half st[2][4];
#pragma unroll
		for (int j = 0; j < 2; ++j) {
			int X = 2 * mx + j;
			int x0 = X >> 2;
			int x1 = min(x0 + 1, w - 1);
			float tx = static_cast<float>(X & 3) * 0.25f;
			uchar4 tl = *reinterpret_cast<const uchar4 *>(row0 + x0 * 4ll);
			uchar4 tr = *reinterpret_cast<const uchar4 *>(row0 + x1 * 4ll);
			uchar4 bl = *reinterpret_cast<const uchar4 *>(row1 + x0 * 4ll);
			uchar4 brr = *reinterpret_cast<const uchar4 *>(row1 + x1 * 4ll);
			const unsigned char *ptl = &tl.x, *ptr_ = &tr.x, *pbl = &bl.x, *pbr = &brr.x;
			unsigned char o8[3];
#pragma unroll
			for (int o = 0; o < 3; ++o) {
				const float *wq = c_w2 + ((i * 2 + j) * 3 + o) * 32;
				float acc = 0.f;
#pragma unroll
				for (int c = 0; c < 32; ++c) acc = fmaf(v[c], wq[c], acc);
				float z = tanhf(acc + c_b2[o]);
				float a = preprocess_px(ptl[o]), bq = preprocess_px(ptr_[o]);
				float c_ = preprocess_px(pbl[o]), d = preprocess_px(pbr[o]);
				float topv = __fadd_rn(a, __fmul_rn(__fsub_rn(bq, a), tx));
				float botv = __fadd_rn(c_, __fmul_rn(__fsub_rn(d, c_), tx));
				float up = __fadd_rn(topv, __fmul_rn(__fsub_rn(botv, topv), ty));
				float r = fminf(fmaxf(__fadd_rn(up, z), -0.5f), 0.5f);
				o8[o] = static_cast<unsigned char>(
				    static_cast<int>(__fmul_rn(__fadd_rn(r, 0.5f), 255.0f)));
				st[j][o] = __float2half_rn(r - br);
				if (out_raw) {
					out_raw[((static_cast<size_t>(b) * oh + Y) * ow + X) * 3 + o] = r;
				}
			}
			st[j][3] = __half(0.f);
			px_out[j] = make_uchar4(o8[0], o8[1], o8[2], 0);
		}
		// two adjacent HR pixels: 2 x 4 bytes of BGRX (caller images are only
		// guaranteed 4-byte aligned), 16 bytes of fp16 state
		uchar4 *orow = reinterpret_cast<uchar4 *>(f.out + Y * f.out_stride + (2 * mx) * 4ll);
		orow[0] = px_out[0];
		orow[1] = px_out[1];
		*reinterpret_cast<uint4 *>(
		    pre_gen_next + ((static_cast<size_t>(b) * oh + Y) * ow + 2 * mx) * 4) =
		    *reinterpret_cast<const uint4 *>(&st[0][0]);
	}
}

}  // namespace

cudaError_t launch_preprocess(const FrameIO *io, const __half *flow_prev, __half *flow_next,
    const float *brightness, int batch, int h, int w, int ph, int pw, int k, int cstride, cudaStream_t s) {
	dim3 block(128);
	dim3 grid((pw + 127) / 128, ph, batch);
	preprocess_kernel<<<grid, block, 0, s>>>(io, flow_prev, flow_next, brightness, h, w, ph, pw, k, cstride);
	return cudaGetLastError();
}

namespace {

// one block per stream; fixed thread-strided order + fixed tree => bit-stable
__global__ void __launch_bounds__(1024) brightness_kernel(const FrameIO *__restrict__ io, float *__restrict__ out,
    int h, int w) {
	__shared__ float part[1024];
	const FrameIO f = io[blockIdx.x];
	float acc = 0.f;
	const int n = h * w;
	for (int p = threadIdx.x; p < n; p += 1024) {
		const int y = p / w, x = p - y * w;
		const uchar4 px = *reinterpret_cast<const uchar4 *>(f.in + y * f.in_stride + x * 4ll);
		// BGR_LUMA = [0.1140, 0.5870, 0.2989] (utils.py:151), times 3 (models.py:776)
		acc += preprocess_px(px.x) * 0.1140f * 3.f + preprocess_px(px.y) * 0.5870f * 3.f +
		       preprocess_px(px.z) * 0.2989f * 3.f;
	}
	part[threadIdx.x] = acc;
	__syncthreads();
	for (int s = 512; s > 0; s >>= 1) {
		if (threadIdx.x < s) part[threadIdx.x] += part[threadIdx.x + s];
		__syncthreads();
	}
	if (threadIdx.x == 0) out[blockIdx.x] = part[0] / static_cast<float>(3 * n);
}

// both formulations of the u8 -> float conversion for every byte value (self-check entry)
__global__ void u8_table_kernel(float *fast, float *ieee) {
	fast[threadIdx.x] = preprocess_px(threadIdx.x);
	ieee[threadIdx.x] = preprocess_px_ieee(threadIdx.x);
}

}  // namespace

cudaError_t launch_u8_table(float *fast256, float *ieee256, cudaStream_t s) {
	u8_table_kernel<<<1, 256, 0, s>>>(fast256, ieee256);
	return cudaGetLastError();
}

cudaError_t launch_brightness(const FrameIO *io, float *out, int batch, int h, int w, cudaStream_t s) {
	brightness_kernel<<<batch, 1024, 0, s>>>(io, out, h, w);
	return cudaGetLastError();
}

cudaError_t launch_final(const __half *mid, const float *w2, const float *bias2, const FrameIO *io,
    __half *pre_gen_next, float *out_raw, const float *brightness, int batch, int h, int w,
    cudaStream_t s) {
	dim3 block(128);
	dim3 grid((2 * w + 127) / 128, 2 * h, batch);
	final_kernel<<<grid, block, 0, s>>>(mid, w2, bias2, io, pre_gen_next, out_raw, brightness, h, w);
	return cudaGetLastError();
}

}  // namespace ju
