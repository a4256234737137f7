// Output temporal filter with scene-cut gate of the deployed graphs
// (scripts/inference/onnx/frame_moving_avg.py:142-307), as two HBM-bound passes
// over tensors the frame already produced:
//
//   out  = generator output after Clip, fp16 [batch, 4h, 4w, 4]   (written by the tail kernel)
//   pw   = warped previous output as fed to the generator, i.e. the space-to-depth
//          channels 3.. of the generator input, fp16 [batch, h, w, 64]
//
//   pass 1 (stats):  d = |out - pw| or (out - pw)^2, optionally luma weighted;
//                    window == 0: deterministic per-stream partial sums
//                    window  > 0: one value per w x w cell (zero padded, centred),
//                                 cond = sign / tanh(mean * gain - threshold * gain)
//   pass 2 (blend):  cond (global scalar, or the cell map resized bilinearly,
//                    "asymmetric" coordinates) -> mask = cond*(-s/2) + s/2,
//                    mask2 = cond*(s/2) + 1 - s/2, final = pw*mask + out*mask2;
//                    u8 BGRX = trunc((final + 0.5) * 255) and the fp16 recurrent
//                    state (minus brightness) are written from `final`.
//
// One thread owns one LOW-RES pixel = a 4x4 block of output pixels, so the 96
// bytes of pw it needs are one contiguous S2D row and its stores are whole
// 16-byte (u8) / 32-byte (state) segments that neighbouring lanes extend to
// full lines.
#include "kernels.h"

namespace ju {

namespace {

// LUMA_NORM (frame_moving_avg.py:95-96), B, G, R, already times 3, rounded to fp32 like numpy does
__device__ __forceinline__ float luma_w(int c, int l2) {
	const float l = c == 0 ? 0.1140f * 3.f : (c == 1 ? 0.5870f * 3.f : 0.2989f * 3.f);
	return l2 ? l * l : l;
}

struct LrBlock {
	float out[4][4][3];
	float pw[4][4][3];
};

// loads the 4x4 output pixels and the 48 warped values of low-res pixel (b, y, x)
__device__ __forceinline__ void load_block(const __half *__restrict__ out_raw, const __half *__restrict__ gen_in,
    int b, int y, int x, int h, int w, int limit, LrBlock &blk) {
	const __half *g = gen_in + ((static_cast<size_t>(b) * h + y) * w + x) * 64;
	__align__(16) __half gv[64];
#pragma unroll
	for (int q = 0; q < 8; ++q) reinterpret_cast<uint4 *>(gv)[q] = reinterpret_cast<const uint4 *>(g)[q];
#pragma unroll
	for (int i = 0; i < 4; ++i) {
		const __half *o = out_raw + ((static_cast<size_t>(b) * 4 * h + 4 * y + i) * 4 * w + 4 * x) * 4;
		__align__(16) __half ov[16];
		reinterpret_cast<uint4 *>(ov)[0] = reinterpret_cast<const uint4 *>(o)[0];
		reinterpret_cast<uint4 *>(ov)[1] = reinterpret_cast<const uint4 *>(o)[1];
#pragma unroll
		for (int j = 0; j < 4; ++j) {
#pragma unroll
			for (int c = 0; c < 3; ++c) {
				blk.out[i][j][c] = __half2float(ov[j * 4 + c]);
				float p = __half2float(gv[3 + (i * 4 + j) * 3 + c]);
				if (limit) p = fmaxf(fminf(p, 0.5f), -0.5f);
				blk.pw[i][j][c] = p;
			}
		}
	}
}

__device__ __forceinline__ float diff_norm(float o, float p, int l2) {
	const float d = __fsub_rn(o, p);
	return l2 ? __fmul_rn(d, d) : fabsf(d);
}

// ---- window == 0: per-stream partial sums, fixed summation order -----------------
constexpr int kStatThreads = 256;

__global__ void __launch_bounds__(kStatThreads)
filter_stats_kernel(const __half *__restrict__ out_raw, const __half *__restrict__ gen_in,
    float *__restrict__ partial, int h, int w, FilterParams fp) {
	const int b = blockIdx.y;
	const int idx = blockIdx.x * kStatThreads + threadIdx.x;
	float acc = 0.f;
	if (idx < h * w) {
		LrBlock blk;
		load_block(out_raw, gen_in, b, idx / w, idx % w, h, w, fp.limit, blk);
#pragma unroll
		for (int i = 0; i < 4; ++i)
#pragma unroll
			for (int j = 0; j < 4; ++j)
#pragma unroll
				for (int c = 0; c < 3; ++c) {
					float d = diff_norm(blk.out[i][j][c], blk.pw[i][j][c], fp.norm_l2);
					if (fp.luma) d = __fmul_rn(d, luma_w(c, fp.norm_l2));
					acc = __fadd_rn(acc, d);
				}
	}
	__shared__ float red[kStatThreads];
	red[threadIdx.x] = acc;
	__syncthreads();
	for (int s = kStatThreads / 2; s > 0; s >>= 1) {
		if (threadIdx.x < s) red[threadIdx.x] = __fadd_rn(red[threadIdx.x], red[threadIdx.x + s]);
		__syncthreads();
	}
	if (threadIdx.x == 0) partial[static_cast<size_t>(b) * gridDim.x + blockIdx.x] = red[0];
}

// one block per stream: cond[b] from the partial sums (fixed order)
__global__ void filter_cond_kernel(const float *__restrict__ partial, float *__restrict__ cond, int n_partial,
    float inv_count, FilterParams fp) {
	const int b = blockIdx.x;
	__shared__ float red[256];
	float acc = 0.f;
	for (int i = threadIdx.x; i < n_partial; i += 256) acc = __fadd_rn(acc, partial[static_cast<size_t>(b) * n_partial + i]);
	red[threadIdx.x] = acc;
	__syncthreads();
	for (int s = 128; s > 0; s >>= 1) {
		if (threadIdx.x < s) red[threadIdx.x] = __fadd_rn(red[threadIdx.x], red[threadIdx.x + s]);
		__syncthreads();
	}
	if (threadIdx.x == 0) {
		const float gain_coef = fp.gain == 0.f ? 1.f : fp.gain;
		float m = __fmul_rn(red[0], inv_count);
		if (fp.luma || fp.gain != 0.f) m = __fmul_rn(m, gain_coef);
		const float th = __fadd_rn(m, -fp.threshold * gain_coef);
		cond[b] = fp.gain == 0.f ? (th > 0.f ? 1.f : (th < 0.f ? -1.f : 0.f)) : tanhf(th);
	}
}

// ---- window > 0: one warp per cell ------------------------------------------------
__global__ void filter_cells_kernel(const __half *__restrict__ out_raw, const __half *__restrict__ gen_in,
    float *__restrict__ cells, int h, int w, int cells_y, int cells_x, int pad_t, int pad_l, FilterParams fp) {
	const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	const int b = blockIdx.y;
	if (warp >= cells_y * cells_x) return;
	const int cy = warp / cells_x, cx = warp % cells_x;
	const int wnd = fp.window;
	const int H4 = 4 * h, W4 = 4 * w;
	float acc = 0.f;
	for (int p = lane; p < wnd * wnd; p += 32) {
		const int Y = cy * wnd + p / wnd - pad_t, X = cx * wnd + p % wnd - pad_l;
		if (Y < 0 || Y >= H4 || X < 0 || X >= W4) continue;  // zero padding of the Conv
		const __half *o = out_raw + ((static_cast<size_t>(b) * H4 + Y) * W4 + X) * 4;
		const __half *g = gen_in + ((static_cast<size_t>(b) * h + (Y >> 2)) * w + (X >> 2)) * 64 + 3 +
		                  ((Y & 3) * 4 + (X & 3)) * 3;
#pragma unroll
		for (int c = 0; c < 3; ++c) {
			float pv = __half2float(g[c]);
			if (fp.limit) pv = fmaxf(fminf(pv, 0.5f), -0.5f);
			float d = diff_norm(__half2float(o[c]), pv, fp.norm_l2);
			if (fp.luma) d = __fmul_rn(d, luma_w(c, fp.norm_l2));
			acc = __fadd_rn(acc, d);
		}
	}
#pragma unroll
	for (int s = 16; s > 0; s >>= 1) acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, s));
	if (lane == 0) {
		const float gain_coef = fp.gain == 0.f ? 1.f : fp.gain;
		const float m = __fmul_rn(acc, gain_coef / (3.f * wnd * wnd));
		const float th = __fadd_rn(m, -fp.threshold * gain_coef);
		cells[(static_cast<size_t>(b) * cells_y + cy) * cells_x + cx] =
		    fp.gain == 0.f ? (th > 0.f ? 1.f : (th < 0.f ? -1.f : 0.f)) : tanhf(th);
	}
}

// ---- blend + pack ------------------------------------------------------------------
__global__ void __launch_bounds__(128)
filter_blend_kernel(const __half *__restrict__ out_raw, const __half *__restrict__ gen_in,
    const float *__restrict__ cond, const float *__restrict__ cells, const FrameIO *__restrict__ io,
    __half *__restrict__ pre_gen_next, const float *__restrict__ brightness, int h, int w, int cells_y,
    int cells_x, int pad_t, int pad_l, FilterParams fp) {
	const int x = blockIdx.x * blockDim.x + threadIdx.x;
	const int y = blockIdx.y, b = blockIdx.z;
	if (x >= w) return;
	const FrameIO f = io[b];
	const float br = brightness ? brightness[b] : 0.f;
	const float c1 = fp.strength * 0.5f, c2 = -fp.strength * 0.5f, c3 = fp.c3;
	LrBlock blk;
	load_block(out_raw, gen_in, b, y, x, h, w, fp.limit, blk);
	const float cglobal = fp.window == 0 ? cond[b] : 0.f;
#pragma unroll
	for (int i = 0; i < 4; ++i) {
		const int Y = 4 * y + i;
		int cy0 = 0, cy1 = 0;
		float ty = 0.f;
		if (fp.window) {
			const float src = __fdiv_rn(static_cast<float>(Y + pad_t), static_cast<float>(fp.window));
			const float lo = floorf(src);
			ty = __fsub_rn(src, lo);
			cy0 = static_cast<int>(lo);
			cy1 = min(cy0 + 1, cells_y - 1);
		}
		__align__(16) unsigned char px[16];
		__align__(16) __half st[16];
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			float cnd = cglobal;
			if (fp.window) {
				const int X = 4 * x + j;
				const float src = __fdiv_rn(static_cast<float>(X + pad_l), static_cast<float>(fp.window));
				const float lo = floorf(src);
				const float tx = __fsub_rn(src, lo);
				const int cx0 = static_cast<int>(lo), cx1 = min(cx0 + 1, cells_x - 1);
				const float *row0 = cells + (static_cast<size_t>(b) * cells_y + cy0) * cells_x;
				const float *row1 = cells + (static_cast<size_t>(b) * cells_y + cy1) * cells_x;
				const float top = __fadd_rn(__fmul_rn(row0[cx0], __fsub_rn(1.f, tx)), __fmul_rn(row0[cx1], tx));
				const float bot = __fadd_rn(__fmul_rn(row1[cx0], __fsub_rn(1.f, tx)), __fmul_rn(row1[cx1], tx));
				cnd = __fadd_rn(__fmul_rn(top, __fsub_rn(1.f, ty)), __fmul_rn(bot, ty));
			}
			const float mask = __fadd_rn(__fmul_rn(cnd, c2), c1);
			const float mask2 = __fadd_rn(__fmul_rn(cnd, c1), c3);
#pragma unroll
			for (int c = 0; c < 3; ++c) {
				const float r = __fadd_rn(__fmul_rn(blk.pw[i][j][c], mask), __fmul_rn(blk.out[i][j][c], mask2));
				px[j * 4 + c] = static_cast<unsigned char>(static_cast<int>(__fmul_rn(__fadd_rn(r, 0.5f), 255.0f)));
				st[j * 4 + c] = __float2half_rn(r - br);
			}
			px[j * 4 + 3] = 0;
			st[j * 4 + 3] = __float2half_rn(0.f);
		}
		// caller images are only guaranteed 4-byte aligned (pointer and stride): four pixel stores
		uchar4 *odst = reinterpret_cast<uchar4 *>(f.out + static_cast<long long>(Y) * f.out_stride + 16ll * x);
#pragma unroll
		for (int j = 0; j < 4; ++j) odst[j] = reinterpret_cast<const uchar4 *>(px)[j];
		uint4 *sdst = reinterpret_cast<uint4 *>(
		    pre_gen_next + ((static_cast<size_t>(b) * 4 * h + Y) * 4 * w + 4 * x) * 4);
		sdst[0] = reinterpret_cast<const uint4 *>(st)[0];
		sdst[1] = reinterpret_cast<const uint4 *>(st)[1];
	}
}

}  // namespace

void filter_geometry(const FilterParams &fp, int h, int w, int *cells_y, int *cells_x, int *pad_t, int *pad_l) {
	if (fp.window <= 0) {
		*cells_y = *cells_x = *pad_t = *pad_l = 0;
		return;
	}
	const int H4 = 4 * h, W4 = 4 * w, wnd = fp.window;
	const int oh = (H4 + wnd - 1) / wnd * wnd, ow = (W4 + wnd - 1) / wnd * wnd;
	*cells_y = oh / wnd;
	*cells_x = ow / wnd;
	*pad_t = (oh - H4) / 2;
	*pad_l = (ow - W4) / 2;
}

int filter_partials_per_stream(int h, int w) { return (h * w + kStatThreads - 1) / kStatThreads; }

cudaError_t launch_frame_filter(const __half *out_raw, const __half *gen_in, const FrameIO *io, __half *pre_gen_next,
    const float *brightness, float *scratch, const FilterParams &fp, int batch, int h, int w, cudaStream_t s) {
	int cells_y, cells_x, pad_t, pad_l;
	filter_geometry(fp, h, w, &cells_y, &cells_x, &pad_t, &pad_l);
	float *cond = scratch;            // [batch]
	float *work = scratch + batch;    // partial sums or the cell map
	if (fp.window == 0) {
		const int np = filter_partials_per_stream(h, w);
		filter_stats_kernel<<<dim3(np, batch), kStatThreads, 0, s>>>(out_raw, gen_in, work, h, w, fp);
		filter_cond_kernel<<<batch, 256, 0, s>>>(work, cond, np, 1.0f / (48.0f * h * w), fp);
	} else {
		const int n_cells = cells_y * cells_x;
		filter_cells_kernel<<<dim3((n_cells * 32 + 255) / 256, batch), 256, 0, s>>>(out_raw, gen_in, work, h, w,
		    cells_y, cells_x, pad_t, pad_l, fp);
	}
	filter_blend_kernel<<<dim3((w + 127) / 128, h, batch), 128, 0, s>>>(out_raw, gen_in, cond, work, io,
	    pre_gen_next, brightness, h, w, cells_y, cells_x, pad_t, pad_l, fp);
	return cudaGetLastError();
}

}  // namespace ju
