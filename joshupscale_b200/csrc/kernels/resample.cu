// MaxPool2D(2) and legacy bilinear x2 of NHWC fp16 feature maps (HBM-bound).
//
// Replaces, inside the reference's TensorRT engine, layers.MaxPool2D(pool_size=2)
// (scripts/training/models.py:406-409) and UpscaleLayer(resize_type="bilinear",
// scale=2, dtype="float32") = tf.compat.v1.image.resize_bilinear(
// align_corners=False, half_pixel_centers=False) (models.py:441-446;
// keras_layers.py:46-52): src = dst/2, hi = min(lo+1, size-1),
// value = top + (bottom - top)*ty with top = tl + (tr - tl)*tx, in fp32.
#include "kernels.h"

namespace ju {

namespace {

// one thread = 8 channels (16 bytes) of one output pixel
__global__ void maxpool2_kernel(const __half *__restrict__ in, __half *__restrict__ out, int h,
    int w, int c8, size_t total) {
	size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
	if (idx >= total) return;
	int cv = idx % c8;
	size_t p = idx / c8;
	int ow = w / 2, oh = h / 2;
	int ox = p % ow;
	size_t q = p / ow;
	int oy = q % oh;
	size_t b = q / oh;
	const uint4 *src = reinterpret_cast<const uint4 *>(in);
	size_t base = ((b * h + 2 * oy) * w + 2 * ox) * c8 + cv;
	uint4 a = src[base], bq = src[base + c8], c = src[base + static_cast<size_t>(w) * c8],
	      d = src[base + static_cast<size_t>(w) * c8 + c8];
	uint4 r;
	const __half2 *pa = reinterpret_cast<const __half2 *>(&a);
	const __half2 *pb = reinterpret_cast<const __half2 *>(&bq);
	const __half2 *pc = reinterpret_cast<const __half2 *>(&c);
	const __half2 *pd = reinterpret_cast<const __half2 *>(&d);
	__half2 *pr = reinterpret_cast<__half2 *>(&r);
#pragma unroll
	for (int e = 0; e < 4; ++e) pr[e] = __hmax2(__hmax2(pa[e], pb[e]), __hmax2(pc[e], pd[e]));
	reinterpret_cast<uint4 *>(out)[idx] = r;
}

__global__ void upscale2_kernel(const __half *__restrict__ in, __half *__restrict__ out, int h,
    int w, int c8, size_t total) {
	size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
	if (idx >= total) return;
	int cv = idx % c8;
	size_t p = idx / c8;
	int ow = 2 * w, oh = 2 * h;
	int ox = p % ow;
	size_t q = p / ow;
	int oy = q % oh;
	size_t b = q / oh;
	int y0 = oy >> 1, x0 = ox >> 1;
	int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
	float ty = (oy & 1) * 0.5f, tx = (ox & 1) * 0.5f;
	const uint4 *src = reinterpret_cast<const uint4 *>(in);
	uint4 tl = src[((b * h + y0) * w + x0) * c8 + cv];
	uint4 tr = src[((b * h + y0) * w + x1) * c8 + cv];
	uint4 bl = src[((b * h + y1) * w + x0) * c8 + cv];
	uint4 br = src[((b * h + y1) * w + x1) * c8 + cv];
	const __half *ptl = reinterpret_cast<const __half *>(&tl);
	const __half *ptr_ = reinterpret_cast<const __half *>(&tr);
	const __half *pbl = reinterpret_cast<const __half *>(&bl);
	const __half *pbr = reinterpret_cast<const __half *>(&br);
	__align__(16) __half r[8];
#pragma unroll
	for (int e = 0; e < 8; ++e) {
		float a = __half2float(ptl[e]), bq = __half2float(ptr_[e]);
		float c = __half2float(pbl[e]), d = __half2float(pbr[e]);
		float topv = __fadd_rn(a, __fmul_rn(__fsub_rn(bq, a), tx));
		float botv = __fadd_rn(c, __fmul_rn(__fsub_rn(d, c), tx));
		r[e] = __float2half_rn(__fadd_rn(topv, __fmul_rn(__fsub_rn(botv, topv), ty)));
	}
	reinterpret_cast<uint4 *>(out)[idx] = *reinterpret_cast<const uint4 *>(r);
}

}  // namespace

cudaError_t launch_maxpool2(const __half *in, __half *out, int batch, int h, int w, int c, cudaStream_t s) {
	if (c % 8 || h % 2 || w % 2) return cudaErrorInvalidValue;
	size_t total = static_cast<size_t>(batch) * (h / 2) * (w / 2) * (c / 8);
	maxpool2_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(in, out, h, w, c / 8, total);
	return cudaGetLastError();
}

cudaError_t launch_upscale2(const __half *in, __half *out, int batch, int h, int w, int c, cudaStream_t s) {
	if (c % 8) return cudaErrorInvalidValue;
	size_t total = static_cast<size_t>(batch) * (2 * h) * (2 * w) * (c / 8);
	upscale2_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(in, out, h, w, c / 8, total);
	return cudaGetLastError();
}

}  // namespace ju
