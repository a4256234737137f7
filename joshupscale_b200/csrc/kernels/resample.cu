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

// One thread = 8 channels of one INPUT pixel -> the 2x2 block of output pixels
// it anchors.  With scale 2 the legacy (asymmetric) bilinear weights are 0 or
// 0.5, so out(2y,2x) = in(y,x) exactly and the other three are single lerps of
// the same four inputs (a b / c d); the reference's operation order
// (top = tl + (tr-tl)*tx ; bot likewise ; top + (bot-top)*ty, fp32) is kept, so
// results are bit-identical to evaluating the general formula per output.
__global__ void upscale2_kernel(const __half *__restrict__ in, __half *__restrict__ out, int h,
    int w, int c8, size_t total) {
	size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
	if (idx >= total) return;
	int cv = idx % c8;
	size_t p = idx / c8;
	int x = p % w;
	size_t q = p / w;
	int y = q % h;
	size_t b = q / h;
	int y1 = min(y + 1, h - 1), x1 = min(x + 1, w - 1);
	const uint4 *src = reinterpret_cast<const uint4 *>(in);
	const uint4 va = src[((b * h + y) * w + x) * c8 + cv];
	const uint4 vb = src[((b * h + y) * w + x1) * c8 + cv];
	const uint4 vc = src[((b * h + y1) * w + x) * c8 + cv];
	const uint4 vd = src[((b * h + y1) * w + x1) * c8 + cv];
	const __half *pa = reinterpret_cast<const __half *>(&va);
	const __half *pb = reinterpret_cast<const __half *>(&vb);
	const __half *pc = reinterpret_cast<const __half *>(&vc);
	const __half *pd = reinterpret_cast<const __half *>(&vd);
	__align__(16) __half r01[8], r10[8], r11[8];
#pragma unroll
	for (int e = 0; e < 8; ++e) {
		const float a = __half2float(pa[e]), bq = __half2float(pb[e]);
		const float c = __half2float(pc[e]), d = __half2float(pd[e]);
		const float top = __fadd_rn(a, __fmul_rn(__fsub_rn(bq, a), 0.5f));   // tx = 0.5
		const float bot = __fadd_rn(c, __fmul_rn(__fsub_rn(d, c), 0.5f));
		r01[e] = __float2half_rn(top);                                        // (2y, 2x+1): ty = 0
		r10[e] = __float2half_rn(__fadd_rn(a, __fmul_rn(__fsub_rn(c, a), 0.5f)));  // (2y+1, 2x): tx = 0
		r11[e] = __float2half_rn(__fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), 0.5f)));
	}
	uint4 *dst = reinterpret_cast<uint4 *>(out);
	const size_t ow = 2 * static_cast<size_t>(w);
	const size_t o00 = ((b * 2 * h + 2 * y) * ow + 2 * x) * c8 + cv;
	dst[o00] = va;                                            // (2y, 2x): tx = ty = 0 -> in(y, x)
	dst[o00 + c8] = *reinterpret_cast<const uint4 *>(r01);
	dst[o00 + ow * c8] = *reinterpret_cast<const uint4 *>(r10);
	dst[o00 + ow * c8 + c8] = *reinterpret_cast<const uint4 *>(r11);
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
	size_t total = static_cast<size_t>(batch) * h * w * (c / 8);  // one thread per input pixel-vector
	upscale2_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(in, out, h, w, c / 8, total);
	return cudaGetLastError();
}

}  // namespace ju
