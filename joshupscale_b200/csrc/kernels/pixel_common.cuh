// Pixel arithmetic shared by the pixel-I/O, warp and tail kernels.
#pragma once

namespace ju {

// PreprocessLayer (scripts/training/keras_layers.py:195-208): float32(u8) / 255 - 0.5.
// The quotient must be the correctly rounded IEEE division (numpy / TF compute x / 255 in fp32),
// but an IEEE divide costs ~10 instructions incl. a slow-path check and these kernels are
// instruction-bound.  q0 = x * fl(1/255) followed by one FMA residual step is the correctly
// rounded quotient for every x in 0..255 (checked exhaustively against __fdiv_rn on the device by
// tests/test_gpu_kernels.py::test_u8_conversion_matches_ieee_division_for_all_bytes).
__device__ __forceinline__ float preprocess_px(unsigned int v) {
	const float x = static_cast<float>(v);
	const float r = __int_as_float(0x3b808081);  // fl(1 / 255)
	const float q = __fmul_rn(x, r);
	const float rem = __fmaf_rn(-q, 255.0f, x);
	return __fsub_rn(__fmaf_rn(rem, r, q), 0.5f);
}

// reference formulation, kept for the exhaustive on-device comparison
__device__ __forceinline__ float preprocess_px_ieee(unsigned int v) {
	return __fsub_rn(__fdiv_rn(static_cast<float>(v), 255.0f), 0.5f);
}

}  // namespace ju
