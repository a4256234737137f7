// C entry points around the REFERENCE's image <-> tensor conversion (core/src/cuda_convert.cc.cu,
// compiled unmodified by oracle/Makefile into oracle/_ref/libref_convert.so) so that tests can
// run it on the GPU next to this repo's pixel kernels.  TEST INFRASTRUCTURE ONLY.
//
//   ref_image_to_tensor : what TensorRTBackend::process does to the input image
//                         (core/src/tensorrt_backend.cc:270-278): BGRX u8, any stride ->
//                         [H, W, 3] float32 / float16 engine input
//   ref_tensor_to_image : engine output [H, W, 3] float32 -> BGRX u8 image
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <exception>
#include <vector>

#include "JoshUpscale/core/cuda_convert.h"

using namespace JoshUpscale::core;

namespace {
template <typename F>
int guarded(F &&f) {
	try {
		f();
		return 0;
	} catch (const std::exception &e) {
		std::fprintf(stderr, "ref_convert: %s\n", e.what());
		return 1;
	} catch (...) {
		return 1;
	}
}
}  // namespace

extern "C" __attribute__((visibility("default"))) int ref_image_to_tensor(const void *host_image, long long stride,
    int width, int height, int half_precision, void *host_out) {
	return guarded([&] {
		Image img{const_cast<void *>(host_image), DataLocation::CPU, static_cast<std::ptrdiff_t>(stride),
		    static_cast<std::size_t>(width), static_cast<std::size_t>(height)};
		const std::size_t n = static_cast<std::size_t>(width) * height * 3;
		cuda::CudaStream stream;
		// the staging buffer holds the 4-byte pixels (cuda_convert.cc.cu:365); the reference also
		// requires width * height to be a multiple of 32 (:186)
		cuda::CudaBuffer<std::uint8_t> internal(static_cast<std::size_t>(width) * height * 4);
		GenericTensor from(img);
		if (half_precision) {
			cuda::CudaBuffer<__half> to(n);
			cuda::cudaConvert(from, to, internal, stream);
			stream.synchronize();
			cuda::cudaCheck(::cudaMemcpy(host_out, to.get(), n * sizeof(__half), ::cudaMemcpyDeviceToHost));
		} else {
			cuda::CudaBuffer<float> to(n);
			cuda::cudaConvert(from, to, internal, stream);
			stream.synchronize();
			cuda::cudaCheck(::cudaMemcpy(host_out, to.get(), n * sizeof(float), ::cudaMemcpyDeviceToHost));
		}
	});
}

extern "C" __attribute__((visibility("default"))) int ref_tensor_to_image(const float *host_in, int width, int height,
    void *host_image, long long stride) {
	return guarded([&] {
		Image img{host_image, DataLocation::CPU, static_cast<std::ptrdiff_t>(stride), static_cast<std::size_t>(width),
		    static_cast<std::size_t>(height)};
		const std::size_t n = static_cast<std::size_t>(width) * height * 3;
		cuda::CudaStream stream;
		cuda::CudaBuffer<std::uint8_t> internal(static_cast<std::size_t>(width) * height * 4);
		cuda::CudaBuffer<float> from(n);
		cuda::cudaCheck(::cudaMemcpy(from.get(), host_in, n * sizeof(float), ::cudaMemcpyHostToDevice));
		GenericTensor to(img);
		cuda::cudaConvert(from, to, internal, stream);
		stream.synchronize();
	});
}
