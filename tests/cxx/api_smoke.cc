// Caller-side smoke test of the C++ API, written the way the reference's
// plugins use it (avisynth_plugin/src/main.cc:40-61, 110-148): unique_ptr
// ownership, getters for sizes, bottom-up RGB32 frames passed as last-row
// pointer + negative stride, exceptions formatted by getExceptionString()
// inside the catch block, a custom LogSink.
//
// usage: api_smoke <model.jup> <frames.bin> <nframes> <out.bin>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <memory>
#include <vector>

#include "JoshUpscale/core.h"

using namespace JoshUpscale;

struct CountingSink : core::LogSink {
	int count = 0;
	void operator()(const char *tag, core::LogLevel level, const std::string &message) override {
		++count;
		std::cerr << "[sink] " << tag << " " << static_cast<int>(level) << " " << message << "\n";
	}
};

int main(int argc, char **argv) {
	if (argc != 5) return 2;
	CountingSink sink;
	core::setLogSink(&sink);

	// error path first: must throw, and getExceptionString must describe it
	try {
		std::unique_ptr<core::Runtime> bad(core::createRuntime(0, "/nonexistent/model.jup"));
		std::cerr << "expected an exception\n";
		return 3;
	} catch (...) {
		std::string msg = core::getExceptionString();
		std::cout << "error: " << msg << "\n";
		if (msg.find("model.jup") == std::string::npos) return 4;
	}

	// graphics interop on a machine without a GL context (this box is headless): both entry points
	// must fail with an exception, like the reference's, never crash or return a dummy
	try {
		std::unique_ptr<core::GraphicsResourceImage> tex(core::getGLImage(1, core::GraphicsResourceImageType::INPUT));
		std::cout << "getGLImage succeeded: a GL context is current\n";
	} catch (...) {
		std::cout << "getGLImage: " << core::getExceptionString() << "\n";
	}
	try {
		std::cout << "GL device " << core::getGLDeviceIndex() << "\n";
	} catch (...) {
		std::cout << "getGLDeviceIndex: " << core::getExceptionString() << "\n";
	}

	std::unique_ptr<core::Runtime> runtime;
	try {
		runtime.reset(core::createRuntime(0, argv[1]));
	} catch (...) {
		std::cerr << core::getExceptionString() << "\n";
		return 5;
	}
	const std::size_t w = runtime->getInputWidth(), h = runtime->getInputHeight();
	const std::size_t ow = runtime->getOutputWidth(), oh = runtime->getOutputHeight();
	if (ow != 4 * w || oh != 4 * h) return 6;
	const int n = std::atoi(argv[3]);
	std::ifstream in(argv[2], std::ios::binary);
	std::ofstream out(argv[4], std::ios::binary);
	std::vector<std::uint8_t> frame(w * h * 4), flipped(w * h * 4), result(ow * oh * 4), unflipped(ow * oh * 4);
	for (int t = 0; t < n; ++t) {
		in.read(reinterpret_cast<char *>(frame.data()), static_cast<std::streamsize>(frame.size()));
		// store bottom-up like AviSynth RGB32
		for (std::size_t y = 0; y < h; ++y)
			std::copy_n(frame.data() + y * w * 4, w * 4, flipped.data() + (h - 1 - y) * w * 4);
		core::Image src{flipped.data() + (h - 1) * w * 4, core::DataLocation::CPU,
		    -static_cast<std::ptrdiff_t>(w * 4), w, h};
		core::Image dst{result.data() + (oh - 1) * ow * 4, core::DataLocation::CPU,
		    -static_cast<std::ptrdiff_t>(ow * 4), ow, oh};
		try {
			runtime->processImage(src, dst);
		} catch (...) {
			std::cerr << core::getExceptionString() << "\n";
			return 7;
		}
		for (std::size_t y = 0; y < oh; ++y)
			std::copy_n(result.data() + (oh - 1 - y) * ow * 4, ow * 4, unflipped.data() + y * ow * 4);
		out.write(reinterpret_cast<const char *>(unflipped.data()), static_cast<std::streamsize>(unflipped.size()));
	}
	core::setLogSink(nullptr);
	std::cout << "ok " << n << " frames, sink messages " << sink.count << "\n";
	return sink.count > 0 ? 0 : 8;
}
