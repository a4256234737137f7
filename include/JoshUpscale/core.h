// Public C++ runtime API of libJoshUpscale (B200 build).
//
// Source-compatible with the reference's core/public/JoshUpscale/core.h: the
// AviSynth plugin (avisynth_plugin/src/main.cc:57,144) and the OBS filter
// (obs_plugin/src/filter.cc:291,384) compile against this header unchanged.
// Every declaration cites the reference line it replaces.  The engine behind
// createRuntime() is hand-written sm_100a CUDA instead of a TensorRT engine;
// `modelPath` therefore names a `.jup` weight container, not a `.trt` file.
//
// Typical caller (what both plugins do, avisynth_plugin/src/main.cc:57-68, 125-148):
//
//     std::unique_ptr<core::Runtime> rt;
//     try {
//         rt.reset(core::createRuntime(/*deviceId=*/0, "model_psp.jup"));
//     } catch (...) {
//         report(core::getExceptionString());        // must be called inside the catch block
//     }
//     // frames must match the model: rt->getInputWidth() x rt->getInputHeight() in,
//     // 4x that out; BGRX, any signed byte stride, host or device memory
//     core::Image in{src, core::DataLocation::CPU, srcPitch, w, h};
//     core::Image out{dst, core::DataLocation::CPU, dstPitch, 4 * w, 4 * h};
//     rt->processImage(in, out);                      // synchronous; advances the recurrent state
//
// Behaviour that differs from the reference build is listed in INTEGRATION.md section 1
// (model file format, graphics-resource images, size mismatch reporting).  The same engine is
// reachable through a plain C ABI, include/joshupscale_c.h, for FFI callers.
#pragma once

#include <cstddef>
#include <cstdint>
#include <filesystem>
#include <string>

#include "JoshUpscale/core/export.h"

#ifdef _WIN32
struct ID3D11Texture2D;
struct ID3D11Device;
#endif

namespace JoshUpscale {
namespace core {

// ---- logging (reference core.h:21-28) ------------------------------------
enum class LogLevel : std::uint8_t { INFO, WARNING, ERROR };

struct LogSink {
	virtual void operator()(const char *tag, LogLevel logLevel, const std::string &message) = 0;
};

// The sink is borrowed, global and unsynchronised; it must outlive its use
// (reference core/src/logging.cc:61-62).  nullptr restores the console sink.
JOSHUPSCALE_EXPORT void setLogSink(LogSink *sink);

// ---- images (reference core.h:30-38) -------------------------------------
// Pixels are 4 bytes B,G,R,X.  `stride` is in bytes and may be negative
// (bottom-up RGB32: ptr addresses the LAST memory row, avisynth main.cc:125-142)
// or larger than width*4.  X is ignored on input and written as 0 on output.
enum class DataLocation : std::uint8_t { CPU, CUDA, GRAPHICS_RESOURCE };

struct Image {
	void *ptr;
	DataLocation location;
	std::ptrdiff_t stride;
	std::size_t width;
	std::size_t height;
};

// ---- graphics interop (reference core.h:40-62) ---------------------------
// getGLImage() registers a GL_TEXTURE_2D of the calling thread's current GL context with CUDA
// (the GL entry points are resolved from the process's libGL at call time, the library itself
// links no GL); the returned image (location GRAPHICS_RESOURCE) can be passed to processImage().
// Without a current context / GL library both functions throw, like the reference's.
enum class GraphicsResourceImageType : std::uint8_t { INPUT, OUTPUT };

struct GraphicsResourceImage {
	virtual ~GraphicsResourceImage() {}
	Image getImage() const { return m_Image; }

protected:
	Image m_Image = {};
};

#ifdef _WIN32
// reference core.h:54-59; the Windows / D3D11 build of this library is not provided
JOSHUPSCALE_EXPORT int getD3D11DeviceIndex(ID3D11Device *d3d11Device);
JOSHUPSCALE_EXPORT GraphicsResourceImage *getD3D11Image(ID3D11Texture2D *d3d11Texture, GraphicsResourceImageType type);
#endif

JOSHUPSCALE_EXPORT int getGLDeviceIndex();
JOSHUPSCALE_EXPORT GraphicsResourceImage *getGLImage(std::uint32_t image, GraphicsResourceImageType type);

// ---- runtime (reference core.h:64-92) ------------------------------------
// One Runtime = one device, one CUDA stream, one recurrent state.  Calls on
// one instance must be serialised; processImage() is synchronous: the output
// image is complete when it returns.  A fresh runtime starts from all-zero
// recurrent state (reference core/include/JoshUpscale/core/cuda.h:69-72).
struct Runtime {
	virtual ~Runtime() {}

	virtual void processImage(const Image &inputImage, const Image &outputImage) = 0;

	std::size_t getInputWidth() const { return m_InputWidth; }
	std::size_t getInputHeight() const { return m_InputHeight; }
	std::size_t getOutputWidth() const { return m_OutputWidth; }
	std::size_t getOutputHeight() const { return m_OutputHeight; }

protected:
	std::size_t m_InputWidth = 0;
	std::size_t m_InputHeight = 0;
	std::size_t m_OutputWidth = 0;
	std::size_t m_OutputHeight = 0;
};

// Creates a runtime on CUDA device `deviceId` from the weight container at `modelPath`:
// loads and folds the weights, allocates the activation / state buffers for one stream and
// captures the frame's CUDA graphs, so the first processImage() call is as fast as any other.
// Throws (std::invalid_argument for a bad device index, ju::ModelException for an unreadable or
// unsupported model - e.g. a TensorRT `.trt` engine -, ju::CudaException for CUDA failures).
// Several runtimes may coexist, also on the same device; each owns its stream and state.
// Caller owns the returned pointer (delete through the virtual destructor).
JOSHUPSCALE_EXPORT Runtime *createRuntime(int deviceId, const std::filesystem::path &modelPath);

// Must be called from inside a catch block: re-inspects the in-flight
// exception and formats nested chains as "Type: what\n  Type: what"
// (reference core/src/core.cc:37-41, core/src/exception.cc:51-79).
JOSHUPSCALE_EXPORT std::string getExceptionString();

}  // namespace core
}  // namespace JoshUpscale
