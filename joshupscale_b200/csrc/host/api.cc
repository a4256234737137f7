// Public C++ API (include/JoshUpscale/core.h) and C-ABI (include/joshupscale_c.h)
// over the sm_100a engine.
#include <cuda_fp16.h>
#include <cxxabi.h>
#include <dlfcn.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <exception>
#include <memory>
#include <typeinfo>

#include "JoshUpscale/core.h"
#include "common.h"
#include "engine.h"
#include "joshupscale_c.h"

namespace ju {

// ---------------------------------------------------------------------------
// exception pretty-printer
// ---------------------------------------------------------------------------
namespace {

std::string demangle(const char *name) {
	int status = 1;
	std::unique_ptr<char, void (*)(void *)> res{abi::__cxa_demangle(name, nullptr, nullptr, &status), std::free};
	return status == 0 ? res.get() : name;
}

void appendCurrent(std::ostringstream &os, int depth);

void appendException(std::ostringstream &os, const std::exception &e, int depth) {
	os << demangle(typeid(e).name()) << ": " << e.what();
	try {
		std::rethrow_if_nested(e);
	} catch (...) {
		os << "\n";
		for (int i = 0; i <= depth; ++i) os << "  ";
		appendCurrent(os, depth + 1);
	}
}

void appendCurrent(std::ostringstream &os, int depth) {
	try {
		throw;
	} catch (const std::exception &e) {
		appendException(os, e, depth);
	} catch (...) {
		os << "Unknown error";
	}
}

}  // namespace

std::string currentExceptionString() {
	std::ostringstream os;
	if (!std::current_exception()) return "No active exception";
	appendCurrent(os, 0);
	return os.str();
}

// ---------------------------------------------------------------------------
// logging
// ---------------------------------------------------------------------------
namespace {

using ::JoshUpscale::core::LogLevel;
using ::JoshUpscale::core::LogSink;

// Default sink, like the reference's console sink (core/src/logging.cc:50-62): every level goes
// to stderr as "<timestamp with ms> <LEVEL> [<tag>] <message>".
struct ConsoleSink : LogSink {
	void operator()(const char *tag, LogLevel level, const std::string &message) override {
		static const char *names[] = {"INFO", "WARNING", "ERROR"};
		const auto now = std::chrono::system_clock::now();
		const std::time_t secs = std::chrono::system_clock::to_time_t(now);
		const int ms = static_cast<int>(
		    std::chrono::duration_cast<std::chrono::milliseconds>(now.time_since_epoch()).count() % 1000);
		char stamp[32];
		std::tm tmv{};
		localtime_r(&secs, &tmv);
		std::strftime(stamp, sizeof(stamp), "%Y-%m-%d %H:%M:%S", &tmv);
		std::fprintf(stderr, "%s.%03d %s [%s] %s\n", stamp, ms, names[static_cast<int>(level)], tag, message.c_str());
	}
};

struct CallbackSink : LogSink {
	ju_log_fn fn = nullptr;
	void *user = nullptr;
	void operator()(const char *tag, LogLevel level, const std::string &message) override {
		if (fn) fn(tag, static_cast<int>(level), message.c_str(), user);
	}
};

ConsoleSink g_DefaultSink;
CallbackSink g_CallbackSink;
LogSink *g_Sink = &g_DefaultSink;

}  // namespace

void setLogSinkInternal(LogSink *sink) { g_Sink = sink ? sink : &g_DefaultSink; }

void logMessage(LogLevel level, const char *tag, const std::string &msg) { (*g_Sink)(tag, level, msg); }

// ---------------------------------------------------------------------------
// Runtime
// ---------------------------------------------------------------------------
namespace {

struct B200Runtime final : ::JoshUpscale::core::Runtime {
	B200Runtime(int deviceId, const std::string &modelPath, int batch)
	    : engine(ModelFile::load(modelPath), deviceId, batch) {
		m_InputWidth = engine.spec().frameW;
		m_InputHeight = engine.spec().frameH;
		m_OutputWidth = 4 * m_InputWidth;
		m_OutputHeight = 4 * m_InputHeight;
	}

	void processImage(const ::JoshUpscale::core::Image &in, const ::JoshUpscale::core::Image &out) override {
		ju_image i{in.ptr, static_cast<std::int32_t>(in.location), in.stride, in.width, in.height};
		ju_image o{out.ptr, static_cast<std::int32_t>(out.location), out.stride, out.width, out.height};
		engine.process(1, &i, &o);
	}

	Engine engine;
};

thread_local std::string t_LastError;

template <typename F>
int guarded(F &&f) {
	try {
		f();
		return 0;
	} catch (...) {
		t_LastError = currentExceptionString();
		return 1;
	}
}

}  // namespace

}  // namespace ju

// ===========================================================================
// C++ API
// ===========================================================================
namespace JoshUpscale {
namespace core {

void setLogSink(LogSink *sink) { ::ju::setLogSinkInternal(sink); }

std::string getExceptionString() { return ::ju::currentExceptionString(); }

Runtime *createRuntime(int deviceId, const std::filesystem::path &modelPath) {
	return new ::ju::B200Runtime(deviceId, modelPath.string(), 1);
}

// ---- OpenGL interop (reference core/src/core.cc:92-149) --------------------
// The build machine has no GL headers and the B200 box no display, so nothing here is linked
// against libGL: the three GL calls the reference makes are resolved from the process's libGL at
// call time, and the CUDA <-> GL entry points of the CUDA runtime are declared by hand (they are
// part of cudart; cuda_gl_interop.h only adds the GL typedefs).  Without a current GL context
// these calls fail exactly like the reference's: with an exception.
}  // namespace core
}  // namespace JoshUpscale

extern "C" {
cudaError_t cudaGraphicsGLRegisterImage(struct cudaGraphicsResource **resource, unsigned int image,
    unsigned int target, unsigned int flags);
cudaError_t cudaGLGetDevices(unsigned int *pCudaDeviceCount, int *pCudaDevices, unsigned int cudaDeviceCount,
    int deviceList);
}

namespace JoshUpscale {
namespace core {

namespace {

constexpr unsigned int kGlTexture2D = 0x0DE1, kGlTextureWidth = 0x1000, kGlTextureHeight = 0x1001;
constexpr int kCudaGLDeviceListAll = 1;  // cudaGLDeviceListAll

struct GlApi {
	void (*bindTexture)(unsigned int, unsigned int) = nullptr;
	unsigned int (*getError)() = nullptr;
	void (*getTexLevelParameteriv)(unsigned int, int, unsigned int, int *) = nullptr;

	static const GlApi &get() {
		static const GlApi api = [] {
			GlApi a;
			void *lib = nullptr;
			for (const char *name : {"libGL.so.1", "libGL.so", "libOpenGL.so.0"}) {
				lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
				if (lib) break;
			}
			if (!lib) throw std::runtime_error("OpenGL library not found (libGL.so.1)");
			a.bindTexture = reinterpret_cast<decltype(a.bindTexture)>(dlsym(lib, "glBindTexture"));
			a.getError = reinterpret_cast<decltype(a.getError)>(dlsym(lib, "glGetError"));
			a.getTexLevelParameteriv =
			    reinterpret_cast<decltype(a.getTexLevelParameteriv)>(dlsym(lib, "glGetTexLevelParameteriv"));
			if (!a.bindTexture || !a.getError || !a.getTexLevelParameteriv) {
				throw std::runtime_error("OpenGL library lacks glBindTexture / glGetError / glGetTexLevelParameteriv");
			}
			return a;
		}();
		return api;
	}
};

struct GLResourceImage final : GraphicsResourceImage {
	GLResourceImage(std::uint32_t image, GraphicsResourceImageType type) {
		const GlApi &gl = GlApi::get();
		const unsigned int flags = type == GraphicsResourceImageType::INPUT ? cudaGraphicsRegisterFlagsReadOnly
		                                                                    : cudaGraphicsRegisterFlagsWriteDiscard;
		gl.bindTexture(kGlTexture2D, image);
		const unsigned int error = gl.getError();
		if (error != 0) throw std::runtime_error("Failed to bind texture: " + std::to_string(error));
		int width = 0, height = 0;
		gl.getTexLevelParameteriv(kGlTexture2D, 0, kGlTextureWidth, &width);
		gl.getTexLevelParameteriv(kGlTexture2D, 0, kGlTextureHeight, &height);
		gl.bindTexture(kGlTexture2D, 0);
		if (width <= 0 || height <= 0) throw std::runtime_error("texture has no level-0 image");
		cudaGraphicsResource_t resource = nullptr;
		::ju::checkCuda(cudaGraphicsGLRegisterImage(&resource, image, kGlTexture2D, flags), "cudaGraphicsGLRegisterImage");
		m_Image.location = DataLocation::GRAPHICS_RESOURCE;
		m_Image.ptr = resource;
		m_Image.width = static_cast<std::size_t>(width);
		m_Image.height = static_cast<std::size_t>(height);
	}
	~GLResourceImage() override { cudaGraphicsUnregisterResource(static_cast<cudaGraphicsResource_t>(m_Image.ptr)); }
};

}  // namespace

int getGLDeviceIndex() {
	int device = -1;
	unsigned int count = 0;
	::ju::checkCuda(cudaGLGetDevices(&count, &device, 1, kCudaGLDeviceListAll), "cudaGLGetDevices");
	if (count != 1) throw std::runtime_error("Failed to determine CUDA device");
	return device;
}

GraphicsResourceImage *getGLImage(std::uint32_t image, GraphicsResourceImageType type) {
	return new GLResourceImage(image, type);
}

}  // namespace core
}  // namespace JoshUpscale

// ===========================================================================
// C-ABI
// ===========================================================================
struct ju_runtime {
	std::unique_ptr<ju::B200Runtime> impl;
};

using ju::guarded;

extern "C" {

int ju_create(const char *model_path, int device, int batch, ju_runtime **out) {
	return guarded([&] {
		if (!model_path || !out) throw std::invalid_argument("null argument");
		auto rt = std::make_unique<ju_runtime>();
		rt->impl = std::make_unique<ju::B200Runtime>(device, model_path, batch);
		*out = rt.release();
	});
}

void ju_destroy(ju_runtime *rt) { delete rt; }

int ju_process(ju_runtime *rt, const ju_image *input, const ju_image *output) {
	return guarded([&] {
		if (!rt || !input || !output) throw std::invalid_argument("null argument");
		rt->impl->engine.process(1, input, output);
	});
}

int ju_process_batch(ju_runtime *rt, int n, const ju_image *inputs, const ju_image *outputs) {
	return guarded([&] {
		if (!rt || !inputs || !outputs) throw std::invalid_argument("null argument");
		rt->impl->engine.process(n, inputs, outputs);
	});
}

int ju_get_info(const ju_runtime *rt, ju_info *info) {
	return guarded([&] {
		if (!rt || !info) throw std::invalid_argument("null argument");
		const ju::Engine &e = rt->impl->engine;
		const ju::ModelSpec &s = e.spec();
		std::memset(info, 0, sizeof(*info));
		info->input_width = s.frameW;
		info->input_height = s.frameH;
		info->output_width = 4 * s.frameW;
		info->output_height = 4 * s.frameH;
		info->padded_width = s.padW;
		info->padded_height = s.padH;
		info->batch = e.batch();
		info->flow_num_inputs = s.flowInputs;
		info->gen_filters = s.genFilters;
		info->gen_blocks = s.genBlocks;
		info->flow_arch = s.flowArch;
		info->conv_impl = e.convImpl();
		info->kernels_per_frame = static_cast<std::uint32_t>(e.kernelsPerFrame());
		info->device = e.device();
		info->gflop_per_frame = 2.0 * (s.flowGmacs() + s.genGmacs());
	});
}

const char *ju_last_error(void) { return ju::t_LastError.c_str(); }

void ju_set_log_sink(ju_log_fn fn, void *user) {
	ju::g_CallbackSink.fn = fn;
	ju::g_CallbackSink.user = user;
	ju::setLogSinkInternal(fn ? &ju::g_CallbackSink : nullptr);
}

int ju_debug_inject_stall(ju_runtime *rt, int kernel_id) {
	return guarded([&] {
		if (!rt) throw std::invalid_argument("null argument");
		rt->impl->engine.injectStall(kernel_id);
	});
}

int ju_reset_state(ju_runtime *rt) {
	return guarded([&] {
		if (!rt) throw std::invalid_argument("null argument");
		rt->impl->engine.resetState();
	});
}

int ju_read_tensor(ju_runtime *rt, const char *name, void *dst, uint64_t capacity, ju_tensor_desc *desc) {
	return guarded([&] {
		if (!rt || !name) throw std::invalid_argument("null argument");
		rt->impl->engine.readTensor(name, dst, capacity, desc);
	});
}

int ju_write_state(ju_runtime *rt, const char *name, const void *src, uint64_t bytes) {
	return guarded([&] {
		if (!rt || !name || !src) throw std::invalid_argument("null argument");
		rt->impl->engine.writeState(name, src, bytes);
	});
}

int ju_profile_ops(ju_runtime *rt, int iters, ju_op_time *ops, int capacity, int *count) {
	return guarded([&] {
		if (!rt || !count) throw std::invalid_argument("null argument");
		auto r = rt->impl->engine.profileOps(iters);
		*count = static_cast<int>(r.size());
		if (ops) {
			for (int i = 0; i < capacity && i < *count; ++i) ops[i] = r[i];
		}
	});
}

int ju_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		(void) cudaGetLastError();
		return 0;
	}
	return n;
}

int ju_set_device(int device) {
	return guarded([&] { JU_CUDA(cudaSetDevice(device)); });
}

const char *ju_version(void) { return "joshupscale-b200 0.1 (sm_100a)"; }

// ---- kernel entry points --------------------------------------------------

namespace {

// device-side FrameIO table for tightly packed test buffers
struct TempIo {
	ju::DeviceBuffer buf;
	TempIo(const uint8_t *frames, uint8_t *outs, int batch, int h, int w) : buf(sizeof(ju::FrameIO) * batch) {
		std::vector<ju::FrameIO> io(batch);
		for (int b = 0; b < batch; ++b) {
			io[b].in = frames ? frames + static_cast<size_t>(b) * h * w * 4 : nullptr;
			io[b].in_stride = static_cast<long long>(w) * 4;
			io[b].out = outs ? outs + static_cast<size_t>(b) * 16 * h * w * 4 : nullptr;
			io[b].out_stride = static_cast<long long>(w) * 16;
		}
		buf.upload(io.data(), sizeof(ju::FrameIO) * batch);
	}
	const ju::FrameIO *get() const { return buf.as<ju::FrameIO>(); }
};

void requireDevice() {
	if (ju_device_count() < 1) throw std::runtime_error("no CUDA device: joshupscale has no CPU fallback");
}

}  // namespace

int ju_launch_preprocess(const uint8_t *frames, const void *flow_prev, void *flow_next, int batch,
    int h, int w, int ph, int pw, int k, int cstride, void *stream) {
	return guarded([&] {
		requireDevice();
		TempIo io(frames, nullptr, batch, h, w);
		auto s = static_cast<cudaStream_t>(stream);
		JU_CUDA(ju::launch_preprocess(io.get(), static_cast<const __half *>(flow_prev),
		    static_cast<__half *>(flow_next), nullptr, batch, h, w, ph, pw, k, cstride, s));
		JU_CUDA(cudaStreamSynchronize(s));
	});
}

int ju_launch_conv(int impl, const void *in, const void *weights, const float *bias, const void *residual,
    void *out, int batch, int h, int w, int cin_stride, int cin, int cout, int cout_stride, int ksize,
    int act, float slope, int out_f32, int shuffle2, void *stream) {
	return guarded([&] {
		requireDevice();
		ju::ConvArgs a{};
		a.in = static_cast<const __half *>(in);
		a.weights = weights;
		a.bias = bias;
		a.residual = static_cast<const __half *>(residual);
		a.out = out;
		a.batch = batch;
		a.h = h;
		a.w = w;
		a.cin_stride = cin_stride;
		a.cin = cin;
		a.cout = cout;
		a.cout_stride = cout_stride;
		a.ksize = ksize;
		a.act = act;
		a.slope = slope;
		a.out_f32 = out_f32;
		a.shuffle2 = shuffle2 == 1 ? 1 : 0;
		a.pool = shuffle2 == 2 ? 1 : 0;
		auto s = static_cast<cudaStream_t>(stream);
		if (impl == 0) {
			JU_CUDA(ju::launch_conv_simt(a, s));
		} else if (impl == 1) {
			ju::ConvTcLaunch l;
			JU_CUDA(ju::conv_tc_prepare(a, ju::conv_tc_default_options(), &l));
			JU_CUDA(ju::conv_tc_launch(l, nullptr, s));
		} else {
			throw std::invalid_argument("unknown conv impl");
		}
	});
}

int ju_set_option(const char *key, int value) {
	return guarded([&] {
		if (!key) throw std::invalid_argument("null key");
		if (std::strcmp(key, "tc_variant") == 0) {
			ju::conv_tc_default_options().variant = value;
		} else if (std::strcmp(key, "tc_tma_epilogue") == 0) {
			ju::conv_tc_default_options().tma_epilogue = value;
		} else if (std::strcmp(key, "tc_pdl") == 0) {
			ju::conv_tc_default_options().pdl = value;
		} else if (std::strcmp(key, "tc_dual") == 0) {
			ju::conv_tc_default_options().dual = value;
		} else {
			throw std::invalid_argument(std::string("unknown option ") + key);
		}
	});
}

int ju_bench_conv(int impl, int batch, int h, int w, int cin, int cout, int ksize, int with_residual,
    int iters, double *usec) {
	return guarded([&] {
		requireDevice();
		if (!usec || iters < 1) throw std::invalid_argument("bad arguments");
		const int cs = (cin + 63) / 64 * 64, os = (cout + 63) / 64 * 64;
		const size_t px = static_cast<size_t>(batch) * h * w;
		ju::DeviceBuffer in(px * cs * 2), out(px * os * 2), res(px * os * 2), bias(cout * 4);
		JU_CUDA(cudaMemset(in.get(), 0x2c, in.bytes()));   // fp16 0x2c2c ~ 0.065
		JU_CUDA(cudaMemset(res.get(), 0x2c, res.bytes()));
		const int cinp = impl >= 1 ? cs : (cin + 15) / 16 * 16;
		ju::DeviceBuffer wts(static_cast<size_t>(ksize) * ksize * cinp * cout * 2);
		JU_CUDA(cudaMemset(wts.get(), 0x24, wts.bytes()));  // ~0.016
		ju::ConvArgs a{};
		a.in = in.as<__half>();
		a.weights = wts.get();
		a.bias = bias.as<float>();
		a.residual = with_residual ? res.as<__half>() : nullptr;
		a.out = out.get();
		a.batch = batch;
		a.h = h;
		a.w = w;
		a.cin_stride = cs;
		a.cin = cinp;
		a.cout = cout;
		a.cout_stride = os;
		a.ksize = ksize;
		a.act = ju::ACT_RELU;
		cudaEvent_t e0, e1;
		JU_CUDA(cudaEventCreate(&e0));
		JU_CUDA(cudaEventCreate(&e1));
		ju::ConvTcLaunch l;
		if (impl == 1) JU_CUDA(ju::conv_tc_prepare(a, ju::conv_tc_default_options(), &l));
		auto launch = [&] {
			if (impl == 1) {
				JU_CUDA(ju::conv_tc_launch(l, nullptr, nullptr));
			} else {
				JU_CUDA(ju::launch_conv_simt(a, nullptr));
			}
		};
		for (int i = 0; i < 3; ++i) launch();
		JU_CUDA(cudaEventRecord(e0, nullptr));
		for (int i = 0; i < iters; ++i) launch();
		JU_CUDA(cudaEventRecord(e1, nullptr));
		JU_CUDA(cudaEventSynchronize(e1));
		float ms = 0.f;
		JU_CUDA(cudaEventElapsedTime(&ms, e0, e1));
		cudaEventDestroy(e0);
		cudaEventDestroy(e1);
		*usec = ms * 1000.0 / iters;
	});
}

int64_t ju_pack_conv_weights(int impl, const float *kernel, const float *scale, int ksize, int cin,
    int cin_padded, int cout, void *dst) {
	if (impl == 0) {
		auto bytes = static_cast<int64_t>(ju::conv_simt_weight_bytes(ksize, cin_padded, cout));
		if (dst) ju::conv_simt_pack_weights(kernel, scale, ksize, cin, cin_padded, cout, static_cast<__half *>(dst));
		return bytes;
	}
	if (impl == 1 && cin_padded % 64 == 0) {
		auto bytes = static_cast<int64_t>(ju::conv_tc_weight_bytes(ksize, cin_padded, cout));
		if (dst) ju::conv_tc_pack_weights(kernel, scale, ksize, cin, cin_padded, cout, static_cast<__half *>(dst));
		return bytes;
	}
	return -1;
}

int ju_launch_maxpool2(const void *in, void *out, int batch, int h, int w, int c, void *stream) {
	return guarded([&] {
		requireDevice();
		JU_CUDA(ju::launch_maxpool2(static_cast<const __half *>(in), static_cast<__half *>(out), batch, h, w, c,
		    static_cast<cudaStream_t>(stream)));
	});
}

int ju_launch_upscale2(const void *in, void *out, int batch, int h, int w, int c, void *stream) {
	return guarded([&] {
		requireDevice();
		JU_CUDA(ju::launch_upscale2(static_cast<const __half *>(in), static_cast<__half *>(out), batch, h, w, c,
		    static_cast<cudaStream_t>(stream)));
	});
}

int ju_launch_warp_s2d(const void *pre_gen, const float *flow_head, const uint8_t *frames, void *gen_in,
    float *taps, int batch, int h, int w, int ph, int pw, int cstride, void *stream) {
	return guarded([&] {
		requireDevice();
		TempIo io(frames, nullptr, batch, h, w);
		auto s = static_cast<cudaStream_t>(stream);
		JU_CUDA(ju::launch_warp_s2d(static_cast<const __half *>(pre_gen), flow_head, io.get(),
		    static_cast<__half *>(gen_in), taps, nullptr, batch, h, w, ph, pw, cstride, s));
		JU_CUDA(cudaStreamSynchronize(s));
	});
}

int ju_launch_final(const void *mid, const float *w2, const float *bias2, const uint8_t *frames,
    uint8_t *out_bgrx, void *pre_gen_next, float *out_raw, int batch, int h, int w, void *stream) {
	return guarded([&] {
		requireDevice();
		TempIo io(frames, out_bgrx, batch, h, w);
		auto s = static_cast<cudaStream_t>(stream);
		JU_CUDA(ju::launch_final(static_cast<const __half *>(mid), w2, bias2, io.get(),
		    static_cast<__half *>(pre_gen_next), out_raw, nullptr, batch, h, w, s));
		JU_CUDA(cudaStreamSynchronize(s));
	});
}

int ju_launch_tail(const void *trunk, const void *w1, const float *bias1, const float *w2, const float *bias2,
    const uint8_t *frames, uint8_t *out_bgrx, void *pre_gen_next, float *out_raw, int batch, int h, int w,
    int act, float slope, void *stream) {
	return guarded([&] {
		requireDevice();
		TempIo io(frames, out_bgrx, batch, h, w);
		auto s = static_cast<cudaStream_t>(stream);
		ju::TailArgs a{};
		a.in = static_cast<const __half *>(trunk);
		a.cin_stride = 64;
		a.weights1 = w1;
		// the kernel takes these small tensors as parameters: fetch them from the device
		std::vector<float> hb1(128), hw2(4 * 3 * 32), hb2(3);
		JU_CUDA(cudaMemcpy(hb1.data(), bias1, hb1.size() * sizeof(float), cudaMemcpyDeviceToHost));
		JU_CUDA(cudaMemcpy(hw2.data(), w2, hw2.size() * sizeof(float), cudaMemcpyDeviceToHost));
		JU_CUDA(cudaMemcpy(hb2.data(), bias2, hb2.size() * sizeof(float), cudaMemcpyDeviceToHost));
		a.bias1_host = hb1.data();
		a.w2_host = hw2.data();
		a.bias2_host = hb2.data();
		a.io = io.get();
		a.pre_gen_next = static_cast<__half *>(pre_gen_next);
		a.out_raw = out_raw;
		a.batch = batch;
		a.h = h;
		a.w = w;
		a.act = act;
		a.slope = slope;
		a.pdl = 0;
		ju::TailTcLaunch l;
		JU_CUDA(ju::tail_tc_prepare(a, &l));
		JU_CUDA(ju::tail_tc_launch(l, nullptr, s));
		JU_CUDA(cudaStreamSynchronize(s));
	});
}

// ---- raw device helpers ---------------------------------------------------

int ju_dev_alloc(void **ptr, uint64_t bytes) {
	return guarded([&] {
		requireDevice();
		JU_CUDA(cudaMalloc(ptr, bytes));
		JU_CUDA(cudaMemset(*ptr, 0, bytes));
	});
}
int ju_dev_free(void *ptr) {
	return guarded([&] { JU_CUDA(cudaFree(ptr)); });
}
int ju_dev_upload(void *dst, const void *src, uint64_t bytes) {
	return guarded([&] { JU_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice)); });
}
int ju_dev_download(void *dst, const void *src, uint64_t bytes) {
	return guarded([&] { JU_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost)); });
}
int ju_dev_memset(void *dst, int value, uint64_t bytes) {
	return guarded([&] { JU_CUDA(cudaMemset(dst, value, bytes)); });
}
int ju_dev_sync(void) {
	return guarded([&] { JU_CUDA(cudaDeviceSynchronize()); });
}

int ju_host_alloc(void **ptr, uint64_t bytes) {
	return guarded([&] {
		requireDevice();
		JU_CUDA(cudaMallocHost(ptr, bytes));
		std::memset(*ptr, 0, bytes);
	});
}
int ju_host_free(void *ptr) {
	return guarded([&] { JU_CUDA(cudaFreeHost(ptr)); });
}

namespace {
cudaEvent_t g_TimerBegin = nullptr, g_TimerEnd = nullptr;
void *g_FlushBuf[8] = {nullptr};
constexpr size_t kFlushBytes = 256u << 20;
}

int ju_l2_flush(void) {
	return guarded([&] {
		requireDevice();
		int dev = 0;
		JU_CUDA(cudaGetDevice(&dev));
		if (dev < 0 || dev >= 8) throw std::invalid_argument("device index out of range");
		if (!g_FlushBuf[dev]) JU_CUDA(cudaMalloc(&g_FlushBuf[dev], kFlushBytes));
		static int toggle = 0;
		JU_CUDA(cudaMemsetAsync(g_FlushBuf[dev], ++toggle & 0xff, kFlushBytes, nullptr));
		JU_CUDA(cudaStreamSynchronize(nullptr));
	});
}

int ju_u8_conversion_table(float *fast256, float *ieee256) {
	return guarded([&] {
		requireDevice();
		if (!fast256 || !ieee256) throw std::invalid_argument("null table pointer");
		ju::DeviceBuffer buf(2 * 256 * sizeof(float));
		float *d = buf.as<float>();
		ju::checkCuda(ju::launch_u8_table(d, d + 256, nullptr), "launch_u8_table");
		JU_CUDA(cudaMemcpy(fast256, d, 256 * sizeof(float), cudaMemcpyDeviceToHost));
		JU_CUDA(cudaMemcpy(ieee256, d + 256, 256 * sizeof(float), cudaMemcpyDeviceToHost));
	});
}

int ju_timer_begin(void) {
	return guarded([&] {
		requireDevice();
		if (!g_TimerBegin) {
			JU_CUDA(cudaEventCreate(&g_TimerBegin));
			JU_CUDA(cudaEventCreate(&g_TimerEnd));
		}
		JU_CUDA(cudaEventRecord(g_TimerBegin, nullptr));
	});
}

int ju_timer_end(double *usec) {
	return guarded([&] {
		if (!g_TimerBegin || !usec) throw std::invalid_argument("timer not started");
		JU_CUDA(cudaEventRecord(g_TimerEnd, nullptr));
		JU_CUDA(cudaEventSynchronize(g_TimerEnd));
		float ms = 0.f;
		JU_CUDA(cudaEventElapsedTime(&ms, g_TimerBegin, g_TimerEnd));
		*usec = ms * 1000.0;
	});
}

}  // extern "C"
