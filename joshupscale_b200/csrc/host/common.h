// Host-side plumbing shared by the engine and the API layers: exceptions,
// logging, CUDA RAII.  Re-creates (idiomatically, not line by line) what the
// reference keeps in core/include/JoshUpscale/core/{cuda,exception,logging}.h.
#pragma once

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>

#include "JoshUpscale/core.h"

namespace ju {

// ---- exceptions (reference cuda.h:24-38: CudaException carries the
// cudaGetErrorString text) ---------------------------------------------
struct CudaException : std::runtime_error {
	explicit CudaException(cudaError_t err, const char *what)
	    : std::runtime_error(std::string(what) + ": " + cudaGetErrorName(err) + " (" +
	                         cudaGetErrorString(err) + ")")
	    , code(err) {}
	cudaError_t code;
};

struct ModelException : std::runtime_error {
	using std::runtime_error::runtime_error;
};

// a tcgen05 pipeline wait expired (kernels.h: TcStatus); the frame is lost, the runtime is not
struct KernelStallException : std::runtime_error {
	using std::runtime_error::runtime_error;
};

inline void checkCuda(cudaError_t err, const char *what) {
	if (err != cudaSuccess) {
		(void) cudaGetLastError();  // clear the sticky-less error state
		throw CudaException(err, what);
	}
}

#define JU_CUDA(expr) ::ju::checkCuda((expr), #expr)

// "Type: what\n  Type: what" for the in-flight exception (reference
// core/src/exception.cc:51-79).  Must be called inside a catch block.
std::string currentExceptionString();

// ---- logging (reference logging.h:28-45, logging.cc:50-62) -------------
void setLogSinkInternal(::JoshUpscale::core::LogSink *sink);
void logMessage(::JoshUpscale::core::LogLevel level, const char *tag, const std::string &msg);

struct LogLine {
	LogLine(::JoshUpscale::core::LogLevel level, const char *tag) : m_Level(level), m_Tag(tag) {}
	~LogLine() { logMessage(m_Level, m_Tag, m_Stream.str()); }
	template <typename T>
	LogLine &operator<<(const T &v) {
		m_Stream << v;
		return *this;
	}

private:
	::JoshUpscale::core::LogLevel m_Level;
	const char *m_Tag;
	std::ostringstream m_Stream;
};

#define JU_LOG_INFO ::ju::LogLine(::JoshUpscale::core::LogLevel::INFO, __func__)
#define JU_LOG_WARN ::ju::LogLine(::JoshUpscale::core::LogLevel::WARNING, __func__)
#define JU_LOG_ERROR ::ju::LogLine(::JoshUpscale::core::LogLevel::ERROR, __func__)

// ---- CUDA RAII -----------------------------------------------------------
// Device allocations are zero-filled, like the reference's CudaBuffer
// (cuda.h:69-72): zero == mid-grey initial recurrent state.
class DeviceBuffer {
public:
	DeviceBuffer() = default;
	explicit DeviceBuffer(std::size_t bytes) : m_Bytes(bytes) {
		if (bytes == 0) return;
		JU_CUDA(cudaMalloc(&m_Ptr, bytes));
		cudaError_t e = cudaMemset(m_Ptr, 0, bytes);
		if (e != cudaSuccess) {
			cudaFree(m_Ptr);
			checkCuda(e, "cudaMemset");
		}
	}
	~DeviceBuffer() { reset(); }
	DeviceBuffer(DeviceBuffer &&o) noexcept : m_Ptr(o.m_Ptr), m_Bytes(o.m_Bytes) {
		o.m_Ptr = nullptr;
		o.m_Bytes = 0;
	}
	DeviceBuffer &operator=(DeviceBuffer &&o) noexcept {
		if (this != &o) {
			reset();
			m_Ptr = o.m_Ptr;
			m_Bytes = o.m_Bytes;
			o.m_Ptr = nullptr;
			o.m_Bytes = 0;
		}
		return *this;
	}
	DeviceBuffer(const DeviceBuffer &) = delete;
	DeviceBuffer &operator=(const DeviceBuffer &) = delete;

	void reset() {
		if (m_Ptr) cudaFree(m_Ptr);
		m_Ptr = nullptr;
		m_Bytes = 0;
	}
	void *get() const { return m_Ptr; }
	template <typename T>
	T *as() const { return static_cast<T *>(m_Ptr); }
	std::size_t bytes() const { return m_Bytes; }
	void upload(const void *src, std::size_t bytes) {
		JU_CUDA(cudaMemcpy(m_Ptr, src, bytes, cudaMemcpyHostToDevice));
	}

private:
	void *m_Ptr = nullptr;
	std::size_t m_Bytes = 0;
};

class PinnedBuffer {
public:
	PinnedBuffer() = default;
	explicit PinnedBuffer(std::size_t bytes) : m_Bytes(bytes) {
		JU_CUDA(cudaMallocHost(&m_Ptr, bytes));
	}
	~PinnedBuffer() {
		if (m_Ptr) cudaFreeHost(m_Ptr);
	}
	PinnedBuffer(PinnedBuffer &&o) noexcept : m_Ptr(o.m_Ptr), m_Bytes(o.m_Bytes) { o.m_Ptr = nullptr; }
	PinnedBuffer &operator=(PinnedBuffer &&o) noexcept {
		std::swap(m_Ptr, o.m_Ptr);
		std::swap(m_Bytes, o.m_Bytes);
		return *this;
	}
	PinnedBuffer(const PinnedBuffer &) = delete;
	PinnedBuffer &operator=(const PinnedBuffer &) = delete;
	template <typename T>
	T *as() const { return static_cast<T *>(m_Ptr); }

private:
	void *m_Ptr = nullptr;
	std::size_t m_Bytes = 0;
};

// Saves / restores the calling thread's current device around a call, like
// the reference's DeviceContext (cuda.h:297-308).
class DeviceGuard {
public:
	explicit DeviceGuard(int device) {
		JU_CUDA(cudaGetDevice(&m_Prev));
		if (m_Prev != device) JU_CUDA(cudaSetDevice(device));
		m_Changed = m_Prev != device;
	}
	~DeviceGuard() {
		if (m_Changed) cudaSetDevice(m_Prev);
	}
	DeviceGuard(const DeviceGuard &) = delete;
	DeviceGuard &operator=(const DeviceGuard &) = delete;

private:
	int m_Prev = 0;
	bool m_Changed = false;
};

}  // namespace ju
