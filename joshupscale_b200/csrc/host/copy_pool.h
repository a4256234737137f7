// Small host-side copy pool: moves image rows between the engine's page-locked staging buffers
// and the caller's PAGEABLE images on several threads.
//
// Why: the reference's callers hand processImage ordinary heap memory (AviSynth frame buffers,
// avisynth_plugin/src/main.cc:125-142).  A cudaMemcpyAsync to / from pageable memory is staged by
// the driver through one internal buffer on one thread and serialises with the frame; copying
// device -> engine-owned pinned memory asynchronously (band by band, overlapped with the tail
// kernel) and finishing with a multi-threaded memcpy into the caller's buffer keeps the
// plugin-visible latency close to the pinned-memory figure.
#pragma once

#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace ju {

struct CopyJob {
	std::uint8_t *dst = nullptr;
	const std::uint8_t *src = nullptr;
	std::ptrdiff_t dstStride = 0, srcStride = 0;  // bytes between rows (may be negative)
	std::size_t rowBytes = 0, rows = 0;
};

class HostCopyPool {
public:
	explicit HostCopyPool(int threads);
	~HostCopyPool();
	HostCopyPool(const HostCopyPool &) = delete;
	HostCopyPool &operator=(const HostCopyPool &) = delete;

	int threads() const { return static_cast<int>(m_Workers.size()); }
	// splits the rows of `job` over the workers; returns immediately
	void submit(const CopyJob &job);
	// blocks until every submitted row has been copied
	void wait();

private:
	void run();

	std::vector<std::thread> m_Workers;
	std::deque<CopyJob> m_Queue;
	std::mutex m_Mutex;
	std::condition_variable m_Wake, m_Idle;
	std::size_t m_Pending = 0;
	bool m_Stop = false;
};

}  // namespace ju
