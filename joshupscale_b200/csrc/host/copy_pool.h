// Small host-side copy pool: moves image rows between the engine's page-locked staging buffers
// and the caller's PAGEABLE images on several threads.
//
// Why: the reference's callers hand processImage ordinary heap memory (AviSynth frame buffers,
// avisynth_plugin/src/main.cc:125-142).  A cudaMemcpyAsync to / from pageable memory is staged by
// the driver through one internal buffer on one thread and serialises with the frame; copying
// device -> engine-owned pinned memory asynchronously (band by band, overlapped with the tail
// kernel) and finishing with a multi-threaded memcpy into the caller's buffer keeps the
// plugin-visible latency close to the pinned-memory figure.
//
// Latency matters more than throughput here (one 1080p frame is 8 MB, a frame takes 0.4 ms), so
// the workers are only parked BETWEEN frames: begin() wakes them (tens of microseconds, hidden
// behind the GPU work), then they and the calling thread take jobs from a lock-free ring by
// spinning, and end() parks them again.  (A first version that woke the workers per job from a
// cudaLaunchHostFunc callback was slower than the driver's own staging: 1.56 vs 0.95 ms per frame.)
#pragma once

#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
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
	// wakes the workers for one frame (idempotent); jobs may be submitted until end()
	void begin();
	// splits the rows of `job` over the workers and the caller; returns immediately
	void submit(const CopyJob &job);
	// the calling thread copies one queued part if there is one (used while it polls for the next band)
	bool help() { return takeOne(); }
	// the calling thread helps until every submitted row has been copied
	void wait();
	// wait(), then park the workers
	void end();

private:
	static constexpr int kSlots = 1024;  // jobs per frame: streams x bands x parts
	void run();
	bool takeOne();
	static void copy(const CopyJob &job);

	std::vector<std::thread> m_Workers;
	CopyJob m_Ring[kSlots];
	std::atomic<int> m_Head{0}, m_Tail{0}, m_Done{0};
	std::atomic<bool> m_Active{false};
	std::mutex m_Mutex;
	std::condition_variable m_Wake;
	bool m_Stop = false;
};

}  // namespace ju
