#include "copy_pool.h"

#include <algorithm>
#include <cstring>

#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define JU_CPU_RELAX() _mm_pause()
#else
#define JU_CPU_RELAX() std::this_thread::yield()
#endif

namespace ju {

HostCopyPool::HostCopyPool(int threads) {
	// the calling thread copies too: `threads` - 1 workers
	for (int i = 1; i < threads; ++i) m_Workers.emplace_back([this] { run(); });
}

HostCopyPool::~HostCopyPool() {
	{
		std::lock_guard<std::mutex> lock(m_Mutex);
		m_Stop = true;
		m_Active.store(false, std::memory_order_release);
	}
	m_Wake.notify_all();
	for (std::thread &t : m_Workers) t.join();
}

void HostCopyPool::begin() {
	if (m_Active.load(std::memory_order_acquire)) return;
	{
		std::lock_guard<std::mutex> lock(m_Mutex);
		m_Head.store(0, std::memory_order_relaxed);
		m_Tail.store(0, std::memory_order_relaxed);
		m_Done.store(0, std::memory_order_relaxed);
		m_Active.store(true, std::memory_order_release);
	}
	m_Wake.notify_all();
}

void HostCopyPool::copy(const CopyJob &job) {
	const std::ptrdiff_t dense = static_cast<std::ptrdiff_t>(job.rowBytes);
	if (job.dstStride == dense && job.srcStride == dense) {
		std::memcpy(job.dst, job.src, job.rowBytes * job.rows);
		return;
	}
	for (std::size_t r = 0; r < job.rows; ++r) {
		std::memcpy(job.dst + static_cast<std::ptrdiff_t>(r) * job.dstStride, job.src + static_cast<std::ptrdiff_t>(r) * job.srcStride,
		    job.rowBytes);
	}
}

void HostCopyPool::submit(const CopyJob &job) {
	if (job.rows == 0 || job.rowBytes == 0) return;
	begin();
	const std::size_t parts = std::min<std::size_t>(m_Workers.size() + 1, job.rows);
	for (std::size_t i = 0; i < parts; ++i) {
		const std::size_t r0 = job.rows * i / parts, r1 = job.rows * (i + 1) / parts;
		CopyJob part = job;
		part.dst = job.dst + static_cast<std::ptrdiff_t>(r0) * job.dstStride;
		part.src = job.src + static_cast<std::ptrdiff_t>(r0) * job.srcStride;
		part.rows = r1 - r0;
		const int slot = m_Tail.load(std::memory_order_relaxed);
		if (slot >= kSlots) {  // ring full (never with the engine's band counts): copy in place
			copy(part);
			continue;
		}
		m_Ring[slot] = part;
		m_Tail.store(slot + 1, std::memory_order_release);  // single producer: the engine's calling thread
	}
}

bool HostCopyPool::takeOne() {
	int i = m_Head.load(std::memory_order_relaxed);
	while (i < m_Tail.load(std::memory_order_acquire)) {
		if (m_Head.compare_exchange_weak(i, i + 1, std::memory_order_acq_rel)) {
			copy(m_Ring[i]);
			m_Done.fetch_add(1, std::memory_order_release);
			return true;
		}
	}
	return false;
}

void HostCopyPool::wait() {
	while (m_Done.load(std::memory_order_acquire) < m_Tail.load(std::memory_order_acquire)) {
		if (!takeOne()) JU_CPU_RELAX();
	}
}

void HostCopyPool::end() {
	if (!m_Active.load(std::memory_order_acquire)) return;
	wait();
	std::lock_guard<std::mutex> lock(m_Mutex);
	m_Active.store(false, std::memory_order_release);
}

void HostCopyPool::run() {
	for (;;) {
		{
			std::unique_lock<std::mutex> lock(m_Mutex);
			m_Wake.wait(lock, [this] { return m_Stop || m_Active.load(std::memory_order_acquire); });
			if (m_Stop) return;
		}
		while (m_Active.load(std::memory_order_acquire)) {
			if (!takeOne()) JU_CPU_RELAX();
		}
	}
}

}  // namespace ju
