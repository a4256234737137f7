#include "copy_pool.h"

#include <cstring>

namespace ju {

HostCopyPool::HostCopyPool(int threads) {
	if (threads < 1) threads = 1;
	for (int i = 0; i < threads; ++i) m_Workers.emplace_back([this] { run(); });
}

HostCopyPool::~HostCopyPool() {
	{
		std::lock_guard<std::mutex> lock(m_Mutex);
		m_Stop = true;
	}
	m_Wake.notify_all();
	for (std::thread &t : m_Workers) t.join();
}

void HostCopyPool::submit(const CopyJob &job) {
	if (job.rows == 0 || job.rowBytes == 0) return;
	const std::size_t parts = std::min<std::size_t>(m_Workers.size(), job.rows);
	{
		std::lock_guard<std::mutex> lock(m_Mutex);
		for (std::size_t i = 0; i < parts; ++i) {
			const std::size_t r0 = job.rows * i / parts, r1 = job.rows * (i + 1) / parts;
			CopyJob part = job;
			part.dst = job.dst + static_cast<std::ptrdiff_t>(r0) * job.dstStride;
			part.src = job.src + static_cast<std::ptrdiff_t>(r0) * job.srcStride;
			part.rows = r1 - r0;
			m_Queue.push_back(part);
			++m_Pending;
		}
	}
	m_Wake.notify_all();
}

void HostCopyPool::wait() {
	std::unique_lock<std::mutex> lock(m_Mutex);
	m_Idle.wait(lock, [this] { return m_Pending == 0; });
}

void HostCopyPool::run() {
	for (;;) {
		CopyJob job;
		{
			std::unique_lock<std::mutex> lock(m_Mutex);
			m_Wake.wait(lock, [this] { return m_Stop || !m_Queue.empty(); });
			if (m_Queue.empty()) return;  // stop requested and nothing left
			job = m_Queue.front();
			m_Queue.pop_front();
		}
		const std::ptrdiff_t dense = static_cast<std::ptrdiff_t>(job.rowBytes);
		if (job.dstStride == dense && job.srcStride == dense) {
			std::memcpy(job.dst, job.src, job.rowBytes * job.rows);
		} else {
			for (std::size_t r = 0; r < job.rows; ++r) {
				std::memcpy(job.dst + static_cast<std::ptrdiff_t>(r) * job.dstStride,
				    job.src + static_cast<std::ptrdiff_t>(r) * job.srcStride, job.rowBytes);
			}
		}
		{
			std::lock_guard<std::mutex> lock(m_Mutex);
			if (--m_Pending == 0) m_Idle.notify_all();
		}
	}
}

}  // namespace ju
