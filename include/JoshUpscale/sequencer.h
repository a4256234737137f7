// Frame sequencing for random-access hosts of a strictly recurrent filter.
//
// Runtime::processImage must see frames in order (the recurrent state lives in
// the runtime), but a frame server may ask for any frame number.  The reference's
// AviSynth plugin solves this inside JoshUpscaleFilter::GetFrame
// (avisynth_plugin/src/main.cc:75-161): serve recent frames from a small ring
// cache, warm the stream up over the previous MAX_BACKTRACK_SIZE frames after a
// seek (mirroring negative frame numbers at the start of the clip), and reset
// when the request is too far away.  This header restates that policy as a
// host-side utility that is independent of AviSynth, so any caller of the
// drop-in runtime (a frame server, a script, a test) gets the same sequencing:
//
//   FrameSequencer<Frame> seq(
//       [&](int sourceIndex) { return clip.frame(sourceIndex); },        // fetch
//       [&](const Frame &src) { Frame dst = alloc(); runtime->processImage(view(src), view(dst)); return dst; });
//   Frame out = seq.get(n);
//
// Header-only, no CUDA, no dependency on the library; `Frame` is any copyable
// handle (PVideoFrame, shared_ptr<Image>, ...).
#pragma once

#include <cstddef>
#include <functional>
#include <utility>
#include <vector>

namespace JoshUpscale {

namespace core {

template <typename Frame>
class FrameSequencer {
public:
	// avisynth_plugin/src/main.cc:17-18
	static constexpr int kMaxBacktrack = 16;
	static constexpr std::size_t kCacheSize = 16;

	using Fetch = std::function<Frame(int sourceIndex)>;
	using Process = std::function<Frame(const Frame &source)>;

	struct Stats {
		std::size_t processed = 0;   // processImage calls
		std::size_t cacheHits = 0;   // requests served from the ring
		std::size_t resets = 0;      // stream restarts (seek out of reach)
		std::size_t backtracks = 0;  // requests that had to process earlier frames first
	};

	FrameSequencer(Fetch fetch, Process process) : m_Fetch(std::move(fetch)), m_Process(std::move(process)) {
		m_Cache.reserve(kCacheSize);
	}

	// Output frame `n`.  Frames before the first one of a stream are the clip mirrored
	// around frame 0 (source index |k|), exactly like the plugin's warm-up.
	Frame get(int n) {
		if (n < m_Next) {
			const std::size_t back = static_cast<std::size_t>(m_Next - n);
			if (back <= m_Cache.size()) {
				++m_Stats.cacheHits;
				return m_Cache[(m_Cache.size() - back + m_Shift) % kCacheSize];
			}
			reset(n);
		}
		if (n > m_Next) {
			if (m_Next + kMaxBacktrack < n) reset(n);
			++m_Stats.backtracks;
		}
		Frame out{};
		for (int k = m_Next; k <= n; ++k) out = step(k);
		return out;
	}

	// Number of the frame the runtime expects next (main.cc:41, 150).
	int next() const { return m_Next; }
	const Stats &stats() const { return m_Stats; }

private:
	void reset(int n) {
		m_Next = n - kMaxBacktrack;
		m_Cache.clear();
		m_Shift = 0;
		m_Uncached = kMaxBacktrack;
		++m_Stats.resets;
	}

	Frame step(int k) {
		Frame out = m_Process(m_Fetch(k >= 0 ? k : -k));
		++m_Stats.processed;
		m_Next = k + 1;
		if (m_Uncached > 0) {
			--m_Uncached;  // warm-up outputs are never served
		} else if (m_Cache.size() == kCacheSize) {
			m_Cache[m_Shift] = out;
			m_Shift = (m_Shift + 1) % kCacheSize;
		} else {
			m_Cache.push_back(out);
		}
		return out;
	}

	Fetch m_Fetch;
	Process m_Process;
	Stats m_Stats;
	int m_Next = -kMaxBacktrack;
	std::vector<Frame> m_Cache;
	std::size_t m_Shift = 0;
	std::size_t m_Uncached = static_cast<std::size_t>(kMaxBacktrack);
};

}  // namespace core

}  // namespace JoshUpscale
