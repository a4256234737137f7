// Drives JoshUpscale::core::FrameSequencer with the request script given on the
// command line and prints one line per request:
//   <n> -> <output id> | processed source indices ... | hits resets backtracks next
// A "frame" is an int: source frames are their index, the fake runtime returns
// 1000 * (running call counter) + source index, so recurrence order is visible.
#include <JoshUpscale/sequencer.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

int main(int argc, char **argv) {
	std::vector<int> processed;
	int calls = 0;
	JoshUpscale::core::FrameSequencer<int> seq([](int idx) { return idx; },
	    [&](const int &src) {
		    processed.push_back(src);
		    return 1000 * (++calls) + src;
	    });
	for (int i = 1; i < argc; ++i) {
		processed.clear();
		const int n = std::atoi(argv[i]);
		const int out = seq.get(n);
		std::printf("%d -> %d |", n, out);
		for (int p : processed) std::printf(" %d", p);
		std::printf(" | %zu %zu %zu %d\n", seq.stats().cacheHits, seq.stats().resets, seq.stats().backtracks,
		    seq.next());
	}
	return 0;
}
