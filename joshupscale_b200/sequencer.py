"""Python twin of include/JoshUpscale/sequencer.h: in-order driving of a
recurrent runtime for callers that request frames by number.

Same policy as the reference's AviSynth filter (JoshUpscaleFilter::GetFrame,
avisynth_plugin/src/main.cc:75-161): a ring cache of the last CACHE_SIZE
outputs, a MAX_BACKTRACK_SIZE-frame warm-up after every (re)start that mirrors
negative frame numbers around frame 0, and a reset when a request is out of
reach.  `process` is normally `Runtime.process`; `fetch(i)` returns source
frame i.
"""

from __future__ import annotations

import dataclasses
from typing import Callable, Generic, List, TypeVar

F = TypeVar("F")

MAX_BACKTRACK_SIZE = 16  # avisynth_plugin/src/main.cc:17
CACHE_SIZE = 16  # avisynth_plugin/src/main.cc:18


@dataclasses.dataclass
class SequencerStats:
    processed: int = 0
    cache_hits: int = 0
    resets: int = 0
    backtracks: int = 0


class FrameSequencer(Generic[F]):
    def __init__(self, fetch: Callable[[int], F], process: Callable[[F], F]):
        self._fetch = fetch
        self._process = process
        self.stats = SequencerStats()
        self.next = -MAX_BACKTRACK_SIZE
        self._cache: List[F] = []
        self._shift = 0
        self._uncached = MAX_BACKTRACK_SIZE

    def _reset(self, n: int) -> None:
        self.next = n - MAX_BACKTRACK_SIZE
        self._cache.clear()
        self._shift = 0
        self._uncached = MAX_BACKTRACK_SIZE
        self.stats.resets += 1

    def _step(self, k: int) -> F:
        out = self._process(self._fetch(abs(k)))
        self.stats.processed += 1
        self.next = k + 1
        if self._uncached > 0:
            self._uncached -= 1
        elif len(self._cache) == CACHE_SIZE:
            self._cache[self._shift] = out
            self._shift = (self._shift + 1) % CACHE_SIZE
        else:
            self._cache.append(out)
        return out

    def get(self, n: int) -> F:
        if n < self.next:
            back = self.next - n
            if back <= len(self._cache):
                self.stats.cache_hits += 1
                return self._cache[(len(self._cache) - back + self._shift) % CACHE_SIZE]
            self._reset(n)
        if n > self.next:
            if self.next + MAX_BACKTRACK_SIZE < n:
                self._reset(n)
            self.stats.backtracks += 1
        out = None
        for k in range(self.next, n + 1):
            out = self._step(k)
        return out
