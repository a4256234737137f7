"""Frame sequencing policy (avisynth_plugin/src/main.cc:75-161): the C++ header,
its Python twin and an independent brute-force model must agree request by request."""

import os
import shutil
import subprocess

import numpy as np
import pytest

from joshupscale_b200 import sequencer as js

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPTS = [
    list(range(0, 40)),                                    # plain playback after the mirrored warm-up
    [0, 1, 2, 1, 0, 3],                                    # frames 0.. are cached (the 16 mirrored warm-up outputs are not)
    list(range(0, 40)) + [39, 30, 24, 23, 40, 41],         # ring cache hits, then a miss beyond the ring
    [100, 101, 90, 150, 149, 166, 167, 200],               # seeks: within reach, out of reach, backwards
    [5, 25, 21, 22, 60, 44, 61],
]


def _model(requests):
    """Straight restatement of the plugin's recursion with explicit lists."""
    nxt, cache, dont = -16, [], 16
    calls = 0
    hits = resets = backtracks = 0
    lines = []
    for n in requests:
        processed = []
        out = None
        if n < nxt and nxt - n <= len(cache):
            hits += 1
            out = cache[len(cache) - (nxt - n)]
        else:
            if n < nxt:
                nxt, cache, dont = n - 16, [], 16
                resets += 1
            if n > nxt:
                if nxt + 16 < n:
                    nxt, cache, dont = n - 16, [], 16
                    resets += 1
                backtracks += 1
            while nxt <= n:
                calls += 1
                out = 1000 * calls + abs(nxt)
                processed.append(abs(nxt))
                nxt += 1
                if dont > 0:
                    dont -= 1
                else:
                    cache.append(out)
                    cache = cache[-16:]
        lines.append(f"{n} -> {out} | {' '.join(map(str, processed))} | {hits} {resets} {backtracks} {nxt}"
                     .replace("|  |", "| |"))
    return lines


def _python(requests):
    calls = [0]
    processed = []

    def process(src):
        calls[0] += 1
        processed.append(src)
        return 1000 * calls[0] + src

    seq = js.FrameSequencer(lambda i: i, process)
    lines = []
    for n in requests:
        processed.clear()
        out = seq.get(n)
        lines.append(f"{n} -> {out} | {' '.join(map(str, processed))} | "
                     f"{seq.stats.cache_hits} {seq.stats.resets} {seq.stats.backtracks} {seq.next}"
                     .replace("|  |", "| |"))
    return lines


@pytest.mark.parametrize("requests", SCRIPTS)
def test_python_twin_matches_model(requests):
    assert _python(requests) == _model(requests)


def test_first_request_warms_up_over_mirrored_frames():
    lines = _python([0])
    # frames -16..-1 are the clip mirrored around 0, then frame 0 itself
    assert lines[0].split(" | ")[1] == " ".join(str(abs(k)) for k in range(-16, 1))


@pytest.fixture(scope="module")
def trace_binary(tmp_path_factory):
    cxx = shutil.which("g++")
    if not cxx:
        pytest.skip("no g++")
    out = str(tmp_path_factory.mktemp("seq") / "sequencer_trace")
    subprocess.run([cxx, "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cxx", "sequencer_trace.cc"), "-o", out], check=True)
    return out


@pytest.mark.parametrize("requests", SCRIPTS)
def test_cxx_header_matches_model(trace_binary, requests):
    got = subprocess.run([trace_binary] + [str(n) for n in requests], check=True, capture_output=True,
                         text=True).stdout.splitlines()
    want = _model(requests)
    assert [" ".join(g.split()) for g in got] == [" ".join(w.split()) for w in want]


# ---- the reference's own AviSynth filter, compiled unmodified ---------------------------------
# oracle/Makefile builds avisynth_plugin/src/main.cc from /root/reference against a stand-in SDK
# header (oracle/ref_avisynth/avisynth.h) and this repo's public include/JoshUpscale/core.h, with
# a recording fake runtime (oracle/ref_avisynth/driver.cc).

REF_MAIN = "/root/reference/avisynth_plugin/src/main.cc"
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "avisynth_trace")


@pytest.fixture(scope="module")
def reference_filter_binary():
    if os.path.exists(REF_MAIN) and shutil.which("make") and shutil.which("g++"):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "_ref/avisynth_trace"], check=True,
                       capture_output=True)
    if not os.path.exists(REF_BIN):
        pytest.skip("reference sources not available and oracle/_ref/avisynth_trace not prebuilt")
    return REF_BIN


@pytest.mark.parametrize("requests", SCRIPTS)
def test_reference_avisynth_filter_matches_sequencer(reference_filter_binary, requests):
    """Request by request, the reference's GetFrame must return the same frame and feed the same
    source frames to processImage as the FrameSequencer policy (model, C++ header, Python twin)."""
    got = subprocess.run([reference_filter_binary] + [str(n) for n in requests], check=True, capture_output=True,
                         text=True).stdout.splitlines()
    def head(line):  # "<n> -> <id> | <processed ...>" without the statistics the reference does not expose
        first, processed = line.split("|")[:2]
        return " ".join((first.strip() + " | " + processed.strip()).split())

    assert [head(g + " |") for g in got] == [head(w) for w in _model(requests)]
    assert [head(g + " |") for g in got] == [head(w) for w in _python(requests)]


def _random_scripts(count=24, seed=2024):
    """Playback with scrubbing: mostly +1, some small steps back / forward, occasional far seeks."""
    rng = np.random.default_rng(seed)
    scripts = []
    for _ in range(count):
        n, out = int(rng.integers(0, 50)), []
        for _ in range(int(rng.integers(20, 70))):
            r = rng.random()
            if r < 0.55:
                n += 1
            elif r < 0.75:
                n -= int(rng.integers(1, 6))
            elif r < 0.88:
                n += int(rng.integers(2, 20))
            elif r < 0.95:
                n -= int(rng.integers(10, 40))
            else:
                n += int(rng.integers(20, 200))
            n = max(n, 0)
            out.append(n)
        scripts.append(out)
    return scripts


def test_random_scrubbing_reference_header_and_twin_agree(reference_filter_binary, trace_binary):
    """Randomised request sequences: the reference's compiled AviSynth filter, the C++ header and
    the Python twin must return the same frames and issue the same processImage calls."""
    def head(line):
        first, processed = line.split("|")[:2]
        return " ".join((first.strip() + " | " + processed.strip()).split())

    for requests in _random_scripts():
        args = [str(n) for n in requests]
        ref = subprocess.run([reference_filter_binary] + args, check=True, capture_output=True, text=True)
        hdr = subprocess.run([trace_binary] + args, check=True, capture_output=True, text=True)
        ref_lines = [head(g + " |") for g in ref.stdout.splitlines()]
        assert ref_lines == [head(g) for g in hdr.stdout.splitlines()], requests
        assert ref_lines == [head(g) for g in _python(requests)], requests
