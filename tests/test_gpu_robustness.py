"""Concurrency and failure behaviour of the persistent kernels (ADVICE high):
two full-size runtimes driven from two threads on one device, and a stalled pipeline surfacing as
a recoverable exception instead of a trap."""

import os
import threading
import zlib

import numpy as np
import pytest

from joshupscale_b200 import runtime as jrt
from joshupscale_b200 import synthetic
from tests.gpu_util import make_model, require_gpu

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, scope="module")
def _gpu():
    require_gpu()


def _crcs(rt, frames, n):
    out = np.empty(rt.out_shape, np.uint8)
    crcs = []
    for t in range(n):
        rt.process(frames[t % len(frames)], out)
        crcs.append(zlib.crc32(out))
    return crcs


def test_two_full_size_runtimes_two_threads_one_device(tmp_path):
    """Two psp_fast runtimes (each trunk launch wants all 148 SMs with 1 CTA/SM and has inter-CTA
    dependencies) hammered from two threads: 2000 frames each, every frame's CRC equal to a solo
    run.  Without the per-device turn-taking the two trunks can interleave and starve each other."""
    n = 2000
    cfg, _, path_a = make_model(tmp_path, "psp_fast", seed=5)
    _, _, path_b = make_model(tmp_path, "psp_fast", seed=6)
    fa = synthetic.frames(270, 480, 12, stream_id=1)
    fb = synthetic.frames(270, 480, 12, stream_id=2)
    with jrt.Runtime(path_a) as ra:
        want_a = _crcs(ra, fa, n)
    with jrt.Runtime(path_b) as rb:
        want_b = _crcs(rb, fb, n)
    got, errors = {}, []

    def worker(name, rt, frames):
        try:
            got[name] = _crcs(rt, frames, n)
        except Exception as e:  # noqa: BLE001 - reported below
            errors.append((name, repr(e)))

    with jrt.Runtime(path_a) as ra, jrt.Runtime(path_b) as rb:
        ths = [threading.Thread(target=worker, args=("a", ra, fa)),
               threading.Thread(target=worker, args=("b", rb, fb))]
        for th in ths:
            th.start()
        for th in ths:
            th.join()
    assert not errors, errors
    assert got["a"] == want_a and got["b"] == want_b


@pytest.mark.parametrize("preset", ["small", "psp_fast"])
def test_injected_stall_is_a_recoverable_exception(tmp_path, preset):
    """A pipeline wait that never completes must end the frame with an exception (the reference
    throws from a failed enqueue, tensorrt_backend.cc:266) - not with a trap that poisons the CUDA
    context.  Afterwards: the recurrent state was not advanced, the same runtime keeps working, and
    a new runtime can be created in the same process."""
    cfg, _, path = make_model(tmp_path, preset)
    frames = synthetic.frames(cfg.frame_height, cfg.frame_width, 4)
    with jrt.Runtime(path) as rt:
        want = [rt.process(f).copy() for f in frames]
    os.environ["JU_WAIT_TIMEOUT_MS"] = "250"
    try:
        with jrt.Runtime(path) as rt:
            np.testing.assert_array_equal(rt.process(frames[0]), want[0])
            rt.inject_stall(1)
            with pytest.raises(jrt.JoshUpscaleError, match="frame aborted.*trunk_df_tc_kernel"):
                rt.process(frames[1])
            # the failed frame did not flip the state: the stream continues where it was
            for t in (1, 2, 3):
                np.testing.assert_array_equal(rt.process(frames[t]), want[t])
            with jrt.Runtime(path) as other:
                np.testing.assert_array_equal(other.process(frames[0]), want[0])
    finally:
        os.environ.pop("JU_WAIT_TIMEOUT_MS", None)
    with jrt.Runtime(path) as rt:
        np.testing.assert_array_equal(rt.process(frames[0]), want[0])
