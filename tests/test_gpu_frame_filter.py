"""frame_moving_avg output filter on the GPU (csrc/kernels/frame_filter.cu) against
the oracle graph with the same filter, through the runtime entry point."""

import os

import numpy as np
import pytest

from joshupscale_b200 import config as jcfg
from joshupscale_b200 import runtime as jrt
from joshupscale_b200 import synthetic
from joshupscale_b200 import weights as jw
from oracle import frame_filter as ff
from oracle import reference_graph as og
from tests.gpu_util import require_gpu, u8_stats

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, scope="module")
def _gpu():
    require_gpu()


def _model(tmp_path, preset, flt, tag):
    cfg = jcfg.preset(preset)
    w = jw.init_weights(cfg, 42, True)
    path = os.path.join(str(tmp_path), f"m_{tag}.jup")
    jw.save_model(path, cfg, jw.with_output_filter(w, flt))
    return cfg, w, path


CASES = [
    ("small", dict()),                                            # script defaults: global L1, sign
    ("small", dict(strength=0.5, threshold=0.02, norm="l2", limit=True)),
    ("small", dict(gain=6.0, luma_normalize=True, threshold=0.05)),
    ("small", dict(window=16, threshold=0.03)),
    ("small", dict(window=24, gain=5.0, norm="l2", luma_normalize=True, limit=True, threshold=0.004)),
    ("tiny", dict(window=7, strength=0.4, threshold=0.02)),       # 7 divides neither 84 nor 108
    ("small_bright", dict(threshold=0.05)),                        # with brightness normalisation
]


@pytest.mark.parametrize("preset,kw", CASES)
def test_filtered_output_matches_oracle(tmp_path, preset, kw):
    cfg, w, path = _model(tmp_path, preset, jcfg.OutputFilter(**kw), "f")
    # pan with a hard cut in the middle: both branches of the scene gate are taken
    frames = synthetic.frames(cfg.frame_height, cfg.frame_width, 8, kind="cut")
    with jrt.Runtime(path, 0, 1) as rt:
        got = np.stack([rt.process(f) for f in frames])
    oflt = ff.FrameFilter(**kw)
    ref, _ = og.Graph(cfg, w, "fp32", output_filter=oflt).run(frames)
    emu, _ = og.Graph(cfg, w, "fp16emu", output_filter=oflt).run(frames)
    plain, _ = og.Graph(cfg, w, "fp32").run(frames)
    assert not got[..., 3].any()
    # the filter must actually change the picture on this sequence
    assert (ref[..., :3] != plain[..., :3]).mean() > 0.05
    for t in range(len(frames)):
        m16, frac16, _ = u8_stats(got[t, ..., :3], emu[t, ..., :3])
        m32, _, psnr32 = u8_stats(got[t, ..., :3], ref[t, ..., :3])
        assert m16 <= 1 and frac16 < 0.08, (t, m16, frac16)
        assert m32 <= 2 and psnr32 >= 45.0, (t, m32, psnr32)


def test_filter_off_is_bit_identical_to_no_filter(tmp_path):
    cfg, w, path_plain = _model(tmp_path, "small", None, "plain")
    off = jw.with_output_filter(w, jcfg.OutputFilter())
    off[jw.FILTER_TENSOR] = off[jw.FILTER_TENSOR].copy()
    off[jw.FILTER_TENSOR][0] = 0.0  # present but disabled
    path_off = os.path.join(str(tmp_path), "m_off.jup")
    jw.save_model(path_off, cfg, off)
    frames = synthetic.frames(cfg.frame_height, cfg.frame_width, 4)
    with jrt.Runtime(path_plain, 0, 1) as a, jrt.Runtime(path_off, 0, 1) as b:
        for f in frames:
            np.testing.assert_array_equal(a.process(f), b.process(f))


def test_filter_batched_streams_and_full_size(tmp_path):
    flt = jcfg.OutputFilter(window=32, threshold=0.03)
    cfg, w, path = _model(tmp_path, "psp_fast", flt, "psp")
    streams = [synthetic.frames(cfg.frame_height, cfg.frame_width, 3, stream_id=s,
                                kind="cut" if s else "pan") for s in range(2)]
    with jrt.Runtime(path, 0, 2) as rt2:
        batched = [rt2.process_batch([streams[0][t], streams[1][t]]) for t in range(3)]
    for s in range(2):
        with jrt.Runtime(path, 0, 1) as rt:
            for t in range(3):
                np.testing.assert_array_equal(batched[t][s], rt.process(streams[s][t]))
    # full-size parity of the last frame of stream 1 against the fp32 oracle
    ref, _ = og.Graph(cfg, w, "fp32", output_filter=ff.FrameFilter(window=32, threshold=0.03)).run(streams[1])
    m32, _, psnr32 = u8_stats(batched[2][1][..., :3], ref[2, ..., :3])
    assert m32 <= 2 and psnr32 >= 45.0, (m32, psnr32)
