"""Output temporal filter (scripts/inference/onnx/frame_moving_avg.py): the oracle
restatement against an independent plain-loop restatement and against the
properties the script's mask arithmetic implies."""

import numpy as np
import pytest
import torch

from joshupscale_b200 import config as jcfg
from joshupscale_b200 import weights as jw
from oracle import frame_filter as ff


def _naive(out, pw, flt):
    """Direct per-pixel loops over NHWC float arrays (N = 1), float64 accumulation."""
    out = out[0].astype(np.float64)
    pw = pw[0].astype(np.float64)
    hh, ww, _ = out.shape
    if flt.limit:
        pw = np.clip(pw, -0.5, 0.5)
    gain_coef = 1.0 if flt.gain == 0 else flt.gain
    luma = ff.LUMA_NORM.astype(np.float64)
    if flt.norm == "l2":
        luma = luma * luma
    d = np.abs(out - pw) if flt.norm == "l1" else (out - pw) ** 2
    if flt.luma_normalize:
        d = d * luma
    fn = np.sign if flt.gain == 0 else np.tanh
    cond = np.zeros((hh, ww))
    if flt.window == 0:
        cond[:] = fn(d.mean() * gain_coef - flt.threshold * gain_coef)
    else:
        w = flt.window
        oh, ow = -(-hh // w) * w, -(-ww // w) * w
        pt, pl = (oh - hh) // 2, (ow - ww) // 2
        cells = np.zeros((oh // w, ow // w))
        for cy in range(oh // w):
            for cx in range(ow // w):
                acc = 0.0
                for y in range(cy * w - pt, cy * w - pt + w):
                    for x in range(cx * w - pl, cx * w - pl + w):
                        if 0 <= y < hh and 0 <= x < ww:
                            acc += d[y, x].sum()
                cells[cy, cx] = fn(acc / (3 * w * w) * gain_coef - flt.threshold * gain_coef)
        for y in range(hh):
            sy = (y + pt) / w
            y0 = int(np.floor(sy)); y1 = min(y0 + 1, cells.shape[0] - 1); ty = sy - y0
            for x in range(ww):
                sx = (x + pl) / w
                x0 = int(np.floor(sx)); x1 = min(x0 + 1, cells.shape[1] - 1); tx = sx - x0
                top = cells[y0, x0] * (1 - tx) + cells[y0, x1] * tx
                bot = cells[y1, x0] * (1 - tx) + cells[y1, x1] * tx
                cond[y, x] = top * (1 - ty) + bot * ty
    s = flt.strength
    mask = cond * (-s / 2) + s / 2
    mask2 = cond * (s / 2) + (1 - s / 2)
    return (pw * mask[..., None] + out * mask2[..., None])[None]


def _pair(h, w, seed, delta):
    rng = np.random.default_rng(seed)
    pw = rng.uniform(-0.6, 0.6, (1, h, w, 3)).astype(np.float32)
    out = np.clip(pw + rng.normal(0, delta, pw.shape), -0.5, 0.5).astype(np.float32)
    return out, pw


@pytest.mark.parametrize("flt", [
    ff.FrameFilter(),
    ff.FrameFilter(strength=0.5, threshold=0.02, norm="l2", limit=True),
    ff.FrameFilter(gain=8.0, luma_normalize=True),
    ff.FrameFilter(window=8, threshold=0.05),
    ff.FrameFilter(window=16, gain=4.0, norm="l2", luma_normalize=True, limit=True, threshold=0.01),
    ff.FrameFilter(window=5, strength=0.4),  # 5 does not divide 36 x 44: centred zero padding
])
def test_oracle_filter_matches_plain_loops(flt):
    for delta in (0.02, 0.3):
        out, pw = _pair(36, 44, 3, delta)
        got = ff.frame_moving_avg(torch.from_numpy(out), torch.from_numpy(pw), flt).numpy()
        want = _naive(out, pw, flt)
        assert np.abs(got - want).max() < 2e-6


def test_scene_cut_passes_output_and_static_scene_blends():
    flt = ff.FrameFilter(strength=0.25, threshold=0.1)
    out, pw = _pair(16, 16, 1, 0.0)
    pw = np.clip(pw, -0.5, 0.5)
    out = pw.copy()
    out[0, 0, 0, 0] += 0.01  # tiny difference: mean << threshold -> cond = -1
    got = ff.frame_moving_avg(torch.from_numpy(out), torch.from_numpy(pw), flt).numpy()
    np.testing.assert_allclose(got, 0.25 * pw + 0.75 * out, atol=1e-7)
    cut = -pw  # large difference everywhere -> cond = +1 -> output passes through
    got = ff.frame_moving_avg(torch.from_numpy(cut), torch.from_numpy(pw), flt).numpy()
    np.testing.assert_allclose(got, cut, atol=1e-7)


def test_streams_are_filtered_independently():
    flt = ff.FrameFilter()
    a_out, a_pw = _pair(12, 12, 5, 0.01)
    b_out, b_pw = _pair(12, 12, 6, 0.4)
    both = ff.frame_moving_avg(torch.from_numpy(np.concatenate([a_out, b_out])),
                               torch.from_numpy(np.concatenate([a_pw, b_pw])), flt).numpy()
    one = ff.frame_moving_avg(torch.from_numpy(a_out), torch.from_numpy(a_pw), flt).numpy()
    np.testing.assert_array_equal(both[:1], one)


def test_filter_round_trips_through_the_model_container(tmp_path):
    cfg = jcfg.preset("tiny")
    w = jw.init_weights(cfg, 1)
    flt = jcfg.OutputFilter(strength=0.3, window=16, threshold=0.05, gain=2.0, norm="l2", limit=True)
    path = str(tmp_path / "m.jup")
    jw.save_model(path, cfg, jw.with_output_filter(w, flt))
    _, back = jw.load_model(path)
    np.testing.assert_array_equal(back[jw.FILTER_TENSOR], np.asarray(flt.as_vector(), np.float32))
    # oracle-side description of the same filter agrees field by field
    ofl = ff.FrameFilter(strength=0.3, window=16, threshold=0.05, gain=2.0, norm="l2", limit=True)
    np.testing.assert_array_equal(ofl.as_vector(), back[jw.FILTER_TENSOR])
    assert jw.FILTER_TENSOR not in jw.with_output_filter(back, None)
    with pytest.raises(ValueError):
        jcfg.OutputFilter(norm="l3").as_vector()


# ---- pinned against the reference script itself -----------------------------------------------
# tests/golden/filter_golden.npz was produced by tests/golden/make_filter_golden.py, which runs
# /root/reference/scripts/inference/onnx/frame_moving_avg.py main() unmodified against recording
# stand-ins for onnx / graph.Graph and evaluates the op list it builds.

_GOLDEN_CASES = {
    "defaults": dict(),
    "l2_limit": dict(strength=0.5, threshold=0.02, norm="l2", limit=True),
    "gain_luma": dict(gain=8.0, luma_normalize=True),
    "window8": dict(window=8, threshold=0.05),
    "window16_all": dict(window=16, gain=4.0, norm="l2", luma_normalize=True, limit=True, threshold=0.01),
    "window5": dict(window=5, strength=0.4),
}


@pytest.mark.parametrize("name", sorted(_GOLDEN_CASES))
def test_oracle_filter_reproduces_the_reference_script(name):
    import os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "filter_golden.npz"))
    flt = ff.FrameFilter(**_GOLDEN_CASES[name])
    for k, delta in enumerate((0.02, 0.3)):
        rng = np.random.default_rng(100 + k)  # same inputs as make_filter_golden.case_inputs
        pw = rng.uniform(-0.6, 0.6, (1, 36, 44, 3)).astype(np.float32)
        out = np.clip(pw + rng.normal(0, delta, pw.shape), -0.5, 0.5).astype(np.float32)
        got = ff.frame_moving_avg(torch.from_numpy(out), torch.from_numpy(pw), flt).numpy()
        want = gold[f"{name}/{k}"]
        assert want.shape == got.shape
        assert np.abs(got - want).max() < 2e-6, (name, k, np.abs(got - want).max())
        # the gate must have an effect in at least one of the two scenes
    steady, cut = gold[f"{name}/0"], gold[f"{name}/1"]
    assert np.abs(steady).max() > 0.1 and np.abs(cut).max() > 0.1
