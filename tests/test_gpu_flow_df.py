"""The persistent flow-net kernel (csrc/kernels/flow_df_tc.cu: every layer of get_flow_autoencoder,
scripts/training/models.py:334-481, in one cooperative launch, layers chained by release / acquire
counters; opt-in with JU_FUSED_FLOW=1 because it measured slower than one launch per layer)
performs the same arithmetic as one conv_tc / upscale2 launch per layer: the flow field and the
frames must be BIT-identical, for every size, batch, chunking and history length."""

import dataclasses
import os

import numpy as np
import pytest

from joshupscale_b200 import config as jcfg
from joshupscale_b200 import runtime as jrt
from joshupscale_b200 import synthetic
from tests.gpu_util import make_model, require_gpu

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, scope="module")
def _gpu():
    require_gpu()


def _run(path, clips, env, nframes):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        with jrt.Runtime(path, 0, len(clips)) as rt:
            frames, heads = [], []
            for t in range(nframes):
                frames.append(np.stack(rt.process_batch([c[t] for c in clips])))
                heads.append(rt.read_tensor("flow_head").copy())
            kernels = rt.info.kernels_per_frame
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return np.stack(frames), np.stack(heads), kernels


@pytest.mark.parametrize("preset,batch,nframes", [
    ("small", 1, 5), ("small", 3, 4), ("small_bright", 2, 4), ("psp_fast", 1, 4), ("psp_fast", 2, 3),
    ("ps2_fast", 1, 2)])
def test_persistent_flow_net_bit_identical_to_per_layer_launches(tmp_path, preset, batch, nframes):
    cfg = jcfg.preset(preset)
    if preset == "ps2_fast":
        cfg = dataclasses.replace(cfg, gen_blocks=2)
    cfg, _, path = make_model(tmp_path, cfg)
    clips = [synthetic.frames(cfg.frame_height, cfg.frame_width, nframes, stream_id=s) for s in range(batch)]
    got, got_head, k_fused = _run(path, clips, {"JU_FUSED_FLOW": "1"}, nframes)
    want, want_head, k_layers = _run(path, clips, {"JU_FUSED_FLOW": "0"}, nframes)
    assert k_fused < k_layers - 10, (k_fused, k_layers)  # 18 flow launches became one
    np.testing.assert_array_equal(got_head.view(np.uint32), want_head.view(np.uint32))
    np.testing.assert_array_equal(got, want)


def test_stream_chunks_and_odd_filter_lists(tmp_path):
    """5 streams in chunks of 1 / 2 / all, and a model without the trailing flow/conv_1 (even filter
    list): the same bytes as the per-layer path."""
    cfg = dataclasses.replace(jcfg.preset("small"), flow_filters=(32, 64, 64, 32))
    cfg, _, path = make_model(tmp_path, cfg)
    clips = [synthetic.frames(cfg.frame_height, cfg.frame_width, 3, stream_id=s) for s in range(5)]
    want, want_head, _ = _run(path, clips, {"JU_FUSED_FLOW": "0"}, 3)
    for chunk in ("1", "2", "0"):
        got, got_head, _ = _run(path, clips, {"JU_FUSED_FLOW": "1", "JU_FLOW_SUBBATCH": chunk}, 3)
        np.testing.assert_array_equal(got_head.view(np.uint32), want_head.view(np.uint32))
        np.testing.assert_array_equal(got, want)


def test_flow_net_race_stress(tmp_path):
    """Many replays at full size: a missing dependency between layers shows up as a frame that
    differs from the first run (and from the per-layer launches)."""
    cfg, _, path = make_model(tmp_path, "psp_fast")
    frames = synthetic.frames(270, 480, 4)
    os.environ["JU_FUSED_FLOW"] = "1"
    try:
        with jrt.Runtime(path) as rt:
            first = None
            for _ in range(25):
                rt.reset_state()
                got = np.stack([rt.process(f) for f in frames])
                if first is None:
                    first = got
                else:
                    np.testing.assert_array_equal(got, first)
    finally:
        os.environ.pop("JU_FUSED_FLOW", None)
    want, _, _ = _run(path, [frames], {"JU_FUSED_FLOW": "0"}, 4)
    np.testing.assert_array_equal(first, want[:, 0])


def test_stalled_flow_net_is_a_recoverable_exception(tmp_path):
    cfg, _, path = make_model(tmp_path, "small")
    frames = synthetic.frames(cfg.frame_height, cfg.frame_width, 3)
    with jrt.Runtime(path) as rt:
        want = [rt.process(f).copy() for f in frames]
    os.environ["JU_WAIT_TIMEOUT_MS"] = "250"
    os.environ["JU_FUSED_FLOW"] = "1"
    try:
        with jrt.Runtime(path) as rt:
            np.testing.assert_array_equal(rt.process(frames[0]), want[0])
            rt.inject_stall(4)
            with pytest.raises(jrt.JoshUpscaleError, match="frame aborted.*flow_df_tc_kernel"):
                rt.process(frames[1])
            for t in (1, 2):
                np.testing.assert_array_equal(rt.process(frames[t]), want[t])
    finally:
        os.environ.pop("JU_WAIT_TIMEOUT_MS", None)
        os.environ.pop("JU_FUSED_FLOW", None)
