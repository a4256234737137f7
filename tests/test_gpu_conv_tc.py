"""The tcgen05 implicit-GEMM convolution (conv_tc.cu) against the oracle and
against the SIMT reference kernel, through the C-ABI (ju_launch_conv impl=1)."""

import os

import numpy as np
import pytest
import torch

from joshupscale_b200 import kernels as jk
from joshupscale_b200 import runtime as jrt
from joshupscale_b200 import synthetic
from oracle import reference_graph as og
from tests.gpu_util import make_model, r16, require_gpu, u8_stats

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, scope="module")
def _gpu():
    require_gpu()


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, np.float32))


TC_CASES = [
    # b, h, w, cin, cout, ks  (tile = 16 rows x 8 cols: ragged edges, multi k-block, multi n-tile)
    (1, 16, 8, 64, 64, 3), (1, 20, 13, 64, 64, 3), (2, 48, 40, 64, 64, 3), (1, 1, 1, 64, 64, 3),
    (1, 17, 9, 51, 64, 3), (1, 16, 16, 128, 64, 3), (1, 32, 24, 64, 128, 3), (1, 16, 16, 12, 32, 3),
    (1, 16, 24, 256, 256, 3), (1, 34, 60, 128, 256, 3), (1, 23, 31, 64, 32, 1), (1, 270, 480, 64, 64, 3),
]


@pytest.mark.parametrize("b,h,w,cin,cout,ks", TC_CASES)
@pytest.mark.parametrize("mode", ["plain", "residual_relu", "lrelu_f32"])
def test_conv_tc_vs_oracle(b, h, w, cin, cout, ks, mode):
    rng = np.random.default_rng(cin * 1000 + cout + ks + h)
    x = r16(rng.standard_normal((b, h, w, cin)) * 0.5)
    k = (rng.standard_normal((ks, ks, cin, cout)) / np.sqrt(ks * ks * cin)).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    bias = rng.standard_normal(cout).astype(np.float32) * 0.1
    res = r16(rng.standard_normal((b, h, w, cout)) * 0.5)
    kw = dict(scale=scale, impl=jk.IMPL_TCGEN05)
    want = og.conv2d_same(_t(x), og.r16(_t(k * scale))).numpy()
    if mode == "residual_relu":
        kw.update(bias=bias, residual=res, act=jk.ACT_RELU)
        want = np.maximum(want + bias + res, 0)
    elif mode == "lrelu_f32":
        kw.update(bias=bias, act=jk.ACT_LRELU, slope=0.3, out_f32=True)
        want = want + bias
        want = np.where(want >= 0, want, want * np.float32(0.3))
    got = jk.conv(x, k, **kw).astype(np.float32)
    if mode != "lrelu_f32":
        np.testing.assert_allclose(got, r16(want), rtol=2e-3, atol=1e-3)
    else:
        np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-4)


def test_conv_tc_equals_simt_kernel_closely():
    """Two independent device implementations of the same layer."""
    rng = np.random.default_rng(11)
    x = r16(rng.standard_normal((1, 40, 56, 64)) * 0.5)
    k = (rng.standard_normal((3, 3, 64, 64)) / 24).astype(np.float32)
    a = jk.conv(x, k, act=jk.ACT_RELU, impl=jk.IMPL_SIMT).astype(np.float32)
    b = jk.conv(x, k, act=jk.ACT_RELU, impl=jk.IMPL_TCGEN05).astype(np.float32)
    np.testing.assert_allclose(a, b, rtol=2e-3, atol=1e-3)


def test_conv_tc_transpose_shuffle():
    rng = np.random.default_rng(5)
    b, h, w, cin, cout = 1, 19, 14, 64, 32
    x = r16(rng.standard_normal((b, h, w, cin)) * 0.5)
    kt = (rng.standard_normal((2, 2, cout, cin)) / 8).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    bias = rng.standard_normal(cout).astype(np.float32) * 0.1
    k1 = np.transpose(kt.reshape(4, cout, cin), (2, 0, 1)).reshape(1, 1, cin, 4 * cout)
    got = jk.conv(x, k1, scale=np.tile(scale, 4), bias=np.tile(bias, 4), act=jk.ACT_RELU,
                  shuffle2=True, impl=jk.IMPL_TCGEN05).astype(np.float32)
    want = og.conv2d_transpose_k2s2(_t(x), og.r16(_t(kt * scale[None, None, :, None]))).numpy() + bias
    np.testing.assert_allclose(got, r16(np.maximum(want, 0)), rtol=2e-3, atol=1e-3)


def test_halo_layout_variants_agree():
    rng = np.random.default_rng(12)
    x = r16(rng.standard_normal((1, 33, 29, 64)) * 0.5)
    k = (rng.standard_normal((3, 3, 64, 64)) / 24).astype(np.float32)
    outs = []
    for v in (0, 1):
        jrt.set_option("tc_variant", v)
        outs.append(jk.conv(x, k, impl=jk.IMPL_TCGEN05))
    jrt.set_option("tc_variant", 0)
    np.testing.assert_array_equal(outs[0].view(np.uint16), outs[1].view(np.uint16))


def test_engine_runs_on_tensor_cores_and_matches_simt_engine(tmp_path):
    """Whole graph with tcgen05 convs vs the whole graph with SIMT convs."""
    cfg, w, path = make_model(tmp_path, "small")
    frames = synthetic.frames(cfg.frame_height, cfg.frame_width, 5)
    outs = {}
    for impl in ("1", "0"):
        os.environ["JU_CONV_IMPL"] = impl
        try:
            with jrt.Runtime(path) as rt:
                assert rt.info.conv_impl == int(impl)
                outs[impl] = np.stack([rt.process(f) for f in frames])
        finally:
            os.environ.pop("JU_CONV_IMPL", None)
    m, frac, psnr = u8_stats(outs["1"], outs["0"])
    assert m <= 1 and frac < 0.06 and psnr > 55
    ref, _ = og.Graph(cfg, w, "fp32").run(frames)
    m32, _, p32 = u8_stats(outs["1"][..., :3], ref[..., :3])
    assert m32 <= 2 and p32 >= 45.0


@pytest.mark.parametrize("h,w,act", [(16, 8, "relu"), (21, 27, "relu"), (33, 20, "lrelu"), (270, 480, "relu")])
def test_fused_tail_vs_oracle(h, w, act):
    """conv_trans_1+BN+act -> conv_trans_2+tanh -> +bilinear x4 -> clip -> u8/state,
    one tcgen05 kernel, against the unfused oracle ops (fp16-rounded intermediate)."""
    rng = np.random.default_rng(h * 100 + w)
    b = 2 if h < 100 else 1
    trunk = r16(rng.standard_normal((b, h, w, 64)) * 0.5)
    kt1 = (rng.standard_normal((2, 2, 32, 64)) / 8).astype(np.float32)
    scale1 = rng.uniform(0.5, 1.5, 32).astype(np.float32)
    bias1 = (rng.standard_normal(32) * 0.1).astype(np.float32)
    w2 = r16(rng.standard_normal((2, 2, 3, 32)) * 0.3)
    b2 = (rng.standard_normal(3) * 0.05).astype(np.float32)
    frames = rng.integers(0, 256, (b, h, w, 4), dtype=np.uint8)
    out, state, raw = jk.tail(trunk, kt1, scale1, bias1, w2, b2, frames,
                              act=jk.ACT_RELU if act == "relu" else jk.ACT_LRELU)
    mid = og.conv2d_transpose_k2s2(_t(trunk), og.r16(_t(kt1 * scale1[None, None, :, None]))) + _t(bias1)
    mid = og.r16(og.activation(mid, act))
    z = torch.tanh(og.conv2d_transpose_k2s2(mid, _t(w2), _t(b2)))
    want_raw = torch.clamp(og.resize_bilinear_legacy(og.preprocess(frames[..., :3]), 4) + z, -0.5, 0.5).numpy()
    # the fp16 rounding of the intermediate can flip by one ulp (accumulation order)
    assert np.abs(raw - want_raw).max() < 2e-3 and np.abs(raw - want_raw).mean() < 2e-5
    own = ((raw + np.float32(0.5)) * np.float32(255)).astype(np.uint8)
    np.testing.assert_array_equal(out[..., :3], own)
    assert int(out[..., 3].max()) == 0
    np.testing.assert_array_equal(state[..., :3].view(np.uint16), raw.astype(np.float16).view(np.uint16))
    d = np.abs(out[..., :3].astype(int) - og.postprocess(torch.from_numpy(want_raw)).numpy().astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 0.02


def test_fused_tail_engine_equals_unfused_engine(tmp_path):
    cfg, w, path = make_model(tmp_path, "small")
    frames = synthetic.frames(cfg.frame_height, cfg.frame_width, 4)
    outs = {}
    for fused in ("1", "0"):
        os.environ["JU_FUSED_TAIL"] = fused
        try:
            with jrt.Runtime(path) as rt:
                outs[fused] = np.stack([rt.process(f) for f in frames])
        finally:
            os.environ.pop("JU_FUSED_TAIL", None)
    m, frac, psnr = u8_stats(outs["1"], outs["0"])
    assert m <= 1 and frac < 0.03


@pytest.mark.parametrize("b,h,w,cin,cout", [(1, 16, 8, 64, 64), (2, 36, 44, 32, 32), (1, 272, 480, 32, 32), (1, 68, 120, 64, 64)])
def test_conv_tc_fused_maxpool(b, h, w, cin, cout):
    """conv + BN + ReLU + MaxPool2D(2) in one kernel == maxpool(conv) (models.py:386-409)."""
    rng = np.random.default_rng(h + w + cin)
    x = r16(rng.standard_normal((b, h, w, cin)) * 0.5)
    k = (rng.standard_normal((3, 3, cin, cout)) / np.sqrt(9 * cin)).astype(np.float32)
    bias = rng.standard_normal(cout).astype(np.float32) * 0.1
    full = jk.conv(x, k, bias=bias, act=jk.ACT_RELU, impl=jk.IMPL_TCGEN05, cout_stride=64)
    pooled = jk.conv(x, k, bias=bias, act=jk.ACT_RELU, impl=jk.IMPL_TCGEN05, cout_stride=64, pool=True)
    want = og.max_pool2(_t(full.astype(np.float32))).numpy().astype(np.float16)
    assert pooled.shape == (b, h // 2, w // 2, cout)
    np.testing.assert_array_equal(pooled.view(np.uint16), want.view(np.uint16))


@pytest.mark.parametrize("preset,batch", [("small", 1), ("small", 3), ("psp_fast", 1), ("psp_fast", 2)])
def test_persistent_trunk_bit_identical_to_per_layer_launches(tmp_path, preset, batch):
    """The persistent ResBlock trunk (one launch, per-wave dataflow counters between layers)
    performs the same arithmetic as one conv_tc launch per layer: bit-identical outputs."""
    cfg, w, path = make_model(tmp_path, preset)
    nframes = 6
    frames = [synthetic.frames(cfg.frame_height, cfg.frame_width, nframes, stream_id=s) for s in range(batch)]
    outs = {}
    for fused in ("1", "0"):
        os.environ["JU_FUSED_TRUNK"] = fused
        try:
            with jrt.Runtime(path, 0, batch) as rt:
                res = []
                for t in range(nframes):
                    res.append(np.stack(rt.process_batch([f[t] for f in frames])))
                outs[fused] = np.stack(res)
        finally:
            os.environ.pop("JU_FUSED_TRUNK", None)
    np.testing.assert_array_equal(outs["1"], outs["0"])


def test_dataflow_trunk_race_stress(tmp_path):
    """Many replays of the dataflow trunk at full size with deep (quality) and shallow models:
    any missing dependency between tiles shows up as a frame that differs from the first run."""
    for preset, nframes in (("psp_quality", 24), ("psp_fast", 40)):
        cfg, w, path = make_model(tmp_path, preset)
        frames = synthetic.frames(270, 480, 4)
        with jrt.Runtime(path) as rt:
            first = None
            for rep in range(nframes // 4):
                rt.reset_state()
                got = np.stack([rt.process(f) for f in frames])
                if first is None:
                    first = got
                else:
                    np.testing.assert_array_equal(got, first)
        os.environ["JU_FUSED_TRUNK"] = "0"
        try:
            with jrt.Runtime(path) as rt:
                ref = np.stack([rt.process(f) for f in frames])
        finally:
            os.environ.pop("JU_FUSED_TRUNK", None)
        np.testing.assert_array_equal(first, ref)


@pytest.mark.parametrize("env", [
    {"JU_NO_GRAPH": "1"},
    {"JU_TC_DUAL": "0"},
    {"JU_TC_PDL": "0"},
    {"JU_FUSED_POOL": "0"},
    {"JU_TRUNK_SUBBATCH": "0"},
    {"JU_TRUNK_SUBBATCH": "1"},
    {"JU_TAIL_BANDS": "1"},
    {"JU_TRUNK_LEAD": "0"},
    {"JU_TAIL_BANDS": "5", "JU_NO_GRAPH": "1"},
    {"JU_FUSED_FLOW": "1"},
    {"JU_FUSED_FLOW": "1", "JU_FLOW_SUBBATCH": "1"},
    {"JU_FUSED_FLOW": "1", "JU_FLOW_SUBBATCH": "2", "JU_TC_DUAL": "0"},
    {"JU_TRUNK_COOP": "0"},
    {"JU_COPY_THREADS": "0"},
    {"JU_TRUNK_PAIR": "1"},
    {"JU_TRUNK_PAIR": "1", "JU_TRUNK_COOP": "0", "JU_TRUNK_SUBBATCH": "1"},
    {"JU_COPY_SPIN_US": "0"},                              # copy pool parks and blocks on every band event
    {"JU_COPY_THREADS": "1", "JU_COPY_SPIN_US": "0"},      # the calling thread alone
    {"JU_TRUNK_SUBBATCH": "1", "JU_TAIL_LAST": "2"},       # tail groups (0), (1, 2)
    {"JU_TRUNK_SUBBATCH": "1", "JU_TAIL_GROUP": "1"},      # one tail per stream
    {"JU_TRUNK_SUBBATCH": "1", "JU_TAIL_GROUP": "3", "JU_TAIL_LAST": "3"},
])
def test_execution_switches_do_not_change_the_bytes(tmp_path, env):
    """Scheduling / fusion switches only change HOW the frame is executed:
    the output bytes of a 3-stream, 3-frame run must equal the default configuration's."""
    cfg, _, path = make_model(tmp_path, "psp_fast")
    clips = [synthetic.frames(cfg.frame_height, cfg.frame_width, 3, stream_id=s) for s in range(3)]

    def run():
        with jrt.Runtime(path, 0, 3) as rt:
            batched = [[o.copy() for o in rt.process_batch([c[t] for c in clips])] for t in range(3)]
        with jrt.Runtime(path, 0, 1) as rt:  # batch 1: banded-tail host path
            single = [rt.process(clips[0][t]).copy() for t in range(3)]
        return batched, single

    want_b, want_s = run()
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        got_b, got_s = run()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    for t in range(3):
        np.testing.assert_array_equal(got_s[t], want_s[t])
        np.testing.assert_array_equal(want_s[t], want_b[t][0])
        for s in range(3):
            np.testing.assert_array_equal(got_b[t][s], want_b[t][s])


def test_flow_resnet_runs_on_the_persistent_trunk_bit_identically(tmp_path):
    """get_flow_resnet is conv_1 + ResBlocks of the generator's shape: the engine runs them with the
    persistent trunk kernel.  Same bytes as one conv launch per layer, also with sub-batches."""
    from joshupscale_b200 import config as jcfg
    cfg = jcfg.ModelConfig(frame_height=136, frame_width=240, gen_blocks=2, flow_arch="resnet",
                           flow_resnet_blocks=4, flow_pad_factor=0)
    cfg2, _, path = make_model(tmp_path, cfg)
    clips = [synthetic.frames(cfg.frame_height, cfg.frame_width, 3, stream_id=s) for s in range(3)]

    def run():
        with jrt.Runtime(path, 0, 3) as rt:
            names = [o["name"] for o in rt.profile_ops(1)]
            rt.reset_state()
            return names, [[o.copy() for o in rt.process_batch([c[t] for c in clips])] for t in range(3)]

    names, want = run()
    assert any(n.startswith("flow/block_*(persistent)") for n in names), names
    os.environ["JU_FUSED_TRUNK"] = "0"
    try:
        names0, got = run()
    finally:
        os.environ.pop("JU_FUSED_TRUNK", None)
    assert not any("persistent" in n for n in names0)
    os.environ["JU_TRUNK_SUBBATCH"] = "1"
    try:
        _, got1 = run()
    finally:
        os.environ.pop("JU_TRUNK_SUBBATCH", None)
    for t in range(3):
        for s in range(3):
            np.testing.assert_array_equal(got[t][s], want[t][s])
            np.testing.assert_array_equal(got1[t][s], want[t][s])
