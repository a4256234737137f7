"""The oracle against (a) the reference's own warp source executed under a
numpy tf-shim (tests/golden/warp_golden.npz), (b) an independent naive numpy
restatement of every primitive, (c) torch cross-checks noted in SURVEY.md
appendix A.  CPU only."""

import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from joshupscale_b200 import config as jcfg
from joshupscale_b200 import synthetic, weights as jw
from oracle import naive
from oracle import reference_graph as og

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _rand(rng, *shape):
    return rng.standard_normal(shape).astype(np.float32)


@pytest.mark.parametrize("case", ["small", "border", "subpixel", "integer"])
def test_warp_matches_reference_source(case):
    """Bit-exact against scripts/training/tfa/dense_image_warp.py outputs."""
    g = np.load(os.path.join(GOLDEN, "warp_golden.npz"))
    img, flow, want = g[f"{case}_image"], g[f"{case}_flow"], g[f"{case}_out"]
    got = og.dense_image_warp(torch.from_numpy(img), torch.from_numpy(flow)).numpy()
    np.testing.assert_array_equal(got, want)
    for b in range(img.shape[0]):
        np.testing.assert_array_equal(naive.dense_image_warp(img[b], flow[b]), want[b])


def test_warp_equals_grid_sample_border():
    """replace_dense_warp.py:100-112 swaps the warp for GridSample(border,
    align_corners=0): same sampling positions."""
    rng = np.random.default_rng(0)
    img, flow = _rand(rng, 1, 12, 17, 3), _rand(rng, 1, 12, 17, 2) * 6
    got = og.dense_image_warp(torch.from_numpy(img), torch.from_numpy(flow))
    h, w = 12, 17
    gy, gx = torch.meshgrid(torch.arange(h, dtype=torch.float32),
                            torch.arange(w, dtype=torch.float32), indexing="ij")
    qy = gy - torch.from_numpy(flow[0, ..., 0])
    qx = gx - torch.from_numpy(flow[0, ..., 1])
    grid = torch.stack([qx / (w / 2) + (-1 + 1 / w), qy / (h / 2) + (-1 + 1 / h)], -1)[None]
    ref = F.grid_sample(torch.from_numpy(img).permute(0, 3, 1, 2), grid, mode="bilinear",
                        padding_mode="border", align_corners=False).permute(0, 2, 3, 1)
    assert (got - ref).abs().max() < 2e-5


def test_conv_same_vs_naive():
    rng = np.random.default_rng(1)
    x, k, b = _rand(rng, 7, 9, 5), _rand(rng, 3, 3, 5, 6), _rand(rng, 6)
    got = og.conv2d_same(torch.from_numpy(x)[None], torch.from_numpy(k),
                         torch.from_numpy(b))[0].numpy()
    np.testing.assert_allclose(got, naive.conv2d_same(x, k, b), rtol=1e-5, atol=1e-5)
    k1 = _rand(rng, 1, 1, 5, 4)
    got = og.conv2d_same(torch.from_numpy(x)[None], torch.from_numpy(k1))[0].numpy()
    np.testing.assert_allclose(got, naive.conv2d_same(x, k1), rtol=1e-5, atol=1e-5)


def test_conv_transpose_vs_naive():
    rng = np.random.default_rng(2)
    x, k, b = _rand(rng, 5, 6, 4), _rand(rng, 2, 2, 3, 4), _rand(rng, 3)
    got = og.conv2d_transpose_k2s2(torch.from_numpy(x)[None], torch.from_numpy(k),
                                   torch.from_numpy(b))[0].numpy()
    np.testing.assert_allclose(got, naive.conv2d_transpose_k2s2(x, k, b), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("scale", [2, 4])
def test_legacy_bilinear_vs_naive(scale):
    rng = np.random.default_rng(3)
    x = _rand(rng, 5, 7, 3)
    got = og.resize_bilinear_legacy(torch.from_numpy(x)[None], scale)[0].numpy()
    np.testing.assert_array_equal(got, naive.resize_bilinear_legacy(x, scale))
    # asymmetric: output row 0 == input row 0; last `scale` rows replicate
    np.testing.assert_array_equal(got[0, ::scale], x[0])
    np.testing.assert_array_equal(got[-1, ::scale], x[-1])
    # and it is NOT torch's half-pixel bilinear
    t = F.interpolate(torch.from_numpy(x).permute(2, 0, 1)[None], scale_factor=scale,
                      mode="bilinear", align_corners=False)[0].permute(1, 2, 0).numpy()
    assert np.abs(t - got).max() > 1e-2


def test_space_depth_vs_naive_and_roundtrip():
    rng = np.random.default_rng(4)
    x = _rand(rng, 8, 12, 3)
    s2d = og.space_to_depth(torch.from_numpy(x)[None], 4)[0].numpy()
    np.testing.assert_array_equal(s2d, naive.space_to_depth(x, 4))
    # channel k = (i*4+j)*3+c  <- x[4h+i, 4w+j, c]   (appendix A item 7)
    assert s2d[1, 2, (2 * 4 + 3) * 3 + 1] == x[4 * 1 + 2, 4 * 2 + 3, 1]
    back = og.depth_to_space(torch.from_numpy(s2d)[None], 4)[0].numpy()
    np.testing.assert_array_equal(back, x)
    y = _rand(rng, 3, 5, 32)
    d2s = og.depth_to_space(torch.from_numpy(y)[None], 4)[0].numpy()
    np.testing.assert_array_equal(d2s, naive.depth_to_space(y, 4))
    assert d2s[4 * 2 + 1, 4 * 3 + 2, 1] == y[2, 3, (1 * 4 + 2) * 2 + 1]
    # torch pixel_shuffle uses CRD order: must differ
    ps = F.pixel_shuffle(torch.from_numpy(y).permute(2, 0, 1)[None], 4)[0].permute(1, 2, 0).numpy()
    assert not np.array_equal(ps, d2s)


def test_maxpool_bn_vs_naive():
    rng = np.random.default_rng(5)
    x = _rand(rng, 6, 8, 4)
    np.testing.assert_array_equal(og.max_pool2(torch.from_numpy(x)[None])[0].numpy(),
                                  naive.max_pool2(x))
    g, b, m, v = _rand(rng, 4), _rand(rng, 4), _rand(rng, 4), np.abs(_rand(rng, 4)) + 0.5
    got = og.batch_norm(*(torch.from_numpy(a) for a in (x, g, b, m, v))).numpy()
    np.testing.assert_allclose(got, naive.batch_norm(x, g, b, m, v), rtol=1e-5, atol=1e-6)


def test_pre_post_process():
    u8 = np.arange(256, dtype=np.uint8).reshape(1, 16, 16, 1).repeat(3, -1)
    cur = og.preprocess(u8)
    assert cur.min() == -0.5 and cur.max() == 0.5
    back = og.postprocess(cur).numpy()
    # truncation: round-trip may lose at most one code, never gains
    assert ((u8.astype(int) - back.astype(int)) >= 0).all()
    assert ((u8.astype(int) - back.astype(int)) <= 1).all()
    x = torch.tensor([[-0.5, -0.4981, 0.0, 0.4999, 0.5]])
    np.testing.assert_array_equal(og.postprocess(x).numpy(), [[0, 0, 127, 254, 255]])
    packed = og.pack_bgrx(torch.from_numpy(u8))
    assert packed.shape[-1] == 4 and int(packed[..., 3].max()) == 0


def test_full_graph_tiny_vs_naive_composition():
    """One frame of the tiny model: torch oracle vs the same graph composed
    from the naive numpy primitives (fp32 semantics, BN unfused)."""
    cfg = jcfg.preset("tiny")
    w = jw.init_weights(cfg, seed=3, conditioned=True)
    frames = synthetic.frames(cfg.frame_height, cfg.frame_width, 2)
    g = og.Graph(cfg, w, "fp32")
    state = g.zero_state()
    out0, state, _ = g.step(frames[0:1], state)
    out1, state1, aux = g.step(frames[1:2], state)

    # naive composition of frame 1 given the oracle's state after frame 0
    def bn(x, name):
        return naive.batch_norm(x, w[f"{name}/gamma"], w[f"{name}/beta"],
                                w[f"{name}/moving_mean"], w[f"{name}/moving_variance"])

    relu = lambda a: np.maximum(a, 0)
    cur = frames[1, ..., :3].astype(np.float32) / np.float32(255) - np.float32(0.5)
    ph, pw, h, wd = cfg.padded_height, cfg.padded_width, cfg.frame_height, cfg.frame_width
    cur_pad = np.zeros((ph, pw, 3), np.float32)
    cur_pad[cfg.pad_top:cfg.pad_top + h, cfg.pad_left:cfg.pad_left + wd] = cur
    x = np.concatenate([cur_pad] + [s[0].numpy() for s in state["last_frames"]], -1)
    f = cfg.flow_filters
    n = len(f) // 2
    for i in range(2 * n):
        p = f"flow/block_{i + 1}"
        x = relu(bn(naive.conv2d_same(x, w[f"{p}/conv_1/kernel"]), f"{p}/bn_1"))
        x = relu(bn(naive.conv2d_same(x, w[f"{p}/conv_2/kernel"]), f"{p}/bn_2"))
        x = naive.max_pool2(x) if i < n else naive.resize_bilinear_legacy(x, 2)
    x = relu(bn(naive.conv2d_same(x, w["flow/conv_1/kernel"]), "flow/bn_1"))
    head = naive.conv2d_same(x, w["flow/conv_2/kernel"], w["flow/conv_2/bias"])
    flow = naive.depth_to_space(head, 4)
    flow = flow[cfg.pad_top * 4:cfg.pad_top * 4 + 4 * h, cfg.pad_left * 4:cfg.pad_left * 4 + 4 * wd]
    np.testing.assert_allclose(flow, aux["flow"][0].numpy(), rtol=1e-3, atol=2e-4)
    pre_warp = naive.dense_image_warp(state["pre_gen"][0].numpy(), aux["flow"][0].numpy())
    np.testing.assert_allclose(pre_warp, aux["pre_warp"][0].numpy(), rtol=0, atol=1e-6)
    x = np.concatenate([cur, naive.space_to_depth(pre_warp, 4)], -1)
    x = relu(bn(naive.conv2d_same(x, w["generator/conv_1/kernel"]), "generator/bn_1"))
    for i in range(cfg.gen_blocks):
        p = f"generator/block_{i + 1}"
        hdn = relu(bn(naive.conv2d_same(x, w[f"{p}/conv_1/kernel"]), f"{p}/bn_1"))
        x = relu(bn(naive.conv2d_same(hdn, w[f"{p}/conv_2/kernel"]), f"{p}/bn_2") + x)
    y = relu(bn(naive.conv2d_transpose_k2s2(x, w["generator/conv_trans_1/kernel"]), "generator/bn_2"))
    z = np.tanh(naive.conv2d_transpose_k2s2(y, w["generator/conv_trans_2/kernel"],
                                            w["generator/conv_trans_2/bias"]))
    out_raw = np.clip(naive.resize_bilinear_legacy(cur, 4) + z, -0.5, 0.5)
    np.testing.assert_allclose(out_raw, aux["out_raw"][0].numpy(), rtol=0, atol=2e-5)
    u8 = ((out_raw + np.float32(0.5)) * np.float32(255)).astype(np.uint8)
    diff = np.abs(u8.astype(int) - out1[0, ..., :3].numpy().astype(int))
    assert diff.max() <= 1 and (diff > 0).mean() < 0.01
    # state shift: last' = [cur_pad, last_0, ...]  (models.py:823)
    np.testing.assert_array_equal(state1["last_frames"][0][0].numpy(), cur_pad)
    np.testing.assert_array_equal(state1["last_frames"][1].numpy(), state["last_frames"][0].numpy())


def test_fp16emu_close_to_fp32():
    cfg = jcfg.preset("tiny")
    w = jw.init_weights(cfg, seed=5, conditioned=True)
    frames = synthetic.frames(cfg.frame_height, cfg.frame_width, 4)
    a, _ = og.Graph(cfg, w, "fp32").run(frames)
    b, _ = og.Graph(cfg, w, "fp16emu").run(frames)
    d = np.abs(a.astype(int) - b.astype(int))
    assert d.max() <= 2
    assert og.psnr_u8(a, b) >= 45.0


def test_weights_container_roundtrip(tmp_path):
    for name in ("tiny", "small_resnet", "psp_fast"):
        cfg = jcfg.preset(name)
        w = jw.init_weights(cfg, seed=1, conditioned=(name != "psp_fast"))
        path = str(tmp_path / f"{name}.jup")
        jw.save_model(path, cfg, w)
        cfg2, w2 = jw.load_model(path)
        assert cfg2.padded_height == cfg.padded_height and cfg2.gen_blocks == cfg.gen_blocks
        assert cfg2.flow_arch == cfg.flow_arch and list(w2) == list(w)
        for k in w:
            np.testing.assert_array_equal(w[k], w2[k])


def test_gmac_table_matches_survey():
    cfg = jcfg.preset("psp_quality")
    assert abs(cfg.flow_gmacs() - 17.90) < 0.01
    assert abs(cfg.gen_gmacs() - 234.39) < 0.01
    assert abs(cfg.gflop_per_frame() - 504.6) < 0.1


def test_brightness_normalisation_oracle_wiring():
    """normalize_brightness (models.py:772-779, 802-810): b is subtracted from the
    flow input and the stored state, added to the warped frame; the output is unchanged
    by a constant state offset only through those paths."""
    cfg = jcfg.preset("small_bright")
    w = jw.init_weights(cfg, seed=9)
    frames = synthetic.frames(cfg.frame_height, cfg.frame_width, 2)
    frames[..., :3] = np.clip(frames[..., :3].astype(int) + 60, 0, 255).astype(np.uint8)  # bright scene
    g = og.Graph(cfg, w, "fp32")
    state = g.zero_state()
    _, state, aux = g.step(frames[0:1], state)
    cur = og.preprocess(frames[0:1, ..., :3])
    b = float((cur * torch.tensor(og.BGR_LUMA) * 3).mean())
    assert b > 0.1
    np.testing.assert_allclose(state["pre_gen"].numpy(), (aux["out_raw"] - b).numpy(), atol=1e-6)
    top, left = cfg.pad_top, cfg.pad_left
    inner = state["last_frames"][0][0, top:top + cfg.frame_height, left:left + cfg.frame_width]
    np.testing.assert_allclose(inner.numpy(), (cur[0] - b).numpy(), atol=1e-6)
    assert len(state["last_frames"]) == cfg.flow_num_inputs - 1 == 2
