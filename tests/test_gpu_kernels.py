"""Per-kernel parity: each hand-written CUDA kernel, driven through the C-ABI,
against the CPU oracle on the same seeded inputs.  Bit-exact for index / byte
/ layout work, fp32-accumulation tolerance for the convolutions."""

import numpy as np
import pytest
import torch

from joshupscale_b200 import kernels as jk
from joshupscale_b200 import runtime as jrt
from oracle import reference_graph as og
from tests.gpu_util import r16, require_gpu

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, scope="module")
def _gpu():
    require_gpu()


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, np.float32))


@pytest.mark.parametrize("h,w,ph,pw,k", [(21, 27, 24, 32, 4), (270, 480, 272, 480, 4), (16, 16, 16, 16, 2)])
def test_preprocess_bit_exact(h, w, ph, pw, k):
    rng = np.random.default_rng(0)
    frames = rng.integers(0, 256, (2, h, w, 4), dtype=np.uint8)
    prev = np.zeros((2, ph, pw, 64), np.float16)
    prev[..., :3 * k] = rng.standard_normal((2, ph, pw, 3 * k)).astype(np.float16)
    got = jk.preprocess(frames, prev, ph, pw, k)
    cur = og.preprocess(frames[..., :3]).numpy()
    want = np.zeros_like(prev)
    top, left = (ph - h) // 2, (pw - w) // 2
    want[:, top:top + h, left:left + w, 0:3] = cur.astype(np.float16)
    want[..., 3:3 * k] = prev[..., 0:3 * k - 3]
    np.testing.assert_array_equal(got.view(np.uint16), want.view(np.uint16))


CONV_CASES = [
    # b, h, w, cin, cout, ksize
    (1, 9, 13, 12, 32, 3), (2, 16, 24, 64, 64, 3), (1, 7, 70, 51, 64, 3),
    (1, 12, 20, 128, 256, 3), (1, 8, 8, 256, 128, 3), (1, 10, 11, 64, 32, 1), (1, 5, 130, 32, 32, 3),
]


@pytest.mark.parametrize("b,h,w,cin,cout,ks", CONV_CASES)
@pytest.mark.parametrize("mode", ["plain", "bias_relu", "residual_relu", "lrelu_f32"])
def test_conv_simt_vs_oracle(b, h, w, cin, cout, ks, mode):
    rng = np.random.default_rng(cin * 1000 + cout + ks)
    x = r16(rng.standard_normal((b, h, w, cin)) * 0.5)
    k = (rng.standard_normal((ks, ks, cin, cout)) / np.sqrt(ks * ks * cin)).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    bias = rng.standard_normal(cout).astype(np.float32) * 0.1
    res = r16(rng.standard_normal((b, h, w, cout)) * 0.5)
    kw = dict(scale=scale)
    want = og.conv2d_same(_t(x), og.r16(_t(k * scale))).numpy()
    if mode == "bias_relu":
        kw.update(bias=bias, act=jk.ACT_RELU)
        want = np.maximum(want + bias, 0)
    elif mode == "residual_relu":
        kw.update(bias=bias, residual=res, act=jk.ACT_RELU)
        want = np.maximum(want + bias + res, 0)
    elif mode == "lrelu_f32":
        kw.update(bias=bias, act=jk.ACT_LRELU, slope=0.3, out_f32=True)
        want = want + bias
        want = np.where(want >= 0, want, want * np.float32(0.3))
    got = jk.conv(x, k, **kw).astype(np.float32)
    if mode != "lrelu_f32":
        want = r16(want)
        # fp32 accumulation order differs: allow one fp16 ulp
        np.testing.assert_allclose(got, want, rtol=2e-3, atol=1e-3)
    else:
        np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-4)


def test_conv_transpose_as_shuffled_1x1():
    """Conv2DTranspose(k2,s2) == 1x1 conv to 4*Cout + pixel shuffle (models.py:559-572)."""
    rng = np.random.default_rng(5)
    b, h, w, cin, cout = 1, 9, 14, 64, 32
    x = r16(rng.standard_normal((b, h, w, cin)) * 0.5)
    kt = (rng.standard_normal((2, 2, cout, cin)) / 8).astype(np.float32)  # keras (kh,kw,Cout,Cin)
    scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    bias = rng.standard_normal(cout).astype(np.float32) * 0.1
    # -> (1,1,Cin,4*Cout) with channel (i*2+j)*Cout + o
    k1 = np.transpose(kt.reshape(4, cout, cin), (2, 0, 1)).reshape(1, 1, cin, 4 * cout)
    got = jk.conv(x, k1, scale=np.tile(scale, 4), bias=np.tile(bias, 4), act=jk.ACT_RELU,
                  shuffle2=True).astype(np.float32)
    want = og.conv2d_transpose_k2s2(_t(x), og.r16(_t(kt * scale[None, None, :, None]))).numpy() + bias
    want = r16(np.maximum(want, 0))
    assert got.shape == (b, 2 * h, 2 * w, cout)
    np.testing.assert_allclose(got, want, rtol=2e-3, atol=1e-3)


def test_maxpool_and_upscale_bit_exact():
    rng = np.random.default_rng(6)
    x = rng.standard_normal((2, 12, 20, 64)).astype(np.float16)
    got = jk.maxpool2(x)
    want = og.max_pool2(_t(x.astype(np.float32))).numpy().astype(np.float16)
    np.testing.assert_array_equal(got.view(np.uint16), want.view(np.uint16))
    got = jk.upscale2(x)
    want = og.resize_bilinear_legacy(_t(x.astype(np.float32)), 2).numpy().astype(np.float16)
    np.testing.assert_array_equal(got.view(np.uint16), want.view(np.uint16))


@pytest.mark.parametrize("h,w,ph,pw,mag", [(21, 27, 24, 32, 3.0), (8, 16, 8, 16, 40.0), (270, 480, 272, 480, 2.0)])
def test_warp_s2d_exact(h, w, ph, pw, mag):
    """warp/indexing exact: floors and alphas bit-identical to the fp32 oracle,
    warped values bit-identical after fp16 rounding, S2D layout exact."""
    rng = np.random.default_rng(7)
    b = 2
    pre_gen = np.zeros((b, 4 * h, 4 * w, 4), np.float16)
    pre_gen[..., :3] = rng.uniform(-0.5, 0.5, (b, 4 * h, 4 * w, 3)).astype(np.float16)
    head = (rng.standard_normal((b, ph, pw, 32)) * mag).astype(np.float32)
    frames = rng.integers(0, 256, (b, h, w, 4), dtype=np.uint8)
    gen_in, taps = jk.warp_s2d(pre_gen, head, frames, want_taps=True)

    flow = og.depth_to_space(_t(head), 4)
    oy, ox = ((ph - h) // 2) * 4, ((pw - w) // 2) * 4
    flow = flow[:, oy:oy + 4 * h, ox:ox + 4 * w].contiguous()
    want_taps = og.warp_taps(flow, 4 * h, 4 * w)
    np.testing.assert_array_equal(taps[..., 0], want_taps.fy.numpy().astype(np.float32))
    np.testing.assert_array_equal(taps[..., 1], want_taps.fx.numpy().astype(np.float32))
    np.testing.assert_array_equal(taps[..., 2].view(np.uint32), want_taps.ay.numpy().view(np.uint32))
    np.testing.assert_array_equal(taps[..., 3].view(np.uint32), want_taps.ax.numpy().view(np.uint32))

    warped = og.dense_image_warp(_t(pre_gen[..., :3].astype(np.float32)), flow)
    want = np.zeros((b, h, w, 64), np.float16)
    want[..., 0:3] = og.preprocess(frames[..., :3]).numpy().astype(np.float16)
    want[..., 3:51] = og.space_to_depth(warped, 4).numpy().astype(np.float16)
    np.testing.assert_array_equal(gen_in.view(np.uint16), want.view(np.uint16))


@pytest.mark.parametrize("h,w", [(5, 7), (21, 27), (270, 480)])
def test_final_epilogue(h, w):
    rng = np.random.default_rng(8)
    b = 1
    mid = (rng.standard_normal((b, 2 * h, 2 * w, 32)) * 0.4).astype(np.float16)
    w2 = r16(rng.standard_normal((2, 2, 3, 32)) * 0.3)
    b2 = (rng.standard_normal(3) * 0.05).astype(np.float32)
    frames = rng.integers(0, 256, (b, h, w, 4), dtype=np.uint8)
    out, state, raw = jk.final(mid, w2, b2, frames)
    cur = og.preprocess(frames[..., :3])
    z = torch.tanh(og.conv2d_transpose_k2s2(_t(mid.astype(np.float32)), _t(w2), _t(b2)))
    want_raw = torch.clamp(og.resize_bilinear_legacy(cur, 4) + z, -0.5, 0.5).numpy()
    np.testing.assert_allclose(raw, want_raw, rtol=0, atol=3e-6)
    # u8 = trunc((x+0.5)*255) of the kernel's OWN fp32 value: exact, X byte 0
    own = ((raw + np.float32(0.5)) * np.float32(255)).astype(np.uint8)
    np.testing.assert_array_equal(out[..., :3], own)
    assert int(out[..., 3].max()) == 0
    want_u8 = og.postprocess(torch.from_numpy(want_raw)).numpy()
    d = np.abs(out[..., :3].astype(int) - want_u8.astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 2e-3
    np.testing.assert_array_equal(state[..., :3].view(np.uint16), raw.astype(np.float16).view(np.uint16))
    assert not state[..., 3].any()


def test_u8_conversion_matches_ieee_division_for_all_bytes():
    """The kernels convert pixels with a division-free sequence (pixel_common.cuh); it must equal
    the IEEE fp32 division of the reference formula for every byte value, bit for bit."""
    import ctypes as C
    lib = jrt.load_library()
    fast = (C.c_float * 256)()
    ieee = (C.c_float * 256)()
    jrt._check(lib.ju_u8_conversion_table(fast, ieee))
    fast, ieee = np.frombuffer(fast, np.float32), np.frombuffer(ieee, np.float32)
    want = np.arange(256, dtype=np.float32) / np.float32(255) - np.float32(0.5)
    np.testing.assert_array_equal(ieee.view(np.uint32), want.view(np.uint32))
    np.testing.assert_array_equal(fast.view(np.uint32), want.view(np.uint32))
