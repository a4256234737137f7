"""The reference's own image <-> tensor conversion (core/src/cuda_convert.cc.cu, compiled
unmodified into oracle/_ref/libref_convert.so by oracle/Makefile) run on the GPU next to this
repo's pixel kernels and the oracle's packing: channel order, the ignored / zeroed X byte, padded
strides and the truncating float -> u8 cast come from the reference binary itself."""

import ctypes as C
import os

import numpy as np
import pytest
import torch

from joshupscale_b200 import kernels as jk
from oracle import reference_graph as og
from tests.gpu_util import require_gpu

pytestmark = pytest.mark.gpu

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref_convert.so")


@pytest.fixture(scope="module")
def ref():
    require_gpu()
    if not os.path.exists(LIB):
        pytest.skip("oracle/_ref/libref_convert.so not built (needs /root/reference at build time)")
    lib = C.CDLL(LIB)
    lib.ref_image_to_tensor.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.ref_tensor_to_image.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_longlong]
    return lib


# the reference asserts width * height % 32 == 0 (cuda_convert.cc.cu:186)
@pytest.mark.parametrize("h,w,pad", [(16, 20, 0), (48, 72, 20), (270, 480, 64)])
def test_input_conversion_and_preprocess(ref, h, w, pad):
    rng = np.random.default_rng(h)
    stride = w * 4 + pad
    buf = rng.integers(0, 256, (h, stride), dtype=np.uint8)          # X bytes and row padding are noise
    image = buf[:, :w * 4].reshape(h, w, 4)
    out32 = np.empty((h, w, 3), np.float32)
    assert ref.ref_image_to_tensor(buf.ctypes.data, stride, w, h, 0, out32.ctypes.data) == 0
    # the engine input is float(B), float(G), float(R): X dropped, no scaling (the model's
    # PreprocessLayer does x / 255 - 0.5 afterwards)
    np.testing.assert_array_equal(out32, image[..., :3].astype(np.float32))
    out16 = np.empty((h, w, 3), np.float16)
    assert ref.ref_image_to_tensor(buf.ctypes.data, stride, w, h, 1, out16.ctypes.data) == 0
    np.testing.assert_array_equal(out16, image[..., :3].astype(np.float16))
    # this repo fuses conversion + PreprocessLayer + padding into one kernel
    ph, pw = (h + 7) // 8 * 8, (w + 7) // 8 * 8
    got = jk.preprocess(np.ascontiguousarray(image)[None], np.zeros((1, ph, pw, 64), np.float16), ph, pw, 4)
    want = (out32 / np.float32(255) - np.float32(0.5)).astype(np.float16)
    top, left = (ph - h) // 2, (pw - w) // 2
    np.testing.assert_array_equal(got[0, top:top + h, left:left + w, :3].view(np.uint16), want.view(np.uint16))


@pytest.mark.parametrize("h,w,pad", [(16, 20, 0), (84, 112, 36)])
def test_output_conversion_truncates_and_zeroes_x(ref, h, w, pad):
    rng = np.random.default_rng(w)
    x = rng.uniform(-0.5, 0.5, (h, w, 3)).astype(np.float32)
    x.flat[:6] = [-0.5, 0.5, 0.0, 0.49999997, -0.49999997, 0.25]
    v = ((x + np.float32(0.5)) * np.float32(255)).astype(np.float32)  # PostprocessLayer before its cast
    stride = w * 4 + pad
    buf = np.full((h, stride), 0xEE, np.uint8)
    assert ref.ref_tensor_to_image(v.ctypes.data, w, h, buf.ctypes.data, stride) == 0
    image = buf[:, :w * 4].reshape(h, w, 4)
    np.testing.assert_array_equal(image[..., :3], np.trunc(v).astype(np.uint8))
    assert not image[..., 3].any()                      # X is written as 0
    assert (buf[:, w * 4:] == 0xEE).all()               # row padding untouched
    # the oracle's postprocess + pack (which the CUDA tail kernel is checked against) agree
    want = og.pack_bgrx(og.postprocess(torch.from_numpy(x)[None]))[0].numpy()
    np.testing.assert_array_equal(image, want)
