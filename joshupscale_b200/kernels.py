"""Kernel-level ctypes wrappers (ju_launch_* in include/joshupscale_c.h).

Each function uploads numpy inputs, launches ONE hand-written sm_100a kernel and
downloads the result - the unit the parity tests compare against the oracle.
No CPU fallback: every call needs the built library and a CUDA device.
"""

from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from .runtime import JoshUpscaleError, _check, load_library

ACT_NONE, ACT_RELU, ACT_LRELU = 0, 1, 2
IMPL_SIMT, IMPL_TCGEN05 = 0, 1


class DeviceArray:
    """A zero-initialised device allocation with numpy shape/dtype metadata."""

    def __init__(self, shape, dtype, data: Optional[np.ndarray] = None):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self._lib = load_library()
        self._ptr = C.c_void_p()
        _check(self._lib.ju_dev_alloc(C.byref(self._ptr), max(self.nbytes, 16)))
        if data is not None:
            self.upload(data)

    @property
    def ptr(self) -> int:
        return self._ptr.value

    def upload(self, data: np.ndarray) -> "DeviceArray":
        a = np.ascontiguousarray(data, dtype=self.dtype)
        if a.nbytes != self.nbytes:
            raise ValueError(f"size mismatch: {a.shape} vs {self.shape}")
        _check(self._lib.ju_dev_upload(self._ptr, a.ctypes.data, a.nbytes))
        return self

    def download(self) -> np.ndarray:
        out = np.empty(self.shape, self.dtype)
        _check(self._lib.ju_dev_download(out.ctypes.data, self._ptr, self.nbytes))
        return out

    def free(self) -> None:
        if self._ptr.value:
            self._lib.ju_dev_free(self._ptr)
            self._ptr = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def to_device(a: np.ndarray) -> DeviceArray:
    return DeviceArray(a.shape, a.dtype, a)


def sync() -> None:
    _check(load_library().ju_dev_sync())


def pack_conv_weights(kernel: np.ndarray, scale: Optional[np.ndarray], cin_padded: int,
                      impl: int = IMPL_SIMT) -> np.ndarray:
    """kernel: Keras (kh, kw, Cin, Cout) fp32 -> packed fp16 bytes for `impl`."""
    lib = load_library()
    k = np.ascontiguousarray(kernel, np.float32)
    ks, _, cin, cout = k.shape
    s = None if scale is None else np.ascontiguousarray(scale, np.float32)
    sp = s.ctypes.data if s is not None else None
    nbytes = lib.ju_pack_conv_weights(impl, k.ctypes.data, sp, ks, cin, cin_padded, cout, None)
    if nbytes < 0:
        raise JoshUpscaleError("unsupported conv impl / shape")
    out = np.zeros(nbytes, np.uint8)
    lib.ju_pack_conv_weights(impl, k.ctypes.data, sp, ks, cin, cin_padded, cout, out.ctypes.data)
    return out


def conv(x: np.ndarray, kernel: np.ndarray, scale=None, bias=None, residual=None,
         act: int = ACT_NONE, slope: float = 0.3, out_f32: bool = False, shuffle2: bool = False,
         cin_stride: Optional[int] = None, cout_stride: Optional[int] = None,
         impl: int = IMPL_SIMT, pool: bool = False) -> np.ndarray:
    """x: [B,H,W,Cin] (any float dtype; stored as fp16, zero-padded to cin_stride).
    Returns [B,H,W,Cout] (or [B,2H,2W,Cout/4] with shuffle2) fp16/fp32."""
    lib = load_library()
    b, h, w, cin = x.shape
    ks = kernel.shape[0]
    cout = kernel.shape[3]
    cin_p = (cin + 15) // 16 * 16
    cin_stride = cin_stride or (cin + 63) // 64 * 64
    cpp = cout // 4 if shuffle2 else cout
    cout_stride = cout_stride or cpp
    xin = np.zeros((b, h, w, cin_stride), np.float16)
    xin[..., :cin] = x
    d_x = to_device(xin)
    if impl == IMPL_TCGEN05:
        cin_p = cin_stride
    d_w = to_device(pack_conv_weights(kernel, scale, cin_p, impl))
    d_b = to_device(np.asarray(bias, np.float32)) if bias is not None else None
    oh, ow = (2 * h, 2 * w) if shuffle2 else ((h // 2, w // 2) if pool else (h, w))
    d_r = None
    if residual is not None:
        r = np.zeros((b, oh, ow, cout_stride), np.float16)
        r[..., :cpp] = residual
        d_r = to_device(r)
    d_o = DeviceArray((b, oh, ow, cout_stride), np.float32 if out_f32 else np.float16)
    _check(lib.ju_launch_conv(impl, d_x.ptr, d_w.ptr, d_b.ptr if d_b else None,
                              d_r.ptr if d_r else None, d_o.ptr, b, h, w, cin_stride, cin_p, cout,
                              cout_stride, ks, act, slope, int(out_f32), 2 if pool else int(shuffle2), None))
    sync()
    return d_o.download()[..., :cpp]


def maxpool2(x: np.ndarray) -> np.ndarray:
    b, h, w, c = x.shape
    d_x = to_device(x.astype(np.float16))
    d_o = DeviceArray((b, h // 2, w // 2, c), np.float16)
    _check(load_library().ju_launch_maxpool2(d_x.ptr, d_o.ptr, b, h, w, c, None))
    sync()
    return d_o.download()


def upscale2(x: np.ndarray) -> np.ndarray:
    b, h, w, c = x.shape
    d_x = to_device(x.astype(np.float16))
    d_o = DeviceArray((b, 2 * h, 2 * w, c), np.float16)
    _check(load_library().ju_launch_upscale2(d_x.ptr, d_o.ptr, b, h, w, c, None))
    sync()
    return d_o.download()


def preprocess(frames: np.ndarray, flow_prev: np.ndarray, ph: int, pw: int, k: int) -> np.ndarray:
    """frames [B,H,W,4] u8; flow_prev [B,PH,PW,cstride] fp16 -> flow_next."""
    b, h, w, _ = frames.shape
    cs = flow_prev.shape[-1]
    d_f = to_device(frames)
    d_p = to_device(flow_prev.astype(np.float16))
    d_n = DeviceArray(flow_prev.shape, np.float16)
    _check(load_library().ju_launch_preprocess(d_f.ptr, d_p.ptr, d_n.ptr, b, h, w, ph, pw, k, cs, None))
    return d_n.download()


def warp_s2d(pre_gen: np.ndarray, flow_head: np.ndarray, frames: np.ndarray,
             want_taps: bool = False) -> Tuple[np.ndarray, Optional[np.ndarray]]:
    """pre_gen [B,4H,4W,4] fp16; flow_head [B,PH,PW,32] fp32; frames [B,H,W,4] u8.
    Returns (gen_in [B,H,W,64] fp16, taps [B,4H,4W,4] fp32 or None)."""
    b, h, w, _ = frames.shape
    _, ph, pw, _ = flow_head.shape
    d_pg = to_device(pre_gen.astype(np.float16))
    d_fh = to_device(flow_head.astype(np.float32))
    d_fr = to_device(frames)
    d_gi = DeviceArray((b, h, w, 64), np.float16)
    d_t = DeviceArray((b, 4 * h, 4 * w, 4), np.float32) if want_taps else None
    _check(load_library().ju_launch_warp_s2d(d_pg.ptr, d_fh.ptr, d_fr.ptr, d_gi.ptr,
                                             d_t.ptr if d_t else None, b, h, w, ph, pw, 64, None))
    return d_gi.download(), (d_t.download() if d_t else None)


def final(mid: np.ndarray, w2: np.ndarray, bias2: np.ndarray, frames: np.ndarray):
    """mid [B,2H,2W,32] fp16; w2 Keras (2,2,3,32); frames [B,H,W,4] u8.
    Returns (out_bgrx u8 [B,4H,4W,4], pre_gen_next fp16 [B,4H,4W,4], out_raw fp32 [B,4H,4W,3])."""
    b, h, w, _ = frames.shape
    d_m = to_device(mid.astype(np.float16))
    d_w = to_device(np.ascontiguousarray(w2, np.float32).reshape(4, 3, 32))
    d_b = to_device(np.asarray(bias2, np.float32))
    d_f = to_device(frames)
    d_o = DeviceArray((b, 4 * h, 4 * w, 4), np.uint8)
    d_s = DeviceArray((b, 4 * h, 4 * w, 4), np.float16)
    d_r = DeviceArray((b, 4 * h, 4 * w, 3), np.float32)
    _check(load_library().ju_launch_final(d_m.ptr, d_w.ptr, d_b.ptr, d_f.ptr, d_o.ptr, d_s.ptr,
                                          d_r.ptr, b, h, w, None))
    return d_o.download(), d_s.download(), d_r.download()


def tail(trunk: np.ndarray, kt1: np.ndarray, scale1: np.ndarray, bias1: np.ndarray, w2: np.ndarray,
         bias2: np.ndarray, frames: np.ndarray, act: int = ACT_RELU, slope: float = 0.3):
    """Fused generator tail.  trunk [B,H,W,64]; kt1 Keras conv_trans_1 kernel (2,2,32,64) with
    per-filter BN scale1[32] / bias1[32]; w2 Keras conv_trans_2 kernel (2,2,3,32); frames [B,H,W,4] u8.
    Returns (out_bgrx, pre_gen_next, out_raw) like final()."""
    b, h, w, _ = frames.shape
    cout, cin = kt1.shape[2], kt1.shape[3]
    k1 = np.transpose(np.ascontiguousarray(kt1, np.float32).reshape(4, cout, cin), (2, 0, 1)).reshape(1, 1, cin, 4 * cout)
    d_t = to_device(trunk.astype(np.float16))
    d_w1 = to_device(pack_conv_weights(k1, np.tile(np.asarray(scale1, np.float32), 4), 64, IMPL_TCGEN05))
    d_b1 = to_device(np.tile(np.asarray(bias1, np.float32), 4))
    d_w2 = to_device(np.ascontiguousarray(w2, np.float32).reshape(4, 3, 32))
    d_b2 = to_device(np.asarray(bias2, np.float32))
    d_f = to_device(frames)
    d_o = DeviceArray((b, 4 * h, 4 * w, 4), np.uint8)
    d_s = DeviceArray((b, 4 * h, 4 * w, 4), np.float16)
    d_r = DeviceArray((b, 4 * h, 4 * w, 3), np.float32)
    _check(load_library().ju_launch_tail(d_t.ptr, d_w1.ptr, d_b1.ptr, d_w2.ptr, d_b2.ptr, d_f.ptr, d_o.ptr,
                                         d_s.ptr, d_r.ptr, b, h, w, act, slope, None))
    return d_o.download(), d_s.download(), d_r.download()
