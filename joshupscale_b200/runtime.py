"""ctypes binding over the C-ABI of libJoshUpscale (include/joshupscale_c.h).

Host-side mirror of the reference's Python drivers: `Session` has the same
shape as scripts/inference/onnx/inference.py:46-94 (Session(model).run(image)
-> upscaled BGR uint8, recurrent state kept inside), so scripts written against
the onnxruntime / TensorRT drivers can switch to the B200 engine.

There is no CPU fallback: importing works anywhere, but creating a Runtime or
launching a kernel without the built library or without a CUDA device raises.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence

import numpy as np

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib",
                         "libJoshUpscale.so")
_lib: Optional[C.CDLL] = None

LOC_CPU, LOC_CUDA, LOC_GRAPHICS_RESOURCE = 0, 1, 2


class JuImage(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("location", C.c_int32), ("stride", C.c_int64),
                ("width", C.c_uint64), ("height", C.c_uint64)]


class JuInfo(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "input_width", "input_height", "output_width", "output_height",
        "padded_width", "padded_height", "batch", "flow_num_inputs", "gen_filters",
        "gen_blocks", "flow_arch", "conv_impl", "kernels_per_frame", "device")] + [
        ("gflop_per_frame", C.c_double)]


class JuTensorDesc(C.Structure):
    _fields_ = [("dtype", C.c_uint32), ("ndim", C.c_uint32), ("dims", C.c_uint64 * 4),
                ("bytes", C.c_uint64)]


class JuOpTime(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("usec", C.c_double), ("flops", C.c_double),
                ("bytes", C.c_double), ("tensor_bound", C.c_int32), ("reserved", C.c_int32)]


class JoshUpscaleError(RuntimeError):
    pass


# every symbol include/joshupscale_c.h declares: (name, restype, argtypes)
_VP, _I, _U64, _F = C.c_void_p, C.c_int, C.c_uint64, C.c_float
SYMBOLS = {
    "ju_create": (_I, [C.c_char_p, _I, _I, C.POINTER(_VP)]),
    "ju_destroy": (None, [_VP]),
    "ju_process": (_I, [_VP, C.POINTER(JuImage), C.POINTER(JuImage)]),
    "ju_process_batch": (_I, [_VP, _I, C.POINTER(JuImage), C.POINTER(JuImage)]),
    "ju_get_info": (_I, [_VP, C.POINTER(JuInfo)]),
    "ju_last_error": (C.c_char_p, []),
    "ju_set_log_sink": (None, [_VP, _VP]),
    "ju_reset_state": (_I, [_VP]),
    "ju_debug_inject_stall": (_I, [_VP, _I]),
    "ju_read_tensor": (_I, [_VP, C.c_char_p, _VP, _U64, C.POINTER(JuTensorDesc)]),
    "ju_write_state": (_I, [_VP, C.c_char_p, _VP, _U64]),
    "ju_profile_ops": (_I, [_VP, _I, C.POINTER(JuOpTime), _I, C.POINTER(_I)]),
    "ju_device_count": (_I, []),
    "ju_set_device": (_I, [_I]),
    "ju_version": (C.c_char_p, []),
    "ju_launch_preprocess": (_I, [_VP, _VP, _VP, _I, _I, _I, _I, _I, _I, _I, _VP]),
    "ju_launch_conv": (_I, [_I, _VP, _VP, _VP, _VP, _VP, _I, _I, _I, _I, _I, _I, _I, _I, _I,
                            _F, _I, _I, _VP]),
    "ju_pack_conv_weights": (C.c_int64, [_I, _VP, _VP, _I, _I, _I, _I, _VP]),
    "ju_set_option": (_I, [C.c_char_p, _I]),
    "ju_bench_conv": (_I, [_I, _I, _I, _I, _I, _I, _I, _I, _I, C.POINTER(C.c_double)]),
    "ju_launch_maxpool2": (_I, [_VP, _VP, _I, _I, _I, _I, _VP]),
    "ju_launch_upscale2": (_I, [_VP, _VP, _I, _I, _I, _I, _VP]),
    "ju_launch_warp_s2d": (_I, [_VP, _VP, _VP, _VP, _VP, _I, _I, _I, _I, _I, _I, _VP]),
    "ju_launch_final": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _I, _I, _I, _VP]),
    "ju_launch_tail": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _I, _I, _I, _I, _F, _VP]),
    "ju_dev_alloc": (_I, [C.POINTER(_VP), _U64]),
    "ju_dev_free": (_I, [_VP]),
    "ju_dev_upload": (_I, [_VP, _VP, _U64]),
    "ju_dev_download": (_I, [_VP, _VP, _U64]),
    "ju_dev_memset": (_I, [_VP, _I, _U64]),
    "ju_dev_sync": (_I, []),
    "ju_host_alloc": (_I, [C.POINTER(_VP), _U64]),
    "ju_host_free": (_I, [_VP]),
    "ju_l2_flush": (_I, []),
    "ju_timer_begin": (_I, []),
    "ju_u8_conversion_table": (_I, [C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "ju_timer_end": (_I, [C.POINTER(C.c_double)]),
}


def library_path() -> str:
    return _LIB_PATH


def load_library() -> C.CDLL:
    """Load libJoshUpscale.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise JoshUpscaleError(
            f"{_LIB_PATH} is missing - build it with `python -m joshupscale_b200.build`; "
            "there is no CPU fallback")
    lib = C.CDLL(_LIB_PATH)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def _check(status: int) -> None:
    if status != 0:
        raise JoshUpscaleError(load_library().ju_last_error().decode(errors="replace"))


def set_option(key: str, value: int) -> None:
    _check(load_library().ju_set_option(key.encode(), int(value)))


def bench_conv(impl: int, batch: int, h: int, w: int, cin: int, cout: int, ksize: int = 3,
               residual: bool = False, iters: int = 20) -> float:
    """Mean device time (usec) of one convolution launch on synthetic buffers."""
    usec = C.c_double()
    _check(load_library().ju_bench_conv(impl, batch, h, w, cin, cout, ksize, int(residual), iters,
                                        C.byref(usec)))
    return usec.value


def device_count() -> int:
    return load_library().ju_device_count()


def _image(arr_or_ptr, width: int, height: int, stride: Optional[int] = None,
           location: int = LOC_CPU) -> JuImage:
    if isinstance(arr_or_ptr, np.ndarray):
        a = arr_or_ptr
        if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 4 or a.strides[2] != 1 \
                or a.strides[1] != 4:
            raise ValueError("expected a uint8 [H, W, 4] BGRX array with packed pixels")
        return JuImage(a.ctypes.data, LOC_CPU, a.strides[0], a.shape[1], a.shape[0])
    return JuImage(int(arr_or_ptr), location, stride if stride is not None else width * 4,
                   width, height)


_DTYPES = {0: np.float32, 1: np.float16, 2: np.uint8}


class Runtime:
    """One engine = one device + one CUDA stream + `batch` recurrent states.

    Mirrors core::Runtime (core/public/JoshUpscale/core.h:64-92): process()
    is synchronous, calls must be serialised per instance, fresh state is zero.
    """

    def __init__(self, model_path: str, device: int = 0, batch: int = 1):
        self._lib = load_library()
        self._h = C.c_void_p()
        _check(self._lib.ju_create(os.fsencode(model_path), device, batch, C.byref(self._h)))
        info = JuInfo()
        _check(self._lib.ju_get_info(self._h, C.byref(info)))
        self.info = info
        self.batch = info.batch
        self.in_shape = (info.input_height, info.input_width, 4)
        self.out_shape = (info.output_height, info.output_width, 4)

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            self._lib.ju_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- frames ---------------------------------------------------------
    def process(self, frame: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
        """BGRX uint8 [H,W,4] -> BGRX uint8 [4H,4W,4] (X written as 0)."""
        if out is None:
            out = np.empty(self.out_shape, np.uint8)
        i, o = _image(frame, 0, 0), _image(out, 0, 0)
        _check(self._lib.ju_process(self._h, C.byref(i), C.byref(o)))
        return out

    def process_batch(self, frames: Sequence[np.ndarray],
                      outs: Optional[Sequence[np.ndarray]] = None) -> List[np.ndarray]:
        n = len(frames)
        if outs is None:
            outs = [np.empty(self.out_shape, np.uint8) for _ in range(n)]
        ins = (JuImage * n)(*[_image(f, 0, 0) for f in frames])
        os_ = (JuImage * n)(*[_image(o, 0, 0) for o in outs])
        _check(self._lib.ju_process_batch(self._h, n, ins, os_))
        return list(outs)

    def process_images(self, inputs: Sequence[JuImage], outputs: Sequence[JuImage]) -> None:
        """Raw ju_image entry (device pointers, custom strides)."""
        n = len(inputs)
        ins = (JuImage * n)(*inputs)
        os_ = (JuImage * n)(*outputs)
        _check(self._lib.ju_process_batch(self._h, n, ins, os_))

    # ---- state / debug --------------------------------------------------
    def reset_state(self) -> None:
        _check(self._lib.ju_reset_state(self._h))

    def inject_stall(self, kernel_id: int = 1) -> None:
        """Test hook: the next frame's kernel `kernel_id` (1 = persistent trunk) stalls."""
        _check(self._lib.ju_debug_inject_stall(self._h, kernel_id))

    def read_tensor(self, name: str) -> np.ndarray:
        desc = JuTensorDesc()
        _check(self._lib.ju_read_tensor(self._h, name.encode(), None, 0, C.byref(desc)))
        arr = np.empty(tuple(desc.dims[i] for i in range(desc.ndim)), _DTYPES[desc.dtype])
        assert arr.nbytes == desc.bytes
        _check(self._lib.ju_read_tensor(self._h, name.encode(), arr.ctypes.data, arr.nbytes,
                                        C.byref(desc)))
        return arr

    def write_state(self, name: str, value: np.ndarray) -> None:
        v = np.ascontiguousarray(value)
        _check(self._lib.ju_write_state(self._h, name.encode(), v.ctypes.data, v.nbytes))

    def profile_ops(self, iters: int = 10) -> List[Dict[str, object]]:
        count = C.c_int(0)
        _check(self._lib.ju_profile_ops(self._h, iters, None, 0, C.byref(count)))
        ops = (JuOpTime * count.value)()
        _check(self._lib.ju_profile_ops(self._h, iters, ops, count.value, C.byref(count)))
        return [dict(name=o.name.decode(), usec=o.usec, flops=o.flops, bytes=o.bytes,
                     tensor_bound=bool(o.tensor_bound), launches=o.reserved) for o in ops]


class Session:
    """Drop-in for the reference's inference `Session`
    (scripts/inference/onnx/inference.py:46-94;
    scripts/inference/tensorrt/inference.py:60-193): run(image) takes a BGR
    uint8 [H,W,3] (cv2.imread layout) or BGRX [H,W,4] frame and returns the
    upscaled BGR uint8 [4H,4W,3]; recurrent state lives inside, zero-initialised.
    """

    def __init__(self, model: str, device: int = 0) -> None:
        self.runtime = Runtime(model, device, 1)
        h, w, _ = self.runtime.in_shape
        self._in = np.zeros((h, w, 4), np.uint8)

    def run(self, image: np.ndarray) -> np.ndarray:
        if image.ndim == 4:
            image = image[0]
        if image.shape[:2] != self._in.shape[:2]:
            raise ValueError(f"expected {self._in.shape[:2]} frame, got {image.shape[:2]}")
        self._in[..., :image.shape[2]] = image
        out = self.runtime.process(self._in)
        return out[..., :3]
