"""Generate golden vectors for the dense warp by EXECUTING the reference's own
vendored source, scripts/training/tfa/dense_image_warp.py, under a small numpy
shim of the TensorFlow ops it touches (TensorFlow itself is not installable
offline).  Run in the build container only (needs /root/reference):

    python tests/golden/make_warp_golden.py

Writes tests/golden/warp_golden.npz (inputs + outputs).  The shim implements
exactly: tf.shape, tf.unstack, tf.cast, tf.constant, tf.math.minimum/maximum/
floor, tf.expand_dims, tf.reshape, tf.range, tf.gather, tf.meshgrid, tf.rank, tf.stack,
tf.convert_to_tensor, tf.name_scope, tf.function, tf.control_dependencies,
tf.debugging.assert_*, tf.dtypes.int32 - all with their documented numpy
equivalents, float32 arithmetic preserved.
"""

import contextlib
import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference/scripts/training/tfa/dense_image_warp.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "warp_golden.npz")


class _Shape(tuple):
    @property
    def ndims(self):
        return len(self)


class T(np.ndarray):
    """ndarray whose .shape has .ndims, like tf.TensorShape."""

    @property
    def shape(self):  # type: ignore[override]
        return _Shape(np.ndarray.shape.__get__(self))

    @shape.setter
    def shape(self, value):
        np.ndarray.shape.__set__(self, value)


def _w(a):
    return np.asarray(a).view(T)


def _make_tf():
    tf = types.ModuleType("tensorflow")
    tf.Tensor = np.ndarray
    tf.float16, tf.float32, tf.float64 = np.float16, np.float32, np.float64
    tf.half = np.float16
    tf.dtypes = types.SimpleNamespace(int32=np.int32, float32=np.float32)
    tf.int32 = np.int32
    tf.shape = lambda x: np.array(np.asarray(x).shape, np.int32)
    tf.unstack = lambda x, axis, num=None: [_w(np.take(x, i, axis=axis))
                                            for i in range(np.asarray(x).shape[axis])]
    tf.cast = lambda x, dt: _w(np.asarray(x).astype(dt))
    tf.constant = lambda v, dtype=None: _w(np.asarray(v, dtype=dtype))
    tf.math = types.SimpleNamespace(
        minimum=lambda a, b: _w(np.minimum(a, b)),
        maximum=lambda a, b: _w(np.maximum(a, b)),
        floor=lambda a: _w(np.floor(a)))
    tf.expand_dims = lambda x, axis: _w(np.expand_dims(x, axis))
    tf.reshape = lambda x, s: _w(np.reshape(x, [int(v) for v in s]))
    tf.range = lambda n: _w(np.arange(int(n), dtype=np.int32))
    tf.gather = lambda p, idx: _w(np.asarray(p)[np.asarray(idx)])
    tf.meshgrid = lambda a, b: [_w(m) for m in np.meshgrid(np.asarray(a), np.asarray(b))]
    tf.rank = lambda x: np.asarray(x).ndim
    tf.stack = lambda xs, axis=0: _w(np.stack(xs, axis=axis))
    tf.convert_to_tensor = lambda x: _w(x)
    tf.name_scope = lambda name=None: contextlib.nullcontext()
    tf.control_dependencies = lambda deps: contextlib.nullcontext()

    def function(fn=None, **kw):
        if fn is None:
            return lambda f: f
        return fn
    tf.function = function
    tf.debugging = types.SimpleNamespace(
        assert_equal=lambda *a, **k: None,
        assert_greater_equal=lambda *a, **k: None,
        assert_rank=lambda *a, **k: None)
    tf.TensorShape = lambda x: x
    return tf


def load_reference_warp():
    sys.modules["tensorflow"] = _make_tf()
    spec = importlib.util.spec_from_file_location("ref_dense_image_warp", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    del sys.modules["tensorflow"]
    return mod


def main():
    mod = load_reference_warp()
    rng = np.random.default_rng(7)
    cases = {}
    for name, (h, w, c, mag) in {
            "small": (9, 13, 3, 2.5), "border": (6, 7, 3, 12.0),
            "subpixel": (16, 20, 3, 0.9), "integer": (8, 8, 2, 3.0)}.items():
        img = rng.standard_normal((2, h, w, c)).astype(np.float32)
        flow = (rng.standard_normal((2, h, w, 2)) * mag).astype(np.float32)
        if name == "integer":
            flow = np.round(flow)
        out = np.asarray(mod.dense_image_warp(_w(img), _w(flow)))
        assert out.dtype == np.float32 and out.shape == img.shape
        cases[f"{name}_image"] = img
        cases[f"{name}_flow"] = flow
        cases[f"{name}_out"] = out
    np.savez_compressed(OUT, **cases)
    print("wrote", OUT, {k: v.shape for k, v in cases.items()})


if __name__ == "__main__":
    main()
