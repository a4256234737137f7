"""Golden vectors for the WIRING of the inference graph, produced by EXECUTING the
reference's own model-building code under a small Keras/TensorFlow shim.

TensorFlow / Keras are not installable offline, so the reference model cannot run
as-is.  But everything that is specific to JoshUpscale - which layers exist, in
which order, with which filters / strides / padding, how frames are padded, what
feeds the flow net, where brightness is subtracted and added back, how the warp
output enters the generator, which tensors become the recurrent state - is
plain Python in

    /root/reference/scripts/training/models.py        (get_flow_autoencoder,
        get_flow_resnet, get_generator_resnet, get_inference_model, res_block)
    /root/reference/scripts/training/keras_layers.py  (custom layers)
    /root/reference/scripts/training/tfa/dense_image_warp.py

This script imports THOSE files unmodified.  `tensorflow`, `tensorflow.keras`
(+ `keras_models`, `utils`, which only matter for training) are replaced by the
shim below: a symbolic tensor class that records the graph the reference code
builds, and numpy implementations of the generic primitives it instantiates
(Conv2D, BatchNormalization, ReLU, MaxPool2D, ... = oracle/naive.py, i.e. the
published TF/Keras semantics of SURVEY appendix A).  The graph is then evaluated
frame by frame exactly like scripts/inference/onnx/inference.py:55-94 drives the
exported model (zero initial state, outputs fed back).

What this pins: the dataflow of the oracle (oracle/reference_graph.py) and of the
CUDA engine against the reference source itself.  What stays unpinned: the
arithmetic inside the generic TF/Keras primitives (no TensorFlow to execute).

    python tests/golden/make_graph_golden.py      # build container only
Writes tests/golden/graph_golden.npz.
"""

import ast
import contextlib
import importlib.util
import os
import re
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_DIR = "/root/reference/scripts/training"
OUT = os.path.join(HERE, "graph_golden.npz")
sys.path.insert(0, ROOT)

from oracle import naive  # noqa: E402
from joshupscale_b200 import config as jcfg  # noqa: E402
from joshupscale_b200 import synthetic  # noqa: E402
from joshupscale_b200 import weights as jw  # noqa: E402

# ---------------------------------------------------------------------------
# symbolic tensors
# ---------------------------------------------------------------------------

_SCOPE = []      # names of the models being evaluated (innermost last)
_WEIGHTS = {}    # "<model>/<layer>/<variable>" -> array, set per run


class Sym:
    """A node of the graph the reference code builds: fn(*parent values)."""

    def __init__(self, fn, parents, name=None, dtype="float32"):
        self.fn, self.parents, self.name, self.dtype = fn, list(parents), name, dtype

    def _bin(self, other, op, swap=False):
        if isinstance(other, Sym):
            return Sym((lambda a, b: op(b, a)) if swap else op, [self, other])
        return Sym((lambda a: op(other, a)) if swap else (lambda a: op(a, other)), [self])

    def __add__(self, o): return self._bin(o, lambda a, b: a + b)
    def __radd__(self, o): return self._bin(o, lambda a, b: a + b, True)
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._bin(o, lambda a, b: a - b, True)
    def __mul__(self, o): return self._bin(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._bin(o, lambda a, b: a * b, True)
    def __truediv__(self, o): return self._bin(o, lambda a, b: a / b)
    def __neg__(self): return Sym(lambda a: -a, [self])


def _flatten(x):
    if isinstance(x, dict):
        return [v for k in x for v in _flatten(x[k])]
    if isinstance(x, (list, tuple)):
        return [v for e in x for v in _flatten(e)]
    return [x]


def _unflatten(like, values):
    it = iter(values)

    def rec(x):
        if isinstance(x, dict):
            return {k: rec(v) for k, v in x.items()}
        if isinstance(x, (list, tuple)):
            return [rec(e) for e in x]
        return next(it)
    return rec(like)


def _evaluate(outputs, env):
    memo = dict(env)

    def ev(s):
        if not isinstance(s, Sym):
            return s
        if id(s) not in memo:
            memo[id(s)] = s.fn(*[ev(p) for p in s.parents])
        return memo[id(s)]
    return _unflatten(outputs, [ev(s) for s in _flatten(outputs)])


# ---------------------------------------------------------------------------
# keras shim
# ---------------------------------------------------------------------------

def _f32(x):
    return np.asarray(x, dtype=np.float32)


def _per_image(fn, x, *args):
    """oracle/naive.py works on one [H, W, C] image; the graph carries a batch axis."""
    x = _f32(x)
    return np.stack([fn(x[n], *args) for n in range(x.shape[0])]).astype(np.float32)


def _var(layer, var):
    # the reference scopes block layers with an underscore (models.py get_scoped_name:
    # "block_3_conv_1"); the container writes the same layer as "block_3/conv_1"
    lname = re.sub(r"^(block_\d+)_", r"\1/", layer.name)
    key = f"{_SCOPE[-1]}/{lname}/{var}"
    if key not in _WEIGHTS:
        raise KeyError(f"the reference graph asks for a variable the container does not have: {key}")
    return _f32(_WEIGHTS[key])


class Layer:
    def __init__(self, name=None, dtype=None, trainable=True, **kwargs):
        self.name, self.dtype = name, dtype

    def __call__(self, inputs, *args, **kwargs):
        flat = _flatten(inputs)
        out = Sym(lambda *vals: self.call(_unflatten(inputs, vals)), flat, name=self.name)
        return out

    def get_config(self):
        return {}


class Conv2D(Layer):
    def __init__(self, filters, kernel_size, strides=1, padding="valid", use_bias=True, **kw):
        kw.pop("kernel_regularizer", None)
        kw.pop("kernel_initializer", None)
        super().__init__(**kw)
        ks = kernel_size if isinstance(kernel_size, int) else kernel_size[0]
        st = strides if isinstance(strides, int) else strides[0]
        assert st == 1 and padding.upper() == "SAME" and ks in (1, 3), (ks, st, padding)
        self.filters, self.use_bias = filters, use_bias

    def call(self, x):
        k = _var(self, "kernel")
        assert k.shape[3] == self.filters and k.shape[2] == x.shape[-1], (self.name, k.shape, x.shape)
        return _per_image(naive.conv2d_same, x, k, _var(self, "bias") if self.use_bias else None)


class Conv2DTranspose(Layer):
    def __init__(self, filters, kernel_size, strides=1, padding="valid", use_bias=True, **kw):
        kw.pop("kernel_regularizer", None)
        super().__init__(**kw)
        assert kernel_size in (2, (2, 2)) and strides in (2, (2, 2)), (kernel_size, strides)
        self.filters, self.use_bias = filters, use_bias

    def call(self, x):
        k = _var(self, "kernel")
        assert k.shape[2] == self.filters
        return _per_image(naive.conv2d_transpose_k2s2, x, k, _var(self, "bias") if self.use_bias else None)


class BatchNormalization(Layer):
    def call(self, x):
        return naive.batch_norm(_f32(x), _var(self, "gamma"), _var(self, "beta"), _var(self, "moving_mean"),
                                _var(self, "moving_variance"))


class ReLU(Layer):
    def call(self, x):
        return np.maximum(_f32(x), np.float32(0))


class LeakyReLU(Layer):
    def __init__(self, negative_slope=0.3, **kw):
        super().__init__(**kw)
        self.slope = np.float32(negative_slope)

    def call(self, x):
        x = _f32(x)
        return np.where(x >= 0, x, x * self.slope).astype(np.float32)


class Activation(Layer):
    def __init__(self, activation, **kw):
        super().__init__(**kw)
        self.activation = activation

    def call(self, x):
        return self.activation(x)


class MaxPool2D(Layer):
    def __init__(self, pool_size=2, **kw):
        super().__init__(**kw)
        assert pool_size in (2, (2, 2))

    def call(self, x):
        return _per_image(naive.max_pool2, x)


class ZeroPadding2D(Layer):
    def __init__(self, padding, **kw):
        super().__init__(**kw)
        self.padding = padding

    def call(self, x):
        (t, b), (l, r) = self.padding
        return np.pad(_f32(x), ((0, 0), (t, b), (l, r), (0, 0)))


class Concatenate(Layer):
    def __init__(self, axis=-1, **kw):
        super().__init__(**kw)
        self.axis = axis

    def call(self, xs):
        return np.concatenate([_f32(x) for x in xs], axis=self.axis)


class Add(Layer):
    def call(self, xs):
        out = _f32(xs[0])
        for x in xs[1:]:
            out = out + _f32(x)
        return out


class Lambda(Layer):
    def __init__(self, function, **kw):
        super().__init__(**kw)
        self.function = function

    def call(self, x):
        return self.function(x)


class Identity(Layer):
    def call(self, x):
        return np.asarray(x).astype(self.dtype) if self.dtype else x


class Model(Layer):
    def __init__(self, inputs, outputs, name=None, **kw):
        super().__init__(name=name)
        self.inputs = list(inputs) if isinstance(inputs, (list, tuple)) else [inputs]
        self.outputs = outputs

    def call(self, values):
        values = values if isinstance(values, (list, tuple)) else [values]
        assert len(values) == len(self.inputs)
        env = {}
        for s, v in zip(self.inputs, values):
            v = np.asarray(v)
            # Keras casts inputs to the layer's compute dtype before the first op
            env[id(s)] = v if s.dtype == "uint8" else v.astype(np.float32)
        _SCOPE.append(self.name)
        try:
            return _evaluate(self.outputs, env)
        finally:
            _SCOPE.pop()

    def __call__(self, inputs, *a, **k):
        flat = _flatten(inputs)
        whole = Sym(lambda *vals: self.call(list(vals)), flat, name=self.name)
        outs = _flatten(self.outputs)
        if len(outs) == 1 and not isinstance(self.outputs, (list, tuple, dict)):
            return whole
        picks = [Sym(lambda res, i=i: _flatten(res)[i], [whole]) for i in range(len(outs))]
        return _unflatten(self.outputs, picks)


def _input(shape=None, name=None, dtype="float32", **kw):
    s = Sym(None, [], name=name, dtype=dtype)
    s.shape = (None,) + tuple(shape)
    return s


def _make_ops():
    ops = types.ModuleType("tensorflow.keras.ops")
    ops.tanh = lambda x: np.tanh(_f32(x)).astype(np.float32)
    ops.expand_dims = lambda x, axis: np.expand_dims(x, tuple(axis) if isinstance(axis, (list, tuple)) else axis)
    ops.mean = lambda x, axis=None: np.mean(_f32(x), axis=tuple(axis) if isinstance(axis, (list, tuple)) else axis,
                                            dtype=np.float32)
    ops.cast = lambda x, dtype: np.asarray(x).astype(dtype)
    ops.clip = lambda x, lo, hi: np.clip(_f32(x), np.float32(lo), np.float32(hi))
    return ops


def _make_tf():
    # the ops dense_image_warp.py touches (see make_warp_golden.py) + the three image ops of keras_layers.py
    from make_warp_golden import _make_tf as warp_tf, _w
    tf = warp_tf()
    tf.shape = lambda x: np.array(np.asarray(x).shape, np.int64)

    def resize_bilinear(images, size, align_corners=False, half_pixel_centers=False):
        assert not align_corners and not half_pixel_centers
        images = _f32(images)
        scale = int(size[0]) // images.shape[1]
        assert int(size[0]) == images.shape[1] * scale and int(size[1]) == images.shape[2] * scale
        return _per_image(naive.resize_bilinear_legacy, images, scale)
    tf.compat = types.SimpleNamespace(v1=types.SimpleNamespace(image=types.SimpleNamespace(
        resize_bilinear=resize_bilinear)))
    tf.nn = types.SimpleNamespace(space_to_depth=lambda x, b: _per_image(naive.space_to_depth, x, b),
                                  depth_to_space=lambda x, b: _per_image(naive.depth_to_space, x, b))
    return tf, _w


class _Permissive(types.ModuleType):
    """Stand-in for training-only Keras namespaces: any attribute is a dummy class, enough for
    the type annotations and lookup tables models.py evaluates at import time."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (), {})


def load_reference_models():
    """Import the reference's models.py / keras_layers.py / tfa with the shim in place."""
    sys.path.insert(0, HERE)
    tf, _w = _make_tf()
    layers = types.ModuleType("tensorflow.keras.layers")
    for cls in (Layer, Conv2D, Conv2DTranspose, BatchNormalization, ReLU, LeakyReLU, Activation, MaxPool2D,
                ZeroPadding2D, Concatenate, Add, Lambda, Identity):
        setattr(layers, cls.__name__, cls)
    keras = types.ModuleType("tensorflow.keras")
    keras.layers, keras.ops = layers, _make_ops()
    keras.Model, keras.Input = Model, _input
    keras.optimizers = _Permissive("tensorflow.keras.optimizers")
    keras.optimizers.schedules = _Permissive("tensorflow.keras.optimizers.schedules")
    keras.applications = _Permissive("tensorflow.keras.applications")
    keras.regularizers = _Permissive("tensorflow.keras.regularizers")
    tf.keras = keras
    mods = {"tensorflow": tf, "tensorflow.keras": keras, "tensorflow.keras.layers": layers,
            "tensorflow.keras.ops": keras.ops,
            "tensorflow.keras.optimizers": keras.optimizers,
            "tensorflow.keras.applications": keras.applications,
            "tensorflow.keras.regularizers": keras.regularizers}
    mods["tensorflow.keras.optimizers.schedules"] = keras.optimizers.schedules
    # training-only modules of the reference: stubbed, except for the constant models.py reads
    km = types.ModuleType("keras_models")
    km.FRVSRModelSingle = km.FRVSRModel = km.GANModel = object
    ut = types.ModuleType("utils")
    src = open(os.path.join(REF_DIR, "utils.py")).read()
    for node in ast.parse(src).body:
        if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", None) == "BGR_LUMA":
            ut.BGR_LUMA = ast.literal_eval(node.value)
    ut.copy_model_variables = None
    mods["keras_models"], mods["utils"] = km, ut
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    sys.path.insert(0, REF_DIR)  # tfa/ and keras_layers.py are imported from the reference tree itself
    try:
        spec = importlib.util.spec_from_file_location("ref_models", os.path.join(REF_DIR, "models.py"))
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
    finally:
        sys.path.remove(REF_DIR)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        for k in ("keras_layers", "tfa", "tfa.dense_image_warp"):
            sys.modules.pop(k, None)
    ref._w = _w
    return ref


def build_reference_inference(ref, cfg: jcfg.ModelConfig):
    """The deployed graph, built by the reference's own factory functions from `cfg`."""
    if cfg.flow_arch == "autoencoder":
        flow = ref.get_flow_autoencoder(name="flow", num_inputs=cfg.flow_num_inputs, filters=list(cfg.flow_filters),
                                        activation=cfg.flow_activation)
    else:
        flow = ref.get_flow_resnet(name="flow", num_inputs=cfg.flow_num_inputs, num_filters=cfg.flow_resnet_filters,
                                   num_res_blocks=cfg.flow_resnet_blocks, activation=cfg.flow_activation)
    gen = ref.get_generator_resnet(name="generator", num_filters=cfg.gen_filters, num_res_blocks=cfg.gen_blocks,
                                   activation=cfg.gen_activation)
    return ref.get_inference_model(generator_model=gen, flow_model=flow, skip_processing=False,
                                   frame_height=cfg.frame_height, frame_width=cfg.frame_width,
                                   flow_pad_factor=cfg.flow_pad_factor or None,
                                   normalize_brightness=cfg.normalize_brightness, name="inference")


def run_reference(model, cfg, weights, frames_bgrx):
    """Recurrent roll-out like scripts/inference/onnx/inference.py:55-94: zero state, feed back."""
    global _WEIGHTS
    _WEIGHTS = weights
    pre_gen = np.zeros((1, cfg.out_height, cfg.out_width, 3), np.float32)
    last = [np.zeros((1, cfg.padded_height, cfg.padded_width, 3), np.float32)
            for _ in range(cfg.flow_num_inputs - 1)]
    outs, raws, warps = [], [], []
    for t in range(frames_bgrx.shape[0]):
        cur = frames_bgrx[t:t + 1, :, :, :3]
        res = model.call([cur, pre_gen] + last)
        outs.append(np.asarray(res["output"])[0])
        raws.append(np.asarray(res["output_raw"])[0])
        warps.append(np.asarray(res["pre_warp"])[0])
        pre_gen = np.asarray(res["output_raw"], np.float32)
        last = [np.asarray(x, np.float32) for x in res["last_frames"]]
    return np.stack(outs), np.stack(raws), np.stack(warps)


CASES = {
    # name: (preset, frames, conditioned weights)
    "tiny": ("tiny", 4, True),
    "tiny_default_init": ("tiny", 2, False),
    "small_bright": ("small_bright", 2, True),
    "small_resnet": ("small_resnet", 2, True),
}


def main():
    ref = load_reference_models()
    out = {}
    for name, (preset, nframes, conditioned) in CASES.items():
        cfg = jcfg.preset(preset)
        w = jw.init_weights(cfg, 42, conditioned)
        frames = synthetic.frames(cfg.frame_height, cfg.frame_width, nframes, kind="cut")
        model = build_reference_inference(ref, cfg)
        u8, raw, warp = run_reference(model, cfg, w, frames)
        assert u8.dtype == np.uint8 and u8.shape == (nframes, cfg.out_height, cfg.out_width, 3)
        out[f"{name}/output"] = u8
        if name == "tiny":  # float tensors only for the smallest case (fixture size)
            out[f"{name}/output_raw"] = raw[1:3].astype(np.float32)   # frames 1..2 (frame 0 warps zeros)
            out[f"{name}/pre_warp"] = warp[1:3].astype(np.float32)
        print(name, "ok", u8.shape, "mean", float(u8.mean()))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
