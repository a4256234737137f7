"""Golden vectors for the output temporal filter, produced by EXECUTING the reference's own
graph-rewriting script, /root/reference/scripts/inference/onnx/frame_moving_avg.py.

The script edits an ONNX model through `graph.Graph` (scripts/inference/onnx/graph.py) and needs
the `onnx` package, which is not installable offline.  Here its `main()` runs UNMODIFIED against
three stand-ins: a fake `onnx` (load / save only), a fake `graph.Graph` that offers the methods
main() calls (find_node_by_name, remove_node, insert_node, create_constant, create_value,
create_node, serialize, inputs) and simply records what main() builds, and an identity
`utils.simplify_model`.  The recorded node list - every op, constant, attribute and tensor name
chosen by the reference for the given CLI arguments - is then evaluated with numpy
implementations of the ONNX operators involved (Min, Max, Sub, Abs, Mul, ReduceMean, Conv, Add,
Sign, Tanh, Resize[linear, asymmetric], Slice), written from the ONNX operator specification.

What this pins: the structure and constants of the filter (oracle/frame_filter.py and
csrc/kernels/frame_filter.cu) against the reference source.  What stays unpinned: the ONNX
operator arithmetic itself (no onnxruntime to execute).

    python tests/golden/make_filter_golden.py      # build container only
Writes tests/golden/filter_golden.npz.
"""

import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/scripts/inference/onnx/frame_moving_avg.py"
OUT = os.path.join(HERE, "filter_golden.npz")

H4, W4 = 36, 44  # "model output" size of the stand-in graph (not a multiple of every window)


class Node:
    def __init__(self, name, op_type, inputs, outputs, attrs):
        self.name, self.op_type, self.input, self.output, self.attrs = name, op_type, list(inputs), list(outputs), attrs


class FakeGraph:
    """Records the edits frame_moving_avg.main() makes."""

    def __init__(self, model):
        dim = lambda v: types.SimpleNamespace(dim_value=v)  # noqa: E731
        vi = lambda dims: types.SimpleNamespace(type=types.SimpleNamespace(tensor_type=types.SimpleNamespace(  # noqa: E731
            shape=types.SimpleNamespace(dim=[dim(d) for d in dims]))))
        # inputs of the exported model: cur_frame, pre_gen (NCHW), ...
        self.inputs = [vi([1, 3, H4 // 4, W4 // 4]), vi([1, 3, H4, W4])]
        self.values = {}
        self.nodes = []
        # the two nodes main() looks up by their hard-coded names
        self._named = {
            mod.INPUT_NODE: Node(mod.INPUT_NODE, "SpaceToDepth", ["pre_warp"], ["s2d"], {}),
            mod.TARGET_NODE: Node(mod.TARGET_NODE, "Clip", ["pre_clip"], ["generator_output"], {}),
        }
        self.removed, self.reinserted = [], []

    def find_node_by_name(self, name):
        return self._named.get(name)

    def remove_node(self, node):
        self.removed.append(node.name)

    def insert_node(self, node):
        self.reinserted.append((node.name, list(node.output)))

    def create_constant(self, name, value):
        self.values[name] = np.array(value)

    def create_value(self, name, value):
        self.values[name] = np.array(value)

    def create_node(self, name, op_type, inputs=None, outputs=None, **attrs):
        node = Node(name, op_type, inputs or [], outputs or [name], attrs)
        self.nodes.append(node)
        return node

    def serialize(self):
        return self


def load_reference_script():
    global mod
    saved_onnx = types.ModuleType("onnx")
    saved_onnx.load = lambda path: "model"
    saved_onnx.save = lambda model, path: CAPTURED.append(model)
    g = types.ModuleType("graph")
    g.Graph = FakeGraph
    u = types.ModuleType("utils")
    u.simplify_model = lambda model, num_checks=0: model
    saved = {k: sys.modules.get(k) for k in ("onnx", "graph", "utils")}
    sys.modules.update({"onnx": saved_onnx, "graph": g, "utils": u})
    try:
        spec = importlib.util.spec_from_file_location("ref_frame_moving_avg", REF)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


CAPTURED = []
mod = None


# ---------------------------------------------------------------------------
# ONNX operator semantics (from the operator specification), NCHW float32
# ---------------------------------------------------------------------------

def _conv(x, w, strides, pads):
    n, c, h, ww = x.shape
    m, cw, kh, kw = w.shape
    assert cw == c and m == 1
    xp = np.pad(x, ((0, 0), (0, 0), (pads[0], pads[2]), (pads[1], pads[3])))
    oh = (xp.shape[2] - kh) // strides[0] + 1
    ow = (xp.shape[3] - kw) // strides[1] + 1
    out = np.zeros((n, 1, oh, ow), np.float32)
    for i in range(oh):
        for j in range(ow):
            patch = xp[:, :, i * strides[0]:i * strides[0] + kh, j * strides[1]:j * strides[1] + kw]
            out[:, 0, i, j] = (patch.astype(np.float64) * w[0].astype(np.float64)).sum(axis=(1, 2, 3))
    return out


def _resize_linear_asymmetric(x, scales):
    assert scales[0] == 1 and scales[1] == 1
    n, c, h, w = x.shape
    oh, ow = int(np.floor(h * scales[2])), int(np.floor(w * scales[3]))
    out = np.empty((n, c, oh, ow), np.float32)
    for oy in range(oh):
        sy = np.float32(oy) / np.float32(scales[2])       # asymmetric: x_orig = x_resized / scale
        y0 = int(np.floor(sy)); y1 = min(y0 + 1, h - 1); ty = np.float32(sy - y0)
        for ox in range(ow):
            sx = np.float32(ox) / np.float32(scales[3])
            x0 = int(np.floor(sx)); x1 = min(x0 + 1, w - 1); tx = np.float32(sx - x0)
            top = x[:, :, y0, x0] * (1 - tx) + x[:, :, y0, x1] * tx
            bot = x[:, :, y1, x0] * (1 - tx) + x[:, :, y1, x1] * tx
            out[:, :, oy, ox] = top * (1 - ty) + bot * ty
    return out


def evaluate(graph, out_nchw, pre_warp_nchw):
    env = dict(graph.values)
    env["pre_warp"] = pre_warp_nchw.astype(np.float32)
    # main() renames the Clip node's output to "output_pre_mask" and re-inserts it
    assert graph.removed == [mod.TARGET_NODE] and graph.reinserted == [(mod.TARGET_NODE, ["output_pre_mask"])]
    env["output_pre_mask"] = out_nchw.astype(np.float32)
    for node in graph.nodes:
        a = [env[i] for i in node.input]
        op = node.op_type
        if op == "Min":
            r = np.minimum(a[0], a[1])
        elif op == "Max":
            r = np.maximum(a[0], a[1])
        elif op == "Sub":
            r = a[0] - a[1]
        elif op == "Add":
            r = a[0] + a[1]
        elif op == "Mul":
            r = a[0] * a[1]
        elif op == "Abs":
            r = np.abs(a[0])
        elif op == "Sign":
            r = np.sign(a[0])
        elif op == "Tanh":
            r = np.tanh(a[0])
        elif op == "ReduceMean":
            assert not node.attrs  # all axes, keepdims = 1
            r = np.mean(a[0], dtype=np.float64, keepdims=True).astype(np.float32)
        elif op == "Conv":
            r = _conv(a[0], a[1], node.attrs["strides"], node.attrs["pads"])
        elif op == "Resize":
            assert node.attrs["coordinate_transformation_mode"] == "asymmetric" and node.attrs["mode"] == "linear"
            assert a[1].size == 0
            r = _resize_linear_asymmetric(a[0], [float(v) for v in a[2]])
        elif op == "Slice":
            starts, ends, axes = [int(v) for v in a[1]], [int(v) for v in a[2]], [int(v) for v in a[3]]
            sl = [slice(None)] * a[0].ndim
            for s, e, ax in zip(starts, ends, axes):
                sl[ax] = slice(s, e)
            r = a[0][tuple(sl)]
        else:
            raise NotImplementedError(op)
        env[node.output[0]] = np.asarray(r, np.float32)
    return env["generator_output"]


# CLI argument sets (frame_moving_avg.py:53-87); the first is the script's defaults
CASES = {
    "defaults": dict(),
    "l2_limit": dict(strength=0.5, threshold=0.02, norm="L2", limit=True),
    "gain_luma": dict(gain=8.0, luma_normalize=True),
    "window8": dict(window=8, threshold=0.05),
    "window16_all": dict(window=16, gain=4.0, norm="L2", luma_normalize=True, limit=True, threshold=0.01),
    "window5": dict(window=5, strength=0.4),
}


def case_inputs(seed, delta):
    rng = np.random.default_rng(seed)
    pw = rng.uniform(-0.6, 0.6, (1, H4, W4, 3)).astype(np.float32)
    out = np.clip(pw + rng.normal(0, delta, pw.shape), -0.5, 0.5).astype(np.float32)
    return out, pw


def main():
    ref = load_reference_script()
    res = {}
    for name, kw in CASES.items():
        args = dict(model_path="in.onnx", output_path="out.onnx", num_checks=0, strength=0.25, window=0,
                    threshold=0.1, gain=0.0, norm="L1", limit=False, luma_normalize=False)
        args.update(kw)
        args["norm"] = ref.NormType[args["norm"]]
        CAPTURED.clear()
        assert ref.main(**args) == 0
        graph = CAPTURED[-1]
        for k, delta in enumerate((0.02, 0.3)):  # a steady scene and a scene cut
            out, pw = case_inputs(100 + k, delta)
            got = evaluate(graph, out.transpose(0, 3, 1, 2), pw.transpose(0, 3, 1, 2))
            res[f"{name}/{k}"] = got.transpose(0, 2, 3, 1).astype(np.float32)
        print(name, "ok:", len(graph.nodes), "nodes,", sorted({n.op_type for n in graph.nodes}))
    np.savez_compressed(OUT, **res)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
