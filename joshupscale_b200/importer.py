"""Offline weight importer: trained Keras weights -> the native `.jup` container.

The reference trains with Keras 3 and checkpoints `*.weights.h5`
(scripts/training/train_local.py:119-128, 188); its layers carry the names used
throughout this package (scripts/training/models.py: `block_<i>_conv_<j>`, `block_<i>_bn_<j>`,
`conv_<j>`, `bn_<j>`, `conv_trans_<j>` inside a flow model and a generator model;
the container spells the block scope with a slash, `block_<i>/conv_<j>`).  This
tool takes either

  * an `.npz` whose keys are Keras variable paths, e.g. written where TensorFlow
    is installed with
        np.savez("w.npz", **{v.path: v.numpy() for v in model.variables})
  * a Keras 3 `.weights.h5`, read by the numpy-only HDF5 reader in `hdf5_lite.py` (no h5py),

recognises the flow / generator variables by their layer-relative path whatever
model scopes precede them (`final/full/generator_1/...`, `:0` suffixes, Keras'
`_<n>` uniquifiers), infers the architecture hyper-parameters from the tensor
shapes, validates every shape and writes the container the C++ runtime loads.

    python -m joshupscale_b200.importer weights.npz model.jup --height 270 --width 480
"""

from __future__ import annotations

import argparse
import re
import sys
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import numpy as np

from .config import ModelConfig, OutputFilter
from .weights import save_model, with_output_filter

_BN_VARS = ("gamma", "beta", "moving_mean", "moving_variance")


class ImportError_(ValueError):
    """The source does not describe a model this runtime implements."""


def _normalise(name: str) -> List[str]:
    name = re.sub(r":\d+$", "", name.strip("/"))
    return [p for p in name.split("/") if p]


def _scope_kind(parts: List[str]) -> Optional[str]:
    """'flow' / 'generator' from the model scopes preceding the layer path."""
    for p in reversed(parts):
        base = re.sub(r"_\d+$", "", p).lower()
        if "flow" in base:
            return "flow"
        if "gen" in base:
            return "generator"
    return None


def _split(parts: List[str]) -> Optional[Tuple[Optional[str], str]]:
    """(scope kind, layer-relative path) of a variable path, or None if it is not a
    conv / bn variable of the two networks (optimizer slots, discriminator, ...)."""
    if not parts:
        return None
    var = parts[-1]
    if var not in ("kernel", "bias") + _BN_VARS:
        return None
    layer = parts[-2] if len(parts) > 1 else ""
    rest = parts[:-2]
    # the reference scopes block layers with an underscore (models.py get_scoped_name):
    # "block_3_conv_1"; "block_3/conv_1" is accepted as well
    m = re.fullmatch(r"(block_\d+)_((?:conv|bn)_\d+)", layer)
    if m:
        rel = [m.group(1), m.group(2), var]
    elif re.fullmatch(r"(conv|bn|conv_trans)_\d+", layer):
        rel = [layer, var]
        if rest and re.fullmatch(r"block_\d+", rest[-1]):
            rel.insert(0, rest[-1])
            rest = rest[:-1]
    else:
        return None
    return _scope_kind(rest), "/".join(rel)


def collect(arrays: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """Map arbitrary Keras variable paths to `flow/...` / `generator/...` keys."""
    out: Dict[str, np.ndarray] = {}
    for name, arr in arrays.items():
        sp = _split(_normalise(name))
        if sp is None:
            continue
        kind, rel = sp
        if kind is None:
            continue  # some other network (discriminator, VGG, ...): no 'flow*' / 'gen*' scope in its path
        key = f"{kind}/{rel}"
        if key in out:
            raise ImportError_(f"two source variables map to {key} (second: '{name}')")
        out[key] = np.asarray(arr, np.float32)
    if not out:
        raise ImportError_("no flow / generator variables found")
    return out


def infer_config(w: Dict[str, np.ndarray], frame_height: int, frame_width: int,
                 flow_pad_factor: Optional[int] = None,
                 flow_activation: str = "relu", gen_activation: str = "relu",
                 normalize_brightness: bool = False) -> ModelConfig:
    """Architecture hyper-parameters from tensor shapes (scripts/training/models.py:257-263,
    334-339, 484-491); frame size and the options that leave no trace in the weights come from
    the caller."""
    def shape(key):
        if key not in w:
            raise ImportError_(f"missing variable {key}")
        return w[key].shape

    def count(prefix):
        n = 0
        while f"{prefix}/block_{n + 1}/conv_1/kernel" in w:
            n += 1
        return n

    gen_blocks = count("generator")
    k1 = shape("generator/conv_1/kernel")
    if len(k1) != 4 or k1[:3] != (3, 3, 51):
        raise ImportError_(f"generator/conv_1/kernel has shape {k1}, expected (3, 3, 51, filters)")
    gen_filters = int(k1[3])
    flow_blocks = count("flow")
    head = shape("flow/conv_2/kernel")
    resnet = head[0] == 1
    if flow_pad_factor is None:
        flow_pad_factor = 0 if resnet else 8  # the autoencoder pools three times
    kw = dict(frame_height=frame_height, frame_width=frame_width, flow_pad_factor=flow_pad_factor,
              flow_activation=flow_activation, gen_activation=gen_activation, gen_filters=gen_filters,
              gen_blocks=gen_blocks, normalize_brightness=normalize_brightness)
    if resnet:
        # get_flow_resnet: conv_1, ResBlocks, 1x1 conv_2
        c1 = shape("flow/conv_1/kernel")
        kw.update(flow_arch="resnet", flow_num_inputs=int(c1[2]) // 3, flow_resnet_filters=int(c1[3]),
                  flow_resnet_blocks=flow_blocks)
    else:
        filters = [int(shape(f"flow/block_{i + 1}/conv_1/kernel")[3]) for i in range(flow_blocks)]
        if "flow/conv_1/kernel" in w:  # odd filter count: trailing conv_1 (models.py:455-468)
            filters.append(int(shape("flow/conv_1/kernel")[3]))
        c1 = shape("flow/block_1/conv_1/kernel")
        kw.update(flow_arch="autoencoder", flow_num_inputs=int(c1[2]) // 3, flow_filters=tuple(filters))
    cfg = ModelConfig(**kw)
    cfg.validate()
    return cfg


def _expected_shapes(cfg: ModelConfig) -> Dict[str, Tuple[int, ...]]:
    from .weights import init_weights
    return {k: v.shape for k, v in init_weights(cfg, 0, False).items()}


def import_weights(arrays: Dict[str, np.ndarray], frame_height: int, frame_width: int,
                   **options) -> Tuple[ModelConfig, "OrderedDict[str, np.ndarray]"]:
    """Full import: collect, infer, validate; returns (config, weights in container order)."""
    found = collect(arrays)
    cfg = infer_config(found, frame_height, frame_width, **options)
    want = _expected_shapes(cfg)
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for key, shp in want.items():
        if key not in found:
            raise ImportError_(f"missing variable {key}")
        if tuple(found[key].shape) != tuple(shp):
            raise ImportError_(f"{key}: shape {tuple(found[key].shape)}, expected {tuple(shp)}")
        if not np.isfinite(found[key]).all():
            raise ImportError_(f"{key} contains non-finite values")
        out[key] = found[key]
    extra = sorted(set(found) - set(want))
    if extra:
        raise ImportError_(f"unexpected variables for this architecture: {extra[:4]}")
    return cfg, out


# ---------------------------------------------------------------------------
# sources
# ---------------------------------------------------------------------------

def read_npz(path: str) -> Dict[str, np.ndarray]:
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def read_keras_h5(path: str) -> Dict[str, np.ndarray]:
    """Keras 3 `.weights.h5`: datasets live at `.../<layer>/vars/<index>`; the index is mapped
    back to a variable name from the layer's tensor ranks (conv: kernel[, bias]; batch norm:
    gamma, beta, moving_mean, moving_variance).  Legacy files that store named datasets
    (`.../kernel:0`) pass through unchanged."""
    from .hdf5_lite import read_datasets  # noqa: PLC0415
    flat = {name.strip("/"): arr for name, arr in read_datasets(path, skip_unsupported=True).items()}
    return rename_indexed_vars(flat)


def rename_indexed_vars(flat: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """`<layer>/vars/<i>` -> `<layer>/<variable name>` (see read_keras_h5)."""
    by_layer: Dict[str, Dict[int, np.ndarray]] = {}
    out: Dict[str, np.ndarray] = {}
    for name, arr in flat.items():
        m = re.fullmatch(r"(.*)/vars/(\d+)", name.strip("/"))
        if m:
            by_layer.setdefault(m.group(1), {})[int(m.group(2))] = arr
        else:
            out[name] = arr
    for layer, vars_ in by_layer.items():
        # `layers/<name>` path components carry no information
        clean = "/".join(p for p in layer.split("/") if p != "layers")
        ordered = [vars_[i] for i in sorted(vars_)]
        ranks = [a.ndim for a in ordered]
        if ranks == [4] or ranks == [4, 1]:
            names = ["kernel", "bias"][:len(ranks)]
        elif ranks == [1, 1, 1, 1]:
            names = list(_BN_VARS)
        else:
            continue  # not a conv / batch-norm layer
        for n, a in zip(names, ordered):
            out[f"{clean}/{n}"] = a
    return out


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("source", help=".npz of Keras variable paths, or a Keras 3 .weights.h5")
    ap.add_argument("output", help="model.jup")
    ap.add_argument("--height", type=int, required=True, help="input frame height (PSP 270, PS2 360)")
    ap.add_argument("--width", type=int, required=True, help="input frame width (480)")
    ap.add_argument("--flow-pad-factor", type=int, default=None)
    ap.add_argument("--flow-activation", default="relu", choices=["relu", "lrelu"])
    ap.add_argument("--gen-activation", default="relu", choices=["relu", "lrelu"])
    ap.add_argument("--normalize-brightness", action="store_true")
    ap.add_argument("--filter", action="store_true", help="attach the frame_moving_avg output filter")
    ap.add_argument("--filter-strength", type=float, default=0.25)
    ap.add_argument("--filter-window", type=int, default=0)
    ap.add_argument("--filter-threshold", type=float, default=0.1)
    ap.add_argument("--filter-gain", type=float, default=0.0)
    ap.add_argument("--filter-norm", default="l1", choices=["l1", "l2"])
    ap.add_argument("--filter-limit", action="store_true")
    ap.add_argument("--filter-luma-normalize", action="store_true")
    a = ap.parse_args(argv)
    arrays = read_keras_h5(a.source) if a.source.endswith((".h5", ".hdf5")) else read_npz(a.source)
    try:
        cfg, w = import_weights(arrays, a.height, a.width, flow_pad_factor=a.flow_pad_factor,
                                flow_activation=a.flow_activation, gen_activation=a.gen_activation,
                                normalize_brightness=a.normalize_brightness)
    except ImportError_ as exc:
        print(f"import failed: {exc}", file=sys.stderr)
        return 1
    if a.filter:
        w = with_output_filter(w, OutputFilter(a.filter_strength, a.filter_window, a.filter_threshold,
                                               a.filter_gain, a.filter_norm, a.filter_limit,
                                               a.filter_luma_normalize))
    save_model(a.output, cfg, w)
    print(f"{a.output}: flow {cfg.flow_arch} K={cfg.flow_num_inputs}, generator {cfg.gen_filters}x{cfg.gen_blocks}, "
          f"{cfg.frame_width}x{cfg.frame_height} -> {cfg.out_width}x{cfg.out_height}, "
          f"{sum(v.size for v in w.values())} parameters")
    return 0


if __name__ == "__main__":
    sys.exit(main())
