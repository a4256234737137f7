"""Random-init weights in Keras layouts and the `.jup` model container.

The reference ships no weights; parity is defined on "the same inputs and
random-init weights" (BASELINE.json north_star).  Tensors are kept in the
layouts Keras uses so that a trained `.weights.h5` could be converted 1:1:

  Conv2D kernel           (kh, kw, Cin, Cout)   scripts/training/models.py:218-225
  Conv2DTranspose kernel  (kh, kw, Cout, Cin)   scripts/training/models.py:559-566
  BatchNormalization      gamma, beta, moving_mean, moving_variance

BN folding and the kernel-native fp16 packing happen inside the C++ loader
(joshupscale_b200/csrc/host/model.cc); this file only writes the container.
The `.jup` file takes the place of the reference's serialized TensorRT engine
at the `modelPath` argument of createRuntime (core/src/core.cc:154-167).
"""

from __future__ import annotations

import struct
from collections import OrderedDict
from typing import Dict, Tuple

import numpy as np

from .config import BN_EPS, LRELU_SLOPE, ModelConfig

MAGIC = b"JUPMDL\x00\x01"
VERSION = 1
_HEADER_FMT = "<8sII" + "IIII" + "II" + "I16I" + "II" + "IfIf" + "I" + "f" + "I"
_ENTRY_FMT = "<96sII4IQQ"
_ARCH = {"autoencoder": 0, "resnet": 1}
_ACT = {"relu": 0, "lrelu": 1}


def _glorot(rng: np.random.Generator, shape: Tuple[int, ...], fan_in: int,
            fan_out: int) -> np.ndarray:
    limit = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-limit, limit, size=shape).astype(np.float32)


def _conv(rng, kh, cin, cout):
    # keras glorot_uniform: fan_in = kh*kw*Cin, fan_out = kh*kw*Cout
    return _glorot(rng, (kh, kh, cin, cout), kh * kh * cin, kh * kh * cout)


def _conv_t(rng, kh, cin, cout):
    # Conv2DTranspose kernel (kh, kw, Cout, Cin); keras computes fans from the
    # kernel shape as (receptive * shape[-2], receptive * shape[-1])
    return _glorot(rng, (kh, kh, cout, cin), kh * kh * cout, kh * kh * cin)


def _bn(rng, c, conditioned, small_gamma=False):
    if not conditioned:
        return OrderedDict(
            gamma=np.ones(c, np.float32), beta=np.zeros(c, np.float32),
            moving_mean=np.zeros(c, np.float32),
            moving_variance=np.ones(c, np.float32))
    if small_gamma:
        gamma = rng.uniform(0.05, 0.15, c)
    else:
        gamma = rng.uniform(0.8, 1.2, c)
    return OrderedDict(
        gamma=gamma.astype(np.float32),
        beta=rng.normal(0, 0.05, c).astype(np.float32),
        moving_mean=rng.normal(0, 0.05, c).astype(np.float32),
        moving_variance=rng.uniform(0.8, 1.2, c).astype(np.float32))


def init_weights(cfg: ModelConfig, seed: int = 42,
                 conditioned: bool = True) -> "OrderedDict[str, np.ndarray]":
    """Seeded random-init weights.

    conditioned=False: set A "Keras-default" (glorot kernels, zero biases,
    identity BN statistics).  conditioned=True: set B - same
    kernels but non-trivial BN statistics, small bn_2.gamma in every ResBlock
    so the residual trunk does not blow up, a damped output head so the tanh
    branch is not saturated, and a flow head that produces sub-pixel to
    few-pixel motion so the warp is exercised away from integer positions.
    """
    cfg.validate()
    rng = np.random.default_rng(seed)
    w: "OrderedDict[str, np.ndarray]" = OrderedDict()

    def put_bn(prefix, c, small_gamma=False):
        for k, v in _bn(rng, c, conditioned, small_gamma).items():
            w[f"{prefix}/{k}"] = v

    cin = 3 * cfg.flow_num_inputs
    if cfg.flow_arch == "autoencoder":
        f = cfg.flow_filters
        n = len(f) // 2
        for i in range(2 * n):
            p = f"flow/block_{i + 1}"
            w[f"{p}/conv_1/kernel"] = _conv(rng, 3, cin, f[i])
            put_bn(f"{p}/bn_1", f[i])
            w[f"{p}/conv_2/kernel"] = _conv(rng, 3, f[i], f[i])
            put_bn(f"{p}/bn_2", f[i])
            cin = f[i]
        if len(f) % 2:
            w["flow/conv_1/kernel"] = _conv(rng, 3, cin, f[-1])
            put_bn("flow/bn_1", f[-1])
            cin = f[-1]
        w["flow/conv_2/kernel"] = _conv(rng, 3, cin, 32)
        w["flow/conv_2/bias"] = np.zeros(32, np.float32)
    else:
        nf = cfg.flow_resnet_filters
        w["flow/conv_1/kernel"] = _conv(rng, 3, cin, nf)
        put_bn("flow/bn_1", nf)
        for i in range(cfg.flow_resnet_blocks):
            p = f"flow/block_{i + 1}"
            w[f"{p}/conv_1/kernel"] = _conv(rng, 3, nf, nf)
            put_bn(f"{p}/bn_1", nf)
            w[f"{p}/conv_2/kernel"] = _conv(rng, 3, nf, nf)
            put_bn(f"{p}/bn_2", nf, small_gamma=True)
        w["flow/conv_2/kernel"] = _conv(rng, 1, nf, 32)
        w["flow/conv_2/bias"] = np.zeros(32, np.float32)
    if conditioned:
        w["flow/conv_2/kernel"] = (w["flow/conv_2/kernel"] * 16.0).astype(np.float32)
        w["flow/conv_2/bias"] = rng.normal(0, 1.5, 32).astype(np.float32)

    nf = cfg.gen_filters
    w["generator/conv_1/kernel"] = _conv(rng, 3, 51, nf)
    put_bn("generator/bn_1", nf)
    for i in range(cfg.gen_blocks):
        p = f"generator/block_{i + 1}"
        w[f"{p}/conv_1/kernel"] = _conv(rng, 3, nf, nf)
        put_bn(f"{p}/bn_1", nf)
        w[f"{p}/conv_2/kernel"] = _conv(rng, 3, nf, nf)
        put_bn(f"{p}/bn_2", nf, small_gamma=True)
    w["generator/conv_trans_1/kernel"] = _conv_t(rng, 2, nf, 32)
    put_bn("generator/bn_2", 32)
    w["generator/conv_trans_2/kernel"] = _conv_t(rng, 2, 32, 3)
    w["generator/conv_trans_2/bias"] = np.zeros(3, np.float32)
    if conditioned:
        w["generator/conv_trans_2/kernel"] = (
            w["generator/conv_trans_2/kernel"] * 4.0).astype(np.float32)
        w["generator/conv_trans_2/bias"] = rng.normal(0, 0.02, 3).astype(np.float32)
    return w


def _pack_header(cfg: ModelConfig, n_tensors: int, header_bytes: int) -> bytes:
    if cfg.flow_arch == "autoencoder":
        filt = list(cfg.flow_filters)
    else:
        filt = [cfg.flow_resnet_filters, cfg.flow_resnet_blocks]
    if len(filt) > 16:
        raise ValueError("too many flow filters")
    nf = len(filt)
    filt = filt + [0] * (16 - nf)
    return struct.pack(
        _HEADER_FMT, MAGIC, VERSION, header_bytes,
        cfg.frame_height, cfg.frame_width, cfg.padded_height, cfg.padded_width,
        _ARCH[cfg.flow_arch], cfg.flow_num_inputs,
        nf, *filt,
        cfg.gen_filters, cfg.gen_blocks,
        _ACT[cfg.flow_activation], LRELU_SLOPE,
        _ACT[cfg.gen_activation], LRELU_SLOPE,
        int(cfg.normalize_brightness), BN_EPS, n_tensors)


FILTER_TENSOR = "meta/frame_moving_avg"


def with_output_filter(weights: Dict[str, np.ndarray], flt) -> "OrderedDict[str, np.ndarray]":
    """Copy of `weights` with the output filter `flt` (config.OutputFilter, or None
    to remove it) attached as the `meta/frame_moving_avg` tensor."""
    out = OrderedDict((k, v) for k, v in weights.items() if k != FILTER_TENSOR)
    if flt is not None:
        out[FILTER_TENSOR] = np.asarray(flt.as_vector(), np.float32)
    return out


def save_model(path: str, cfg: ModelConfig, weights: Dict[str, np.ndarray]) -> None:
    """Write a `.jup` container: header, tensor table, 64-byte aligned fp32 data."""
    cfg.validate()
    hsize = struct.calcsize(_HEADER_FMT)
    esize = struct.calcsize(_ENTRY_FMT)
    table_end = hsize + esize * len(weights)
    offset = (table_end + 63) // 64 * 64
    entries, blobs = [], []
    for name, arr in weights.items():
        a = np.ascontiguousarray(arr, dtype="<f4")
        if a.ndim > 4:
            raise ValueError("rank > 4")
        dims = list(a.shape) + [1] * (4 - a.ndim)
        entries.append(struct.pack(_ENTRY_FMT, name.encode(), 0, a.ndim, *dims,
                                   offset, a.nbytes))
        blobs.append((offset, a.tobytes()))
        offset = (offset + a.nbytes + 63) // 64 * 64
    with open(path, "wb") as f:
        f.write(_pack_header(cfg, len(weights), hsize))
        for e in entries:
            f.write(e)
        for off, blob in blobs:
            f.seek(off)
            f.write(blob)
        f.truncate(offset)


def load_model(path: str) -> Tuple[ModelConfig, "OrderedDict[str, np.ndarray]"]:
    """Read a `.jup` container back (round-trip check for the C++ reader)."""
    with open(path, "rb") as f:
        data = f.read()
    hsize = struct.calcsize(_HEADER_FMT)
    h = struct.unpack_from(_HEADER_FMT, data, 0)
    if h[0] != MAGIC or h[1] != VERSION:
        raise ValueError("not a .jup model")
    (fh, fw, ph, pw, arch, k, nfilt) = h[3:10]
    filt = h[10:26][:nfilt]
    gen_filters, gen_blocks, act_f, _, act_g, _, nb, _, n_tensors = h[26:35]
    arch_name = {v: k_ for k_, v in _ARCH.items()}[arch]
    act_name = {v: k_ for k_, v in _ACT.items()}
    pad = 0
    if ph != fh or pw != fw:
        # the container stores the padded size, not the factor: take the reference's factor (8) when
        # it reproduces the padded size, else the smallest that does
        for cand in (8, *range(2, 257)):
            if (fh + cand - 1) // cand * cand == ph and (fw + cand - 1) // cand * cand == pw:
                pad = cand
                break
    kw = dict(frame_height=fh, frame_width=fw, flow_pad_factor=pad,
              flow_arch=arch_name, flow_num_inputs=k,
              flow_activation=act_name[act_f], gen_filters=gen_filters,
              gen_blocks=gen_blocks, gen_activation=act_name[act_g],
              normalize_brightness=bool(nb))
    if arch_name == "autoencoder":
        kw["flow_filters"] = tuple(filt)
    else:
        kw["flow_resnet_filters"], kw["flow_resnet_blocks"] = filt
    cfg = ModelConfig(**kw)
    weights: "OrderedDict[str, np.ndarray]" = OrderedDict()
    esize = struct.calcsize(_ENTRY_FMT)
    for i in range(n_tensors):
        e = struct.unpack_from(_ENTRY_FMT, data, h[2] + i * esize)
        name = e[0].rstrip(b"\x00").decode()
        ndim = e[2]
        dims = e[3:7][:ndim]
        off, nbytes = e[7], e[8]
        weights[name] = np.frombuffer(data, "<f4", nbytes // 4, off).reshape(dims).copy()
    return cfg, weights
