"""ORACLE - CPU restatement of JoshUpscale's per-frame recurrent upscaling graph.

TEST INFRASTRUCTURE ONLY.  Nothing under joshupscale_b200/ may import this
module; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs use it, as the checker / reported baseline.

PARITY STATUS: "parity unpinned" for everything that executes inside
TensorFlow/Keras in the reference (Conv2D, BatchNormalization, MaxPool2D,
Conv2DTranspose, resize_bilinear, depth_to_space, space_to_depth): the
reference pins tensorflow-cpu==2.18.0 (scripts/training/requirements.txt:1),
which is not installable here, and ships no tests, golden vectors or weights
(CONTRIBUTING.md:24-26).  Those ops are restated from their published TF/Keras
semantics and cross-checked against an independent
naive numpy restatement (oracle/naive.py).  The dense warp is the one op whose
source is vendored in the reference (scripts/training/tfa/dense_image_warp.py);
it is pinned by executing that file under a numpy shim of the handful of tf ops
it uses (tests/golden/make_warp_golden.py -> tests/golden/warp_*.npz).

Conventions: tensors are torch float32 NHWC unless noted, channel order BGR.

Two precisions:
  "fp32"    - the reference semantics, BN applied as a separate op.
  "fp16emu" - the storage contract of the B200 engine: BN folded into fp16
              weights, activations rounded to fp16 wherever the engine stores
              them, fp32 accumulation everywhere.  Used to separate "kernel is
              wrong" from "fp16 is fp16" in the parity tests.
"""

from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

BGR_LUMA = (0.1140, 0.5870, 0.2989)  # scripts/training/utils.py:151
BN_EPS = 1e-3
LRELU_SLOPE = 0.3


# --------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------

def _t(x) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x
    return torch.from_numpy(np.ascontiguousarray(x))


def r16(x: torch.Tensor) -> torch.Tensor:
    """Round to fp16 storage and back (fp16emu mode)."""
    return x.to(torch.float16).to(torch.float32)


def conv2d_same(x: torch.Tensor, kernel: torch.Tensor,
                bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """keras.layers.Conv2D(strides=1, padding="same"): cross-correlation,
    kernel (kh, kw, Cin, Cout), zero padding (kh-1)/2 on every side.
    Call sites: scripts/training/models.py:218-225, 237-244, 377-396, 531-537.
    """
    kh = kernel.shape[0]
    w = kernel.permute(3, 2, 0, 1).contiguous()
    y = F.conv2d(x.permute(0, 3, 1, 2), w, bias, padding=(kh - 1) // 2)
    return y.permute(0, 2, 3, 1).contiguous()


def batch_norm(x, gamma, beta, mean, var, eps=BN_EPS):
    """keras BatchNormalization at inference, axis=-1 (models.py:226-228)."""
    return (x - mean) / torch.sqrt(var + eps) * gamma + beta


def activation(x, kind: str):
    """models.py:24-27: "relu" -> ReLU, "lrelu" -> LeakyReLU (slope 0.3)."""
    if kind == "relu":
        return torch.clamp_min(x, 0.0)
    if kind == "lrelu":
        return torch.where(x >= 0, x, x * LRELU_SLOPE)
    raise ValueError(kind)


def max_pool2(x):
    """keras MaxPool2D(pool_size=2): stride 2, VALID (models.py:406-409)."""
    n, h, w, c = x.shape
    x = x[:, :h // 2 * 2, :w // 2 * 2, :]
    return x.reshape(n, h // 2, 2, w // 2, 2, c).amax(dim=(2, 4))


def resize_bilinear_legacy(x, scale: int):
    """tf.compat.v1.image.resize_bilinear(align_corners=False,
    half_pixel_centers=False) by an integer factor (keras_layers.py:46-52):
    src = dst / scale, lo = floor(src), hi = min(lo + 1, size - 1),
    value = top + (bottom - top) * ty with top = tl + (tr - tl) * tx.
    """
    n, h, w, c = x.shape
    oy = torch.arange(h * scale, dtype=torch.float32) / scale
    ox = torch.arange(w * scale, dtype=torch.float32) / scale
    y0 = oy.floor().long()
    x0 = ox.floor().long()
    y1 = torch.clamp(y0 + 1, max=h - 1)
    x1 = torch.clamp(x0 + 1, max=w - 1)
    ty = (oy - y0).view(1, -1, 1, 1)
    tx = (ox - x0).view(1, 1, -1, 1)
    rows0 = x[:, y0]
    rows1 = x[:, y1]
    top = rows0[:, :, x0] + (rows0[:, :, x1] - rows0[:, :, x0]) * tx
    bot = rows1[:, :, x0] + (rows1[:, :, x1] - rows1[:, :, x0]) * tx
    return top + (bot - top) * ty


def depth_to_space(x, b: int):
    """tf.nn.depth_to_space NHWC ("DCR"): out[n, h*b+i, w*b+j, c] =
    in[n, h, w, (i*b + j)*C' + c] (keras_layers.py:175)."""
    n, h, w, c = x.shape
    cp = c // (b * b)
    x = x.reshape(n, h, w, b, b, cp).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(n, h * b, w * b, cp).contiguous()


def space_to_depth(x, b: int):
    """tf.nn.space_to_depth NHWC: out[n, h, w, (i*b + j)*C + c] =
    in[n, h*b+i, w*b+j, c] (keras_layers.py:129)."""
    n, h, w, c = x.shape
    x = x.reshape(n, h // b, b, w // b, b, c).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(n, h // b, w // b, b * b * c).contiguous()


def conv2d_transpose_k2s2(x, kernel, bias=None):
    """keras Conv2DTranspose(kernel_size=2, strides=2, padding="same"),
    kernel (kh, kw, Cout, Cin): out[2h+i, 2w+j, o] = sum_c in[h,w,c]*K[i,j,o,c]
    (models.py:559-566, 573-579)."""
    w = kernel.permute(3, 2, 0, 1).contiguous()  # (Cin, Cout, kh, kw)
    y = F.conv_transpose2d(x.permute(0, 3, 1, 2), w, bias, stride=2)
    return y.permute(0, 2, 3, 1).contiguous()


@dataclasses.dataclass
class WarpTaps:
    fy: torch.Tensor  # int32 [N,H,W] clamped floor of the y query
    fx: torch.Tensor
    ay: torch.Tensor  # float32 [N,H,W] clamped alpha
    ax: torch.Tensor


def warp_taps(flow, height: int, width: int) -> WarpTaps:
    """Query points and interpolation taps of dense_image_warp
    (scripts/training/tfa/dense_image_warp.py:113-139, 232-237):
    q = grid - flow with flow[...,0] = dy, flow[...,1] = dx; per axis
    floor = min(max(0, floor(q)), size - 2); alpha = clamp(q - floor, 0, 1).
    All arithmetic in flow's dtype (float32)."""
    gy = torch.arange(height, dtype=torch.float32).view(1, -1, 1)
    gx = torch.arange(width, dtype=torch.float32).view(1, 1, -1)
    qy = gy - flow[..., 0]
    qx = gx - flow[..., 1]
    fy = torch.clamp(torch.floor(qy), 0.0, float(height - 2))
    fx = torch.clamp(torch.floor(qx), 0.0, float(width - 2))
    ay = torch.clamp(qy - fy, 0.0, 1.0)
    ax = torch.clamp(qx - fx, 0.0, 1.0)
    return WarpTaps(fy.to(torch.int32), fx.to(torch.int32), ay, ax)


def dense_image_warp(image, flow):
    """tfa dense_image_warp (dense_image_warp.py:154-171, 222-245): gather the
    four corners, lerp along x then along y:
      top = ax*(TR-TL)+TL ; bot = ax*(BR-BL)+BL ; out = ay*(bot-top)+top."""
    n, h, w, c = image.shape
    taps = warp_taps(flow, h, w)
    fy = taps.fy.long()
    fx = taps.fx.long()
    flat = image.reshape(n, h * w, c)
    base = (fy * w + fx).reshape(n, -1)

    def gather(idx):
        return torch.gather(flat, 1, idx.unsqueeze(-1).expand(-1, -1, c))

    tl = gather(base)
    tr = gather(base + 1)
    bl = gather(base + w)
    br = gather(base + w + 1)
    ax = taps.ax.reshape(n, -1, 1)
    ay = taps.ay.reshape(n, -1, 1)
    top = ax * (tr - tl) + tl
    bot = ax * (br - bl) + bl
    out = ay * (bot - top) + top
    return out.reshape(n, h, w, c)


def preprocess(frame_u8):
    """PreprocessLayer (keras_layers.py:208): x / 255 - 0.5 in float32."""
    return _t(frame_u8).to(torch.float32) / 255.0 - 0.5


def postprocess(x):
    """PostprocessLayer (keras_layers.py:227-230): uint8((x + 0.5) * 255);
    the float -> uint8 cast truncates toward zero."""
    return ((x + 0.5) * 255.0).to(torch.uint8)


def pack_bgrx(u8_bgr):
    """C++ output packing: [B, G, R, 0] (core/src/cuda_convert.cc.cu:39-45)."""
    n, h, w, _ = u8_bgr.shape
    out = torch.zeros((n, h, w, 4), dtype=torch.uint8)
    out[..., :3] = u8_bgr
    return out


# --------------------------------------------------------------------------
# networks
# --------------------------------------------------------------------------

class Graph:
    """The inference model of scripts/training/models.py:680-829 for one
    ModelConfig-like object `cfg` and a dict of Keras-layout weights."""

    def __init__(self, cfg, weights: Dict[str, np.ndarray],
                 precision: str = "fp32", output_filter=None):
        assert precision in ("fp32", "fp16emu")
        self.cfg = cfg
        self.precision = precision
        # optional oracle.frame_filter.FrameFilter (scripts/inference/onnx/frame_moving_avg.py)
        self.output_filter = output_filter
        self.w = {k: _t(np.asarray(v, np.float32)) for k, v in weights.items()}
        self.taps: Dict[str, torch.Tensor] = {}
        self.keep_taps = False

    # ---- helpers -------------------------------------------------------
    def _tap(self, name, value):
        if self.keep_taps:
            self.taps[name] = value

    def _round(self, x):
        return r16(x) if self.precision == "fp16emu" else x

    def _bn_params(self, name):
        return (self.w[f"{name}/gamma"], self.w[f"{name}/beta"],
                self.w[f"{name}/moving_mean"], self.w[f"{name}/moving_variance"])

    def _conv_bn_act(self, x, conv, bn, act, residual=None, kernel=None):
        """conv (no bias) -> BN -> [+ residual] -> activation."""
        k = self.w[f"{conv}/kernel"] if kernel is None else kernel
        gamma, beta, mean, var = self._bn_params(bn)
        if self.precision == "fp32":
            y = batch_norm(conv2d_same(x, k), gamma, beta, mean, var)
        else:
            s = gamma / torch.sqrt(var + BN_EPS)
            y = conv2d_same(x, r16(k * s)) + (beta - mean * s)
        if residual is not None:
            y = y + residual
        if act is not None:
            y = activation(y, act)
        return self._round(y)

    def _res_block(self, x, prefix, act):
        """res_block (models.py:193-254): conv-BN-act-conv-BN-add-act."""
        h = self._conv_bn_act(x, f"{prefix}/conv_1", f"{prefix}/bn_1", act)
        return self._conv_bn_act(h, f"{prefix}/conv_2", f"{prefix}/bn_2", act,
                                 residual=x)

    # ---- flow nets -----------------------------------------------------
    def flow_autoencoder(self, frames: List[torch.Tensor]):
        """get_flow_autoencoder (models.py:334-481)."""
        cfg = self.cfg
        act = cfg.flow_activation
        x = torch.cat(frames, dim=-1)  # [current, previous, ...] models.py:373-375
        f = cfg.flow_filters
        n = len(f) // 2
        for i in range(2 * n):
            p = f"flow/block_{i + 1}"
            x = self._conv_bn_act(x, f"{p}/conv_1", f"{p}/bn_1", act)
            x = self._conv_bn_act(x, f"{p}/conv_2", f"{p}/bn_2", act)
            if i < n:
                x = max_pool2(x)
            else:
                # UpscaleLayer(resize_type="bilinear", scale=2, dtype="float32")
                x = self._round(resize_bilinear_legacy(x, 2))
            self._tap(f"{p}", x)
        if len(f) % 2:
            x = self._conv_bn_act(x, "flow/conv_1", "flow/bn_1", act)
        k = self.w["flow/conv_2/kernel"]
        if self.precision == "fp16emu":
            k = r16(k)
        head = conv2d_same(x, k) + self.w["flow/conv_2/bias"]  # stays fp32
        self._tap("flow/head", head)
        return depth_to_space(head, 4)

    def flow_resnet(self, frames: List[torch.Tensor]):
        """get_flow_resnet (models.py:257-331)."""
        cfg = self.cfg
        act = cfg.flow_activation
        x = torch.cat(frames, dim=-1)
        x = self._conv_bn_act(x, "flow/conv_1", "flow/bn_1", act)
        for i in range(cfg.flow_resnet_blocks):
            x = self._res_block(x, f"flow/block_{i + 1}", act)
        k = self.w["flow/conv_2/kernel"]
        if self.precision == "fp16emu":
            k = r16(k)
        head = conv2d_same(x, k) + self.w["flow/conv_2/bias"]
        self._tap("flow/head", head)
        return depth_to_space(head, 4)

    def flow_model(self, frames):
        if self.cfg.flow_arch == "autoencoder":
            return self.flow_autoencoder(frames)
        return self.flow_resnet(frames)

    # ---- generator -----------------------------------------------------
    def generator(self, cur, pre_warp):
        """get_generator_resnet (models.py:484-595).  `cur` is float32
        [N,H,W,3] exactly as preprocessed; `pre_warp` [N,4H,4W,3]."""
        cfg = self.cfg
        act = cfg.gen_activation
        x = torch.cat([self._round(cur), space_to_depth(pre_warp, 4)], dim=-1)
        self._tap("generator/input", x)
        x = self._conv_bn_act(x, "generator/conv_1", "generator/bn_1", act)
        self._tap("generator/conv_1", x)
        for i in range(cfg.gen_blocks):
            x = self._res_block(x, f"generator/block_{i + 1}", act)
            self._tap(f"generator/block_{i + 1}", x)
        # conv_trans_1 (no bias) + bn_2 + act
        k1 = self.w["generator/conv_trans_1/kernel"]  # (2,2,32,nf)
        gamma, beta, mean, var = self._bn_params("generator/bn_2")
        if self.precision == "fp32":
            y = batch_norm(conv2d_transpose_k2s2(x, k1), gamma, beta, mean, var)
        else:
            s = gamma / torch.sqrt(var + BN_EPS)
            y = conv2d_transpose_k2s2(x, r16(k1 * s.view(1, 1, -1, 1))) + (beta - mean * s)
        y = self._round(activation(y, act))
        self._tap("generator/conv_trans_1", y)
        k2 = self.w["generator/conv_trans_2/kernel"]
        if self.precision == "fp16emu":
            k2 = r16(k2)
        z = conv2d_transpose_k2s2(y, k2) + self.w["generator/conv_trans_2/bias"]
        z = torch.tanh(z)
        self._tap("generator/tanh", z)
        up = resize_bilinear_legacy(cur, 4)  # UpscaleLayer(scale=4), fp32 input
        out = torch.clamp(up + z, -0.5, 0.5)  # Add + ClipLayer
        return out

    # ---- one recurrent step -------------------------------------------
    def zero_state(self, batch: int = 1):
        """Zero initial recurrent state (keras_models.py:58-60;
        scripts/inference/onnx/inference.py:67-70)."""
        cfg = self.cfg
        return {
            "pre_gen": torch.zeros(batch, cfg.out_height, cfg.out_width, 3),
            "last_frames": [
                torch.zeros(batch, cfg.padded_height, cfg.padded_width, 3)
                for _ in range(cfg.flow_num_inputs - 1)],
        }

    def step(self, frame_bgrx_u8, state):
        """One call of the deployed model (models.py:766-828 wiring).

        frame_bgrx_u8: uint8 [N,H,W,4] (X ignored) or [N,H,W,3].
        Returns (output_bgrx_u8 [N,4H,4W,4], new_state, aux dict).
        """
        cfg = self.cfg
        frame = _t(frame_bgrx_u8)[..., :3]
        cur = preprocess(frame)
        cur_pad = cur
        brightness = None
        if cfg.normalize_brightness:
            luma = torch.tensor(BGR_LUMA, dtype=torch.float32)
            brightness = (cur * luma * 3).mean(dim=(1, 2, 3), keepdim=True)
            cur_pad = cur_pad - brightness
        ph, pw = cfg.padded_height, cfg.padded_width
        h, w = cfg.frame_height, cfg.frame_width
        if ph != h or pw != w:
            top, left = (ph - h) // 2, (pw - w) // 2
            cur_pad = F.pad(cur_pad, (0, 0, left, pw - w - left, top, ph - h - top))
        cur_pad = self._round(cur_pad)
        flow = self.flow_model([cur_pad] + list(state["last_frames"]))
        if ph != h or pw != w:
            oy, ox = ((ph - h) // 2) * 4, ((pw - w) // 2) * 4
            flow = flow[:, oy:oy + 4 * h, ox:ox + 4 * w, :]
        flow = flow.contiguous()
        pre_warp = dense_image_warp(state["pre_gen"], flow)
        if brightness is not None:
            pre_warp = pre_warp + brightness
        pre_warp = self._round(pre_warp)
        out_raw = self.generator(cur, pre_warp)
        if self.output_filter is not None:
            from .frame_filter import frame_moving_avg
            if self.precision == "fp16emu":
                out_raw = r16(out_raw)  # the product blends the fp16-stored generator output
            filter_debug = {}
            out_raw = frame_moving_avg(out_raw, pre_warp, self.output_filter, filter_debug)
        else:
            filter_debug = None
        output = pack_bgrx(postprocess(out_raw))
        new_pre_gen = out_raw - brightness if brightness is not None else out_raw
        new_state = {
            "pre_gen": self._round(new_pre_gen),
            "last_frames": [cur_pad] + list(state["last_frames"][:-1]),
        }
        aux = {"flow": flow, "pre_warp": pre_warp, "out_raw": out_raw,
               "cur_pad": cur_pad, "filter": filter_debug}
        return output, new_state, aux

    def run(self, frames_u8, state=None):
        """Recurrent roll-out over [T,H,W,4] frames of ONE stream."""
        if state is None:
            state = self.zero_state(1)
        outs = []
        for t in range(frames_u8.shape[0]):
            o, state, _ = self.step(frames_u8[t:t + 1], state)
            outs.append(o[0].numpy())
        return np.stack(outs), state


def psnr_u8(a: np.ndarray, b: np.ndarray) -> float:
    d = a.astype(np.float64) - b.astype(np.float64)
    mse = float((d * d).mean())
    if mse == 0:
        return float("inf")
    return 10.0 * np.log10(255.0 * 255.0 / mse)
