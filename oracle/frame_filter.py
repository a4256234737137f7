"""CPU restatement of the deployed graphs' output temporal filter.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg, never by the product.

Restates scripts/inference/onnx/frame_moving_avg.py:142-307, an ONNX graph
rewrite that replaces the generator's clipped output `out` by a blend with the
warped previous output `pre_warp` unless a scene change is detected:

    pw    = clip(pre_warp, -0.5, 0.5) if limit else pre_warp          (:159-170)
    d     = |out - pw|  (L1)   or   (out - pw)^2  (L2)                 (:174-186)
    m     = mean(d [* luma]) [* gain]                  window == 0    (:187-213)
          = conv(d, ones/(3 w^2) * gain [* luma], stride w, zero pad) window  > 0  (:214-235)
    cond  = sign(m - threshold*gain_coef)  if gain == 0 else tanh(...) (:236-245)
            (window > 0: bilinear "asymmetric" resize by w, crop the padding)  (:246-279)
    mask  = cond * (-s/2) + s/2 ;  mask2 = cond * (s/2) + (1 - s/2)    (:280-293)
    final = pw * mask + out * mask2                                    (:294-303)

`final` replaces the clip output, so it feeds both Postprocess (u8 image) and
output_raw (the recurrent pre_gen state).

Pinning: tests/golden/make_filter_golden.py runs the reference script's main()
unmodified against recording stand-ins for `onnx` / `graph.Graph` and evaluates
the op list it builds; tests/test_frame_filter.py checks this module against
those vectors (structure and constants pinned).  The arithmetic of the ONNX
operators themselves (Conv, ReduceMean, Resize) is restated from the operator
specification - onnx / onnxruntime are not installed here.
"""

from __future__ import annotations

import dataclasses

import numpy as np
import torch
import torch.nn.functional as F

# LUMA_NORM, frame_moving_avg.py:95-96 (channel order B, G, R), already times 3
LUMA_NORM = np.array([0.1140, 0.5870, 0.2989], dtype=np.float32) * 3


@dataclasses.dataclass(frozen=True)
class FrameFilter:
    """CLI arguments of frame_moving_avg.py:53-87 with the same defaults."""
    strength: float = 0.25
    window: int = 0
    threshold: float = 0.1
    gain: float = 0.0
    norm: str = "l1"  # "l1" | "l2"
    limit: bool = False
    luma_normalize: bool = False

    def as_vector(self) -> np.ndarray:
        """float32[8] stored in the model container as `meta/frame_moving_avg`."""
        return np.array([1.0, self.strength, float(self.window), self.threshold, self.gain,
                         1.0 if self.norm.lower() == "l2" else 0.0,
                         1.0 if self.limit else 0.0,
                         1.0 if self.luma_normalize else 0.0], dtype=np.float32)


def _resize_asymmetric(x: torch.Tensor, scale: int) -> torch.Tensor:
    """ONNX Resize(mode=linear, coordinate_transformation_mode=asymmetric) of
    [N,H,W] by an integer scale: src = dst / scale, neighbour clamped to the
    last sample (frame_moving_avg.py:258-266)."""
    n, h, w = x.shape

    def axis(size):
        o = torch.arange(size * scale, dtype=torch.float32)
        src = o / np.float32(scale)
        lo = torch.floor(src)
        t = src - lo
        lo = lo.to(torch.int64)
        hi = torch.clamp(lo + 1, max=size - 1)
        return lo, hi, t

    y0, y1, ty = axis(h)
    x0, x1, tx = axis(w)
    top = x[:, y0][:, :, x0] * (1 - tx) + x[:, y0][:, :, x1] * tx
    bot = x[:, y1][:, :, x0] * (1 - tx) + x[:, y1][:, :, x1] * tx
    return top * (1 - ty)[None, :, None] + bot * ty[None, :, None]


def frame_moving_avg(out: torch.Tensor, pre_warp: torch.Tensor, flt: FrameFilter,
                     debug: dict = None) -> torch.Tensor:
    """out, pre_warp: float32 [N, 4H, 4W, 3] (NHWC, BGR).  Each batch entry is
    an independent stream (the reference graph has N = 1).  `debug`, if given,
    receives "th": the scene-detection value before sign / tanh ([N] or [N, cells_y,
    cells_x]) so that tests can tell decisions that sit on the threshold."""
    s = float(flt.strength)
    gain_coef = 1.0 if flt.gain == 0 else float(flt.gain)
    pw = torch.clamp(pre_warp, -0.5, 0.5) if flt.limit else pre_warp
    diff = out - pw
    d = diff.abs() if flt.norm.lower() == "l1" else diff * diff
    luma = torch.from_numpy(LUMA_NORM.copy())
    if flt.norm.lower() == "l2":
        luma2 = luma * luma
    else:
        luma2 = luma
    n, hh, ww, _ = out.shape
    if flt.window == 0:
        if flt.luma_normalize:
            m = (d * (luma2 * np.float32(gain_coef))).mean(dim=(1, 2, 3))
        else:
            m = d.mean(dim=(1, 2, 3))
            if flt.gain != 0:
                m = m * np.float32(gain_coef)
        th = m + np.float32(-flt.threshold * gain_coef)
        if debug is not None:
            debug["th"] = th.clone()
        cond = torch.sign(th) if flt.gain == 0 else torch.tanh(th)
        cond = cond.view(n, 1, 1, 1)
    else:
        wnd = flt.window
        oh, ow = [((x + wnd - 1) // wnd) * wnd for x in (hh, ww)]
        pad_t, pad_l = (oh - hh) // 2, (ow - ww) // 2
        kernel = torch.full((1, 3, wnd, wnd), 1.0 / 3 / wnd / wnd * gain_coef, dtype=torch.float32)
        if flt.luma_normalize:
            kernel = kernel * luma2.view(1, 3, 1, 1)
        dn = d.permute(0, 3, 1, 2)
        dn = F.pad(dn, (pad_l, ow - ww - pad_l, pad_t, oh - hh - pad_t))
        m = F.conv2d(dn, kernel, stride=wnd)[:, 0]  # [N, oh/w, ow/w]
        th = m + np.float32(-flt.threshold * gain_coef)
        if debug is not None:
            debug["th"] = th.clone()
            debug["pad"] = (pad_t, pad_l)
        cond = torch.sign(th) if flt.gain == 0 else torch.tanh(th)
        cond = _resize_asymmetric(cond, wnd)[:, pad_t:pad_t + hh, pad_l:pad_l + ww]
        cond = cond.unsqueeze(-1)
    c1, c2, c3 = np.float32(s / 2), np.float32(-s / 2), np.float32(1 - s / 2)
    mask = cond * c2 + c1
    mask2 = cond * c1 + c3
    return pw * mask + out * mask2
