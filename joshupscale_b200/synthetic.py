"""Seeded synthetic BGRX frame sequences.

The reference has no dataset or fixture; these are the frames every parity
test, the bench and the CPU baseline use.  Pixel format is the one the
plugins hand to processImage: 4 bytes per pixel B,G,R,X
(core/public/JoshUpscale/core.h:32-38, avisynth_plugin/src/main.cc:125-142).
The X byte is filled with 255 on purpose: it must be ignored on input.
"""

from __future__ import annotations

import numpy as np


def _box_blur5(img: np.ndarray) -> np.ndarray:
    acc = np.zeros_like(img, dtype=np.float32)
    for dy in range(-2, 3):
        for dx in range(-2, 3):
            acc += np.roll(np.roll(img, dy, axis=0), dx, axis=1)
    return acc / 25.0


def base_texture(stream_id: int = 0, size: int = 1024) -> np.ndarray:
    rng = np.random.default_rng(1234 + stream_id)
    noise = rng.integers(0, 256, size=(size, size, 3)).astype(np.float32)
    tex = _box_blur5(noise)
    # stretch the blurred noise back to the full u8 range
    tex = (tex - tex.mean()) * 4.0 + 127.5
    return np.clip(tex, 0, 255).astype(np.uint8)


def frames(height: int, width: int, count: int, stream_id: int = 0,
           kind: str = "pan") -> np.ndarray:
    """Return [count, height, width, 4] u8 BGRX frames.

    kind: "pan" (moving crop + noise), "black", "white", "checker",
          "cut" (hard scene cut at count//2).
    """
    out = np.empty((count, height, width, 4), np.uint8)
    out[..., 3] = 255
    if kind == "black":
        out[..., :3] = 0
        return out
    if kind == "white":
        out[..., :3] = 255
        return out
    if kind == "checker":
        yy, xx = np.mgrid[0:height, 0:width]
        for t in range(count):
            out[t, ..., :3] = ((((yy + t) // 4 + xx // 4) & 1) * 255)[..., None]
        return out
    rng = np.random.default_rng(99 + stream_id)
    tex = base_texture(stream_id)
    size = tex.shape[0]
    y0 = int(rng.integers(0, size // 4))
    x0 = int(rng.integers(0, size // 4))
    for t in range(count):
        if kind == "cut" and t == count // 2:
            tex = base_texture(stream_id + 1000)
        yy = (y0 + t + np.arange(height)) % size
        xx = (x0 + 2 * t + np.arange(width)) % size
        crop = tex[yy][:, xx].astype(np.float32)
        crop += rng.normal(0, 2.0, crop.shape)
        out[t, ..., :3] = np.clip(crop, 0, 255).astype(np.uint8)
    return out
