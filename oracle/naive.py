"""ORACLE (second, independent restatement) - naive numpy versions of every
primitive, written from the formulas in SURVEY.md appendix A with explicit
index loops / einsum, sharing no code with oracle/reference_graph.py.

TEST INFRASTRUCTURE ONLY (see oracle/reference_graph.py header).  Used on
tiny shapes to cross-validate the torch-based oracle, because TensorFlow
(the reference's arithmetic provider) cannot be run here.
"""

from __future__ import annotations

import numpy as np


def conv2d_same(x, k, bias=None):
    """x [H,W,Cin], k [kh,kw,Cin,Cout]; y[h,w,o] = sum x[h+i-p, w+j-p, c] k[i,j,c,o]."""
    kh, kw, cin, cout = k.shape
    ph, pw = (kh - 1) // 2, (kw - 1) // 2
    h, w, _ = x.shape
    xp = np.zeros((h + kh - 1, w + kw - 1, cin), np.float64)
    xp[ph:ph + h, pw:pw + w] = x
    y = np.zeros((h, w, cout), np.float64)
    for i in range(kh):
        for j in range(kw):
            y += np.einsum("hwc,co->hwo", xp[i:i + h, j:j + w], k[i, j].astype(np.float64))
    if bias is not None:
        y += bias
    return y.astype(np.float32)


def batch_norm(x, gamma, beta, mean, var, eps=1e-3):
    return (gamma * (x - mean) / np.sqrt(var + eps) + beta).astype(np.float32)


def max_pool2(x):
    h, w, c = x.shape
    out = np.empty((h // 2, w // 2, c), x.dtype)
    for i in range(h // 2):
        for j in range(w // 2):
            out[i, j] = x[2 * i:2 * i + 2, 2 * j:2 * j + 2].reshape(4, c).max(axis=0)
    return out


def resize_bilinear_legacy(x, s):
    h, w, c = x.shape
    out = np.empty((h * s, w * s, c), np.float32)
    for oy in range(h * s):
        sy = np.float32(oy) / np.float32(s)
        y0 = int(np.floor(sy))
        y1 = min(y0 + 1, h - 1)
        ty = np.float32(sy - y0)
        for ox in range(w * s):
            sx = np.float32(ox) / np.float32(s)
            x0 = int(np.floor(sx))
            x1 = min(x0 + 1, w - 1)
            tx = np.float32(sx - x0)
            top = x[y0, x0] + (x[y0, x1] - x[y0, x0]) * tx
            bot = x[y1, x0] + (x[y1, x1] - x[y1, x0]) * tx
            out[oy, ox] = top + (bot - top) * ty
    return out


def depth_to_space(x, b):
    h, w, c = x.shape
    cp = c // (b * b)
    out = np.empty((h * b, w * b, cp), x.dtype)
    for i in range(b):
        for j in range(b):
            out[i::b, j::b] = x[:, :, (i * b + j) * cp:(i * b + j + 1) * cp]
    return out


def space_to_depth(x, b):
    h, w, c = x.shape
    out = np.empty((h // b, w // b, b * b * c), x.dtype)
    for i in range(b):
        for j in range(b):
            out[:, :, (i * b + j) * c:(i * b + j + 1) * c] = x[i::b, j::b]
    return out


def conv2d_transpose_k2s2(x, k, bias=None):
    """x [H,W,Cin], k [2,2,Cout,Cin]; out[2h+i, 2w+j, o] = sum_c x[h,w,c] k[i,j,o,c]."""
    h, w, _ = x.shape
    cout = k.shape[2]
    out = np.zeros((2 * h, 2 * w, cout), np.float64)
    for i in range(2):
        for j in range(2):
            out[i::2, j::2] = np.einsum("hwc,oc->hwo", x.astype(np.float64),
                                        k[i, j].astype(np.float64))
    if bias is not None:
        out += bias
    return out.astype(np.float32)


def dense_image_warp(img, flow):
    """img [H,W,C] float32, flow [H,W,2] float32 (dy, dx)."""
    h, w, c = img.shape
    out = np.empty_like(img)
    for y in range(h):
        for x in range(w):
            qy = np.float32(y) - flow[y, x, 0]
            qx = np.float32(x) - flow[y, x, 1]
            fy = np.float32(min(max(0.0, np.floor(qy)), h - 2))
            fx = np.float32(min(max(0.0, np.floor(qx)), w - 2))
            ay = np.float32(min(max(np.float32(0), np.float32(qy - fy)), np.float32(1)))
            ax = np.float32(min(max(np.float32(0), np.float32(qx - fx)), np.float32(1)))
            iy, ix = int(fy), int(fx)
            tl, tr = img[iy, ix], img[iy, ix + 1]
            bl, br = img[iy + 1, ix], img[iy + 1, ix + 1]
            top = ax * (tr - tl) + tl
            bot = ax * (br - bl) + bl
            out[y, x] = ay * (bot - top) + top
    return out
