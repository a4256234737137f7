"""Shared helpers for the GPU parity tests."""

import os

import numpy as np
import pytest
import torch

from joshupscale_b200 import config as jcfg
from joshupscale_b200 import runtime as jrt
from joshupscale_b200 import synthetic
from joshupscale_b200 import weights as jw
from oracle import reference_graph as og


def require_gpu():
    if not os.path.exists(jrt.library_path()):
        pytest.fail("libJoshUpscale.so is not built (run __graft_entry__.build())")
    if jrt.device_count() < 1:
        pytest.fail("no CUDA device visible: GPU tests cannot fall back to CPU")


def make_model(tmp_path, preset, seed=42, conditioned=True):
    cfg = jcfg.preset(preset) if isinstance(preset, str) else preset
    w = jw.init_weights(cfg, seed, conditioned)
    path = os.path.join(str(tmp_path), f"model_{seed}_{int(conditioned)}.jup")
    jw.save_model(path, cfg, w)
    return cfg, w, path


def r16(a):
    return np.asarray(a, np.float32).astype(np.float16).astype(np.float32)


def u8_stats(a, b):
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    return int(d.max()), float((d > 0).mean()), og.psnr_u8(a, b)
