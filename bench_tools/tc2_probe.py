"""Probe of the CTA-pair (cta_group::2) convolution: correctness vs the oracle
and timing, in a subprocess (a protocol bug traps the context).
    python bench_tools/tc2_probe.py
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [(1, 16, 8, "plain"), (1, 32, 16, "plain"), (1, 20, 13, "bias_relu"), (2, 48, 40, "residual_relu"),
         (1, 270, 480, "residual_relu")]


def child():
    import numpy as np
    import torch
    from joshupscale_b200 import kernels as jk
    from joshupscale_b200 import runtime as jrt
    from oracle import reference_graph as og
    for (b, h, w, mode) in CASES:
        rng = np.random.default_rng(h * w)
        x = (rng.standard_normal((b, h, w, 64)) * 0.5).astype(np.float16).astype(np.float32)
        k = (rng.standard_normal((3, 3, 64, 64)) / 24).astype(np.float32)
        bias = (rng.standard_normal(64) * 0.1).astype(np.float32)
        kw = {}
        want = og.conv2d_same(torch.from_numpy(x), og.r16(torch.from_numpy(k))).numpy()
        if mode == "bias_relu":
            kw.update(bias=bias, act=jk.ACT_RELU)
            want = np.maximum(want + bias, 0)
        elif mode == "residual_relu":
            res = (rng.standard_normal((b, h, w, 64)) * 0.5).astype(np.float16).astype(np.float32)
            kw.update(bias=bias, residual=res, act=jk.ACT_RELU)
            want = np.maximum(want + bias + res, 0)
        got = jk.conv(x, k, impl=jk.IMPL_TCGEN05_2CTA, **kw).astype(np.float32)
        err = float(np.abs(got - want).max())
        print(json.dumps(dict(case=[b, h, w, mode], max_err=err, ok=bool(err < 5e-3))), flush=True)
    for name, args in {"res64_b1": (1, 270, 480), "res64_b16": (16, 270, 480)}.items():
        t2 = jrt.bench_conv(2, args[0], args[1], args[2], 64, 64, 3, residual=True, iters=50)
        t1 = jrt.bench_conv(1, args[0], args[1], args[2], 64, 64, 3, residual=True, iters=50)
        fl = 2.0 * args[0] * args[1] * args[2] * 64 * 64 * 9
        print(json.dumps({name: dict(usec_2cta=t2, usec_1cta=t1, tflops_2cta=fl / t2 / 1e6, tflops_1cta=fl / t1 / 1e6)}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child()
    else:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], capture_output=True, text=True, timeout=300)
        print(r.stdout[-3000:])
        print("rc", r.returncode, r.stderr[-1500:])
