"""Hardware probe for the tcgen05 convolution: runs every halo/descriptor
variant of conv_tc.cu in its own subprocess (a wrong UMMA descriptor can trap
the context) and records correctness vs the oracle plus kernel timings.

    python bench_tools/tc_probe.py            # on the GPU box; writes gpurun_out/tc_probe.json
"""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [
    # b, h, w, cin, cout, ks, mode
    (1, 16, 8, 64, 64, 3, "plain"),
    (1, 20, 13, 64, 64, 3, "plain"),
    (2, 48, 40, 64, 64, 3, "residual_relu"),
    (1, 16, 16, 128, 64, 3, "bias_relu"),
    (1, 32, 24, 64, 128, 3, "bias_relu"),
    (1, 16, 16, 64, 32, 3, "f32"),
    (1, 16, 24, 256, 256, 3, "bias_relu"),
    (1, 16, 8, 64, 128, 1, "shuffle"),
    (1, 270, 480, 64, 64, 3, "residual_relu"),
]


def child(variant: int, epi: int = 1, pdl: int = 1):
    import numpy as np
    import torch
    from joshupscale_b200 import kernels as jk
    from joshupscale_b200 import runtime as jrt
    from oracle import reference_graph as og
    jrt.set_option("tc_variant", variant)
    jrt.set_option("tc_tma_epilogue", epi)
    jrt.set_option("tc_pdl", pdl)
    results = []
    for (b, h, w, cin, cout, ks, mode) in CASES:
        rng = np.random.default_rng(cin + cout + h)
        x = (rng.standard_normal((b, h, w, cin)) * 0.5).astype(np.float16).astype(np.float32)
        k = (rng.standard_normal((ks, ks, cin, cout)) / np.sqrt(ks * ks * cin)).astype(np.float32)
        bias = (rng.standard_normal(cout) * 0.1).astype(np.float32)
        kw = {}
        want = og.conv2d_same(torch.from_numpy(x), og.r16(torch.from_numpy(k))).numpy()
        if mode == "bias_relu":
            kw.update(bias=bias, act=jk.ACT_RELU)
            want = np.maximum(want + bias, 0)
        elif mode == "residual_relu":
            res = (rng.standard_normal((b, h, w, cout)) * 0.5).astype(np.float16).astype(np.float32)
            kw.update(bias=bias, residual=res, act=jk.ACT_RELU)
            want = np.maximum(want + bias + res, 0)
        elif mode == "f32":
            kw.update(bias=bias, out_f32=True)
            want = want + bias
        elif mode == "shuffle":
            kw.update(bias=bias, act=jk.ACT_RELU, shuffle2=True)
            want = np.maximum(want + bias, 0)
            cpp = cout // 4
            want = want.reshape(b, h, w, 2, 2, cpp).transpose(0, 1, 3, 2, 4, 5).reshape(b, 2 * h, 2 * w, cpp)
        try:
            got = jk.conv(x, k, impl=jk.IMPL_TCGEN05, **kw).astype(np.float32)
            err = float(np.abs(got - want).max())
            ok = bool(err < 5e-3)
        except Exception as e:  # noqa: BLE001
            err, ok = str(e)[:200], False
        results.append(dict(case=[b, h, w, cin, cout, ks, mode], max_err=err, ok=ok))
        print(json.dumps(results[-1]), flush=True)
        if not ok and isinstance(err, str):
            break
    timings = {}
    if all(r["ok"] for r in results):
        for name, args in {"res64_270x480": (1, 270, 480, 64, 64, 3, True),
                           "res64_b16": (16, 270, 480, 64, 64, 3, True),
                           "flow256_34x60": (1, 34, 60, 256, 256, 3, False),
                           "flow128to64_136x240": (1, 136, 240, 128, 64, 3, False),
                           "ct1_1x1": (1, 270, 480, 64, 128, 1, False)}.items():
            us = jrt.bench_conv(1, *args[:6], residual=args[6], iters=50)
            flops = 2.0 * args[0] * args[1] * args[2] * args[3] * args[4] * args[5] ** 2
            timings[name] = dict(usec=us, tflops=flops / us / 1e6)
        print(json.dumps(dict(timings=timings)), flush=True)
    return results, timings


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
        return
    out = {}
    for v, epi, pdl in ((0, 0, 0), (0, 1, 0), (0, 1, 1), (0, 0, 1), (1, 1, 1)):
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(v), str(epi), str(pdl)],
                           capture_output=True, text=True, timeout=300)
        lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
        key = f"variant{v}_epi{epi}_pdl{pdl}"
        out[key] = dict(rc=r.returncode, lines=lines, stderr=r.stderr[-600:])
        print(key, "rc", r.returncode, [l.get("ok", l) for l in lines], r.stderr[-300:], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "tc_probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
