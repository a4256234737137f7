"""Launch one convolution shape repeatedly (for ncu / quick timing).
    python bench_tools/conv_micro.py <impl> <batch> <h> <w> <cin> <cout> <ks> <residual> [iters]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from joshupscale_b200 import runtime as jrt  # noqa: E402

a = [int(x) for x in sys.argv[1:9]]
iters = int(sys.argv[9]) if len(sys.argv) > 9 else 20
us = jrt.bench_conv(a[0], a[1], a[2], a[3], a[4], a[5], a[6], residual=bool(a[7]), iters=iters)
flops = 2.0 * a[1] * a[2] * a[3] * a[4] * a[5] * a[6] ** 2
print(f"{us:.2f} us  {flops / us / 1e6:.1f} TFLOP/s")
