"""CPU feasibility study for an FP8 (E4M3) ResBlock trunk.

Emulates, inside the oracle's fp16-storage graph, what a tcgen05 kind::f8f6f4 trunk would compute:
conv inputs and BN-folded weights of the generator's ResBlock convolutions quantised to E4M3
(per-tensor activation scale, per-output-channel weight scale, fp32 accumulate), residual stream
kept in fp16.  Reports max-abs / PSNR of the u8 output against the fp32 graph: the north-star gate
(max-abs <= 2, PSNR >= 45 dB) decides whether such a path could ever be enabled.

    python bench_tools/fp8_feasibility.py [preset] [frames] [conditioned 0/1]
"""

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from joshupscale_b200 import config as jcfg  # noqa: E402
from joshupscale_b200 import synthetic  # noqa: E402
from joshupscale_b200 import weights as jw  # noqa: E402
from oracle import reference_graph as og  # noqa: E402

E4M3_MAX = 448.0


def q8(x, scale):
    """fake-quantise: x / scale -> e4m3 -> * scale (scale broadcastable)."""
    y = torch.clamp(x / scale, -E4M3_MAX, E4M3_MAX).to(torch.float8_e4m3fn).to(torch.float32)
    return y * scale


class Fp8TrunkGraph(og.Graph):
    def __init__(self, cfg, weights, act_headroom=2.0, quant_second_conv_only=False):
        super().__init__(cfg, weights, "fp16emu")
        self.headroom = act_headroom
        self.second_only = quant_second_conv_only

    def _conv8(self, x, conv, bn, act, residual=None):
        k = self.w[f"{conv}/kernel"]
        gamma, beta, mean, var = self._bn_params(bn)
        s = gamma / torch.sqrt(var + og.BN_EPS)
        kf = k * s  # (kh, kw, cin, cout)
        wscale = kf.abs().amax(dim=(0, 1, 2), keepdim=True).clamp_min(1e-12) / E4M3_MAX
        kq = q8(kf, wscale)
        ascale = x.abs().max().clamp_min(1e-12) * self.headroom / E4M3_MAX
        xq = q8(x, ascale)
        y = og.conv2d_same(xq, kq) + (beta - mean * s)
        if residual is not None:
            y = y + residual
        y = og.activation(y, act)
        return og.r16(y)

    def _res_block(self, x, prefix, act):
        if not prefix.startswith("generator/"):
            return super()._res_block(x, prefix, act)
        if self.second_only:
            h = self._conv_bn_act(x, f"{prefix}/conv_1", f"{prefix}/bn_1", act)
        else:
            h = self._conv8(x, f"{prefix}/conv_1", f"{prefix}/bn_1", act)
        return self._conv8(h, f"{prefix}/conv_2", f"{prefix}/bn_2", act, residual=x)


def main():
    preset = sys.argv[1] if len(sys.argv) > 1 else "psp_quality"
    nframes = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    conditioned = bool(int(sys.argv[3])) if len(sys.argv) > 3 else True
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = jcfg.preset(preset)
    w = jw.init_weights(cfg, 42, conditioned)
    frames = synthetic.frames(cfg.frame_height, cfg.frame_width, nframes)
    ref, _ = og.Graph(cfg, w, "fp32").run(frames)
    rows = []
    for name, g in (("fp16 storage (shipping contract)", og.Graph(cfg, w, "fp16emu")),
                    ("fp8 e4m3 trunk convs (both convs of every block)", Fp8TrunkGraph(cfg, w)),
                    ("fp8 e4m3, conv_2 of every block only", Fp8TrunkGraph(cfg, w, quant_second_conv_only=True))):
        out, _ = g.run(frames)
        worst_abs, worst_psnr = 0, 1e9
        for t in range(nframes):
            d = np.abs(out[t, ..., :3].astype(int) - ref[t, ..., :3].astype(int))
            worst_abs = max(worst_abs, int(d.max()))
            worst_psnr = min(worst_psnr, og.psnr_u8(out[t, ..., :3], ref[t, ..., :3]))
        rows.append((name, worst_abs, worst_psnr))
        print(f"{preset} set {'B' if conditioned else 'A'}, {nframes} frames | {name}: max-abs {worst_abs}, "
              f"min PSNR {worst_psnr:.2f} dB -> gate (<=2, >=45 dB) "
              f"{'PASS' if worst_abs <= 2 and worst_psnr >= 45 else 'FAIL'}", flush=True)


if __name__ == "__main__":
    main()
