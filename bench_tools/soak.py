"""Soak / determinism run: many recurrent frames back to back, twice, comparing output checksums.
    python bench_tools/soak.py [preset] [streams] [frames]
A protocol bug in the persistent kernels would show up as a trap (bounded waits) or as a checksum
that differs between the two passes."""
import os
import sys
import tempfile
import time
import zlib

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from joshupscale_b200 import config as jcfg, runtime as jrt, synthetic, weights as jw  # noqa: E402


def one_pass(path, cfg, streams, frames, clips):
    sums = []
    with jrt.Runtime(path, 0, streams) as rt:
        outs = [np.empty(rt.out_shape, np.uint8) for _ in range(streams)]
        t0 = time.perf_counter()
        for t in range(frames):
            rt.process_batch([c[t % len(c)] for c in clips], outs)
            if (t + 1) % max(frames // 8, 1) == 0:
                sums.append(zlib.crc32(b"".join(o.tobytes() for o in outs)))
        dt = time.perf_counter() - t0
    return sums, frames * streams / dt


def main():
    preset = sys.argv[1] if len(sys.argv) > 1 else "psp_fast"
    streams = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    frames = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
    cfg = jcfg.preset(preset)
    w = jw.init_weights(cfg, 42, True)
    clips = [synthetic.frames(cfg.frame_height, cfg.frame_width, 16, stream_id=s) for s in range(streams)]
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "m.jup")
        jw.save_model(path, cfg, w)
        a, fps_a = one_pass(path, cfg, streams, frames, clips)
        b, fps_b = one_pass(path, cfg, streams, frames, clips)
    ok = a == b
    print(f"soak {preset} x{streams}: {frames} frames twice, {fps_a:.0f} / {fps_b:.0f} fps through host buffers, "
          f"checksums {'identical' if ok else 'DIFFER'}: {[hex(x) for x in a]}")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
