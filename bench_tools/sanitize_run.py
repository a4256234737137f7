"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool memcheck python bench_tools/sanitize_run.py [preset] [frames] [batch]
Runs a few recurrent frames (plain and with the output filter) through the runtime entry point."""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from joshupscale_b200 import config as jcfg, runtime as jrt, synthetic, weights as jw  # noqa: E402


def main():
    preset = sys.argv[1] if len(sys.argv) > 1 else "small"
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    batch = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    cfg = jcfg.preset(preset)
    w = jw.init_weights(cfg, 42, True)
    with tempfile.TemporaryDirectory() as d:
        for tag, flt in (("plain", None), ("filter", jcfg.OutputFilter(window=16, threshold=0.2))):
            path = os.path.join(d, f"{tag}.jup")
            jw.save_model(path, cfg, jw.with_output_filter(w, flt))
            clips = [synthetic.frames(cfg.frame_height, cfg.frame_width, frames, stream_id=s) for s in range(batch)]
            with jrt.Runtime(path, 0, batch) as rt:
                acc = 0
                for t in range(frames):
                    outs = rt.process_batch([c[t] for c in clips])
                    acc += int(np.sum(outs[0][::7, ::7, :3]))
            print(tag, "checksum", acc)


if __name__ == "__main__":
    main()
