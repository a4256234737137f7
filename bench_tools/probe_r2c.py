"""Round-2 probe (GPU): a trunk variant selected by environment switches vs the default trunk.

    python bench_tools/probe_r2c.py                      # JU_TRUNK_PAIR=1 (CTA pairs), with and without cooperative launch
    python bench_tools/probe_r2c.py JU_TRUNK_DIRECT=1    # any KEY=VALUE[,KEY=VALUE] sets, one run each
"""

import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from joshupscale_b200 import config as jcfg  # noqa: E402
from joshupscale_b200 import runtime as jrt  # noqa: E402
from joshupscale_b200 import synthetic  # noqa: E402
from joshupscale_b200 import weights as jw  # noqa: E402


def run(preset, batch, env):
    for k, v in env.items():
        os.environ[k] = v
    try:
        cfg = jcfg.preset(preset)
        w = jw.init_weights(cfg, 42, True)
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "m.jup")
            jw.save_model(path, cfg, w)
            frames = [synthetic.frames(cfg.frame_height, cfg.frame_width, 2, stream_id=s) for s in range(batch)]
            with jrt.Runtime(path, 0, batch) as rt:
                outs = [np.stack(rt.process_batch([f[t] for f in frames])) for t in range(2)]
                ops = rt.profile_ops(20)
        groups = {o["name"][6:]: round(o["usec"], 1) for o in ops if o["name"].startswith("group:")}
        return np.stack(outs), groups, None
    except Exception as e:  # noqa: BLE001
        return None, None, repr(e)
    finally:
        for k in env:
            os.environ.pop(k, None)


def main():
    sets = [dict(kv.split("=", 1) for kv in arg.split(",")) for arg in sys.argv[1:]]
    if not sets:
        sets = [{"JU_TRUNK_PAIR": "1"}, {"JU_TRUNK_PAIR": "1", "JU_TRUNK_COOP": "0"}]
    for preset, batch in (("small", 1), ("psp_fast", 1), ("psp_quality", 1), ("psp_quality", 2), ("psp_quality", 16)):
        base, g0, err = run(preset, batch, {})
        print(json.dumps({"preset": preset, "batch": batch, "env": {}, "groups": g0, "error": err}), flush=True)
        for env in sets:
            out, g, err = run(preset, batch, env)
            same = None if out is None or base is None else bool(np.array_equal(out, base))
            diff = None if out is None or base is None else int(np.abs(out.astype(np.int32) - base.astype(np.int32)).max())
            print(json.dumps({"preset": preset, "batch": batch, "env": env, "groups": g, "bit_identical": same,
                              "max_abs_diff_u8": diff, "error": err}), flush=True)


if __name__ == "__main__":
    main()
