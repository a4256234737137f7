"""Round-2 probe (GPU): the persistent flow-net kernel vs one launch per layer, and the stream
chunk size (JU_FLOW_SUBBATCH) at larger batches.  Prints one JSON line per configuration."""

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
                kernels = rt.info.kernels_per_frame
        groups = {o["name"][6:]: round(o["usec"], 1) for o in ops if o["name"].startswith("group:")}
        return np.stack(outs), groups, kernels, None
    except Exception as e:  # noqa: BLE001
        return None, None, None, repr(e)
    finally:
        for k in env:
            os.environ.pop(k, None)


def main():
    for preset, batch, envs in (
            ("psp_fast", 1, [{"JU_FUSED_FLOW": "1"}, {"JU_FUSED_FLOW": "1", "JU_TC_DUAL": "0"}]),
            ("psp_quality", 2, [{"JU_FUSED_FLOW": "1"}, {"JU_FUSED_FLOW": "1", "JU_FLOW_SUBBATCH": "1"}]),
            ("psp_quality", 16, [{"JU_FUSED_FLOW": "1", "JU_FLOW_SUBBATCH": "2"}, {"JU_FUSED_FLOW": "1", "JU_FLOW_SUBBATCH": "8"},
                                 {"JU_FUSED_FLOW": "1", "JU_FLOW_SUBBATCH": "0"}])):
        base, g0, k0, err = run(preset, batch, {})
        print(json.dumps({"preset": preset, "batch": batch, "env": {}, "kernels": k0, "groups": g0, "error": err}), flush=True)
        for env in envs:
            out, g, k, err = run(preset, batch, env)
            same = None if out is None or base is None else bool(np.array_equal(out, base))
            print(json.dumps({"preset": preset, "batch": batch, "env": env, "kernels": k, "groups": g,
                              "bit_identical": same, "error": err}), flush=True)


if __name__ == "__main__":
    main()
