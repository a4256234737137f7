"""Round-2 probe (GPU): what-if timing of the stacked-rows trunk (JU_TRUNK_TS=1).  Each run switches parts of
the per-unit work off (JU_TRUNK_DEBUG_SKIP bits; results are garbage, only the time is read):
1 = no dependency polls, 2 = no TMA stores / publish, 4 = no halo loads, 8 = no shortcut, 16 = no MMAs,
32 = no weight writes, 64 = no epilogue math / stmatrix, 128 = no tcgen05.ld."""

import json
import sys

import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from probe_r2c import run  # noqa: E402


def main():
    import os
    batch = int(os.environ.get("PROBE_BATCH", "1"))
    masks = [int(a) for a in sys.argv[1:]] or [0, 1, 2, 4, 8, 16, 32, 64, 192, 255, 255 - 16]
    for m in masks:
        env = {"JU_TRUNK_TS": "1", "JU_TRUNK_SUBBATCH": str(batch)}
        if m:
            env["JU_TRUNK_DEBUG_SKIP"] = str(m)
        _, g, err = run("psp_quality", batch, env)
        print(json.dumps({"batch": batch, "skip": m, "resblocks_us": g and g.get("resblocks"), "error": err}), flush=True)


if __name__ == "__main__":
    main()
