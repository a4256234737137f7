#!/usr/bin/env python
"""Headline benchmark: 1080p output frames/s of the per-frame recurrent
upscaling path (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload psp_fast_b1]
    python bench.py --impl reference ...      # CPU restatement of the reference graph

A "step" advances every stream of the workload by one frame through the public
entry point (ju_process_batch == the C-ABI twin of Runtime::processImage).
Default workload = BASELINE.json configs[1]: PSP fast generator, 1 stream,
batch 1 (the latency path).  Workload names are <preset>_b<streams per GPU>.

  value : frames/s with the input frames already resident in HBM
          (DataLocation::CUDA images), per-step CUDA-event timing, L2 flushed
          between steps (the per-frame working set is smaller than the L2)
  e2e   : the same through HOST (pinned) BGRX buffers - H2D of the frame and
          D2H of the upscaled frame inside the timed region
  roofline : dominant kernel (ResBlock 3x3 conv) algorithmic FLOP/s over the
          measured bf16 peak, timed live with CUDA events on the engine stream
  cpu_baseline : the CPU restatement of the reference graph (TensorFlow is not
          installable offline), timed on this box's host cores on a bounded sample

Multi-GPU: streams are sharded across ranks (weak scaling, no collective on the
frame path); NCCL is only used for the barrier and the max-over-ranks time.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from joshupscale_b200 import config as jcfg  # noqa: E402
from joshupscale_b200 import sharding  # noqa: E402
from joshupscale_b200 import synthetic  # noqa: E402
from joshupscale_b200 import weights as jw  # noqa: E402

METRIC = "1080p_output_frames_per_sec"
UNIT = "frames/s"
FRAME_POOL = 24  # distinct synthetic frames per stream, cycled


def parse_workload(name: str):
    preset, _, b = name.rpartition("_b")
    return jcfg.preset(preset), int(b), preset


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tensor_burst=p["bf16_tflops"],
                    tensor_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sustained=1400.0, source="fallback")


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: NVML polled every 5 ms from a
    thread (the recipe's nvidia-smi -lms 200 line yields too few samples for a 0.1 s region);
    falls back to the nvidia-smi loop when pynvml is unavailable."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap,power.draw")
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
               "sw_power_cap": 0x4}  # nvmlClocksEventReason* bit masks

    def __init__(self, device: int):
        self.samples = []  # (sm_mhz, reasons bitmask, power_w)
        self.max_mhz = None
        self.lines = []
        self.proc = None
        self.stop_flag = threading.Event()
        self.nvml = None
        try:
            import pynvml  # noqa: PLC0415
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            index = int(visible.split(",")[device]) if visible and visible.split(",")[device].isdigit() else device
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:  # noqa: BLE001 - any NVML problem: use the nvidia-smi loop
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "200", "-i", str(device)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag.is_set():
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                try:
                    mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:  # noqa: BLE001 - older binding name
                    mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                try:
                    power = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                except Exception:  # noqa: BLE001
                    power = None
                self.samples.append((mhz, mask, power))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            sm = [s[0] for s in self.samples]
            mask = 0
            for s in self.samples:
                mask |= s[1]
            powers = [s[2] for s in self.samples if s[2] is not None]
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_min_mhz": min(sm) if sm else None,
                    "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(k for k, bit in self.REASONS.items() if mask & bit),
                    "power_w_max": max(powers) if powers else None,
                    "samples": len(sm), "source": "nvml, 5 ms period over both timed regions"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi -lms 200"}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ---------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle (port of the reference graph)
# ---------------------------------------------------------------------------

def cpu_reference_fps(cfg, streams, weights, max_steps, warmup, budget_s):
    import torch
    from oracle import reference_graph as og
    torch.set_num_threads(os.cpu_count() or 1)
    g = og.Graph(cfg, weights, "fp32")
    frames = np.stack([synthetic.frames(cfg.frame_height, cfg.frame_width, 4, stream_id=s)
                       for s in range(streams)], axis=1)  # [T, S, H, W, 4]
    state = g.zero_state(streams)
    for t in range(warmup):
        _, state, _ = g.step(frames[t % 4], state)
    times = []
    start = time.perf_counter()
    for t in range(max_steps):
        t0 = time.perf_counter()
        _, state, _ = g.step(frames[t % 4], state)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - start > budget_s:
            break
    total = sum(times)
    return dict(fps=streams * len(times) / total, steps=len(times), seconds=total,
                cores=torch.get_num_threads())


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    cfg, streams, preset = parse_workload(args.workload)
    weights = jw.init_weights(cfg, 42, True)
    r = cpu_reference_fps(cfg, streams, weights, args.steps, min(args.warmup, 2), budget_s=150.0)
    sample = (f"{r['steps']} of {args.steps} requested steps ({r['seconds']:.1f} s), {streams} stream(s), "
              "fp32 torch-CPU restatement of the reference Keras graph (TensorFlow unavailable offline)")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["fps"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": min(args.warmup, 2),
        "ms_per_step": 1000.0 * r["seconds"] / r["steps"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": args.workload, "preset": preset, "streams": streams,
                   "frame": [cfg.frame_width, cfg.frame_height]},
        "cpu_baseline": {"value": r["fps"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                         "sample": sample},
        "e2e": {"value": r["fps"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------

def run_ours(args):
    import ctypes as C

    from joshupscale_b200 import kernels as jk
    from joshupscale_b200 import runtime as jrt

    rank, world, local = dist_env()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    use_dist = world > 1
    torch = None
    dist = None
    if use_dist:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = jrt.load_library()
    if lib.ju_device_count() <= local:
        raise SystemExit("no CUDA device for this rank; there is no CPU fallback")

    cfg, streams, preset = parse_workload(args.workload)
    weights = jw.init_weights(cfg, 42, True)
    tmp = tempfile.mkdtemp(prefix="jubench_")
    model_path = os.path.join(tmp, f"{preset}.jup")
    jw.save_model(model_path, cfg, weights)

    # the runtime sets/restores its device per call; the ju_dev_*/ju_timer_* helpers
    # below act on the calling thread's current device
    jrt._check(lib.ju_set_device(local))
    rt = jrt.Runtime(model_path, local, streams)
    h, w = cfg.frame_height, cfg.frame_width
    in_bytes, out_bytes = h * w * 4, 16 * h * w * 4

    # global stream ids owned by this rank (stream s lives on rank s mod world)
    my_streams = sharding.streams_for_rank(streams * world, world, rank)
    pool = [synthetic.frames(h, w, FRAME_POOL, stream_id=sid) for sid in my_streams]

    # --- device-resident inputs / outputs (value) ---
    d_in = [[jk.to_device(pool[s][t]) for s in range(streams)] for t in range(FRAME_POOL)]
    d_out = [jk.DeviceArray((4 * h, 4 * w, 4), np.uint8) for _ in range(streams)]
    dev_imgs_in = [[jrt.JuImage(d_in[t][s].ptr, jrt.LOC_CUDA, w * 4, w, h) for s in range(streams)]
                   for t in range(FRAME_POOL)]
    dev_imgs_out = [jrt.JuImage(d_out[s].ptr, jrt.LOC_CUDA, w * 16, 4 * w, 4 * h) for s in range(streams)]

    # --- pinned host buffers (e2e) ---
    host_in, host_out = [], []
    for s in range(streams):
        p_in, p_out = C.c_void_p(), C.c_void_p()
        jrt._check(lib.ju_host_alloc(C.byref(p_in), in_bytes * FRAME_POOL))
        jrt._check(lib.ju_host_alloc(C.byref(p_out), out_bytes))
        arr = np.ctypeslib.as_array(C.cast(p_in, C.POINTER(C.c_uint8)), shape=(FRAME_POOL, h, w, 4))
        arr[...] = pool[s]
        host_in.append((p_in, arr))
        host_out.append(p_out)
    host_imgs_in = [[jrt.JuImage(host_in[s][0].value + t * in_bytes, jrt.LOC_CPU, w * 4, w, h)
                     for s in range(streams)] for t in range(FRAME_POOL)]
    host_imgs_out = [jrt.JuImage(host_out[s].value, jrt.LOC_CPU, w * 16, 4 * w, 4 * h)
                     for s in range(streams)]

    def barrier():
        if use_dist:
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
        else:
            jrt._check(lib.ju_dev_sync())

    usec = C.c_double()

    def timed_steps(imgs_in, imgs_out, steps, warmup, flush):
        for t in range(warmup):
            rt.process_images(imgs_in[t % FRAME_POOL], imgs_out)
        barrier()
        per_step = []
        for t in range(steps):
            if flush:
                jrt._check(lib.ju_l2_flush())
            jrt._check(lib.ju_timer_begin())
            rt.process_images(imgs_in[(warmup + t) % FRAME_POOL], imgs_out)
            jrt._check(lib.ju_timer_end(C.byref(usec)))
            per_step.append(usec.value)
        barrier()
        return per_step

    def global_max(x):
        return sharding.max_over_ranks(x, dist if use_dist else None, "cuda" if use_dist else None)

    sampler = ClockSampler(local) if rank == 0 else None
    steps, warmup = args.steps, max(args.warmup, 3)
    t_dev = timed_steps(dev_imgs_in, dev_imgs_out, steps, warmup, flush=True)
    t_e2e = timed_steps(host_imgs_in, host_imgs_out, steps, warmup, flush=True)
    clocks = sampler.stop() if sampler else None

    # steady state: back-to-back frames, L2 warm (deployment regime), wall clock
    barrier()
    t0 = time.perf_counter()
    for t in range(steps):
        rt.process_images(dev_imgs_in[t % FRAME_POOL], dev_imgs_out)
    barrier()
    steady_s = time.perf_counter() - t0

    total_dev = global_max(sum(t_dev) * 1e-6)
    total_e2e = global_max(sum(t_e2e) * 1e-6)
    steady_s = global_max(steady_s)
    n_streams_total = streams * world

    if rank == 0:
        peaks = load_peaks()
        all_ops = rt.profile_ops(10)
        rt.reset_state()
        ops = [o for o in all_ops if not o["name"].startswith("group:") and not o["name"].startswith("sync:")]
        groups = {o["name"][6:]: o for o in all_ops if o["name"].startswith("group:")}
        # dominant kernel: the generator's ResBlock convs.  Launch duration =
        # CUDA-event time around the back-to-back run of all ResBlock launches
        # (as replayed by the graph) / number of launches.
        res = [o for o in ops if o["name"].startswith("generator/block_")]
        grp = groups["resblocks"]
        res_usec = grp["usec"]
        frame_usec = sum(g["usec"] for g in groups.values())
        layers = grp["launches"]                      # ResBlock conv layers per frame
        flops_per_layer = grp["flops"] / layers       # 9.555 GFLOP at PSP batch 1
        bytes_per_layer = grp["bytes"] / layers
        mean_usec = res_usec / layers
        achieved = grp["flops"] / (res_usec * 1e-6) / 1e12
        peak = peaks["tensor_sustained"]
        persistent = any("(persistent)" in o["name"] for o in res)
        # the persistent trunk runs as ceil(streams / chunk) launches of `chunk` streams (engine.cc)
        chunk = int(os.environ.get("JU_TRUNK_SUBBATCH", "-1")) if persistent else 0
        if chunk < 0:  # engine default: streams whose three trunk tensors fit 85 % of the 126 MB L2
            chunk = max(1, int(0.85 * 126 * 2 ** 20 / (3 * h * w * 64 * 2)))
        if chunk <= 0 or chunk > streams:
            chunk = streams
        trunk_launches = len(res)  # one op per sub-batch launch (persistent) or per layer
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath) and (h, w) == (270, 480):
            tj = json.load(open(tpath))
            if persistent:
                per_layer = tj.get(f"trunk_df_per_layer_{chunk}_streams")
                traffic = per_layer * layers if per_layer else None
            else:
                traffic = tj.get(f"batch{streams}_270x480")
        roofline = {
            "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak, "traffic": traffic,
            "traffic_note": "dram bytes per launch (all layers, %d stream(s)) from ncu --set full of this kernel inside "
                            "bench.py (profiles/ncu_traffic.json); algorithmic bytes per launch = %d"
                            % (chunk if persistent else streams, int(bytes_per_layer * layers / max(trunk_launches, 1))),
            "kernel": ("trunk_df_tc_kernel: generator conv_1 + all ResBlock conv3x3 64->64 layers of %d stream(s) in one "
                       "persistent launch" % chunk
                       if persistent else "conv_tc_kernel<3,1>: ResBlock conv3x3 64->64 (generator/block_*/conv_*)"),
            "layers_per_step": layers, "kernel_launches_per_step": trunk_launches, "usec_per_layer": mean_usec,
            "usec_per_launch": res_usec / trunk_launches,
            "flops_per_layer": flops_per_layer, "share_of_step": res_usec / frame_usec,
            "frac_of_burst_peak": achieved / peaks["tensor_burst"],
            "peak_source": f"{peaks['source']} bf16 dense, sustained (burst {peaks['tensor_burst']})",
        }
        hbm_ops = {}
        for o in ops:  # ops that run once per sub-batch (tail, filter) are summed per name
            if not o["tensor_bound"]:
                acc = hbm_ops.setdefault(o["name"], {"usec": 0.0, "bytes": 0.0})
                acc["usec"] += o["usec"]
                acc["bytes"] += o["bytes"]
        for name, acc in hbm_ops.items():
            gbs = acc["bytes"] / (acc["usec"] * 1e-6) / 1e9 if acc["usec"] > 0 else 0.0
            hbm_ops[name] = {"usec": acc["usec"], "gbs": gbs, "frac_of_hbm_peak": gbs / peaks["hbm"]}
        flow_usec = groups["flow"]["usec"]
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"ops_{args.workload}.json"), "w") as f:
            json.dump(all_ops, f, indent=1)

        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_fps(cfg, 1, weights, 12, 1, budget_s=20.0)
            cpu = {"value": r["fps"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                   "sample": f"{r['steps']} frames of 1 {preset} stream ({r['seconds']:.1f} s), fp32 torch-CPU "
                             "restatement of the reference Keras graph (TensorFlow unavailable offline)"}

        info = rt.info
        fps = sharding.aggregate_fps(streams, world, steps, total_dev)
        line = {
            "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": 1000.0 * total_dev / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
            "config": {
                "workload": args.workload, "preset": preset, "streams_per_gpu": streams,
                "frame": [w, h], "output": [4 * w, 4 * h], "gen_blocks": cfg.gen_blocks,
                "flow_arch": cfg.flow_arch, "weights": "seeded random-init (set B, conditioned)",
                "l2": "flushed between steps (256 MiB memset, untimed)",
                "gflop_per_frame": info.gflop_per_frame,
            },
            "e2e": {"value": n_streams_total * steps / total_e2e, "unit": UNIT,
                    "h2d_bytes_per_step": streams * in_bytes, "d2h_bytes_per_step": streams * out_bytes},
            "gpu_launches": int(info.kernels_per_frame) * steps,
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "latency_ms": {"p50": statistics.median(t_dev) / 1000.0,
                           "p95": sorted(t_dev)[int(0.95 * (len(t_dev) - 1))] / 1000.0,
                           "p50_e2e": statistics.median(t_e2e) / 1000.0},
            "steady_fps": n_streams_total * steps / steady_s,
            "conv_impl": "tcgen05" if info.conv_impl == 1 else "simt",
            "step_breakdown_usec": {"flow": flow_usec, "resblocks": res_usec, "frame_total": frame_usec},
            "hbm_kernels": hbm_ops,
            "end_to_end_tflops": info.gflop_per_frame * fps / 1000.0,
        }
        print(json.dumps(line), flush=True)
    rt.close()
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="psp_fast_b1")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
