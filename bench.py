#!/usr/bin/env python
"""Headline benchmark: 1080p output frames/s of the per-frame recurrent
upscaling path (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload psp_quality_b16]
    python bench.py --impl reference ...      # CPU restatement of the reference graph

A "step" advances every stream of the workload by one frame through the public
entry point (ju_process_batch == the C-ABI twin of Runtime::processImage).
Workload names are <preset>_b<streams per GPU>.  Default workload = the largest
single-GPU configuration of BASELINE.json (configs[2]: PSP quality, the
reference's only pinned model, 16 independent streams batched on one B200).

  value    : frames/s with the input frames already resident in HBM
             (DataLocation::CUDA images), per-step CUDA-event timing, L2 flushed
             between steps
  e2e      : the same through HOST (pinned) BGRX buffers - H2D of the frames and
             D2H of the upscaled frames inside the timed region
  e2e_pageable : the same through pageable host buffers (what the AviSynth
             plugin passes, avisynth_plugin/src/main.cc:125-142)
  latency_ms : p50 / p95 per-frame latency at batch 1 (the other half of the
             metric) for psp_quality_b1 and psp_fast_b1 (configs[1], 300 frames)
  roofline : dominant kernel (persistent ResBlock trunk) algorithmic FLOP/s, timed
             with CUDA events on the engine stream inside a >= 1 s region of
             back-to-back launches -> sustained peak (roofline.timed_alone: the same
             kernel over 10 iterations -> burst peak)
  sustained : >= 2 s of back-to-back frames, no L2 flush, NVML clocks / power
  other_configs : short runs of BASELINE configs 4 (PS2 quality) and 5 (quality +
             fast runtimes side by side on every GPU)
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
FRAME_POOL = 12  # distinct synthetic frames per stream, cycled
DEFAULT_WORKLOAD = "psp_quality_b16"
LATENCY_WORKLOADS = ("psp_quality_b1", "psp_fast_b1")
LATENCY_FRAMES = 300  # BASELINE.json configs[1]: 300-frame sequence, first 20 discarded
CPU_NOTE = "fp32 torch-CPU restatement of the reference Keras graph (TensorFlow unavailable offline)"


def parse_workload(name: str):
    preset, _, b = name.rpartition("_b")
    return jcfg.preset(preset), int(b), preset


def workload_config(name: str):
    """The `config` object of the JSON line: identical in both arms (ours / reference)."""
    cfg, streams, preset = parse_workload(name)
    return {
        "workload": name, "preset": preset, "streams_per_gpu": streams,
        "frame": [cfg.frame_width, cfg.frame_height],
        "output": [4 * cfg.frame_width, 4 * cfg.frame_height],
        "gen_blocks": cfg.gen_blocks, "flow_arch": cfg.flow_arch,
        "weights": "seeded random-init (set B, conditioned)",
        "l2": "GPU arm: flushed between steps (256 MiB memset, untimed)",
        "gflop_per_frame": cfg.gflop_per_frame(),
    }


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tensor_burst=p["bf16_tflops"],
                    tensor_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    source="MEASURED_PEAKS.json")
    # /opt/skills/guides/B200_PROFILING.md fallback figures
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sustained=1400.0, source="B200_PROFILING.md fallback")


class ClockSampler:
    """SM clock, power and throttle reasons DURING a timed region: NVML polled every 5 ms from a
    thread (the recipe's nvidia-smi -lms 200 line yields too few samples for a 0.1 s region);
    falls back to the nvidia-smi loop when pynvml is unavailable."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap,power.draw")
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
               "sw_power_cap": 0x4}  # nvmlClocksEventReason* bit masks

    def __init__(self, device: int, note: str):
        self.samples = []  # (sm_mhz, reasons bitmask, power_w)
        self.max_mhz = None
        self.lines = []
        self.proc = None
        self.note = note
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
                    "power_w_median": statistics.median(powers) if powers else None,
                    "samples": len(sm), "source": f"nvml, 5 ms period over {self.note}"}
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
                "samples": len(sm), "source": f"nvidia-smi -lms 200 over {self.note}"}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ---------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle (port of the reference graph)
# ---------------------------------------------------------------------------

def cpu_reference_fps(cfg, weights, steps, warmup, budget_s):
    """`steps` timed steps after `warmup` untimed ones; one step = one frame of one stream of the
    workload (a bounded sample: every stream runs the same graph)."""
    import torch
    from oracle import reference_graph as og
    torch.set_num_threads(os.cpu_count() or 1)
    g = og.Graph(cfg, weights, "fp32")
    frames = synthetic.frames(cfg.frame_height, cfg.frame_width, 4)[:, None]  # [T, 1, H, W, 4]
    state = g.zero_state(1)
    for t in range(warmup):
        _, state, _ = g.step(frames[t % 4], state)
    times = []
    start = time.perf_counter()
    for t in range(steps):
        t0 = time.perf_counter()
        _, state, _ = g.step(frames[(warmup + t) % 4], state)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - start > budget_s:
            break
    total = sum(times)
    return dict(fps=len(times) / total, steps=len(times), seconds=total, cores=torch.get_num_threads())


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    cfg, streams, preset = parse_workload(args.workload)
    weights = jw.init_weights(cfg, 42, True)
    r = cpu_reference_fps(cfg, weights, args.steps, args.warmup, budget_s=240.0)
    sample = (f"{r['steps']} steps ({r['seconds']:.1f} s); one step = one frame of ONE of the workload's "
              f"{streams} {preset} stream(s) (all streams run the same graph), {CPU_NOTE}")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["fps"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": args.warmup,
        "ms_per_step": 1000.0 * r["seconds"] / r["steps"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": workload_config(args.workload),
        "cpu_baseline": {"value": r["fps"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                         "sample": sample},
        "e2e": {"value": r["fps"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------

_FRAME_CACHE = {}


def stream_frames(h, w, stream_id):
    key = (h, w, stream_id)
    if key not in _FRAME_CACHE:
        _FRAME_CACHE[key] = synthetic.frames(h, w, FRAME_POOL, stream_id=stream_id)
    return _FRAME_CACHE[key]


class Workload:
    """A runtime plus device-resident, pinned-host and pageable-host image sets for one workload."""

    def __init__(self, name, device, rank, world, tmpdir, host=True):
        import ctypes as C

        from joshupscale_b200 import kernels as jk
        from joshupscale_b200 import runtime as jrt
        self.C, self.jrt = C, jrt
        self.lib = jrt.load_library()
        self.name = name
        self.cfg, self.streams, self.preset = parse_workload(name)
        self.weights = jw.init_weights(self.cfg, 42, True)
        path = os.path.join(tmpdir, f"{self.preset}.jup")
        if not os.path.exists(path):
            jw.save_model(path, self.cfg, self.weights)
        self.rt = jrt.Runtime(path, device, self.streams)
        h, w = self.cfg.frame_height, self.cfg.frame_width
        self.h, self.w = h, w
        self.in_bytes, self.out_bytes = h * w * 4, 16 * h * w * 4
        s_n = self.streams
        # global stream ids owned by this rank (stream s lives on rank s mod world)
        ids = sharding.streams_for_rank(s_n * world, world, rank)
        pool = [stream_frames(h, w, sid) for sid in ids]
        self._keep = []
        # device-resident images
        d_in = [[jk.to_device(pool[s][t]) for s in range(s_n)] for t in range(FRAME_POOL)]
        d_out = [jk.DeviceArray((4 * h, 4 * w, 4), np.uint8) for _ in range(s_n)]
        self._keep += [d_in, d_out]
        self.dev_in = [[jrt.JuImage(d_in[t][s].ptr, jrt.LOC_CUDA, w * 4, w, h) for s in range(s_n)]
                       for t in range(FRAME_POOL)]
        self.dev_out = [jrt.JuImage(d_out[s].ptr, jrt.LOC_CUDA, w * 16, 4 * w, 4 * h) for s in range(s_n)]
        self.pin_in = self.pin_out = self.page_in = self.page_out = None
        self._pinned = []
        if host:
            self.pin_in, self.pin_out = [], []
            frames_in = [[] for _ in range(FRAME_POOL)]
            for s in range(s_n):
                p_in, p_out = C.c_void_p(), C.c_void_p()
                jrt._check(self.lib.ju_host_alloc(C.byref(p_in), self.in_bytes * FRAME_POOL))
                jrt._check(self.lib.ju_host_alloc(C.byref(p_out), self.out_bytes))
                self._pinned += [p_in, p_out]
                arr = np.ctypeslib.as_array(C.cast(p_in, C.POINTER(C.c_uint8)), shape=(FRAME_POOL, h, w, 4))
                arr[...] = pool[s]
                for t in range(FRAME_POOL):
                    frames_in[t].append(jrt.JuImage(p_in.value + t * self.in_bytes, jrt.LOC_CPU, w * 4, w, h))
                self.pin_out.append(jrt.JuImage(p_out.value, jrt.LOC_CPU, w * 16, 4 * w, 4 * h))
            self.pin_in = frames_in
            # pageable numpy buffers, as a plugin host would hand over
            page_out = [np.zeros((4 * h, 4 * w, 4), np.uint8) for _ in range(s_n)]
            page_in = [[np.ascontiguousarray(pool[s][t]) for s in range(s_n)] for t in range(FRAME_POOL)]
            self._keep += [page_in, page_out]
            self.page_in = [[jrt._image(page_in[t][s], 0, 0) for s in range(s_n)] for t in range(FRAME_POOL)]
            self.page_out = [jrt._image(o, 0, 0) for o in page_out]

    def images(self, kind):
        return {"device": (self.dev_in, self.dev_out), "pinned": (self.pin_in, self.pin_out),
                "pageable": (self.page_in, self.page_out)}[kind]

    def timed(self, kind, steps, warmup, flush, barrier):
        """Per-step device time (usec, CUDA events around each call) of `steps` frames."""
        C, jrt, lib = self.C, self.jrt, self.lib
        imgs_in, imgs_out = self.images(kind)
        usec = C.c_double()
        for t in range(warmup):
            self.rt.process_images(imgs_in[t % FRAME_POOL], imgs_out)
        barrier()
        per_step = []
        for t in range(steps):
            if flush:
                jrt._check(lib.ju_l2_flush())
            jrt._check(lib.ju_timer_begin())
            self.rt.process_images(imgs_in[(warmup + t) % FRAME_POOL], imgs_out)
            jrt._check(lib.ju_timer_end(C.byref(usec)))
            per_step.append(usec.value)
        barrier()
        return per_step

    def back_to_back(self, kind, seconds, min_steps, barrier):
        """Wall-clock run of at least `seconds` (and `min_steps`) without L2 flushes."""
        imgs_in, imgs_out = self.images(kind)
        barrier()
        t0 = time.perf_counter()
        n = 0
        while True:
            self.rt.process_images(imgs_in[n % FRAME_POOL], imgs_out)
            n += 1
            if n >= min_steps and time.perf_counter() - t0 >= seconds:
                break
        wall = time.perf_counter() - t0
        barrier()
        return n, wall

    def close(self):
        self.rt.close()
        for p in self._pinned:
            self.lib.ju_host_free(p)
        self._pinned = []
        self._keep = []


def pct(values, q):
    s = sorted(values)
    return s[int(q * (len(s) - 1))]


def run_ours(args):
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
    # the runtime sets/restores its device per call; the ju_dev_*/ju_timer_* helpers act on the
    # calling thread's current device
    jrt._check(lib.ju_set_device(local))
    tmp = tempfile.mkdtemp(prefix="jubench_")

    def barrier():
        if use_dist:
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
        else:
            jrt._check(lib.ju_dev_sync())

    def global_max(x):
        return sharding.max_over_ranks(x, dist if use_dist else None, "cuda" if use_dist else None)

    def global_sum(x):
        if not use_dist:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    steps, warmup = args.steps, max(args.warmup, 3)
    main = Workload(args.workload, local, rank, world, tmp)
    streams, cfg, preset = main.streams, main.cfg, main.preset
    n_streams_total = streams * world

    # ---- headline: device-resident value, pinned e2e, pageable e2e --------------------------
    sampler = ClockSampler(local, "the value / e2e / e2e_pageable timed regions") if rank == 0 else None
    t_dev = main.timed("device", steps, warmup, True, barrier)
    t_e2e = main.timed("pinned", steps, warmup, True, barrier)
    t_page = main.timed("pageable", steps, warmup, True, barrier)
    clocks = sampler.stop() if sampler else None
    total_dev = global_max(sum(t_dev) * 1e-6)
    total_e2e = global_max(sum(t_e2e) * 1e-6)
    total_page = global_max(sum(t_page) * 1e-6)

    # ---- sustained: >= 2 s of back-to-back frames, L2 warm (deployment regime), wall clock ----
    sampler = ClockSampler(local, "the sustained region") if rank == 0 else None
    s_n, s_wall = main.back_to_back("device", args.sustained_seconds, steps, barrier)
    sustained_clocks = sampler.stop() if sampler else None
    s_wall = global_max(s_wall)
    s_frames = global_sum(s_n * streams)

    # ---- per-kernel timing (rank 0): short = kernels timed alone (burst peak), long = inside a
    # >= 1 s region of back-to-back launches (sustained peak) ---------------------------------
    roofline = hbm_ops = breakdown = None
    info = main.rt.info
    if rank == 0:
        peaks = load_peaks()
        short_roof, hbm_ops, breakdown, all_ops = kernel_report(main, peaks, 10)
        frame_usec = breakdown["frame_total"]
        # The roofline entry: the dominant kernel timed INSIDE a >= 1 s region of back-to-back
        # launches (no L2 flush, clocks and power sampled) against the SUSTAINED peak - both sides
        # of the fraction then run under the same power / clock regime.  The same kernel timed
        # alone over 10 iterations against the BURST peak is reported next to it.
        long_iters = int(max(20, min(2000, 1.2e6 / max(frame_usec, 1.0))))
        sampler = ClockSampler(local, "the >= 1 s per-kernel timing pass")
        roofline, _, long_breakdown, _ = kernel_report(main, peaks, long_iters)
        long_clocks = sampler.stop()
        roofline["peak"] = peaks["tensor_sustained"]
        roofline["frac"] = roofline["achieved"] / peaks["tensor_sustained"]
        roofline["peak_source"] = (f"{peaks['source']} bf16 dense, SUSTAINED: the kernel is timed inside a "
                                   f"{long_iters * long_breakdown['frame_total'] * 1e-6:.2f} s region of back-to-back "
                                   "launches (second pass of ju_profile_ops), like the sustained peak itself")
        roofline["clocks"] = long_clocks
        roofline["timed_alone"] = {
            "achieved": short_roof["achieved"], "peak": peaks["tensor_burst"], "frac": short_roof["frac"],
            "usec_per_launch": short_roof["usec_per_launch"],
            "peak_source": f"{peaks['source']} bf16 dense, BURST; 10 iterations after 2 warm-ups"}
        for key in ("frac_of_sustained_peak",):
            roofline.pop(key, None)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"ops_{args.workload}.json"), "w") as f:
            json.dump(all_ops, f, indent=1)
    main.rt.reset_state()

    # ---- latency half of the metric: batch-1 runtimes ---------------------------------------
    latency = {}
    if not args.no_extras:
        for name in LATENCY_WORKLOADS:
            if name == args.workload:
                continue
            try:
                wl = Workload(name, local, rank, world, tmp)
                d = wl.timed("device", LATENCY_FRAMES, 20, True, barrier)
                e = wl.timed("pinned", LATENCY_FRAMES, 20, True, barrier)
                g = wl.timed("pageable", LATENCY_FRAMES, 20, True, barrier)
                latency[name] = {
                    "p50": pct(d, 0.5) / 1000.0, "p95": pct(d, 0.95) / 1000.0,
                    "p50_e2e": pct(e, 0.5) / 1000.0, "p95_e2e": pct(e, 0.95) / 1000.0,
                    "p50_e2e_pageable": pct(g, 0.5) / 1000.0, "frames": LATENCY_FRAMES,
                    "fps": 1e6 * len(d) / sum(d), "fps_e2e": 1e6 * len(e) / sum(e),
                    "fps_e2e_pageable": 1e6 * len(g) / sum(g),
                    "gflop_per_frame": wl.cfg.gflop_per_frame(),
                }
                if rank == 0:
                    lr, _, lb, _ = kernel_report(wl, load_peaks(), 10)
                    latency[name]["step_breakdown_usec"] = lb
                    latency[name]["trunk_frac_of_burst_peak"] = lr["frac"]
                    latency[name]["kernels_per_frame"] = int(wl.rt.info.kernels_per_frame)
                wl.close()
            except Exception as exc:  # noqa: BLE001 - extras never cost the headline line
                latency[name] = {"error": repr(exc)}
    latency[args.workload] = {"p50": pct(t_dev, 0.5) / 1000.0, "p95": pct(t_dev, 0.95) / 1000.0,
                              "p50_e2e": pct(t_e2e, 0.5) / 1000.0, "p50_e2e_pageable": pct(t_page, 0.5) / 1000.0,
                              "note": f"per step of {streams} stream(s)"}

    # ---- BASELINE configs 4 and 5 (short runs) ----------------------------------------------
    other = {}
    if not args.no_extras:
        other = other_configs(local, rank, world, tmp, barrier, global_max, max(10, min(steps, 30)))

    # per-rank e2e detail (multi-GPU e2e is bound by the hosts' PCIe / memory system, not the GPUs)
    per_rank = None
    if use_dist:
        mine = torch.tensor([pct(t_e2e, 0.5) / 1000.0, streams * main.out_bytes * len(t_e2e) / (sum(t_e2e) * 1e-6) / 1e9,
                             pct(t_dev, 0.5) / 1000.0], dtype=torch.float64, device="cuda")
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        per_rank = [{"rank": r, "p50_e2e_ms": float(g[0]), "d2h_gbs": float(g[1]), "p50_ms": float(g[2])}
                    for r, g in enumerate(gathered)]

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_fps(cfg, main.weights, 24, 1, budget_s=20.0)
            cpu = {"value": r["fps"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                   "sample": f"{r['steps']} frames of 1 {preset} stream ({r['seconds']:.1f} s), {CPU_NOTE}"}
        fps = sharding.aggregate_fps(streams, world, steps, total_dev)
        line = {
            "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": 1000.0 * total_dev / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
            "config": workload_config(args.workload),
            "e2e": {"value": n_streams_total * steps / total_e2e, "unit": UNIT,
                    "h2d_bytes_per_step": streams * main.in_bytes, "d2h_bytes_per_step": streams * main.out_bytes,
                    "host_memory": "pinned"},
            "e2e_pageable": {"value": n_streams_total * steps / total_page, "unit": UNIT,
                             "host_memory": "pageable (numpy arrays)"},
            "gpu_launches": int(info.kernels_per_frame) * steps,
            "clocks": clocks,
            "roofline": roofline,
            "sustained": {"seconds": s_wall, "frames": s_frames, "fps": s_frames / s_wall,
                          "l2": "not flushed", "clocks": sustained_clocks,
                          "end_to_end_tflops": info.gflop_per_frame * s_frames / s_wall / 1000.0},
            "cpu_baseline": cpu,
            "latency_ms": latency,
            "other_configs": other,
            "conv_impl": "tcgen05" if info.conv_impl == 1 else "simt",
            "step_breakdown_usec": breakdown,
            "hbm_kernels": hbm_ops,
            "end_to_end_tflops": info.gflop_per_frame * fps / 1000.0,
        }
        if per_rank:
            line["per_rank"] = per_rank
        print(json.dumps(line), flush=True)
    main.close()
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()


def kernel_report(wl, peaks, iters):
    """Per-kernel CUDA-event timing of one frame's launches (ju_profile_ops): the dominant kernel's
    roofline entry, the HBM-bound kernels and the step breakdown."""
    h, w, streams = wl.h, wl.w, wl.streams
    all_ops = wl.rt.profile_ops(iters)
    ops = [o for o in all_ops if not o["name"].startswith("group:") and not o["name"].startswith("sync:")]
    groups = {o["name"][6:]: o for o in all_ops if o["name"].startswith("group:")}
    # dominant kernel: the generator's ResBlock convs.  Launch duration = CUDA-event time around
    # the back-to-back run of all its launches (captured as a graph, like the frame) / launches.
    res = [o for o in ops if o["name"].startswith("generator/block_")]
    grp = groups["resblocks"]
    res_usec = grp["usec"]
    frame_usec = sum(g["usec"] for g in groups.values())
    layers = grp["launches"]
    persistent = any("(persistent)" in o["name"] for o in res)
    launches = len(res)  # one op per sub-batch launch (persistent) or per layer
    chunk = max(1, -(-streams // max(launches, 1))) if persistent else streams
    achieved = grp["flops"] / (res_usec * 1e-6) / 1e12
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
        "bound": "tensor", "achieved": achieved, "peak": peaks["tensor_burst"], "unit": "TFLOP/s",
        "frac": achieved / peaks["tensor_burst"], "traffic": traffic,
        "traffic_note": "dram bytes per launch from ncu --set full of this kernel inside bench.py "
                        "(profiles/ncu_traffic.json); algorithmic bytes per launch = %d"
                        % int(grp["bytes"] / max(launches, 1)),
        "kernel": ("trunk_df_tc_kernel: generator conv_1 + all ResBlock conv3x3 64->64 layers of %d stream(s) in "
                   "one persistent launch" % chunk if persistent
                   else "conv_tc_kernel<3,1>: ResBlock conv3x3 64->64 (generator/block_*/conv_*)"),
        "layers_per_step": layers, "kernel_launches_per_step": launches,
        "flops_per_launch": grp["flops"] / max(launches, 1),
        "usec_per_launch": res_usec / max(launches, 1), "usec_per_layer_and_stream": res_usec / max(layers, 1) / streams,
        "share_of_step": res_usec / frame_usec,
        "timing": f"CUDA events on the engine stream around the kernel's back-to-back launches, {iters} "
                  "iterations after 2 warm-ups",
        "peak_source": f"{peaks['source']} bf16 dense, burst (the kernel is timed alone); the same kernel "
                       "inside a >= 1 s region is under sustained.dominant_kernel",
        "frac_of_sustained_peak": achieved / peaks["tensor_sustained"],
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
    flow = groups.get("flow")
    breakdown = {"flow": flow["usec"] if flow else None, "resblocks": res_usec, "frame_total": frame_usec}
    if flow and flow["usec"] > 0:
        breakdown["flow_tflops"] = flow["flops"] / (flow["usec"] * 1e-6) / 1e12
        breakdown["flow_frac_of_burst_peak"] = breakdown["flow_tflops"] / peaks["tensor_burst"]
    return roofline, hbm_ops, breakdown, all_ops


def other_configs(local, rank, world, tmp, barrier, global_max, steps):
    """Short runs of the BASELINE.json configurations the headline does not cover."""
    out = {}
    # config 4: PS2 quality at its native shape, 8 streams per GPU
    try:
        wl = Workload("ps2_quality_b8", local, rank, world, tmp)
        d = wl.timed("device", steps, 3, True, barrier)
        e = wl.timed("pinned", steps, 3, True, barrier)
        td, te = global_max(sum(d) * 1e-6), global_max(sum(e) * 1e-6)
        n = wl.streams * world * steps
        out["ps2_quality_b8"] = {
            "config": "BASELINE configs[3]: PS2 quality 360x480 -> 1440x1920, 8 streams per GPU, stream-sharded",
            "ps2_fps": n / td, "ps2_fps_e2e": n / te, "fps_1080p_equivalent": n / td * (1440 * 1920) / (1080 * 1920),
            "steps": steps, "gflop_per_frame": wl.cfg.gflop_per_frame()}
        wl.close()
    except Exception as exc:  # noqa: BLE001
        out["ps2_quality_b8"] = {"error": repr(exc)}
    # config 5's per-GPU shape: 8 streams, half quality half fast = two runtimes on one device,
    # advanced alternately (frames of different runtimes on one device take turns anyway)
    try:
        wq = Workload("psp_quality_b4", local, rank, world, tmp, host=False)
        wf = Workload("psp_fast_b4", local, rank, world, tmp, host=False)
        for wl in (wq, wf):
            wl.timed("device", 0, 3, False, barrier)
        qi, qo = wq.images("device")
        fi, fo = wf.images("device")
        n = 3 * steps
        barrier()
        t0 = time.perf_counter()
        for t in range(n):
            wq.rt.process_images(qi[t % FRAME_POOL], qo)
            wf.rt.process_images(fi[t % FRAME_POOL], fo)
        barrier()
        wall = global_max(time.perf_counter() - t0)
        out["mixed_4_quality_4_fast"] = {
            "config": "BASELINE configs[4] shape: per GPU 4 quality + 4 fast PSP streams (8 streams / GPU, 64 on 8 "
                      "GPUs), two runtimes on one device advanced alternately, same number of frames per stream",
            "fps": n * (wq.streams + wf.streams) * world / wall, "frames_per_stream": n, "wall_s": wall,
            "timing": "wall clock around the loop, max over ranks, L2 not flushed"}
        wq.close()
        wf.close()
    except Exception as exc:  # noqa: BLE001
        out["mixed_4_quality_4_fast"] = {"error": repr(exc)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--sustained-seconds", type=float, default=2.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the batch-1 latency runtimes and the config 4 / 5 runs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
