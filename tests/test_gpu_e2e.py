"""End-to-end parity of the runtime entry point (ju_process == the C-ABI twin
of Runtime::processImage) against the CPU oracle, plus the boundary contract:
strides, CUDA-resident images, batching, determinism, state handling."""

import ctypes

import numpy as np
import pytest
import torch

from joshupscale_b200 import kernels as jk
from joshupscale_b200 import runtime as jrt
from joshupscale_b200 import synthetic
from oracle import reference_graph as og
from tests.gpu_util import make_model, require_gpu, u8_stats

pytestmark = pytest.mark.gpu

# north star: final image within max-abs 2/255 and >= 45 dB PSNR of the fp32
# reference graph, fp16 storage on the GPU
MAX_ABS_FP32 = 2
MIN_PSNR_DB = 45.0


@pytest.fixture(autouse=True, scope="module")
def _gpu():
    require_gpu()


def _run_gpu(path, frames, batch=1):
    with jrt.Runtime(path, 0, batch) as rt:
        return np.stack([rt.process(f) for f in frames])


@pytest.mark.parametrize("preset,conditioned,nframes", [
    ("tiny", True, 6), ("small", True, 6), ("small", False, 4), ("small_resnet", True, 4),
    ("small_bright", True, 5)])
def test_recurrent_parity_small(tmp_path, preset, conditioned, nframes):
    cfg, w, path = make_model(tmp_path, preset, conditioned=conditioned)
    frames = synthetic.frames(cfg.frame_height, cfg.frame_width, nframes)
    got = _run_gpu(path, frames)
    emu, _ = og.Graph(cfg, w, "fp16emu").run(frames)
    ref, _ = og.Graph(cfg, w, "fp32").run(frames)
    assert not got[..., 3].any()
    for t in range(nframes):
        m16, frac16, _ = u8_stats(got[t, ..., :3], emu[t, ..., :3])
        m32, _, psnr32 = u8_stats(got[t, ..., :3], ref[t, ..., :3])
        # against the fp16-emulating oracle only truncation flips remain
        assert m16 <= 1 and frac16 < 0.06, (t, m16, frac16)
        assert m32 <= MAX_ABS_FP32 and psnr32 >= MIN_PSNR_DB, (t, m32, psnr32)


def test_intermediate_taps_match_oracle(tmp_path):
    """flow head, generator input and recurrent state after 3 frames."""
    cfg, w, path = make_model(tmp_path, "small")
    frames = synthetic.frames(cfg.frame_height, cfg.frame_width, 3)
    g = og.Graph(cfg, w, "fp16emu")
    g.keep_taps = True
    state = g.zero_state()
    with jrt.Runtime(path) as rt:
        for t in range(3):
            out = rt.process(frames[t])
            want, state, aux = g.step(frames[t:t + 1], state)
            head = rt.read_tensor("flow_head")
            np.testing.assert_allclose(head, g.taps["flow/head"].numpy(), rtol=0, atol=5e-3)
            gen_in = rt.read_tensor("gen_in").astype(np.float32)
            want_in = g.taps["generator/input"].numpy()
            assert not gen_in[..., 51:].any()
            # warp of a slightly different flow: compare loosely, LR channels exactly
            np.testing.assert_array_equal(gen_in[..., :3], want_in[..., :3])
            assert np.abs(gen_in[..., 3:51] - want_in[..., 3:]).mean() < 2e-3
            pre_gen = rt.read_tensor("pre_gen").astype(np.float32)
            assert np.abs(pre_gen[..., :3] - state["pre_gen"].numpy()).max() < 6e-3
            flow_in = rt.read_tensor("flow_in").astype(np.float32)
            want_fi = torch.cat(list(state["last_frames"]), -1).numpy()
            np.testing.assert_array_equal(flow_in[..., :want_fi.shape[-1]], want_fi)


def test_full_size_psp_quality_16_frames(tmp_path):
    """BASELINE.json config 1: PSP quality, 1 stream, 16 synthetic 480x270 frames."""
    cfg, w, path = make_model(tmp_path, "psp_quality")
    frames = synthetic.frames(270, 480, 16)
    got = _run_gpu(path, frames)
    ref, _ = og.Graph(cfg, w, "fp32").run(frames)
    worst_abs, worst_psnr = 0, 1e9
    for t in range(16):
        m, _, p = u8_stats(got[t, ..., :3], ref[t, ..., :3])
        worst_abs, worst_psnr = max(worst_abs, m), min(worst_psnr, p)
    print(f"PSP quality 16 frames vs fp32 oracle: max-abs {worst_abs}, min PSNR {worst_psnr:.2f} dB")
    assert worst_abs <= MAX_ABS_FP32 and worst_psnr >= MIN_PSNR_DB


def test_adversarial_inputs(tmp_path):
    cfg, w, path = make_model(tmp_path, "small")
    for kind in ("black", "white", "checker", "cut"):
        frames = synthetic.frames(cfg.frame_height, cfg.frame_width, 6, kind=kind)
        got = _run_gpu(path, frames)
        ref, _ = og.Graph(cfg, w, "fp32").run(frames)
        m, _, p = u8_stats(got[..., :3], ref[..., :3])
        assert m <= MAX_ABS_FP32 and p >= MIN_PSNR_DB, (kind, m, p)


def test_bit_stable_and_reset(tmp_path):
    cfg, w, path = make_model(tmp_path, "small")
    frames = synthetic.frames(cfg.frame_height, cfg.frame_width, 5)
    a = _run_gpu(path, frames)
    b = _run_gpu(path, frames)
    np.testing.assert_array_equal(a, b)
    with jrt.Runtime(path) as rt:
        for f in frames[:3]:
            rt.process(f)
        rt.reset_state()
        c = np.stack([rt.process(f) for f in frames])
    np.testing.assert_array_equal(a, c)


def test_strides_and_bottom_up_images(tmp_path):
    """Padded and negative strides (AviSynth passes the last row + negative
    stride, avisynth_plugin/src/main.cc:125-142)."""
    cfg, w, path = make_model(tmp_path, "small")
    h, wd = cfg.frame_height, cfg.frame_width
    frames = synthetic.frames(h, wd, 3)
    want = _run_gpu(path, frames)
    lib = jrt.load_library()
    with jrt.Runtime(path) as rt:
        for t in range(3):
            # input: bottom-up with row padding; output: bottom-up with padding
            in_pitch, out_pitch = wd * 4 + 32, wd * 16 + 64
            ibuf = np.zeros((h, in_pitch), np.uint8)
            ibuf[::-1, :wd * 4] = frames[t].reshape(h, wd * 4)
            obuf = np.full((4 * h, out_pitch), 0xAB, np.uint8)
            i = jrt.JuImage(ibuf.ctypes.data + (h - 1) * in_pitch, jrt.LOC_CPU, -in_pitch, wd, h)
            o = jrt.JuImage(obuf.ctypes.data + (4 * h - 1) * out_pitch, jrt.LOC_CPU, -out_pitch, 4 * wd, 4 * h)
            rt.process_images([i], [o])
            np.testing.assert_array_equal(obuf[::-1, :wd * 16].reshape(4 * h, 4 * wd, 4), want[t])
            assert (obuf[:, wd * 16:] == 0xAB).all()  # padding untouched


def test_cuda_resident_images(tmp_path):
    cfg, w, path = make_model(tmp_path, "small")
    h, wd = cfg.frame_height, cfg.frame_width
    frames = synthetic.frames(h, wd, 3)
    want = _run_gpu(path, frames)
    with jrt.Runtime(path) as rt:
        d_in = jk.DeviceArray((h, wd, 4), np.uint8)
        d_out = jk.DeviceArray((4 * h, 4 * wd, 4), np.uint8)
        for t in range(3):
            d_in.upload(frames[t])
            i = jrt.JuImage(d_in.ptr, jrt.LOC_CUDA, wd * 4, wd, h)
            o = jrt.JuImage(d_out.ptr, jrt.LOC_CUDA, wd * 16, 4 * wd, 4 * h)
            rt.process_images([i], [o])
            np.testing.assert_array_equal(d_out.download(), want[t])


def test_batched_streams_equal_single_streams(tmp_path):
    """N independent streams in lockstep == each stream alone (bit-exact)."""
    cfg, w, path = make_model(tmp_path, "small")
    h, wd = cfg.frame_height, cfg.frame_width
    streams = [synthetic.frames(h, wd, 4, stream_id=s) for s in range(3)]
    singles = [_run_gpu(path, s) for s in streams]
    with jrt.Runtime(path, 0, 3) as rt:
        for t in range(4):
            outs = rt.process_batch([s[t] for s in streams])
            for s in range(3):
                np.testing.assert_array_equal(outs[s], singles[s][t])


def test_errors_leave_state_untouched(tmp_path):
    cfg, w, path = make_model(tmp_path, "small")
    h, wd = cfg.frame_height, cfg.frame_width
    frames = synthetic.frames(h, wd, 3)
    want = _run_gpu(path, frames)
    with jrt.Runtime(path) as rt:
        np.testing.assert_array_equal(rt.process(frames[0]), want[0])
        with pytest.raises(jrt.JoshUpscaleError, match="size"):
            rt.process(np.zeros((h + 1, wd, 4), np.uint8))
        bad = jrt.JuImage(1, 7, 0, wd, h)
        good = jrt._image(np.empty(rt.out_shape, np.uint8), 0, 0)
        with pytest.raises(jrt.JoshUpscaleError, match="unknown image location"):
            rt.process_images([bad], [good])
        # CUDA images: pointer / stride must be 4-byte aligned and at least one row wide
        d_in = jk.to_device(frames[1])
        d_out = jk.DeviceArray((4 * h, 4 * wd, 4), np.uint8)
        ok_in = jrt.JuImage(d_in.ptr, jrt.LOC_CUDA, wd * 4, wd, h)
        for bad_out in (jrt.JuImage(d_out.ptr + 2, jrt.LOC_CUDA, wd * 16, 4 * wd, 4 * h),
                        jrt.JuImage(d_out.ptr, jrt.LOC_CUDA, wd * 16 - 4, 4 * wd, 4 * h)):
            with pytest.raises(jrt.JoshUpscaleError, match="aligned|stride"):
                rt.process_images([ok_in], [bad_out])
        np.testing.assert_array_equal(rt.process(frames[1]), want[1])


def test_session_mirror_of_reference_driver(tmp_path):
    cfg, w, path = make_model(tmp_path, "small")
    frames = synthetic.frames(cfg.frame_height, cfg.frame_width, 2)
    want = _run_gpu(path, frames)
    sess = jrt.Session(path)
    for t in range(2):
        np.testing.assert_array_equal(sess.run(frames[t][..., :3]), want[t][..., :3])


def test_cxx_api_like_the_plugins(tmp_path):
    """Compile a caller against include/JoshUpscale/core.h (the header the
    AviSynth/OBS plugins include) and check it produces the same bytes."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg, w, path = make_model(tmp_path, "small")
    frames = synthetic.frames(cfg.frame_height, cfg.frame_width, 3)
    want = _run_gpu(path, frames)
    exe = str(tmp_path / "api_smoke")
    libdir = os.path.dirname(jrt.library_path())
    subprocess.run(["g++", "-std=c++20", "-O1", "-I", os.path.join(root, "include"),
                    os.path.join(root, "tests", "cxx", "api_smoke.cc"), "-o", exe,
                    "-L", libdir, "-lJoshUpscale", f"-Wl,-rpath,{libdir}"], check=True)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    frames.tofile(fin)
    r = subprocess.run([exe, path, fin, "3", fout], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ModelException" in r.stdout and "model.jup" in r.stdout
    got = np.fromfile(fout, np.uint8).reshape(want.shape)
    np.testing.assert_array_equal(got, want)


def test_long_sequence_does_not_drift(tmp_path):
    """40 recurrent frames: the fp16 engine tracks the fp32 graph without accumulating error."""
    cfg, w, path = make_model(tmp_path, "small")
    frames = synthetic.frames(cfg.frame_height, cfg.frame_width, 40)
    got = _run_gpu(path, frames)
    ref, _ = og.Graph(cfg, w, "fp32").run(frames)
    psnrs = [u8_stats(got[t, ..., :3], ref[t, ..., :3]) for t in range(40)]
    assert max(p[0] for p in psnrs) <= MAX_ABS_FP32
    assert min(p[2] for p in psnrs[30:]) >= MIN_PSNR_DB
    assert min(p[2] for p in psnrs[30:]) > min(p[2] for p in psnrs[:10]) - 3.0


def test_ps2_size_and_fast_model(tmp_path):
    """BASELINE configs 2 and 4: PS2 frame size (360x480 -> 1440x1920) and the fast generator."""
    import dataclasses
    from joshupscale_b200 import config as jcfg
    cfg = dataclasses.replace(jcfg.preset("ps2_fast"), gen_blocks=2)
    cfg, w, path = make_model(tmp_path, cfg)
    frames = synthetic.frames(360, 480, 3)
    got = _run_gpu(path, frames)
    assert got.shape == (3, 1440, 1920, 4)
    ref, _ = og.Graph(cfg, w, "fp32").run(frames)
    for t in range(3):
        m, _, p = u8_stats(got[t, ..., :3], ref[t, ..., :3])
        assert m <= MAX_ABS_FP32 and p >= MIN_PSNR_DB


def test_two_runtimes_interleaved_and_threaded(tmp_path):
    """Different models in one process (the OBS plugin switches presets; config 5 mixes
    quality and fast streams on one GPU): instances must not share mutable state."""
    import threading
    cfg_a, w_a, path_a = make_model(tmp_path, "small", seed=1)
    cfg_b, w_b, path_b = make_model(tmp_path, "small_resnet", seed=2)
    fa = synthetic.frames(cfg_a.frame_height, cfg_a.frame_width, 4, stream_id=1)
    fb = synthetic.frames(cfg_b.frame_height, cfg_b.frame_width, 4, stream_id=2)
    want_a, want_b = _run_gpu(path_a, fa), _run_gpu(path_b, fb)
    with jrt.Runtime(path_a) as ra, jrt.Runtime(path_b) as rb:
        for t in range(2):  # interleaved on one thread
            np.testing.assert_array_equal(ra.process(fa[t]), want_a[t])
            np.testing.assert_array_equal(rb.process(fb[t]), want_b[t])
        results = {}

        def worker(name, rt, frames):
            results[name] = [rt.process(f) for f in frames[2:]]

        ths = [threading.Thread(target=worker, args=("a", ra, fa)), threading.Thread(target=worker, args=("b", rb, fb))]
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        for t in range(2):
            np.testing.assert_array_equal(results["a"][t], want_a[2 + t])
            np.testing.assert_array_equal(results["b"][t], want_b[2 + t])


def test_full_size_batch_equals_single(tmp_path):
    """PSP fast at full size, 3 streams batched == each stream alone (persistent trunk, batch > 1)."""
    cfg, w, path = make_model(tmp_path, "psp_fast")
    streams = [synthetic.frames(270, 480, 2, stream_id=s) for s in range(3)]
    singles = [_run_gpu(path, s) for s in streams]
    with jrt.Runtime(path, 0, 3) as rt:
        for t in range(2):
            outs = rt.process_batch([s[t] for s in streams])
            for s in range(3):
                np.testing.assert_array_equal(outs[s], singles[s][t])


def test_sequencer_drives_the_runtime_like_the_avisynth_filter(tmp_path):
    """FrameSequencer (avisynth_plugin/src/main.cc:75-161 policy) over the real runtime: a seek
    warms the recurrent state up over the 16 previous frames, mirrored at the clip start."""
    from joshupscale_b200 import sequencer as js
    cfg, _, path = make_model(tmp_path, "tiny")
    clip = synthetic.frames(cfg.frame_height, cfg.frame_width, 40)
    with jrt.Runtime(path, 0, 1) as rt, jrt.Runtime(path, 0, 1) as manual:
        seq = js.FrameSequencer(lambda i: clip[i], lambda f: rt.process(f).copy())
        got0 = seq.get(0)
        for k in range(-16, 1):
            want = manual.process(clip[abs(k)])
        np.testing.assert_array_equal(got0, want)
        got1 = seq.get(1)
        np.testing.assert_array_equal(got1, manual.process(clip[1]))
        assert seq.stats.processed == 18 and seq.next == 2
        # far seek: fresh warm-up over frames 14..30 - but the RUNTIME state is not reset by the
        # plugin either (there is no reset call in the reference API); outputs converge with warm-up
        far = seq.get(30)
        assert seq.stats.resets == 1 and seq.stats.processed == 18 + 17
        for k in range(14, 31):
            want = manual.process(clip[k])
        np.testing.assert_array_equal(far, want)


def test_pinned_and_pageable_host_outputs_agree(tmp_path):
    """Page-locked and pageable output buffers must receive the same bytes, for top-down and
    bottom-up images with padded rows, and bytes outside the image rows must stay untouched.
    (Storing the image straight into page-locked memory from the last kernel was measured 3.4x
    slower end to end than the staged copy - 16-byte PCIe writes - and is not done.)"""
    import ctypes as C
    cfg, _, path = make_model(tmp_path, "small")
    h, w = cfg.frame_height, cfg.frame_width
    oh, ow = 4 * h, 4 * w
    frames = synthetic.frames(h, w, 3)
    lib = jrt.load_library()
    pitch = ow * 4 + 64  # padded rows
    p_out = C.c_void_p()
    jrt._check(lib.ju_host_alloc(C.byref(p_out), pitch * oh))
    try:
        pinned = np.ctypeslib.as_array(C.cast(p_out, C.POINTER(C.c_uint8)), shape=(oh, pitch))
        with jrt.Runtime(path, 0, 1) as a, jrt.Runtime(path, 0, 1) as b:
            for t, f in enumerate(frames):
                want = a.process(f)  # pageable numpy output
                pinned[...] = 0xAB
                src = np.ascontiguousarray(f)
                i = jrt.JuImage(src.ctypes.data, jrt.LOC_CPU, w * 4, w, h)
                if t % 2 == 0:
                    o = jrt.JuImage(p_out.value, jrt.LOC_CPU, pitch, ow, oh)
                    b.process_images([i], [o])
                    got = pinned[:, :ow * 4].reshape(oh, ow, 4)
                else:  # bottom-up: pointer to the last memory row, negative stride
                    o = jrt.JuImage(p_out.value + (oh - 1) * pitch, jrt.LOC_CPU, -pitch, ow, oh)
                    b.process_images([i], [o])
                    got = pinned[::-1, :ow * 4].reshape(oh, ow, 4)
                np.testing.assert_array_equal(got, want)
                assert (pinned[:, ow * 4:] == 0xAB).all()
    finally:
        jrt._check(lib.ju_host_free(p_out))


@pytest.mark.parametrize("with_filter", [False, True])
def test_sub_batched_streams_with_overlapped_copies_equal_single_streams(tmp_path, with_filter):
    """Batch 5 at full size runs as trunk/tail sub-batches (2 + 2 + 1 streams) whose images are copied
    to the host on a second stream as each sub-batch finishes; every stream must still equal its own
    single-stream run, frame after frame (host images, so the staged-copy path is exercised)."""
    from joshupscale_b200 import config as jcfg, weights as jw
    import os
    cfg = jcfg.preset("psp_fast")
    w = jw.init_weights(cfg, 42, True)
    flt = jcfg.OutputFilter(window=32, threshold=0.2) if with_filter else None
    path = os.path.join(str(tmp_path), "m.jup")
    jw.save_model(path, cfg, jw.with_output_filter(w, flt))
    n, frames = 5, 3
    clips = [synthetic.frames(cfg.frame_height, cfg.frame_width, frames, stream_id=s,
                              kind="cut" if s % 2 else "pan") for s in range(n)]
    with jrt.Runtime(path, 0, n) as rt:
        batched = [rt.process_batch([c[t] for c in clips]) for t in range(frames)]
        batched = [[o.copy() for o in outs] for outs in batched]
    for s in (0, 3, 4):
        with jrt.Runtime(path, 0, 1) as one:
            for t in range(frames):
                np.testing.assert_array_equal(batched[t][s], one.process(clips[s][t]))


def test_banded_tail_for_host_images_matches_device_resident_run(tmp_path):
    """At batch 1 frames with host images run the tail kernel in bands with overlapped row-band
    copies; device-resident images run it in one pass.  Same bytes either way, also for bottom-up
    host images, and the recurrent state stays in step when the two kinds of call alternate."""
    import ctypes as C
    cfg, _, path = make_model(tmp_path, "psp_fast")
    h, w = cfg.frame_height, cfg.frame_width
    oh, ow = 4 * h, 4 * w
    frames = synthetic.frames(h, w, 4)
    lib = jrt.load_library()
    d_in, d_out = jk.DeviceArray((h, w, 4), np.uint8), jk.DeviceArray((oh, ow, 4), np.uint8)
    with jrt.Runtime(path, 0, 1) as host_rt, jrt.Runtime(path, 0, 1) as dev_rt:
        for t, f in enumerate(frames):
            d_in.upload(np.ascontiguousarray(f))
            dev_rt.process_images([jrt.JuImage(d_in.ptr, jrt.LOC_CUDA, w * 4, w, h)],
                                  [jrt.JuImage(d_out.ptr, jrt.LOC_CUDA, ow * 4, ow, oh)])
            want = d_out.download()
            if t % 2 == 0:
                got = host_rt.process(f)
            else:  # bottom-up host image: pointer to the last memory row, negative stride
                buf = np.full((oh, ow, 4), 0xCD, np.uint8)
                src = np.ascontiguousarray(f[::-1])
                host_rt.process_images(
                    [jrt.JuImage(src.ctypes.data + (h - 1) * w * 4, jrt.LOC_CPU, -w * 4, w, h)],
                    [jrt.JuImage(buf.ctypes.data + (oh - 1) * ow * 4, jrt.LOC_CPU, -ow * 4, ow, oh)])
                got = buf[::-1]
            np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("name", ["tiny", "tiny_default_init", "small_bright", "small_resnet"])
def test_engine_matches_vectors_from_the_reference_built_graph(tmp_path, name):
    """tests/golden/graph_golden.npz comes from executing the reference's own models.py /
    keras_layers.py / tfa under a shim (tests/golden/make_graph_golden.py): the CUDA engine must
    match it within the north-star tolerance, without the oracle in between."""
    import os
    from tests.test_graph_golden import CASES, GOLDEN, golden_inputs
    from joshupscale_b200 import weights as jw
    gold = np.load(GOLDEN)
    cfg, w, frames = golden_inputs(name)
    path = os.path.join(str(tmp_path), "m.jup")
    jw.save_model(path, cfg, w)
    got = _run_gpu(path, frames)
    for t in range(frames.shape[0]):
        m, _, psnr = u8_stats(got[t, ..., :3], gold[f"{name}/output"][t])
        assert m <= MAX_ABS_FP32 and psnr >= MIN_PSNR_DB, (name, t, m, psnr)


def test_alternating_host_and_device_images_on_one_runtime(tmp_path):
    """One runtime, calls alternating between host and device-resident images: the two frame-plan
    variants (banded tail with overlapped copies / single pass) share the recurrent state."""
    cfg, _, path = make_model(tmp_path, "psp_fast")
    h, w = cfg.frame_height, cfg.frame_width
    frames = synthetic.frames(h, w, 5)
    want = _run_gpu(path, frames)
    d_in, d_out = jk.DeviceArray((h, w, 4), np.uint8), jk.DeviceArray((4 * h, 4 * w, 4), np.uint8)
    with jrt.Runtime(path, 0, 1) as rt:
        for t, f in enumerate(frames):
            if t % 2:
                d_in.upload(np.ascontiguousarray(f))
                rt.process_images([jrt.JuImage(d_in.ptr, jrt.LOC_CUDA, w * 4, w, h)],
                                  [jrt.JuImage(d_out.ptr, jrt.LOC_CUDA, w * 16, 4 * w, 4 * h)])
                got = d_out.download()
            else:
                got = rt.process(f)
            np.testing.assert_array_equal(got, want[t])
