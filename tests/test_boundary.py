"""Boundary checks that need no GPU: the C-ABI library builds, loads, exports
every symbol include/joshupscale_c.h declares and the C++ API symbols of
include/JoshUpscale/core.h; failures are reported through ju_last_error."""

import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from joshupscale_b200 import build as jbuild
from joshupscale_b200 import config as jcfg
from joshupscale_b200 import runtime as jrt
from joshupscale_b200 import weights as jw

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    jbuild.build()
    return jrt.load_library()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "joshupscale_c.h")).read()
    return sorted(set(re.findall(r"JU_API[^;]*?\b(ju_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in joshupscale_c.h but not exported"
    # and the python binding table covers exactly the header
    assert sorted(jrt.SYMBOLS) == names


def test_cxx_api_symbols_exported(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", "-C", jrt.library_path()],
                         capture_output=True, text=True, check=True).stdout
    for sym in ("JoshUpscale::core::createRuntime(int, std::filesystem",
                "JoshUpscale::core::getExceptionString",
                "JoshUpscale::core::setLogSink(JoshUpscale::core::LogSink*)",
                "JoshUpscale::core::getGLDeviceIndex()",
                "JoshUpscale::core::getGLImage(unsigned int"):
        assert sym in out, sym
    # hidden visibility: nothing from the ju:: implementation namespace leaks
    assert " ju::" not in out


def test_library_is_self_contained(lib):
    ldd = subprocess.run(["ldd", jrt.library_path()], capture_output=True, text=True).stdout
    assert "libcuda.so" not in ldd and "nvinfer" not in ldd and "cudnn" not in ldd


def test_version_and_device_count(lib):
    assert b"sm_100a" in lib.ju_version()
    assert lib.ju_device_count() >= 0


def test_missing_model_reports_error(lib, tmp_path):
    h = ctypes.c_void_p()
    rc = lib.ju_create(str(tmp_path / "nope.jup").encode(), 0, 1, ctypes.byref(h))
    assert rc != 0 and not h.value
    msg = lib.ju_last_error().decode()
    assert "nope.jup" in msg or "CUDA" in msg or "cuda" in msg


def test_trt_engine_is_rejected_with_clear_message(lib, tmp_path):
    """The reference passes a serialized TensorRT engine at modelPath
    (core/src/core.cc:154-167); this build must say it cannot consume one."""
    p = tmp_path / "model_psp.trt"
    p.write_bytes(b"ptrt" + bytes(4096))
    h = ctypes.c_void_p()
    rc = lib.ju_create(str(p).encode(), 0, 1, ctypes.byref(h))
    assert rc != 0
    msg = lib.ju_last_error().decode()
    if lib.ju_device_count() > 0:
        assert ".jup" in msg


@pytest.mark.skipif(jrt.load_library().ju_device_count() > 0 if os.path.exists(jrt.library_path()) else False,
                    reason="only meaningful without a GPU")
def test_no_cpu_fallback_without_gpu(lib, tmp_path):
    cfg = jcfg.preset("tiny")
    path = str(tmp_path / "tiny.jup")
    jw.save_model(path, cfg, jw.init_weights(cfg))
    with pytest.raises(jrt.JoshUpscaleError):
        jrt.Runtime(path)
    from joshupscale_b200 import kernels as jk
    with pytest.raises(jrt.JoshUpscaleError):
        jk.maxpool2(np.zeros((1, 4, 4, 8), np.float16))


def test_python_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "joshupscale_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cc", ".cu", ".h")):
                text = open(os.path.join(dp, f), errors="replace").read()
                assert "import oracle" not in text and "from oracle" not in text, f
