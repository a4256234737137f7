"""The `.jup` reader (csrc/host/model.cc) must turn every corrupt or crafted container into an
error from ju_create - never an out-of-bounds read or a giant allocation.  No GPU needed: the
file is parsed before the device is touched."""

import ctypes as C
import os
import struct

import pytest

from joshupscale_b200 import config as jcfg
from joshupscale_b200 import runtime as jrt
from joshupscale_b200 import weights as jw

HSIZE = struct.calcsize(jw._HEADER_FMT)
ESIZE = struct.calcsize(jw._ENTRY_FMT)


@pytest.fixture(scope="module")
def good(tmp_path_factory):
    cfg = jcfg.preset("tiny")
    path = str(tmp_path_factory.mktemp("jup") / "tiny.jup")
    jw.save_model(path, cfg, jw.init_weights(cfg, 1, True))
    return open(path, "rb").read()


def _create_error(tmp_path, blob):
    path = os.path.join(tmp_path, "m.jup")
    with open(path, "wb") as f:
        f.write(blob)
    lib = jrt.load_library()
    h = C.c_void_p()
    rc = lib.ju_create(path.encode(), 0, 1, C.byref(h))
    assert rc != 0 and not h.value
    return lib.ju_last_error().decode()


def _patch(blob, offset, fmt, *values):
    b = bytearray(blob)
    struct.pack_into(fmt, b, offset, *values)
    return bytes(b)


def test_header_fields_are_bounded(tmp_path, good):
    # offsets inside the packed header: magic 8, version 4, headerBytes 4, then u32 fields
    field = lambda i: 16 + 4 * i  # noqa: E731 - frameH, frameW, padH, padW, arch, K, nFilters, filters[16], ...
    cases = {
        "headerBytes too small": _patch(good, 12, "<I", 8),
        "headerBytes past the file": _patch(good, 12, "<I", len(good) + 1),
        "huge frame": _patch(good, field(0), "<I", 1 << 30),
        "pad smaller than frame": _patch(good, field(2), "<I", 1),
        "too many flow inputs": _patch(good, field(5), "<I", 4000),
        "too many filters": _patch(good, field(6), "<I", 17),
        "zero filter": _patch(good, field(7), "<I", 0),
        "resnet without blocks entry": _patch(_patch(good, field(4), "<I", 1), field(6), "<I", 1),
        "giant generator": _patch(good, field(7 + 16), "<I", 1 << 20),
        "too many blocks": _patch(good, field(7 + 17), "<I", 1 << 20),
        "tensor count": _patch(good, HSIZE - 4, "<I", 1 << 28),
    }
    for name, blob in cases.items():
        msg = _create_error(str(tmp_path), blob)
        assert "ModelException" in msg, (name, msg)


def test_tensor_entries_cannot_point_outside_the_file(tmp_path, good):
    entry0 = HSIZE  # name[96], dtype, ndim, dims[4], offset u64, nbytes u64
    off_pos, nbytes_pos, dims_pos = entry0 + 96 + 8 + 16, entry0 + 96 + 8 + 16 + 8, entry0 + 96 + 8
    cases = {
        "offset + nbytes wraps": _patch(good, off_pos, "<QQ", 2 ** 64 - 16, 64),
        "offset past the end": _patch(good, off_pos, "<Q", len(good) + 64),
        "nbytes past the end": _patch(good, nbytes_pos, "<Q", len(good)),
        "dims product overflows": _patch(good, dims_pos, "<IIII", 2 ** 24, 2 ** 24, 2 ** 24, 2 ** 24),
        "zero dimension": _patch(good, dims_pos, "<I", 0),
    }
    for name, blob in cases.items():
        msg = _create_error(str(tmp_path), blob)
        assert "ModelException" in msg, (name, msg)


def test_truncated_files(tmp_path, good):
    for cut in (0, 7, HSIZE - 1, HSIZE + ESIZE // 2, len(good) // 2):
        msg = _create_error(str(tmp_path), good[:cut])
        assert "ModelException" in msg, (cut, msg)
