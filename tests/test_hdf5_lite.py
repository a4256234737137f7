"""numpy-only HDF5 reader (joshupscale_b200/hdf5_lite.py) and the importer's `.weights.h5` route.

Pinned two ways: (1) against a file written by libhdf5 itself - scipy ships MATLAB 7.3 test data
(`testhdf5_7.4_GLNX86.mat`: a 512-byte user block, then a symbol-table root group holding the
contiguous float64 dataset `testdouble = 0:pi/4:2*pi`); (2) against files laid out per the
format specification by tests/hdf5_fixture.py in the structure Keras 3 / h5py write
(reference scripts/training/train_local.py:116-129 saves `*.weights.h5`).
"""

import os

import numpy as np
import pytest

from joshupscale_b200 import config as jcfg
from joshupscale_b200 import importer
from joshupscale_b200 import weights as jw
from joshupscale_b200.hdf5_lite import Hdf5Error, Hdf5File, read_datasets

from tests.hdf5_fixture import write_hdf5


def _libhdf5_file():
    try:
        import scipy.io.matlab  # noqa: PLC0415
    except ImportError:
        return None
    p = os.path.join(os.path.dirname(scipy.io.matlab.__file__), "tests", "data", "testhdf5_7.4_GLNX86.mat")
    return p if os.path.exists(p) else None


@pytest.mark.skipif(_libhdf5_file() is None, reason="scipy's MATLAB 7.3 test file is not installed")
def test_reads_a_file_written_by_libhdf5():
    got = read_datasets(_libhdf5_file())
    assert list(got) == ["/testdouble"]
    np.testing.assert_array_equal(got["/testdouble"], (np.arange(9) * (np.pi / 4)).reshape(9, 1))
    assert got["/testdouble"].dtype == np.float64


def _sample_tree(rng):
    many = {f"d{i:03d}": np.full((2,), i, np.float32) for i in range(300)}  # two B-tree levels
    return {
        "a": rng.standard_normal((3, 3, 4, 8)).astype(np.float32),
        "grp": {
            "compact": (np.arange(6, dtype=np.int32).reshape(2, 3), "compact"),
            "chunked": (rng.standard_normal((5, 7, 3)).astype(np.float32), "chunked"),
            "big_endian": np.arange(4, dtype=">f8"),
            "half": rng.standard_normal((4,)).astype(np.float16),
            "u8": np.arange(5, dtype=np.uint8),
            "scalar": np.float32(2.5),
            "empty_group": {},
            "nested": {"deeper": {"x": np.ones((1,), np.float32)}},
        },
        "many": many,
    }


def _flatten(tree, prefix=""):
    out = {}
    for k, v in tree.items():
        if isinstance(v, dict):
            out.update(_flatten(v, f"{prefix}/{k}"))
        else:
            out[f"{prefix}/{k}"] = np.asarray(v[0] if isinstance(v, tuple) else v)
    return out


@pytest.mark.parametrize("userblock,new_style", [(0, False), (512, False), (2048, False), (0, True), (1024, True)])
def test_reads_every_supported_structure(tmp_path, userblock, new_style):
    rng = np.random.default_rng(3)
    tree = _sample_tree(rng)
    p = str(tmp_path / "t.h5")
    write_hdf5(p, tree, userblock=userblock, new_style=new_style)
    got = read_datasets(p)
    want = _flatten(tree)
    assert sorted(got) == sorted(want)
    for k, v in want.items():
        assert got[k].shape == v.shape, k
        assert got[k].dtype == v.dtype.newbyteorder("="), k
        np.testing.assert_array_equal(got[k], v)
    groups = [path for path, ds in Hdf5File(p).walk() if ds is None]
    assert "/grp/empty_group" in groups and "/grp/nested/deeper" in groups


def _keras3_tree(w):
    """`layers/<model>/layers/<layer>/vars/<i>`, the structure Keras 3's save_weights writes."""
    idx = {"kernel": 0, "bias": 1, "gamma": 0, "beta": 1, "moving_mean": 2, "moving_variance": 3}
    tree = {"layers": {}, "vars": {}}
    for k, v in w.items():
        net, rel = k.split("/", 1)
        layer, var = rel.rsplit("/", 1)
        layer = layer.replace("/", "_")
        model = tree["layers"].setdefault("flow_model" if net == "flow" else "generator", {"layers": {}, "vars": {}})
        model["layers"].setdefault(layer, {"vars": {}})["vars"][str(idx[var])] = v
    tree["layers"]["dense"] = {"vars": {"0": np.zeros((4, 1), np.float32)}}  # not part of the graph: dropped
    return tree


@pytest.mark.parametrize("preset", ["tiny", "small_bright"])
def test_weights_h5_to_container(tmp_path, preset):
    cfg = jcfg.preset(preset)
    w = jw.init_weights(cfg, 11, True)
    src, dst = str(tmp_path / "ckpt.weights.h5"), str(tmp_path / "m.jup")
    write_hdf5(src, _keras3_tree(w))
    args = [src, dst, "--height", str(cfg.frame_height), "--width", str(cfg.frame_width)]
    if cfg.normalize_brightness:
        args.append("--normalize-brightness")
    assert importer.main(args) == 0
    back_cfg, back = jw.load_model(dst)
    assert back_cfg == cfg
    for k in w:
        np.testing.assert_array_equal(back[k], w[k])


def test_damaged_files_are_refused(tmp_path):
    rng = np.random.default_rng(4)
    p = str(tmp_path / "t.h5")
    write_hdf5(p, _sample_tree(rng))
    blob = open(p, "rb").read()
    bad = str(tmp_path / "bad.h5")

    open(bad, "wb").write(b"not hdf5 at all" * 10)
    with pytest.raises(Hdf5Error, match="signature"):
        read_datasets(bad)

    open(bad, "wb").write(blob[:len(blob) // 2])  # truncated
    with pytest.raises(Hdf5Error):
        read_datasets(bad)

    broken = bytearray(blob)
    at = broken.index(b"SNOD")
    broken[at:at + 4] = b"XXXX"
    open(bad, "wb").write(bytes(broken))
    with pytest.raises(Hdf5Error, match="unexpected signature"):
        read_datasets(bad)

    # superblock version the reader does not know
    broken = bytearray(blob)
    broken[8] = 9
    open(bad, "wb").write(bytes(broken))
    with pytest.raises(Hdf5Error, match="superblock version"):
        read_datasets(bad)

    # every single-byte corruption either still parses or raises Hdf5Error - never another exception
    for at in rng.integers(0, min(len(blob), 4096), 200):
        broken = bytearray(blob)
        broken[at] ^= 0xFF
        open(bad, "wb").write(bytes(broken))
        try:
            read_datasets(bad)
        except Hdf5Error:
            pass


def test_unsupported_datatype_is_named(tmp_path):
    p = str(tmp_path / "t.h5")
    write_hdf5(p, {"x": np.ones((2,), np.float32)})
    blob = bytearray(open(p, "rb").read())
    # turn the float datatype message (class 1, version 1 -> 0x11) into a string type (class 3)
    at = blob.index(bytes([0x11, 0x20, 31, 0]))
    blob[at] = 0x13
    open(p, "wb").write(bytes(blob))
    with pytest.raises(Hdf5Error, match="string"):
        read_datasets(p)
    assert read_datasets(p, skip_unsupported=True) == {}
