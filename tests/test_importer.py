"""Weight importer: Keras variable paths (npz / Keras-3 h5 layout) -> .jup."""

import re

import numpy as np
import pytest

from joshupscale_b200 import config as jcfg
from joshupscale_b200 import importer
from joshupscale_b200 import weights as jw


def _keras_paths(w, flow_scope="final/full/flow_model_1", gen_scope="final/full/generator_2", suffix=":0"):
    out = {}
    for k, v in w.items():
        net, rel = k.split("/", 1)
        # the reference names block layers "block_3_conv_1" (models.py get_scoped_name)
        rel = re.sub(r"^(block_\d+)/", r"\1_", rel)
        out[f"{flow_scope if net == 'flow' else gen_scope}/{rel}{suffix}"] = v
    # things a training checkpoint also holds and the importer must ignore
    out["discriminator/block_1/conv_1/kernel:0"] = np.zeros((3, 3, 6, 8), np.float32)
    out["adam/final_full_generator_conv_1_kernel_momentum"] = np.zeros((3,), np.float32)
    out["final/full/generator_2/fade/alpha:0"] = np.zeros((), np.float32)
    return out


@pytest.mark.parametrize("preset", ["tiny", "small", "small_resnet", "small_bright"])
def test_import_recovers_config_and_weights(tmp_path, preset):
    cfg = jcfg.preset(preset)
    w = jw.init_weights(cfg, 7, True)
    got_cfg, got_w = importer.import_weights(
        _keras_paths(w), cfg.frame_height, cfg.frame_width,
        normalize_brightness=cfg.normalize_brightness)
    assert got_cfg == cfg
    assert list(got_w) == list(w)
    for k in w:
        np.testing.assert_array_equal(got_w[k], w[k])


def test_cli_npz_to_container_round_trip(tmp_path):
    cfg = jcfg.preset("tiny")
    w = jw.init_weights(cfg, 3, False)
    src, dst = str(tmp_path / "w.npz"), str(tmp_path / "m.jup")
    np.savez(src, **_keras_paths(w, suffix=""))
    rc = importer.main([src, dst, "--height", "21", "--width", "27", "--filter", "--filter-window", "16"])
    assert rc == 0
    back_cfg, back = jw.load_model(dst)
    assert back_cfg == cfg
    assert back[jw.FILTER_TENSOR][2] == 16.0
    for k in w:
        np.testing.assert_array_equal(back[k], w[k])


def test_keras3_h5_layout_is_renamed_by_rank():
    cfg = jcfg.preset("tiny")
    w = jw.init_weights(cfg, 5, True)
    # Keras 3 stores <layer>/vars/<index> instead of variable names
    idx = {"kernel": 0, "bias": 1, "gamma": 0, "beta": 1, "moving_mean": 2, "moving_variance": 3}
    flat = {}
    for k, v in w.items():
        net, rel = k.split("/", 1)
        layer, var = rel.rsplit("/", 1)
        layer = layer.replace("/", "_")  # block_3_conv_1, as the reference names its layers
        scope = "layers/flow_model" if net == "flow" else "layers/generator"
        flat[f"{scope}/layers/{layer}/vars/{idx[var]}"] = v
    flat["layers/dense/vars/0"] = np.zeros((4, 1), np.float32)  # not conv / bn: dropped
    named = importer.rename_indexed_vars(flat)
    got_cfg, got_w = importer.import_weights(named, cfg.frame_height, cfg.frame_width)
    assert got_cfg == cfg
    for k in w:
        np.testing.assert_array_equal(got_w[k], w[k])


def test_import_errors_are_specific(tmp_path):
    cfg = jcfg.preset("tiny")
    w = _keras_paths(jw.init_weights(cfg, 1, True))
    missing = {k: v for k, v in w.items() if "generator_2/block_2_bn_1/beta" not in k}
    with pytest.raises(importer.ImportError_, match="missing variable generator/block_2/bn_1/beta"):
        importer.import_weights(missing, 21, 27)
    bad = dict(w)
    k = "final/full/flow_model_1/block_2_conv_2/kernel:0"
    bad[k] = bad[k][..., :-1]
    with pytest.raises(importer.ImportError_, match="shape"):
        importer.import_weights(bad, 21, 27)
    dup = dict(w)
    dup["other/flow_model_9/block_1/conv_1/kernel"] = w["final/full/flow_model_1/block_1_conv_1/kernel:0"]
    with pytest.raises(importer.ImportError_, match="two source variables"):
        importer.import_weights(dup, 21, 27)
    with pytest.raises(importer.ImportError_, match="no flow / generator"):
        importer.import_weights({"x/y": np.zeros(3)}, 21, 27)
    nan = dict(w)
    k = "final/full/generator_2/conv_1/kernel:0"
    nan[k] = nan[k].copy()
    nan[k][0, 0, 0, 0] = np.nan
    with pytest.raises(importer.ImportError_, match="non-finite"):
        importer.import_weights(nan, 21, 27)
    # frame size that the architecture cannot pool
    with pytest.raises(ValueError):
        importer.import_weights(w, 21, 27, flow_pad_factor=3)
