"""Known-answer vectors PUBLISHED by TensorFlow / ONNX for the generic primitives whose arithmetic
lives outside /root/reference (TensorFlow 2.18, tf2onnx -> ONNX Resize), pinned
against both oracle restatements (oracle/reference_graph.py, oracle/naive.py).

There is no network in the build container, so the vectors are transcribed from the published
sources named in each test; every one of them is small enough to verify by hand against the
operator definition quoted next to it."""

import numpy as np
import pytest
import torch

from oracle import naive
from oracle import reference_graph as og


def _both(fn_name, x, *args):
    a = getattr(og, fn_name)(torch.from_numpy(np.asarray(x, np.float32)), *args).numpy()
    # the naive restatement works on one [H, W, C] image
    b = np.stack([np.asarray(getattr(naive, fn_name)(img, *args)) for img in np.asarray(x, np.float32)])
    return a, b


# tf.nn.depth_to_space API documentation (TensorFlow 2.x "tf.nn.depth_to_space", NHWC, the three
# worked examples of the doc string in tensorflow/core/api_def/base_api/api_def_DepthToSpace.pbtxt)
DEPTH_TO_SPACE_DOC = [
    # [1,1,1,4], block 2 -> [1,2,2,1]
    ([[[[1, 2, 3, 4]]]], 2, [[[[1], [2]], [[3], [4]]]]),
    # [1,1,1,12], block 2 -> [1,2,2,3]
    ([[[[1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12]]]], 2,
     [[[[1, 2, 3], [4, 5, 6]], [[7, 8, 9], [10, 11, 12]]]]),
    # [1,2,2,4], block 2 -> [1,4,4,1]
    ([[[[1, 2, 3, 4], [5, 6, 7, 8]], [[9, 10, 11, 12], [13, 14, 15, 16]]]], 2,
     [[[[1], [2], [5], [6]], [[3], [4], [7], [8]], [[9], [10], [13], [14]], [[11], [12], [15], [16]]]]),
]


@pytest.mark.parametrize("x,block,want", DEPTH_TO_SPACE_DOC)
def test_depth_to_space_matches_the_tf_documentation(x, block, want):
    # used by the flow head: DepthToSpace(4), scripts/training/keras_layers.py:162-175
    a, b = _both("depth_to_space", x, block)
    np.testing.assert_array_equal(a, np.asarray(want, np.float32))
    np.testing.assert_array_equal(b, np.asarray(want, np.float32))


@pytest.mark.parametrize("want,block,x", DEPTH_TO_SPACE_DOC)
def test_space_to_depth_matches_the_tf_documentation(want, block, x):
    # tf.nn.space_to_depth API documentation (api_def_SpaceToDepth.pbtxt): the same three examples
    # read in the other direction; used by the generator input: SpaceToDepth(4), keras_layers.py:116-129
    a, b = _both("space_to_depth", x, block)
    np.testing.assert_array_equal(a, np.asarray(want, np.float32))
    np.testing.assert_array_equal(b, np.asarray(want, np.float32))


def test_legacy_bilinear_matches_the_tf_unit_test_vector():
    """tensorflow/python/ops/image_ops_test.py, ResizeImagesTest.testResizeUp (ResizeMethodV1.BILINEAR,
    i.e. tf.compat.v1.image.resize_bilinear with align_corners=False, half_pixel_centers=False - the
    call keras_layers.py:47-52 makes): a 3x2 image resized to 6x4."""
    img = np.array([64, 32, 32, 64, 50, 100], np.float32).reshape(1, 3, 2, 1)
    want = np.array([64.0, 48.0, 32.0, 32.0, 48.0, 48.0, 48.0, 48.0, 32.0, 48.0, 64.0, 64.0,
                     41.0, 61.5, 82.0, 82.0, 50.0, 75.0, 100.0, 100.0, 50.0, 75.0, 100.0, 100.0],
                    np.float32).reshape(1, 6, 4, 1)
    a, b = _both("resize_bilinear_legacy", img, 2)
    np.testing.assert_array_equal(a, want)
    np.testing.assert_array_equal(b, want)


def test_legacy_bilinear_matches_the_onnx_asymmetric_resize_example():
    """ONNX Resize-10 documentation, example `resize_upsample_linear` (opset 10 'linear' = what later
    opsets call coordinate_transformation_mode="asymmetric", the mode tf2onnx exports the legacy TF
    resize to - the deployed engines run this operator): [[1,2],[3,4]] scaled by 2."""
    img = np.array([[1, 2], [3, 4]], np.float32).reshape(1, 2, 2, 1)
    want = np.array([[1.0, 1.5, 2.0, 2.0], [2.0, 2.5, 3.0, 3.0], [3.0, 3.5, 4.0, 4.0], [3.0, 3.5, 4.0, 4.0]],
                    np.float32).reshape(1, 4, 4, 1)
    a, b = _both("resize_bilinear_legacy", img, 2)
    np.testing.assert_array_equal(a, want)
    np.testing.assert_array_equal(b, want)


def test_scale_4_weights_are_quarters_and_the_far_edge_replicates():
    # SURVEY appendix A.5 restated as a property of the x4 input upscale (models.py:584-587)
    row = np.array([0.0, 4.0, 8.0], np.float32).reshape(1, 1, 3, 1)
    a, b = _both("resize_bilinear_legacy", row, 4)
    want = np.array([0, 1, 2, 3, 4, 5, 6, 7, 8, 8, 8, 8], np.float32)
    for got in (a, b):
        assert got.shape == (1, 4, 12, 1)
        for r in range(4):
            np.testing.assert_array_equal(got[0, r, :, 0], want)


def test_maxpool_and_truncating_cast_examples():
    # tf.keras.layers.MaxPool2D documentation example (pool 2x2, strides 2, valid) on a 4x4 ramp:
    # [[1,2,3,4],[5,6,7,8],[9,10,11,12],[13,14,15,16]] -> [[6, 8], [14, 16]]
    x = np.arange(1, 17, dtype=np.float32).reshape(1, 4, 4, 1)
    a, b = _both("max_pool2", x)
    want = np.array([[6, 8], [14, 16]], np.float32).reshape(1, 2, 2, 1)
    np.testing.assert_array_equal(a, want)
    np.testing.assert_array_equal(b, want)
    # tf.cast float -> uint8 truncates toward zero (PostprocessLayer, keras_layers.py:227-230)
    v = torch.tensor([[-0.5, -0.4981, 0.0, 0.4999, 0.5]])
    got = og.postprocess(v).numpy().ravel().tolist()
    assert got == [0, 0, 127, 254, 255]
