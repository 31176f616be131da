"""Known-answer tests from TensorFlow's own unit tests for the kernel semantics the reference's source does
not show (they live in the un-vendored tensorflow dependency): legacy bilinear resize, exclude-pad SAME
average pooling, VALID/SAME max pooling.  Both CPU implementations used by this repo's tests are checked:
the numpy kernels of oracle/graph_interpreter.py and the torch oracle.

Values: tensorflow/python/ops/image_ops_test.py (ResizeImagesTest.testResizeDown / testResizeUp, BILINEAR rows)
and tensorflow/python/kernel_tests/pooling_ops_test.py (_testAvgPoolValidPadding, _testAvgPoolSamePadding,
_testAvgPoolSameNonSquareWindow, _testMaxPoolValidPadding, _testMaxPoolSamePadding) of the r0.9 tree; each is
also re-derivable by hand from the op definition, which the comments do.
"""
import numpy as np
import torch

from oracle import graph_interpreter as GI
from oracle import inception_oracle as O


def _iota(shape):
    return np.arange(1, int(np.prod(shape)) + 1, dtype=np.float32).reshape(shape)


def test_resize_bilinear_tf_image_ops_test_values():
    down_in = np.array([128, 128, 64, 64, 128, 128, 64, 64, 64, 64, 128, 128, 64, 64, 128, 128,
                        50, 50, 100, 100, 50, 50, 100, 100], np.float32).reshape(1, 6, 4, 1)
    down_out = np.array([128, 64, 64, 128, 50, 100], np.float32).reshape(1, 3, 2, 1)
    up_in = down_out
    up_out = np.array([128.0, 96.0, 64.0, 64.0, 96.0, 96.0, 96.0, 96.0, 64.0, 96.0, 128.0, 128.0,
                       57.0, 85.5, 114.0, 114.0, 50.0, 75.0, 100.0, 100.0, 50.0, 75.0, 100.0, 100.0],
                      np.float32).reshape(1, 6, 4, 1)
    np.testing.assert_array_equal(GI.resize_bilinear(down_in, (3, 2)), down_out)
    np.testing.assert_array_equal(GI.resize_bilinear(up_in, (6, 4)), up_out)
    np.testing.assert_array_equal(O.resize_bilinear_legacy(down_in[0], 3, 2), down_out[0])
    np.testing.assert_array_equal(O.resize_bilinear_legacy(up_in[0], 6, 4), up_out[0])
    # a half-pixel-centre resize (torch / cv2 / PIL / TF2) does NOT give these values
    hp = torch.nn.functional.interpolate(torch.from_numpy(up_in.transpose(0, 3, 1, 2).copy()), size=(6, 4), mode="bilinear",
                                         align_corners=False).numpy().transpose(0, 2, 3, 1)
    assert np.abs(hp - up_out).max() > 1.0


def test_avg_pool_tf_pooling_ops_test_values():
    # VALID 2x2/2 on [1,3,3,3] of 1..27: mean of pixels (0,0),(0,1),(1,0),(1,1) -> 7, 8, 9
    np.testing.assert_array_equal(GI.avg_pool(_iota((1, 3, 3, 3)), [1, 2, 2, 1], [1, 2, 2, 1], "VALID").ravel(), [7.0, 8.0, 9.0])
    # SAME 2x2/2 on [1,2,4,3]
    np.testing.assert_array_equal(GI.avg_pool(_iota((1, 2, 4, 3)), [1, 2, 2, 1], [1, 2, 2, 1], "SAME").ravel(),
                                  [8.5, 9.5, 10.5, 14.5, 15.5, 16.5])
    # SAME 1x2/1 on [[1,2],[3,4]]: the right-hand windows hold ONE in-bounds tap and divide by 1, not 2
    np.testing.assert_array_equal(GI.avg_pool(_iota((1, 2, 2, 1)), [1, 1, 2, 1], [1, 1, 1, 1], "SAME").ravel(),
                                  [1.5, 2.0, 3.5, 4.0])
    # the tower's only SAME pool, 3x3/1, on ones: every output is exactly 1 (corners 4/4, edges 6/6)
    ones = np.ones((2, 5, 7, 8), np.float32)
    np.testing.assert_array_equal(GI.avg_pool(ones, [1, 3, 3, 1], [1, 1, 1, 1], "SAME"), ones)
    t = O.Tower({}).avg_pool_same(torch.ones(2, 8, 5, 7))
    np.testing.assert_array_equal(t.numpy(), 1.0)
    # ... and the torch oracle's 3x3 SAME pool equals the numpy kernel on random data
    x = np.random.default_rng(0).standard_normal((2, 9, 11, 16)).astype(np.float32)
    a = GI.avg_pool(x, [1, 3, 3, 1], [1, 1, 1, 1], "SAME")
    b = O.Tower({}).avg_pool_same(torch.from_numpy(x.transpose(0, 3, 1, 2).copy())).numpy().transpose(0, 2, 3, 1)
    np.testing.assert_allclose(a, b, atol=1e-6)


def test_max_pool_tf_pooling_ops_test_values():
    np.testing.assert_array_equal(GI.max_pool(_iota((1, 3, 3, 3)), [1, 2, 2, 1], [1, 2, 2, 1], "VALID").ravel(),
                                  [13.0, 14.0, 15.0])
    np.testing.assert_array_equal(GI.max_pool(_iota((1, 2, 3, 3)), [1, 2, 2, 1], [1, 2, 2, 1], "SAME").ravel(),
                                  [13.0, 14.0, 15.0, 16.0, 17.0, 18.0])
    x = np.random.default_rng(1).standard_normal((2, 17, 17, 8)).astype(np.float32)
    a = GI.max_pool(x, [1, 3, 3, 1], [1, 2, 2, 1], "VALID")
    b = O.Tower({}).max_pool(torch.from_numpy(x.transpose(0, 3, 1, 2).copy())).numpy().transpose(0, 2, 3, 1)
    np.testing.assert_array_equal(a, b)


def test_conv2d_same_and_valid_against_direct_sums():
    """tf.nn.conv2d is a cross-correlation (no kernel flip) with the TF SAME rule; checked against scalar loops."""
    rng = np.random.default_rng(2)
    x = rng.standard_normal((1, 5, 6, 3)).astype(np.float32)
    w = rng.standard_normal((3, 3, 3, 4)).astype(np.float32)
    for padding, stride in (("SAME", 1), ("VALID", 1), ("VALID", 2)):
        got = GI.conv2d(x, w, [1, stride, stride, 1], padding)
        pad = 1 if padding == "SAME" else 0
        xp = np.pad(x, ((0, 0), (pad, pad), (pad, pad), (0, 0)))
        oh, ow = got.shape[1:3]
        ref = np.zeros_like(got)
        for oy in range(oh):
            for ox in range(ow):
                for r in range(3):
                    for s in range(3):
                        ref[0, oy, ox] += xp[0, oy * stride + r, ox * stride + s] @ w[r, s]
        np.testing.assert_allclose(got, ref, atol=1e-5)
