"""Known-answer tests of the CPU oracle's TF-semantics and its pin to the committed fixtures."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import inception_oracle as O


def test_central_crop_box_sample_jpg_geometry():
    # the reference's data/sample.jpg is 500x375: crop 438x329 at (x0=31, y0=23)  
    assert O.central_crop_box(375, 500) == (23, 31, 329, 438)
    assert O.central_crop_box(299, 299) == (18, 18, 263, 263)
    assert O.central_crop_box(16, 17) == (1, 1, 14, 15)


def _resize_loops(img, oh, ow):
    ih, iw, c = img.shape
    out = np.zeros((oh, ow, c), np.float32)
    sy, sx = np.float32(ih) / np.float32(oh), np.float32(iw) / np.float32(ow)
    for y in range(oh):
        fy = np.float32(y) * sy
        y0 = int(np.floor(fy)); y1 = min(y0 + 1, ih - 1); ly = np.float32(fy - y0)
        for x in range(ow):
            fx = np.float32(x) * sx
            x0 = int(np.floor(fx)); x1 = min(x0 + 1, iw - 1); lx = np.float32(fx - x0)
            top = img[y0, x0] + (img[y0, x1] - img[y0, x0]) * lx
            bot = img[y1, x0] + (img[y1, x1] - img[y1, x0]) * lx
            out[y, x] = top + (bot - top) * ly
    return out


def test_resize_bilinear_legacy_matches_scalar_loops():
    rng = np.random.default_rng(0)
    img = rng.uniform(0, 1, (13, 9, 3)).astype(np.float32)
    np.testing.assert_array_equal(O.resize_bilinear_legacy(img, 7, 11), _resize_loops(img, 7, 11))
    # identity when sizes match; no half-pixel shift: output[0,0] is input[0,0]
    np.testing.assert_array_equal(O.resize_bilinear_legacy(img, 13, 9), img)
    up = O.resize_bilinear_legacy(img, 26, 18)
    np.testing.assert_array_equal(up[::2, ::2], img)           # src = dst * 0.5 lands on pixels
    assert not np.allclose(up, torch.nn.functional.interpolate(
        torch.from_numpy(img.transpose(2, 0, 1))[None], (26, 18), mode="bilinear",
        align_corners=False)[0].numpy().transpose(1, 2, 0))    # differs from half-pixel resize


def test_preprocess_range_and_shape():
    img = np.random.default_rng(1).integers(0, 256, (375, 500, 3), dtype=np.uint8)
    x = O.preprocess_u8(img)
    assert x.shape == (1, 299, 299, 3) and x.dtype == np.float32
    assert -1.0 <= x.min() and x.max() <= 1.0
    white = O.preprocess_u8(np.full((64, 80, 3), 255, np.uint8))
    np.testing.assert_allclose(white, 1.0, atol=1e-6)
    black = O.preprocess_u8(np.zeros((64, 80, 3), np.uint8))
    np.testing.assert_array_equal(black, -1.0)


def test_avg_pool_same_excludes_padding():
    # tf.nn.avg_pool SAME divides by the in-bounds tap count: ones stay ones (corner 4/4, not 4/9)
    tw = O.Tower({})
    y = tw.avg_pool_same(torch.ones(1, 2, 5, 5))
    np.testing.assert_array_equal(y.numpy(), 1.0)
    x = torch.arange(9, dtype=torch.float32).view(1, 1, 3, 3)
    y = tw.avg_pool_same(x)[0, 0]
    assert y[0, 0].item() == pytest.approx((0 + 1 + 3 + 4) / 4)
    assert y[1, 1].item() == pytest.approx(4.0)
    assert y[0, 1].item() == pytest.approx((0 + 1 + 2 + 3 + 4 + 5) / 6)


def test_batch_norm_inference_has_no_gamma_and_frozen_stats():
    # ops.py:122-128 + ops_test.py:623-655: y = (x - moving_mean) * rsqrt(moving_var + 0.001) + beta
    w = {"s/BatchNorm/beta": np.array([0.5, -1.0], np.float32),
         "s/BatchNorm/moving_mean": np.array([1.0, 2.0], np.float32),
         "s/BatchNorm/moving_variance": np.array([4.0, 0.25], np.float32)}
    x = torch.tensor([[[[3.0]], [[1.0]]]])
    y = O.Tower(w).bn_relu(x, "s")
    exp0 = (3.0 - 1.0) / np.sqrt(4.0 + 0.001) + 0.5
    exp1 = max((1.0 - 2.0) / np.sqrt(0.25 + 0.001) - 1.0, 0.0)
    np.testing.assert_allclose(y.numpy().ravel(), [exp0, exp1], rtol=1e-6)


def test_conv_is_cross_correlation_hwio():
    w = np.zeros((1, 3, 1, 1), np.float32); w[0, 0, 0, 0] = 1.0        # picks the LEFT neighbour
    W = {"c/weights": w, "c/BatchNorm/beta": np.zeros(1, np.float32),
         "c/BatchNorm/moving_mean": np.zeros(1, np.float32),
         "c/BatchNorm/moving_variance": np.full(1, 1.0 - 0.001, np.float32)}
    x = torch.arange(1, 6, dtype=torch.float32).view(1, 1, 1, 5)
    y = O.Tower(W).conv(x, "c", 1, (1, 3))
    np.testing.assert_allclose(y.numpy().ravel(), [0, 1, 2, 3, 4], rtol=1e-6)   # SAME pad 1 on W only


def test_top_k_order_and_ties():
    p = np.array([[0.1, 0.4, 0.4, 0.05, 0.05, 0.0]], np.float32)
    v, i = O.top_k(p, 5)
    assert i.tolist() == [[1, 2, 0, 3, 4]] and v.dtype == np.float32 and i.dtype == np.int32
    np.testing.assert_array_equal(v[0], p[0, [1, 2, 0, 3, 4]])


def test_oracle_reproduces_committed_golden(synth6):
    _, W = synth6
    g = np.load(os.path.join(GOLDEN, "oracle_c6.npz"))
    x = np.random.default_rng(0).uniform(-1, 1, (2, 299, 299, 3)).astype(np.float32)
    tw = O.Tower(W)
    logits, pred = tw.forward(x, 6, with_aux=True)
    # thread count / oneDNN kernel choice may reorder fp32 sums: allow 1e-4 on O(1) logits
    np.testing.assert_allclose(logits, g["logits"], atol=2e-4)
    np.testing.assert_allclose(pred, g["predictions"], atol=5e-5)
    np.testing.assert_allclose(tw.end_points["aux_logits"], g["aux_logits"], atol=2e-4)
    v, i = O.top_k(pred, 5)
    np.testing.assert_array_equal(i, g["topk_idx"])
    for k in ("conv0", "pool2", "mixed_35x35x288b", "mixed_17x17x768e", "mixed_8x8x2048b"):
        a = tw.end_points[k]
        assert tuple(g["ep_shape/" + k]) == a.shape
        idx = (np.arange(64, dtype=np.int64) * 2654435761 % a.size)
        np.testing.assert_allclose(a.reshape(-1)[idx], g["ep_samples/" + k], atol=2e-3, rtol=1e-4)
    img = g["pre_img"]
    pre = O.preprocess_u8(img)
    idx = (np.arange(256, dtype=np.int64) * 2654435761 % pre.size)
    np.testing.assert_array_equal(pre.reshape(-1)[idx], g["pre_samples"])
