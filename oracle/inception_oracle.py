"""ORACLE — test infrastructure only, never the product path.

CPU (torch fp32 / numpy) restatement of the reference's classify_image hot path:

    JPEG-decoded uint8 HWC -> convert_image_dtype -> central_crop(0.875)
    -> resize_bilinear(299,299, align_corners=False) -> (x-0.5)*2
    -> Inception-v3 tower (inference mode) -> softmax -> top_k(5)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product package
(tensorflow_web_deploy_b200/) never does; it fails loudly without the CUDA
library.

PARITY PINNING STATUS: **pinned to the reference's own graph** (round 2).  The reference cannot be RUN here (the
arithmetic lives in the un-vendored, un-pinned `tensorflow` ~r0.9 wheel; Python 2; no checkpoint, reference
README.md:14) and its own tests assert shapes / variable counts only.  But its data/model.meta is a serialized
MetaGraphDef of the very graph PyClassification.py:148-164 builds, and tests/test_oracle_pin.py executes THAT graph
(tests/golden/reference_graph.json, decoded by tools/export_reference_graph.py) with the independent numpy
interpreter oracle/graph_interpreter.py and holds this oracle equal to it
  * bit-exactly on the preprocessing output (decode -> x/255 -> crop dim//16 -> legacy bilinear -> (x-.5)*2) for five
    image geometries, incl. the reference's 500x375 sample.jpg,
  * to 2e-5 relative on EVERY end point (conv0 ... mixed_8x8x2048b, aux_logits, logits, predictions) at C = 6 and 1001.
Wiring, op attributes (strides / padding / window sizes), the BatchNorm expansion, concat orders and every constant
therefore come from the reference's artefact.  The kernel semantics the graph cannot carry (legacy bilinear resize,
exclude-pad SAME average pooling, VALID/SAME max pooling) are held to known-answer values from TensorFlow's own unit
tests (tests/test_tf_kats.py).  Also pinned: the 388 variable names/shapes (tests/golden/model_meta_variables.txt),
collections_test.py:43-49,76-92 counts, inception_test.py:27-55,73-85 end-point shapes, ops_test.py:623-655 frozen BN.
What remains unpinned: TF's floating-point summation ORDER inside Conv2D/MatMul (Eigen), i.e. last-bit fp32 differences.

Each function cites the reference file:line it restates.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 0.001                     # inception/inception_model.py:73
U8_TO_F32 = np.float32(1.0 / 255)  # convert_image_dtype scale, PyClassification.py:152


# --------------------------------------------------------------------------
# preprocessing  (PyClassification.py:150-156, eval_image :180-203)
# --------------------------------------------------------------------------
def central_crop_box(h: int, w: int) -> Tuple[int, int, int, int]:
    """tf.image.central_crop(image, 0.875) (PyClassification.py:196):
    offset = dim // 16 (int((dim - dim*0.875)/2) with the 1/16 folded, as the
    graph in data/model.meta shows: Div/y=16, mul/y=2); size = dim - 2*offset.
    Returns (y0, x0, crop_h, crop_w)."""
    y0, x0 = h // 16, w // 16
    return y0, x0, h - 2 * y0, w - 2 * x0


def resize_bilinear_legacy(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """tf.image.resize_bilinear(align_corners=False) of TF1 (PyClassification.py:199-201):
    src = dst * (in/out), lo = floor(src), hi = min(lo+1, in-1), lerp = src-lo;
    no half-pixel centres, no antialias.  img: float32 [H,W,C]."""
    in_h, in_w = img.shape[:2]
    sy = np.float32(in_h) / np.float32(out_h)
    sx = np.float32(in_w) / np.float32(out_w)
    ys = np.arange(out_h, dtype=np.float32) * sy
    xs = np.arange(out_w, dtype=np.float32) * sx
    y_lo = np.floor(ys).astype(np.int64)
    x_lo = np.floor(xs).astype(np.int64)
    y_hi = np.minimum(y_lo + 1, in_h - 1)
    x_hi = np.minimum(x_lo + 1, in_w - 1)
    fy = (ys - y_lo.astype(np.float32))[:, None, None]
    fx = (xs - x_lo.astype(np.float32))[None, :, None]
    tl = img[y_lo][:, x_lo]
    tr = img[y_lo][:, x_hi]
    bl = img[y_hi][:, x_lo]
    br = img[y_hi][:, x_hi]
    top = tl + (tr - tl) * fx
    bot = bl + (br - bl) * fx
    return (top + (bot - top) * fy).astype(np.float32)


def preprocess_u8(rgb_u8: np.ndarray, size: int = 299) -> np.ndarray:
    """uint8 [H,W,3] (decoded JPEG) -> float32 [1,size,size,3] in [-1,1]
    (PyClassification.py:151-156)."""
    assert rgb_u8.dtype == np.uint8 and rgb_u8.ndim == 3 and rgb_u8.shape[2] == 3
    x = rgb_u8.astype(np.float32) * U8_TO_F32
    y0, x0, ch, cw = central_crop_box(*x.shape[:2])
    x = x[y0:y0 + ch, x0:x0 + cw]
    x = resize_bilinear_legacy(x, size, size)
    x = (x - np.float32(0.5)) * np.float32(2.0)
    return x[None]


# --------------------------------------------------------------------------
# layer math  (inception/slim/ops.py)
# --------------------------------------------------------------------------
class Tower:
    """Walks inception/slim/inception_model.py:86-329 on NCHW torch tensors.

    `weights`: dict name -> np.float32 array in the reference's layouts
    (conv HWIO, fc [in,out]).  `quant`: optional callable applied to every
    stored activation and to conv weights (used only to *simulate* the bf16/fp16
    storage of the CUDA path when studying tolerance; parity tests use None).
    """

    def __init__(self, weights: Dict[str, np.ndarray], quant=None, record: Optional[list] = None):
        self.w = weights
        self.q = quant if quant is not None else (lambda t: t)
        self.record = record
        self.end_points: Dict[str, torch.Tensor] = {}

    # ops.conv2d (ops.py:229-245) = tf.nn.conv2d, no bias -> batch_norm -> relu
    def conv(self, x, scope, cout, k, stride=1, same=True):
        kh, kw = k
        W = self._get(scope + "/weights", (kh, kw, x.shape[1], cout))
        Wt = self.q(torch.from_numpy(np.ascontiguousarray(W.transpose(3, 2, 0, 1))))
        pad = ((kh - 1) // 2, (kw - 1) // 2) if same else (0, 0)
        assert not (same and stride != 1), "tower only uses SAME with stride 1"
        y = F.conv2d(x, Wt, None, stride, pad)
        return self.q(self.bn_relu(y, scope))

    # ops.batch_norm inference branch (ops.py:122-128): frozen moving stats, no gamma
    def bn_relu(self, y, scope):
        c = y.shape[1]
        beta = torch.from_numpy(self._get(scope + "/BatchNorm/beta", (c,)))
        mean = torch.from_numpy(self._get(scope + "/BatchNorm/moving_mean", (c,)))
        var = torch.from_numpy(self._get(scope + "/BatchNorm/moving_variance", (c,)))
        r = torch.rsqrt(var + BN_EPS)
        y = y * r.view(1, c, 1, 1) + (beta - mean * r).view(1, c, 1, 1)
        return torch.relu(y)

    def max_pool(self, x):          # ops.py:367-370, always 3x3/2 VALID here
        return F.max_pool2d(x, 3, 2)

    def avg_pool_same(self, x):     # ops.py:397-400, 3x3/1 SAME: divisor = in-bounds taps
        return self.q(F.avg_pool2d(x, 3, 1, 1, count_include_pad=False))

    def fc(self, x, scope, cout):   # ops.fc -> tf.nn.xw_plus_b (ops.py:314)
        W = torch.from_numpy(self._get(scope + "/weights", (x.shape[1], cout)))
        b = torch.from_numpy(self._get(scope + "/biases", (cout,)))
        return x @ W + b

    def _get(self, name, shape):
        if self.record is not None:
            self.record.append((name, tuple(int(s) for s in shape)))
            if name not in self.w:
                return np.zeros(shape, np.float32) if not name.endswith("moving_variance") else np.ones(shape, np.float32)
        a = self.w[name]
        assert tuple(a.shape) == tuple(shape), (name, a.shape, shape)
        return a

    # ---------------------------------------------------------------- tower
    def forward(self, images_nhwc: np.ndarray, num_classes: int, with_aux: bool = True):
        """images_nhwc float32 [N,H,W,3] in [-1,1] -> (logits, predictions) np arrays;
        fills self.end_points (NHWC numpy) like the reference's end_points dict."""
        x = torch.from_numpy(np.ascontiguousarray(images_nhwc.transpose(0, 3, 1, 2)))
        ep = self.end_points
        with torch.no_grad():
            # stem, inception_model.py:86-107 (default stride 1, VALID)
            x = ep["conv0"] = self.conv(x, "conv0", 32, (3, 3), 2, False)
            x = ep["conv1"] = self.conv(x, "conv1", 32, (3, 3), 1, False)
            x = ep["conv2"] = self.conv(x, "conv2", 64, (3, 3), 1, True)
            x = ep["pool1"] = self.max_pool(x)
            x = ep["conv3"] = self.conv(x, "conv3", 80, (1, 1), 1, False)
            x = ep["conv4"] = self.conv(x, "conv4", 192, (3, 3), 1, False)
            x = ep["pool2"] = self.max_pool(x)
            # 35x35 blocks, :112-158
            for name, pool_c in (("mixed_35x35x256a", 32), ("mixed_35x35x288a", 64), ("mixed_35x35x288b", 64)):
                x = ep[name] = self._block35(x, name, pool_c)
            x = ep["mixed_17x17x768a"] = self._block_red_a(x, "mixed_17x17x768a")          # :160-171
            for name, m in (("mixed_17x17x768b", 128), ("mixed_17x17x768c", 160),
                            ("mixed_17x17x768d", 160), ("mixed_17x17x768e", 192)):         # :173-247
                x = ep[name] = self._block17(x, name, m)
            if with_aux:
                ep["aux_logits"] = self._aux(x, num_classes)                               # :249-261
            x = ep["mixed_17x17x1280a"] = self._block_red_b(x, "mixed_17x17x1280a")        # :266-280
            x = ep["mixed_8x8x2048a"] = self._block8(x, "mixed_8x8x2048a")                 # :282-298
            x = ep["mixed_8x8x2048b"] = self._block8(x, "mixed_8x8x2048b")                 # :300-316
            # head, :318-329: avg_pool(shape[1:3], VALID) -> dropout(identity) -> flatten -> fc -> softmax
            feat = x.mean(dim=(2, 3))
            logits = ep["logits"] = self.fc(feat, "logits/logits", num_classes)
            pred = ep["predictions"] = torch.softmax(logits, dim=1)
        out = {}
        for k, v in ep.items():
            a = v.numpy()
            out[k] = np.ascontiguousarray(a.transpose(0, 2, 3, 1)) if a.ndim == 4 else a
        self.end_points = out
        return out["logits"], out["predictions"]

    def _block35(self, x, s, pool_c):
        b0 = self.conv(x, s + "/branch1x1/Conv", 64, (1, 1))
        b1 = self.conv(x, s + "/branch5x5/Conv", 48, (1, 1))
        b1 = self.conv(b1, s + "/branch5x5/Conv_1", 64, (5, 5))
        b2 = self.conv(x, s + "/branch3x3dbl/Conv", 64, (1, 1))
        b2 = self.conv(b2, s + "/branch3x3dbl/Conv_1", 96, (3, 3))
        b2 = self.conv(b2, s + "/branch3x3dbl/Conv_2", 96, (3, 3))
        b3 = self.conv(self.avg_pool_same(x), s + "/branch_pool/Conv", pool_c, (1, 1))
        return torch.cat([b0, b1, b2, b3], 1)

    def _block_red_a(self, x, s):
        b0 = self.conv(x, s + "/branch3x3/Conv", 384, (3, 3), 2, False)
        b1 = self.conv(x, s + "/branch3x3dbl/Conv", 64, (1, 1))
        b1 = self.conv(b1, s + "/branch3x3dbl/Conv_1", 96, (3, 3))
        b1 = self.conv(b1, s + "/branch3x3dbl/Conv_2", 96, (3, 3), 2, False)
        return torch.cat([b0, b1, self.max_pool(x)], 1)

    def _block17(self, x, s, m):
        b0 = self.conv(x, s + "/branch1x1/Conv", 192, (1, 1))
        b1 = self.conv(x, s + "/branch7x7/Conv", m, (1, 1))
        b1 = self.conv(b1, s + "/branch7x7/Conv_1", m, (1, 7))
        b1 = self.conv(b1, s + "/branch7x7/Conv_2", 192, (7, 1))
        b2 = self.conv(x, s + "/branch7x7dbl/Conv", m, (1, 1))
        b2 = self.conv(b2, s + "/branch7x7dbl/Conv_1", m, (7, 1))
        b2 = self.conv(b2, s + "/branch7x7dbl/Conv_2", m, (1, 7))
        b2 = self.conv(b2, s + "/branch7x7dbl/Conv_3", m, (7, 1))
        b2 = self.conv(b2, s + "/branch7x7dbl/Conv_4", 192, (1, 7))
        b3 = self.conv(self.avg_pool_same(x), s + "/branch_pool/Conv", 192, (1, 1))
        return torch.cat([b0, b1, b2, b3], 1)

    def _aux(self, x, num_classes):
        a = self.q(F.avg_pool2d(x, 5, 3))
        a = self.conv(a, "aux_logits/proj", 128, (1, 1))
        a = self.conv(a, "aux_logits/Conv", 768, (a.shape[2], a.shape[3]), 1, False)
        return self.fc(a.flatten(1), "aux_logits/FC", num_classes)

    def _block_red_b(self, x, s):
        b0 = self.conv(x, s + "/branch3x3/Conv", 192, (1, 1))
        b0 = self.conv(b0, s + "/branch3x3/Conv_1", 320, (3, 3), 2, False)
        b1 = self.conv(x, s + "/branch7x7x3/Conv", 192, (1, 1))
        b1 = self.conv(b1, s + "/branch7x7x3/Conv_1", 192, (1, 7))
        b1 = self.conv(b1, s + "/branch7x7x3/Conv_2", 192, (7, 1))
        b1 = self.conv(b1, s + "/branch7x7x3/Conv_3", 192, (3, 3), 2, False)
        return torch.cat([b0, b1, self.max_pool(x)], 1)

    def _block8(self, x, s):
        b0 = self.conv(x, s + "/branch1x1/Conv", 320, (1, 1))
        b1 = self.conv(x, s + "/branch3x3/Conv", 384, (1, 1))
        b1 = torch.cat([self.conv(b1, s + "/branch3x3/Conv_1", 384, (1, 3)),
                        self.conv(b1, s + "/branch3x3/Conv_2", 384, (3, 1))], 1)
        b2 = self.conv(x, s + "/branch3x3dbl/Conv", 448, (1, 1))
        b2 = self.conv(b2, s + "/branch3x3dbl/Conv_1", 384, (3, 3))
        b2 = torch.cat([self.conv(b2, s + "/branch3x3dbl/Conv_2", 384, (1, 3)),
                        self.conv(b2, s + "/branch3x3dbl/Conv_3", 384, (3, 1))], 1)
        b3 = self.conv(self.avg_pool_same(x), s + "/branch_pool/Conv", 192, (1, 1))
        return torch.cat([b0, b1, b2, b3], 1)


def top_k(pred: np.ndarray, k: int = 5):
    """tf.nn.top_k(predictions, 5) (PyClassification.py:163-164): descending,
    ties -> lower index first.  pred [N,C] -> (values [N,k] f32, indices [N,k] i32)."""
    order = np.argsort(-pred, axis=1, kind="stable")[:, :k]
    vals = np.take_along_axis(pred, order, axis=1)
    return vals.astype(np.float32), order.astype(np.int32)


def recorded_variables(num_classes: int, size: int = 299) -> List[Tuple[str, Tuple[int, ...]]]:
    """Names/shapes this restatement asks for, in graph order (structure check)."""
    rec: list = []
    Tower({}, record=rec).forward(np.zeros((1, size, size, 3), np.float32), num_classes)
    return rec


class OracleClassifier:
    """CPU mirror of ImagenetClassifier (PyClassification.py:133-218) on the oracle."""

    def __init__(self, weights: Dict[str, np.ndarray], label_names: List[str], num_classes: int):
        self.weights, self.label_names, self.num_classes = weights, label_names, num_classes

    def predictions(self, images_nhwc: np.ndarray) -> np.ndarray:
        return Tower(self.weights).forward(images_nhwc, self.num_classes, with_aux=False)[1]

    def classify_decoded(self, rgb_u8: np.ndarray):
        pred = self.predictions(preprocess_u8(rgb_u8))
        vals, idx = top_k(pred, min(5, self.num_classes))
        vals, idx = vals[0], idx[0]
        return [self.label_names[i] for i in idx[:3]], vals[:3], float(vals.sum())
