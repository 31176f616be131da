"""ORACLE PIN — test infrastructure only, never the product path.

A numpy interpreter for the reference's OWN inference graph.

tests/golden/reference_graph.json is the subgraph `DecodeJpeg -> ... -> inception_v3/logits/predictions`
decoded from the reference's data/model.meta (tools/export_reference_graph.py): 1742 NodeDefs with
their inputs, attributes and Const tensors — the very graph /root/reference/PyClassification.py:148-164
builds (its 4001-class variant).  Running it gives a second, independently produced answer for every
end point: the WIRING (op order, concat order, strides / paddings / window sizes, the BatchNorm
expansion Add -> Rsqrt -> Mul, Mul -> Sub -> Add, the central-crop arithmetic Div 16 / Mul 2 / Sub /
Pack / Slice, the constants 1/255, 0.5, 2.0, eps) comes from the reference's artefact, not from a
reading of its Python; the KERNELS below are plain numpy written from the TensorFlow (r0.9) op
definitions, sharing no code with oracle/inception_oracle.py (which is torch):

  Conv2D         NHWC x HWIO cross-correlation, VALID / SAME (TF pad rule: total = max((o-1)*s + k - i, 0),
                 before = total // 2), one matmul per filter tap
  MaxPool/AvgPool  windows clipped to the image; AvgPool divides by the number of in-bounds taps
  ResizeBilinear align_corners=False legacy kernel: src = dst * in/out, no half-pixel centres
  Softmax        max-subtracted
and the trivial ones (Cast, Mul, Sub, Add, Div, Rsqrt, Relu, Shape, Slice, Squeeze, Pack, ExpandDims,
Concat, Reshape, MatMul, BiasAdd, Identity, Const, Variable).

tests/test_oracle_pin.py asserts oracle == this interpreter on the preprocessing output and on every
end point of the tower; tests/test_tf_kats.py checks the kernels here against known-answer values
published in TensorFlow's own unit tests.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Iterable, Optional

import numpy as np

GRAPH_JSON = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                          "reference_graph.json")

_NP = {"float32": np.float32, "float64": np.float64, "int32": np.int32, "int64": np.int64, "uint8": np.uint8,
       "bool": np.bool_}


# ------------------------------------------------------------------------------------------ kernels
def _same_pads(i: int, k: int, s: int):
    o = -(-i // s)
    total = max((o - 1) * s + k - i, 0)
    return o, total // 2, total - total // 2


def conv2d(x: np.ndarray, w: np.ndarray, strides, padding: str) -> np.ndarray:
    """tf.nn.conv2d, NHWC x HWIO, no kernel flip."""
    n, ih, iw, ci = x.shape
    kh, kw, ci2, co = w.shape
    assert ci == ci2 and strides[0] == 1 and strides[3] == 1
    sh, sw = strides[1], strides[2]
    if padding == "SAME":
        oh, pt, pb = _same_pads(ih, kh, sh)
        ow, pl, pr = _same_pads(iw, kw, sw)
        x = np.pad(x, ((0, 0), (pt, pb), (pl, pr), (0, 0)))
    else:
        oh, ow = (ih - kh) // sh + 1, (iw - kw) // sw + 1
    out = np.zeros((n * oh * ow, co), np.float32)
    for r in range(kh):
        for s in range(kw):
            patch = x[:, r:r + (oh - 1) * sh + 1:sh, s:s + (ow - 1) * sw + 1:sw, :]
            out += np.ascontiguousarray(patch).reshape(-1, ci) @ w[r, s]
    return out.reshape(n, oh, ow, co)


def _pool(x: np.ndarray, ksize, strides, padding: str, kind: str) -> np.ndarray:
    n, ih, iw, c = x.shape
    kh, kw, sh, sw = ksize[1], ksize[2], strides[1], strides[2]
    if padding == "SAME":
        oh, pt, _ = _same_pads(ih, kh, sh)
        ow, pl, _ = _same_pads(iw, kw, sw)
    else:
        oh, ow, pt, pl = (ih - kh) // sh + 1, (iw - kw) // sw + 1, 0, 0
    out = np.empty((n, oh, ow, c), np.float32)
    for oy in range(oh):
        y0, y1 = max(oy * sh - pt, 0), min(oy * sh - pt + kh, ih)
        for ox in range(ow):
            x0, x1 = max(ox * sw - pl, 0), min(ox * sw - pl + kw, iw)
            win = x[:, y0:y1, x0:x1, :]
            if kind == "max":
                out[:, oy, ox, :] = win.max(axis=(1, 2))
            else:                                   # divisor = taps inside the image (TF AvgPool)
                out[:, oy, ox, :] = win.sum(axis=(1, 2), dtype=np.float32) / np.float32((y1 - y0) * (x1 - x0))
    return out


def max_pool(x, ksize, strides, padding):
    return _pool(x, ksize, strides, padding, "max")


def avg_pool(x, ksize, strides, padding):
    return _pool(x, ksize, strides, padding, "avg")


def resize_bilinear(x: np.ndarray, size, align_corners: bool = False) -> np.ndarray:
    """TF1 ResizeBilinear (legacy, no half-pixel centres).  x float32 [N,H,W,C]."""
    n, ih, iw, c = x.shape
    oh, ow = int(size[0]), int(size[1])
    if align_corners and oh > 1:
        hs, ws = np.float32(ih - 1) / np.float32(oh - 1), np.float32(iw - 1) / np.float32(ow - 1)
    else:
        hs, ws = np.float32(ih) / np.float32(oh), np.float32(iw) / np.float32(ow)
    out = np.empty((n, oh, ow, c), np.float32)
    for oy in range(oh):
        fy = np.float32(oy) * hs
        y0 = int(np.floor(fy))
        y1 = min(y0 + 1, ih - 1)
        ly = np.float32(fy - np.float32(y0))
        for ox in range(ow):
            fx = np.float32(ox) * ws
            x0 = int(np.floor(fx))
            x1 = min(x0 + 1, iw - 1)
            lx = np.float32(fx - np.float32(x0))
            top = x[:, y0, x0] + (x[:, y0, x1] - x[:, y0, x0]) * lx
            bot = x[:, y1, x0] + (x[:, y1, x1] - x[:, y1, x0]) * lx
            out[:, oy, ox] = top + (bot - top) * ly
    return out


def softmax(x: np.ndarray) -> np.ndarray:
    e = np.exp(x - x.max(axis=-1, keepdims=True))
    return (e / e.sum(axis=-1, keepdims=True)).astype(np.float32)


# ------------------------------------------------------------------------------------------ interpreter
class Graph:
    def __init__(self, path: str = GRAPH_JSON):
        doc = json.load(open(path))
        self.input, self.fetches = doc["input"], doc["fetches"]
        self.nodes = {n["name"]: n for n in doc["nodes"]}
        self.order = [n["name"] for n in doc["nodes"]]

    def variables(self):
        """[(name, shape)] of every Variable node, in graph order."""
        return [(n, tuple(self.nodes[n]["attr"]["shape"])) for n in self.order if self.nodes[n]["op"] == "Variable"]

    def run(self, fetches: Iterable[str], feeds: Dict[str, np.ndarray], weights: Dict[str, np.ndarray],
            cache: Optional[dict] = None) -> Dict[str, np.ndarray]:
        """Evaluates `fetches` (node names).  `feeds` overrides any node's output (e.g. 'DecodeJpeg' with a
        decoded uint8 image, or 'ExpandDims_1' with a preprocessed batch); `weights` maps Variable names to arrays
        (a Variable whose declared shape differs in the class dimension takes the array's shape: the reference
        graph is the 4001-class variant of the 6-class one the app builds)."""
        vals = dict(feeds) if cache is None else cache
        if cache is not None:
            vals.update(feeds)
        out = {}
        for f in fetches:
            out[f] = self._eval(f, vals, weights)
        return out

    def _eval(self, ref: str, vals: dict, weights) -> np.ndarray:
        name = ref.split(":")[0]
        if name in vals:
            return vals[name]
        # explicit stack: the graph is ~500 ops deep
        stack = [name]
        while stack:
            cur = stack[-1]
            if cur in vals:
                stack.pop()
                continue
            node = self.nodes[cur]
            deps = [i.split(":")[0] for i in node["inputs"]]
            missing = [d for d in deps if d not in vals]
            if missing:
                stack.extend(missing)
                continue
            vals[cur] = self._apply(node, [vals[d] for d in deps], weights)
            stack.pop()
        return vals[name]

    @staticmethod
    def _apply(node, a, weights):
        op, at = node["op"], node.get("attr", {})
        if op == "Const":
            t = at["value"]
            return np.array(t["values"], dtype=_NP[t["dtype"]]).reshape(t["shape"])
        if op == "Variable":
            w = np.asarray(weights[node["name"]], np.float32)
            want = tuple(at["shape"])
            assert w.ndim == len(want) and all(x == y for x, y in zip(w.shape[:-1], want[:-1])), (node["name"], w.shape, want)
            return w
        if op == "Identity":
            return a[0]
        if op == "Cast":
            return a[0].astype(_NP[at["DstT"]])
        if op == "Mul":
            return a[0] * a[1]
        if op == "Sub":
            return a[0] - a[1]
        if op == "Add":
            return a[0] + a[1]
        if op == "Div":
            if np.issubdtype(np.asarray(a[0]).dtype, np.integer):
                return (np.asarray(a[0]) // np.asarray(a[1])).astype(np.asarray(a[0]).dtype)   # operands are sizes >= 0
            return a[0] / a[1]
        if op == "Rsqrt":
            return (np.float32(1.0) / np.sqrt(a[0])).astype(np.float32)
        if op == "Relu":
            return np.maximum(a[0], np.float32(0))
        if op == "Shape":
            return np.array(a[0].shape, np.int32)
        if op == "Slice":
            begin, size = [int(v) for v in a[1]], [int(v) for v in a[2]]
            idx = tuple(slice(b, None if s < 0 else b + s) for b, s in zip(begin, size))
            return a[0][idx]
        if op == "Squeeze":
            return np.squeeze(a[0], axis=tuple(at["squeeze_dims"]))
        if op == "Pack":
            return np.stack([np.asarray(v) for v in a], axis=0)
        if op == "ExpandDims":
            return np.expand_dims(a[0], int(a[1]))
        if op == "ResizeBilinear":
            return resize_bilinear(a[0], a[1], bool(at.get("align_corners", False)))
        if op == "Conv2D":
            assert at.get("data_format", "NHWC") == "NHWC"
            return conv2d(a[0], a[1], at["strides"], at["padding"])
        if op == "MaxPool":
            return max_pool(a[0], at["ksize"], at["strides"], at["padding"])
        if op == "AvgPool":
            return avg_pool(a[0], at["ksize"], at["strides"], at["padding"])
        if op == "Concat":                      # TF r0.9 Concat: inputs = [concat_dim, values...]
            return np.concatenate(a[1:], axis=int(a[0]))
        if op == "Reshape":
            return a[0].reshape([int(v) for v in a[1]])
        if op == "MatMul":
            x = a[0].T if at.get("transpose_a") else a[0]
            y = a[1].T if at.get("transpose_b") else a[1]
            return (x @ y).astype(np.float32)
        if op == "BiasAdd":
            return a[0] + a[1]
        if op == "Softmax":
            return softmax(a[0])
        raise NotImplementedError("op %s (%s)" % (op, node["name"]))


# end-point name of the reference's end_points dict -> node of its graph
ENDPOINT_NODES = {
    "conv0": "inception_v3/conv0/Relu", "conv1": "inception_v3/conv1/Relu", "conv2": "inception_v3/conv2/Relu",
    "pool1": "inception_v3/pool1/MaxPool", "conv3": "inception_v3/conv3/Relu", "conv4": "inception_v3/conv4/Relu",
    "pool2": "inception_v3/pool2/MaxPool",
    **{m: "inception_v3/%s/concat" % m for m in (
        "mixed_35x35x256a", "mixed_35x35x288a", "mixed_35x35x288b", "mixed_17x17x768a", "mixed_17x17x768b",
        "mixed_17x17x768c", "mixed_17x17x768d", "mixed_17x17x768e", "mixed_17x17x1280a", "mixed_8x8x2048a",
        "mixed_8x8x2048b")},
    "aux_logits": "inception_v3/aux_logits/FC/xw_plus_b",
    "logits": "inception_v3/logits/logits/xw_plus_b",
    "predictions": "inception_v3/logits/predictions",
}
PREPROCESSED_NODE = "ExpandDims_1"      # [1,299,299,3] float32 in [-1,1]: the tensor `images` of PyClassification.py:156
