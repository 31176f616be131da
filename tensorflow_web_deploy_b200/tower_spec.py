"""Static description of the Inception-v3 tower served by the reference.

The reference builds this graph imperatively with TF-Slim
(/root/reference/inception/slim/inception_model.py:86-329).  Here the same
topology is data: a stem list plus eleven block tables, walked by
``conv_layers()`` / ``variables()``.  The CUDA engine builds its launch plan from the same
topology in csrc/engine.cu (``Builder::build``: it asks the weight blob for exactly these names and
shapes and fails loudly on any mismatch); tests/test_spec.py checks this table against the oracle's
recorded variable requests and against the 388 variable names/shapes decoded from the reference's
data/model.meta (tests/golden/model_meta_variables.txt); the GPU parity tests check the engine.

Naming follows the reference's variable scopes: a conv inside
``with tf.variable_scope('mixed_35x35x256a'): with tf.variable_scope('branch5x5')``
created without an explicit scope gets the auto-scope ``Conv``, ``Conv_1`` ...
(/root/reference/inception/slim/ops.py:213).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Iterator, List, Tuple

BN_EPS = 0.001  # /root/reference/inception/inception_model.py:73


@dataclass(frozen=True)
class ConvLayer:
    name: str      # variable scope, e.g. "mixed_35x35x256a/branch5x5/Conv_1"
    in_h: int
    in_w: int
    cin: int
    kh: int
    kw: int
    stride: int
    same: bool     # True = SAME, False = VALID
    cout: int

    @property
    def out_h(self) -> int:
        return -(-self.in_h // self.stride) if self.same else (self.in_h - self.kh) // self.stride + 1

    @property
    def out_w(self) -> int:
        return -(-self.in_w // self.stride) if self.same else (self.in_w - self.kw) // self.stride + 1

    @property
    def macs(self) -> int:
        return self.out_h * self.out_w * self.cout * self.kh * self.kw * self.cin


# (op, args).  "c" = conv(cout, kh, kw, stride, same, explicit_scope|None)
STEM = [
    ("c", "conv0", 32, 3, 3, 2, False),
    ("c", "conv1", 32, 3, 3, 1, False),
    ("c", "conv2", 64, 3, 3, 1, True),
    ("maxpool", "pool1"),
    ("c", "conv3", 80, 1, 1, 1, False),
    ("c", "conv4", 192, 3, 3, 1, False),
    ("maxpool", "pool2"),
]

# Each block: name -> ordered list of (branch_scope, [steps]); a step is
# ("c", cout, kh, kw[, stride, same]) | ("avg",) | ("max",) | ("fork", [steps], [steps])
# where "fork" writes both sub-results side by side (the nested tf.concat at
# inception_model.py:287-288,292-293).  Branch order = concat order.
def _b35(pool_out):
    return [
        ("branch1x1", [("c", 64, 1, 1)]),
        ("branch5x5", [("c", 48, 1, 1), ("c", 64, 5, 5)]),
        ("branch3x3dbl", [("c", 64, 1, 1), ("c", 96, 3, 3), ("c", 96, 3, 3)]),
        ("branch_pool", [("avg",), ("c", pool_out, 1, 1)]),
    ]


def _b17(m):
    return [
        ("branch1x1", [("c", 192, 1, 1)]),
        ("branch7x7", [("c", m, 1, 1), ("c", m, 1, 7), ("c", 192, 7, 1)]),
        ("branch7x7dbl", [("c", m, 1, 1), ("c", m, 7, 1), ("c", m, 1, 7), ("c", m, 7, 1), ("c", 192, 1, 7)]),
        ("branch_pool", [("avg",), ("c", 192, 1, 1)]),
    ]


def _b8():
    return [
        ("branch1x1", [("c", 320, 1, 1)]),
        ("branch3x3", [("c", 384, 1, 1), ("fork", ("c", 384, 1, 3), ("c", 384, 3, 1))]),
        ("branch3x3dbl", [("c", 448, 1, 1), ("c", 384, 3, 3), ("fork", ("c", 384, 1, 3), ("c", 384, 3, 1))]),
        ("branch_pool", [("avg",), ("c", 192, 1, 1)]),
    ]


BLOCKS: List[Tuple[str, list]] = [
    ("mixed_35x35x256a", _b35(32)),
    ("mixed_35x35x288a", _b35(64)),
    ("mixed_35x35x288b", _b35(64)),
    ("mixed_17x17x768a", [
        ("branch3x3", [("c", 384, 3, 3, 2, False)]),
        ("branch3x3dbl", [("c", 64, 1, 1), ("c", 96, 3, 3), ("c", 96, 3, 3, 2, False)]),
        ("branch_pool", [("max",)]),
    ]),
    ("mixed_17x17x768b", _b17(128)),
    ("mixed_17x17x768c", _b17(160)),
    ("mixed_17x17x768d", _b17(160)),
    ("mixed_17x17x768e", _b17(192)),
    ("mixed_17x17x1280a", [
        ("branch3x3", [("c", 192, 1, 1), ("c", 320, 3, 3, 2, False)]),
        ("branch7x7x3", [("c", 192, 1, 1), ("c", 192, 1, 7), ("c", 192, 7, 1), ("c", 192, 3, 3, 2, False)]),
        ("branch_pool", [("max",)]),
    ]),
    ("mixed_8x8x2048a", _b8()),
    ("mixed_8x8x2048b", _b8()),
]

AUX_AFTER = "mixed_17x17x768e"   # inception_model.py:249-261

END_POINTS = (["conv0", "conv1", "conv2", "pool1", "conv3", "conv4", "pool2"]
              + [b[0] for b in BLOCKS] + ["aux_logits", "logits", "predictions"])


def _pool_out(h, k, s):
    return (h - k) // s + 1


def conv_layers(height: int = 299, width: int = 299, include_aux: bool = True) -> List[ConvLayer]:
    """All conv layers in graph-construction order with resolved geometry."""
    out: List[ConvLayer] = []
    h, w, c = height, width, 3
    for step in STEM:
        if step[0] == "c":
            _, name, cout, kh, kw, s, same = step
            L = ConvLayer(name, h, w, c, kh, kw, s, same, cout)
            out.append(L)
            h, w, c = L.out_h, L.out_w, cout
        else:
            h, w = _pool_out(h, 3, 2), _pool_out(w, 3, 2)
    for bname, branches in BLOCKS:
        oh = ow = None
        ctot = 0
        for scope, steps in branches:
            bh, bw, bc = h, w, c
            idx = 0

            def emit(st, bh, bw, bc):
                nonlocal idx
                cout, kh, kw = st[1], st[2], st[3]
                s = st[4] if len(st) > 4 else 1
                same = st[5] if len(st) > 5 else True
                nm = "%s/%s/%s" % (bname, scope, "Conv" if idx == 0 else "Conv_%d" % idx)
                idx += 1
                L = ConvLayer(nm, bh, bw, bc, kh, kw, s, same, cout)
                out.append(L)
                return L

            for st in steps:
                if st[0] == "c":
                    L = emit(st, bh, bw, bc)
                    bh, bw, bc = L.out_h, L.out_w, L.cout
                elif st[0] == "avg":
                    pass                                   # 3x3/1 SAME keeps geometry
                elif st[0] == "max":
                    bh, bw = _pool_out(bh, 3, 2), _pool_out(bw, 3, 2)
                elif st[0] == "fork":
                    La = emit(st[1], bh, bw, bc)
                    Lb = emit(st[2], bh, bw, bc)
                    bh, bw, bc = La.out_h, La.out_w, La.cout + Lb.cout
            oh, ow = bh, bw
            ctot += bc
        h, w, c = oh, ow, ctot
        if include_aux and bname == AUX_AFTER:
            ah, aw = _pool_out(h, 5, 3), _pool_out(w, 5, 3)
            out.append(ConvLayer("aux_logits/proj", ah, aw, c, 1, 1, 1, True, 128))
            out.append(ConvLayer("aux_logits/Conv", ah, aw, 128, ah, aw, 1, False, 768))
    return out


def variables(num_classes: int, height: int = 299, width: int = 299) -> List[Tuple[str, Tuple[int, ...]]]:
    """(name, shape) of every restorable variable, reference naming
    (PyClassification.py:166-172 restores exactly these, under their EMA keys)."""
    v: List[Tuple[str, Tuple[int, ...]]] = []
    for L in conv_layers(height, width, include_aux=True):
        v.append((L.name + "/weights", (L.kh, L.kw, L.cin, L.cout)))
        for s in ("beta", "moving_mean", "moving_variance"):
            v.append(("%s/BatchNorm/%s" % (L.name, s), (L.cout,)))
        if L.name == "aux_logits/Conv":
            v.append(("aux_logits/FC/weights", (768, num_classes)))
            v.append(("aux_logits/FC/biases", (num_classes,)))
    v.append(("logits/logits/weights", (2048, num_classes)))
    v.append(("logits/logits/biases", (num_classes,)))
    return v


def main_tower_macs(height: int = 299, width: int = 299) -> int:
    return sum(L.macs for L in conv_layers(height, width, include_aux=False))


FLOP_PER_IMAGE = 2 * 5_711_168_096  # SURVEY.md Appendix A total, asserted in tests/test_spec.py
