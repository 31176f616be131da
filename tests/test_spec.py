"""Structure of the tower: package spec == oracle == the reference's own artefacts.

Reference anchors: data/model.meta (variable names + shapes + Conv2D/pool attrs decoded into
tests/golden/ by tools/extract_model_meta_names.py), inception/slim/collections_test.py:35-92,
inception/slim/inception_test.py:27-85."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from tensorflow_web_deploy_b200 import tower_spec as T


def _meta_variables():
    out = {}
    for line in open(os.path.join(GOLDEN, "model_meta_variables.txt")):
        if line.startswith("#") or not line.strip():
            continue
        name, shape = line.split()
        out[name] = tuple(int(d) for d in shape.split("x"))
    return out


def test_variable_names_and_shapes_match_model_meta():
    meta = _meta_variables()
    assert len(meta) == 388
    spec = dict(T.variables(4001))          # model.meta is the 4001-class export
    assert set(spec) == set(meta)
    for k, shp in spec.items():
        assert tuple(shp) == meta[k], k


def test_conv_strides_and_padding_match_model_meta():
    ops = json.load(open(os.path.join(GOLDEN, "model_meta_ops.json")))
    convs = {k[len("inception_v3/"):-len("/Conv2D")]: v for k, v in ops.items() if v["op"] == "Conv2D"}
    layers = T.conv_layers()
    assert len(convs) == len(layers) == 96
    for L in layers:
        rec = convs[L.name]
        assert rec["strides"] == [1, L.stride, L.stride, 1], L.name
        assert rec["padding"] == ("SAME" if L.same else "VALID"), L.name
    pools = [v for v in ops.values() if v["op"] != "Conv2D"]
    assert sum(1 for p in pools if p["op"] == "MaxPool") == 4       # pool1, pool2, 768a, 1280a
    assert sum(1 for p in pools if p["op"] == "AvgPool") == 11      # 9 block pools + aux 5x5/3 + global 8x8
    for p in pools:
        if p["op"] == "MaxPool":
            assert p["ksize"] == [1, 3, 3, 1] and p["strides"] == [1, 2, 2, 1] and p["padding"] == "VALID"


def test_variable_counts_collections_test():
    # collections_test.py:43-49: 388 = 98 weights + 2 biases + 96 beta + 0 gamma + 96 mean + 96 var
    v = [n for n, _ in T.variables(1000)]
    assert len(v) == 388
    leaf = lambda s: sum(1 for n in v if n.endswith("/" + s))
    assert (leaf("weights"), leaf("biases"), leaf("beta"), leaf("gamma"), leaf("moving_mean"),
            leaf("moving_variance")) == (98, 2, 96, 0, 96, 96)
    # collections_test.py:76-92 per-scope totals
    per = lambda p: sum(1 for n in v if n.startswith(p + "/"))
    for i in range(5):
        assert per("conv%d" % i) == 4
    for s in ("mixed_35x35x256a", "mixed_35x35x288a", "mixed_35x35x288b"):
        assert per(s) == 28
    assert per("mixed_17x17x768a") == 16
    for s in "bcde":
        assert per("mixed_17x17x768" + s) == 40
    assert per("mixed_17x17x1280a") == 24
    assert per("mixed_8x8x2048a") == per("mixed_8x8x2048b") == 36
    assert per("logits") == 2 and per("aux_logits") == 10


def test_macs_total_and_geometry():
    assert T.main_tower_macs() == 5_711_168_096 and T.FLOP_PER_IMAGE == 2 * 5_711_168_096
    last = T.conv_layers()[-1]
    assert (last.out_h, last.out_w) == (8, 8)                       # inception_test.py:44-46
    last150 = T.conv_layers(150, 150, include_aux=False)[-1]
    assert (last150.out_h, last150.out_w) == (3, 3)                 # inception_test.py:73-85


def test_oracle_requests_exactly_the_spec_variables():
    from oracle import inception_oracle as O
    assert O.recorded_variables(6) == T.variables(6)


def test_oracle_endpoint_shapes_inception_test():
    from oracle import inception_oracle as O
    tw = O.Tower({}, record=[])
    logits, pred = tw.forward(np.zeros((2, 299, 299, 3), np.float32), 1000)
    assert logits.shape == (2, 1000) and pred.shape == (2, 1000)          # inception_test.py:36-38
    assert tw.end_points["aux_logits"].shape == (2, 1000)                 # :39-42
    assert tw.end_points["mixed_8x8x2048b"].shape == (2, 8, 8, 2048)      # :43-46
    assert set(T.END_POINTS) == set(tw.end_points)
    tw = O.Tower({}, record=[])
    tw.forward(np.zeros((1, 150, 150, 3), np.float32), 1000, with_aux=False)
    assert tw.end_points["mixed_8x8x2048b"].shape == (1, 3, 3, 2048)      # :73-85
