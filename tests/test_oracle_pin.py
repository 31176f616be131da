"""Pins the hand-written oracle (oracle/inception_oracle.py, torch) to the reference's OWN graph.

tests/golden/reference_graph.json is the inference subgraph decoded from the reference's
data/model.meta (tools/export_reference_graph.py); oracle/graph_interpreter.py executes it with
plain-numpy kernels.  Wiring, attributes and constants therefore come from the reference's artefact;
the two implementations share no code.  CPU only.
"""
import io
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import graph_interpreter as GI
from oracle import inception_oracle as O
from tensorflow_web_deploy_b200 import weights


@pytest.fixture(scope="module")
def graph():
    return GI.Graph()


def test_graph_fixture_is_the_graph_the_app_builds(graph):
    """PyClassification.py:150-164: placeholder -> decode_jpeg -> convert_image_dtype -> eval_image -> sub .5 -> mul 2
    -> expand_dims -> inception_v3 -> softmax; 388 variables with the names the weight blob uses."""
    doc = json.load(open(GI.GRAPH_JSON))
    ops = {}
    for n in doc["nodes"]:
        ops[n["op"]] = ops.get(n["op"], 0) + 1
    assert ops["Conv2D"] == 96 and ops["Relu"] == 96 and ops["Rsqrt"] == 96       # every conv has BN + ReLU
    assert ops["Concat"] == 15 and ops["MaxPool"] == 4 and ops["AvgPool"] == 11    # 9 SAME 3x3 + aux 5x5/3 + global 8x8
    assert ops["Variable"] == 388
    names = dict(graph.variables())
    W = weights.synthetic_weights(4001)
    assert {k: tuple(v.shape) for k, v in W.items()} == names                       # incl. the [2048,4001] / [768,4001] FCs
    n = graph.nodes
    assert n["convert_image/y"]["attr"]["value"]["values"] == [0.003921568859368563]   # 1/255 as float32
    assert n["Div/y"]["attr"]["value"]["values"] == [16] and n["mul/y"]["attr"]["value"]["values"] == [2]
    assert n["ResizeBilinear"]["attr"]["align_corners"] is False
    assert n["ResizeBilinear/size"]["attr"]["value"]["values"] == [299, 299]
    assert n["Sub/y"]["attr"]["value"]["values"] == [0.5] and n["Mul/y"]["attr"]["value"]["values"] == [2.0]
    eps = n["inception_v3/conv0/BatchNorm/batchnorm/add/y"]["attr"]["value"]["values"][0]
    assert abs(eps - 0.001) < 1e-9


@pytest.mark.parametrize("shape", [(375, 500), (16, 16), (299, 299), (1000, 37), (333, 517)])
def test_preprocessing_equals_reference_graph(graph, shape):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    rgb = rng.integers(0, 256, shape + (3,), dtype=np.uint8)
    got = O.preprocess_u8(rgb)
    ref = graph.run([GI.PREPROCESSED_NODE], {"DecodeJpeg": rgb}, {})[GI.PREPROCESSED_NODE]
    assert got.shape == ref.shape == (1, 299, 299, 3)
    np.testing.assert_array_equal(got, ref)          # same float32 operations in the same order: bit-identical


@pytest.mark.parametrize("num_classes", [6, 1001])
def test_every_end_point_equals_reference_graph(graph, num_classes):
    W = weights.synthetic_weights(num_classes)
    x = np.random.default_rng(num_classes).uniform(-1, 1, (1, 299, 299, 3)).astype(np.float32)
    out = graph.run(list(GI.ENDPOINT_NODES.values()), {GI.PREPROCESSED_NODE: x}, W, cache={})
    tw = O.Tower(W)
    logits, pred = tw.forward(x, num_classes, with_aux=True)
    assert set(GI.ENDPOINT_NODES) == set(tw.end_points)
    for key, node in GI.ENDPOINT_NODES.items():
        a, b = out[node], tw.end_points[key]
        assert a.shape == b.shape, key
        scale = float(np.abs(a).max())
        assert float(np.abs(a - b).max()) <= 2e-5 * max(scale, 1.0), key      # fp32, different summation order only
    np.testing.assert_array_equal(np.argsort(-out[GI.ENDPOINT_NODES["predictions"]], kind="stable")[:, :5],
                                  O.top_k(pred, 5)[1])


def test_reference_sample_jpg_end_to_end(graph):
    """BASELINE configs[0] on the reference's own fixtures: data/sample.jpg (500x375) and data/synset.txt through
    decode -> the reference graph (numpy) == the oracle's classify path; return contract of classify_image."""
    from PIL import Image
    jpeg = open(os.path.join(GOLDEN, "reference_sample.jpg"), "rb").read()
    labels = ["none"] + open(os.path.join(GOLDEN, "reference_synset.txt")).read().splitlines()
    assert labels == ["none", "daisy", "dandelion", "roses", "sunflowers", "tulips"]      # PyClassification.py:175-178
    rgb = np.asarray(Image.open(io.BytesIO(jpeg)).convert("RGB"))
    assert rgb.shape == (375, 500, 3)
    assert O.central_crop_box(375, 500) == (23, 31, 329, 438)
    W = weights.synthetic_weights(6)
    pred = graph.run(["inception_v3/logits/predictions"], {"DecodeJpeg": rgb}, W)["inception_v3/logits/predictions"]
    names, probs, acc = O.OracleClassifier(W, labels, 6).classify_decoded(rgb)
    v, i = O.top_k(pred, 5)
    assert names == [labels[j] for j in i[0, :3]]
    np.testing.assert_allclose(probs, v[0, :3], atol=2e-6)
    assert abs(acc - float(v[0].sum())) < 1e-5
