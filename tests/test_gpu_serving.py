"""The serving stack end to end on a GPU: pre-fork workers + GPU owner + micro-batcher behind real HTTP
(SURVEY 8(f)1 / BASELINE configs[4]).  Concurrent POSTs of the reference's sample.jpg must all get the answer a
single in-process classify_image call gives."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu


def test_concurrent_http_posts_get_the_single_request_answer(synth6, built_lib):
    from tensorflow_web_deploy_b200 import PyClassification as P
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "http_bench.py"), "--workers", "2", "--gpus", "0", "--procs",
                        "2", "--threads", "6", "--requests", "240", "--max-batch", "16", "--window-ms", "3"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["errors"] == 0 and line["all_responses_equal_single_request"] is True
    assert line["curl_contract_ok"] in (True, None)
    assert line["micro_batching"]["mean_batch"] is not None
    assert max(int(k) for k in line["micro_batching"]["batch_histogram"]) > 1          # requests really were merged
    # the same picture through the in-process drop-in class (synthetic C=6 weights, seed 1234, as the server uses)
    clf = P.ImagenetClassifier(synth6[0], os.path.join(GOLDEN, "reference_synset.txt"))
    names, probs, secs, acc = clf.classify_image(open(os.path.join(GOLDEN, "reference_sample.jpg"), "rb").read())
    assert ",".join(names) == line["answer"]
    assert abs(float(acc) - line["ClassificationAccuracy"]) <= 1e-3
