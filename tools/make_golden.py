#!/usr/bin/env python
"""Mints tests/golden/oracle_c6.npz — fixtures for the hot path (SURVEY 8(c) item iii).

The reference cannot run here (no TensorFlow, no checkpoint), so these vectors come from the CPU
oracle on seeded inputs/weights; they pin the oracle against accidental change and give the GPU
tests a fixture that does not need the oracle's 10+ seconds per run.  Per end point we keep summary
statistics and 64 strided samples (the full tensors are ~250 MB).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import inception_oracle as O          # noqa: E402
from tensorflow_web_deploy_b200 import weights    # noqa: E402

C, N, SEED_X = 6, 2, 0


def sample_idx(size, k=64):
    return (np.arange(k, dtype=np.int64) * 2654435761 % size).astype(np.int64)


def main():
    W = weights.synthetic_weights(C)
    x = np.random.default_rng(SEED_X).uniform(-1, 1, (N, 299, 299, 3)).astype(np.float32)
    tw = O.Tower(W)
    logits, pred = tw.forward(x, C, with_aux=True)
    vals, idx = O.top_k(pred, 5)
    out = {"logits": logits, "predictions": pred, "topk_vals": vals, "topk_idx": idx,
           "aux_logits": tw.end_points["aux_logits"]}
    for k, a in tw.end_points.items():
        if a.ndim != 4:
            continue
        flat = a.reshape(-1)
        out["ep_shape/" + k] = np.array(a.shape)
        out["ep_stats/" + k] = np.array([flat.mean(dtype=np.float64), np.abs(flat).mean(dtype=np.float64),
                                         flat.max(), np.sqrt((flat.astype(np.float64) ** 2).mean())])
        out["ep_samples/" + k] = flat[sample_idx(flat.size)]
    # preprocessing fixture: a synthetic 500x375 "photo" (same geometry as the reference's data/sample.jpg)
    rng = np.random.default_rng(5)
    yy, xx = np.mgrid[0:375, 0:500]
    img = np.stack([127 + 120 * np.sin(xx / 37.0 + yy / 91.0), 127 + 120 * np.cos(xx / 53.0 - yy / 29.0),
                    (xx * yy) % 255], -1) + rng.normal(0, 6, (375, 500, 3))
    img = np.clip(img, 0, 255).astype(np.uint8)
    pre = O.preprocess_u8(img)
    out["pre_img"] = img
    out["pre_samples"] = pre.reshape(-1)[sample_idx(pre.size, 256)]
    out["pre_stats"] = np.array([pre.mean(dtype=np.float64), pre.min(), pre.max()])
    pvals, pidx = O.top_k(O.Tower(W).forward(pre, C, with_aux=False)[1], 5)
    out["pre_topk_vals"], out["pre_topk_idx"] = pvals, pidx
    path = os.path.join(ROOT, "tests", "golden", "oracle_c6.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
