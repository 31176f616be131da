#!/usr/bin/env python
"""GPU bring-up diagnostics (run under gpurun).  Not a test suite: prints per-case error statistics
so a failing descriptor / swizzle / im2col setting can be diagnosed from one log.

  python tools/gpu_bringup.py conv [--force-im2col] [--only NAME]
  python tools/gpu_bringup.py pools
  python tools/gpu_bringup.py tower [--precision fp16|bf16] [--n 2]
  python tools/gpu_bringup.py speed [--precision bf16] [--n 128] [--graph]
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def round16(a, precision):
    import torch
    dt = torch.float16 if precision == "fp16" else torch.bfloat16
    return torch.from_numpy(np.ascontiguousarray(a)).to(dt).to(torch.float32).numpy()


def ref_conv(x, w, scale, shift, stride, same, relu, precision):
    import torch
    import torch.nn.functional as F
    # the engine folds the BatchNorm scale into the 16-bit weights (and adds shift inside the accumulator)
    xq, wq = round16(x, precision), round16(w * scale.reshape(1, 1, 1, -1), precision)
    kh, kw = w.shape[:2]
    pad = ((kh - 1) // 2, (kw - 1) // 2) if same else (0, 0)
    y = F.conv2d(torch.from_numpy(xq.transpose(0, 3, 1, 2).copy()), torch.from_numpy(wq.transpose(3, 2, 0, 1).copy()),
                 None, stride, pad)
    y = y + torch.from_numpy(shift).view(1, -1, 1, 1)
    if relu:
        y = torch.relu(y)
    return round16(y.numpy().transpose(0, 2, 3, 1), precision)


CONV_CASES = [
    # name, n, h, w, cin, kh, kw, cout, stride, same
    ("1x1_k64_n64_m128", 1, 8, 16, 64, 1, 1, 64, 1, True),
    ("1x1_k192_n64", 2, 35, 35, 192, 1, 1, 64, 1, True),
    ("1x1_k192_n48", 2, 35, 35, 192, 1, 1, 48, 1, True),
    ("1x1_k64_n80_valid", 1, 73, 73, 64, 1, 1, 80, 1, False),
    ("1x1_k768_n192", 3, 17, 17, 768, 1, 1, 192, 1, True),
    ("1x1_k1280_n320", 4, 8, 8, 1280, 1, 1, 320, 1, True),
    ("1x1_k2048_n448", 2, 8, 8, 2048, 1, 1, 448, 1, True),
    ("1x1_k288_n64", 1, 35, 35, 288, 1, 1, 64, 1, True),
    ("3x3_same_k64_n96", 2, 35, 35, 64, 3, 3, 96, 1, True),
    ("3x3_same_k96_n96", 1, 35, 35, 96, 3, 3, 96, 1, True),
    ("3x3_valid_k32_n32", 1, 37, 37, 32, 3, 3, 32, 1, False),
    ("3x3_same_k32_n64", 1, 37, 37, 32, 3, 3, 64, 1, True),
    ("3x3_valid_k80_n192", 1, 25, 25, 80, 3, 3, 192, 1, False),
    ("5x5_same_k48_n64", 2, 35, 35, 48, 5, 5, 64, 1, True),
    ("3x3_s2_valid_k288_n384", 2, 35, 35, 288, 3, 3, 384, 2, False),
    ("3x3_s2_valid_k96_n96", 1, 35, 35, 96, 3, 3, 96, 2, False),
    ("1x7_same_k128_n128", 2, 17, 17, 128, 1, 7, 128, 1, True),
    ("7x1_same_k160_n192", 2, 17, 17, 160, 7, 1, 192, 1, True),
    ("1x3_same_k384_n384", 2, 8, 8, 384, 1, 3, 384, 1, True),
    ("3x1_same_k384_n384", 2, 8, 8, 384, 3, 1, 384, 1, True),
    ("3x3_same_k448_n384", 2, 8, 8, 448, 3, 3, 384, 1, True),
    ("3x3_s2_valid_k192_n320", 2, 17, 17, 192, 3, 3, 320, 2, False),
    ("5x5_full_valid_k128_n768", 3, 5, 5, 128, 5, 5, 768, 1, False),
    ("3x3_same_tiny", 1, 3, 3, 64, 3, 3, 32, 1, True),
    # halo-resident kernel (large stride-1 maps)
    ("halo_conv1_149_k32_n32_valid", 1, 149, 149, 32, 3, 3, 32, 1, False),
    ("halo_conv2_147_k32_n64_same", 2, 147, 147, 32, 3, 3, 64, 1, True),
    ("halo_conv4_73_k80_n192_valid", 2, 73, 73, 80, 3, 3, 192, 1, False),
    ("halo_35_k48_5x5_n64", 3, 35, 35, 48, 5, 5, 64, 1, True),
    ("halo_35_k96_3x3_n96", 3, 35, 35, 96, 3, 3, 96, 1, True),
    ("halo_40x33_k64_3x3_n128", 2, 40, 33, 64, 3, 3, 128, 1, True),
    ("halo_31x50_k24_1x7_n48", 2, 31, 50, 24, 1, 7, 48, 1, True),
    ("halo_50x31_k136_7x1_n64", 2, 50, 31, 136, 7, 1, 64, 1, True),
]


def cmd_conv(args):
    from tensorflow_web_deploy_b200 import engine
    rng = np.random.default_rng(7)
    results = []
    bad = 0
    for case in CONV_CASES:
        name, n, h, w, cin, kh, kw, cout, stride, same = case
        if args.only and args.only != name:
            continue
        x = rng.standard_normal((n, h, w, cin), dtype=np.float32)
        wt = rng.standard_normal((kh, kw, cin, cout), dtype=np.float32) * np.float32(np.sqrt(2.0 / (kh * kw * cin)))
        scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
        shift = (rng.standard_normal(cout) * 0.1).astype(np.float32)
        for precision in ("fp16", "bf16"):
            t0 = time.time()
            try:
                y, _ = engine.conv2d_bn_relu(x, wt, scale, shift, stride, same, True, precision)
            except Exception as ex:  # noqa
                print("CASE %-28s %s EXC %s" % (name, precision, ex), flush=True)
                results.append({"case": name, "precision": precision, "error": str(ex)})
                bad += 1
                continue
            ref = ref_conv(x, wt, scale, shift, stride, same, True, precision)
            err = np.abs(y - ref)
            tol = (2e-3 if precision == "fp16" else 1.6e-2) * (1.0 + np.abs(ref))
            nbad = int((~(err <= tol)).sum())
            nan = int(np.isnan(y).sum())
            rec = {"case": name, "precision": precision, "max_abs": float(np.nanmax(err)) if err.size else 0.0,
                   "n_bad": nbad, "n_nan": nan, "numel": int(y.size), "secs": round(time.time() - t0, 3)}
            results.append(rec)
            status = "ok " if nbad == 0 else "BAD"
            print("CASE %-28s %s %s max_abs=%.4g bad=%d/%d nan=%d" % (name, precision, status, rec["max_abs"], nbad,
                                                                         y.size, nan), flush=True)
            if nbad:
                bad += 1
                if bad <= 3:
                    np.save(os.path.join(OUT, "conv_fail_%s_%s_got.npy" % (name, precision)), y)
                    np.save(os.path.join(OUT, "conv_fail_%s_%s_ref.npy" % (name, precision)), ref)
                    # where are the mismatches?  rows (pixels) vs columns (channels)
                    m = ~(err <= tol)
                    rows = m.reshape(-1, cout).any(axis=1)
                    cols = m.reshape(-1, cout).any(axis=0)
                    print("   bad rows: %d/%d first=%s ; bad cols: %d/%d first=%s" % (
                        rows.sum(), rows.size, np.nonzero(rows)[0][:8], cols.sum(), cols.size, np.nonzero(cols)[0][:8]),
                        flush=True)
    tag = "im2col" if os.environ.get("ICP_FORCE_IM2COL") else "default"
    with open(os.path.join(OUT, "bringup_conv_%s.json" % tag), "w") as f:
        json.dump(results, f, indent=1)
    print("conv bring-up: %d failing case/precision pairs" % bad)
    return 1 if bad else 0


def cmd_pools(args):
    import torch
    import torch.nn.functional as F
    from tensorflow_web_deploy_b200 import engine
    rng = np.random.default_rng(3)
    bad = 0
    for precision in ("fp16", "bf16"):
        x = rng.standard_normal((2, 17, 17, 64), dtype=np.float32)
        xq = torch.from_numpy(round16(x, precision).transpose(0, 3, 1, 2).copy())
        y = engine.pool2d(x, "max", precision=precision)
        ref = F.max_pool2d(xq, 3, 2).numpy().transpose(0, 2, 3, 1)
        e1 = float(np.abs(y - ref).max())
        y = engine.pool2d(x, "avg", 3, 1, True, precision=precision)
        ref = round16(F.avg_pool2d(xq, 3, 1, 1, count_include_pad=False).numpy().transpose(0, 2, 3, 1), precision)
        e2 = float(np.abs(y - ref).max())
        y = engine.pool2d(x, "avg", 5, 3, False, precision=precision)
        ref = round16(F.avg_pool2d(xq, 5, 3).numpy().transpose(0, 2, 3, 1), precision)
        e3 = float(np.abs(y - ref).max())
        ones = engine.pool2d(np.ones((1, 5, 5, 8), np.float32), "avg", 3, 1, True, precision=precision)
        e4 = float(np.abs(ones - 1.0).max())
        print("POOLS %s max=%.3g avg_same=%.3g avg5s3=%.3g ones=%.3g" % (precision, e1, e2, e3, e4), flush=True)
        lim = 2e-3 if precision == "fp16" else 1.6e-2
        bad += int(e1 > 0) + int(e2 > lim) + int(e3 > lim) + int(e4 > 0)
    return 1 if bad else 0


def _weights_file(num_classes):
    from tensorflow_web_deploy_b200 import weights
    path = os.path.join(tempfile.gettempdir(), "icp_synthetic_c%d.icpw" % num_classes)
    W = weights.synthetic_weights(num_classes)
    if not os.path.exists(path):
        weights.save_blob(path, W)
    return path, W


def cmd_tower(args):
    from oracle import inception_oracle as O
    from tensorflow_web_deploy_b200 import engine
    C = 6
    path, W = _weights_file(C)
    if getattr(args, "calibrated", False):      # trained-like BatchNorm statistics (tests/conftest.py)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from conftest import calibrated_weights
        from tensorflow_web_deploy_b200 import weights as _w
        W = calibrated_weights(C)
        W["logits/logits/weights"] = (W["logits/logits/weights"] * args.fc_scale).astype(np.float32)
        path = os.path.join(tempfile.gettempdir(), "icp_calibrated_c%d.icpw" % C)
        _w.save_blob(path, W)
    x = np.random.default_rng(0).uniform(-1, 1, (args.n, 299, 299, 3)).astype(np.float32)
    t0 = time.time()
    tw = O.Tower(W)
    _, pred = tw.forward(x, C, with_aux=True)
    print("oracle forward n=%d: %.2fs" % (args.n, time.time() - t0), flush=True)
    eng = engine.Engine(path, C, max_batch=max(args.n, 2), precision=args.precision, enable_aux=True,
                        use_cuda_graph=args.graph)
    probs, vals, idx = eng.forward_topk(x, 5)
    print("device ms (first call): %.3f  launches=%d" % (eng.last_forward_ms, eng.launches_per_forward), flush=True)
    rep = {}
    order = ["conv0", "conv1", "conv2", "pool1", "conv3", "conv4", "pool2", "mixed_35x35x256a", "mixed_35x35x288a",
             "mixed_35x35x288b", "mixed_17x17x768a", "mixed_17x17x768b", "mixed_17x17x768c", "mixed_17x17x768d",
             "mixed_17x17x768e", "mixed_17x17x1280a", "mixed_8x8x2048a", "mixed_8x8x2048b", "aux_logits", "logits",
             "predictions"]
    for k in order:
        got = eng.endpoint(k)
        ref = tw.end_points[k]
        err = np.abs(got - ref)
        rel = float(np.sqrt((err ** 2).mean()) / (np.sqrt((ref ** 2).mean()) + 1e-12))
        rep[k] = {"max_abs": float(np.nanmax(err)), "rel_rms": rel, "nan": int(np.isnan(got).sum())}
        print("EP %-20s shape=%s max_abs=%.4g rel_rms=%.4g nan=%d" % (k, got.shape, rep[k]["max_abs"], rel,
                                                                      rep[k]["nan"]), flush=True)
    ov, oi = O.top_k(pred, 5)
    print("probs gpu   ", probs[0])
    print("probs oracle", pred[0])
    print("max |dprob| = %.3g ; topk idx equal: %s ; max |dtopk| = %.3g" % (
        np.abs(probs - pred).max(), bool((idx == oi).all()), np.abs(vals - ov).max()), flush=True)
    rep["max_dprob"] = float(np.abs(probs - pred).max())
    rep["topk_equal"] = bool((idx == oi).all())
    with open(os.path.join(OUT, "bringup_tower_%s%s.json" % (args.precision, "_graph" if args.graph else "")), "w") as f:
        json.dump(rep, f, indent=1)
    lim = 1e-3 if args.precision == "fp16" else 6e-3
    return 0 if rep["max_dprob"] <= lim and rep["topk_equal"] else 1


def cmd_speed(args):
    from tensorflow_web_deploy_b200 import engine
    C = 6
    path, _ = _weights_file(C)
    n = args.n
    x = np.random.default_rng(1).uniform(-1, 1, (n, 299, 299, 3)).astype(np.float32)
    eng = engine.Engine(path, C, max_batch=n, precision=args.precision, use_cuda_graph=args.graph)
    ms = []
    if args.u8:      # the uint8 entry (what bench.py and the serving path use: normalisation folded into conv0)
        x = np.random.default_rng(1).integers(0, 256, (n, 299, 299, 3), dtype=np.uint8)
    for i in range(args.iters + 3):
        if args.u8:
            eng.forward_u8(x)
        else:
            eng.forward(x)
        if i >= 3:
            ms.append(eng.last_forward_ms)
    ms = np.array(ms)
    flops = 2 * 5711168096 * n
    print("SPEED n=%d %s graph=%s%s: device ms median=%.3f min=%.3f -> %.0f img/s, %.1f TFLOP/s (%.1f%% of 1667)" % (
        n, args.precision, args.graph, " u8" if args.u8 else "", np.median(ms), ms.min(), n / (np.median(ms) * 1e-3),
        flops / (np.median(ms) * 1e-3) / 1e12, 100 * flops / (np.median(ms) * 1e-3) / 1667.1e12), flush=True)
    return 0


def cmd_profile(args):
    """Per-launch device times (CUDA events between launches) + achieved TFLOP/s / GB/s per op."""
    from tensorflow_web_deploy_b200 import engine
    C = 6
    path, _ = _weights_file(C)
    n = args.n
    x = np.random.default_rng(1).uniform(-1, 1, (n, 299, 299, 3)).astype(np.float32)
    eng = engine.Engine(path, C, max_batch=n, precision=args.precision)
    for _ in range(2):
        eng.forward(x)
    ms = eng.profile_forward(n, args.reps)
    table = eng.op_table()
    rows = []
    tot = float(ms.sum())
    conv_ms = conv_fl = 0.0
    for (name, kind, fl, by), t in zip(table, ms):
        tf = fl * n / (t * 1e-3) / 1e12 if t > 0 else 0.0
        gb = by * n / (t * 1e-3) / 1e9 if t > 0 else 0.0
        rows.append({"name": name, "kind": int(kind), "ms": float(t), "tflops": float(tf), "gbs": float(gb), "share": float(t) / tot})
        if kind == 1:
            conv_ms += float(t); conv_fl += fl * n
        print("OP %-44s kind=%d ms=%8.4f share=%5.1f%% %8.1f TF/s %8.1f GB/s" % (name, kind, t, 100 * t / tot, tf, gb),
              flush=True)
    print("PROFILE n=%d %s: sum=%.3f ms ; conv_igemm %.3f ms (%.1f%%) -> %.1f TFLOP/s" % (
        n, args.precision, tot, conv_ms, 100 * conv_ms / tot, conv_fl / (conv_ms * 1e-3) / 1e12), flush=True)
    with open(os.path.join(OUT, "profile_ops_n%d_%s.json" % (n, args.precision)), "w") as f:
        json.dump({"n": n, "precision": args.precision, "sum_ms": float(tot), "conv_igemm_ms": float(conv_ms), "ops": rows}, f, indent=1)
    return 0


LAYERS = {
    "conv4": (73, 73, 80, 3, 3, 192, 1, False),
    "c17_1x7_192": (17, 17, 192, 1, 7, 192, 1, True),
    "c17_1x1_768_704": (17, 17, 768, 1, 1, 704, 1, True),
    "c35_3x3_96": (35, 35, 96, 3, 3, 96, 1, True),
    "c8_3x3_448": (8, 8, 448, 3, 3, 384, 1, True),
    "c35_1x1_288_256": (35, 35, 288, 1, 1, 256, 1, True),
    "conv1": (149, 149, 32, 3, 3, 32, 1, False),
    "conv2": (147, 147, 32, 3, 3, 64, 1, True),
    "conv3": (73, 73, 64, 1, 1, 80, 1, False),
    "c35_5x5_48_64": (35, 35, 48, 5, 5, 64, 1, True),
    "c35_3x3_64_96": (35, 35, 64, 3, 3, 96, 1, True),
    # synthetic shapes that isolate the main loop (many full waves of tiles, long K)
    "gemm_k1024_n256": (64, 64, 1024, 1, 1, 256, 1, True),
    "gemm_k1024_n128": (64, 64, 1024, 1, 1, 128, 1, True),
    "gemm_k1024_n192": (64, 64, 1024, 1, 1, 192, 1, True),
    "conv3x3_k256_n256": (64, 64, 256, 3, 3, 256, 1, True),
}


def cmd_layer(args):
    """One layer at batch n through the stateless entry, timed over `reps` launches (for ncu captures)."""
    from tensorflow_web_deploy_b200 import engine
    h, w, cin, kh, kw, cout, stride, same = LAYERS[args.name]
    rng = np.random.default_rng(0)
    x = rng.standard_normal((args.n, h, w, cin), dtype=np.float32)
    wt = rng.standard_normal((kh, kw, cin, cout), dtype=np.float32) * np.float32(np.sqrt(2.0 / (kh * kw * cin)))
    y, ms = engine.conv2d_bn_relu(x, wt, np.ones(cout, np.float32), np.zeros(cout, np.float32), stride, same, True,
                                  args.precision, reps=args.reps)
    oh, ow = y.shape[1:3]
    fl = 2.0 * args.n * oh * ow * cout * kh * kw * cin
    print("LAYER %s n=%d: %.4f ms -> %.1f TFLOP/s" % (args.name, args.n, ms, fl / (ms * 1e-3) / 1e12), flush=True)
    return 0


def main():
    if os.environ.get("ICP_AB_LIB"):      # A/B timing on the same box against a saved build of an earlier commit
        from tensorflow_web_deploy_b200 import engine as _e
        _e.LIB_PATH = os.path.abspath(os.environ["ICP_AB_LIB"])
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest="cmd", required=True)
    p = sub.add_parser("conv"); p.add_argument("--only", default=None)
    sub.add_parser("pools")
    p = sub.add_parser("tower"); p.add_argument("--precision", default="fp16"); p.add_argument("--n", type=int, default=2)
    p.add_argument("--graph", action="store_true"); p.add_argument("--calibrated", action="store_true")
    p.add_argument("--fc-scale", type=float, default=4.0)
    p = sub.add_parser("speed"); p.add_argument("--precision", default="bf16"); p.add_argument("--n", type=int, default=128)
    p.add_argument("--iters", type=int, default=10); p.add_argument("--graph", action="store_true")
    p.add_argument("--u8", action="store_true")
    p = sub.add_parser("profile"); p.add_argument("--precision", default="bf16"); p.add_argument("--n", type=int, default=128)
    p.add_argument("--reps", type=int, default=5)
    p = sub.add_parser("layer"); p.add_argument("--name", required=True); p.add_argument("--n", type=int, default=128)
    p.add_argument("--precision", default="bf16"); p.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    rc = {"layer": cmd_layer, "profile": cmd_profile, "conv": cmd_conv, "pools": cmd_pools, "tower": cmd_tower, "speed": cmd_speed}[args.cmd](args)
    sys.exit(rc)


if __name__ == "__main__":
    main()
