#!/usr/bin/env python
"""Benchmark of the hot path: images/sec of the Inception-v3 classify path on synthetic 299x299x3.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--precision bf16|fp16]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one batch of 128 images per GPU (BASELINE.json configs[2]/[3]: batch=128 bf16 per B200,
1024 over 8) through the whole tower: conv0 + 93 tcgen05 implicit-GEMM convs + pools + softmax/top-5.
Multi-GPU = pure data parallel, weights replicated, one NCCL all-gather of the [128,C] probabilities
per step.  Prints ONE JSON line on rank 0.

value     images/s, whole job, inputs resident in HBM, CUDA events on the launching stream, max over ranks
e2e       images/s through Engine.forward_u8 (public API): pinned-host uint8 batch -> H2D -> tower -> D2H probs
roofline  conv_igemm_kernel: algorithmic conv FLOPs / its summed launch durations (CUDA events between
          launches, icp_profile_forward) vs the measured bf16 peak in MEASURED_PEAKS.json
cpu_baseline / --impl reference : the CPU oracle (torch fp32 restatement of the reference's TF graph —
          the reference itself cannot run: no TensorFlow, Python 2, no checkpoint) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NUM_CLASSES = 6           # reference: NUM_CLASSES + 1 (PyClassification.py:20,160)
FLOP_PER_IMAGE = 2 * 5_711_168_096
BATCH_PER_GPU = 128
METRIC = "images/sec (299x299x3)"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "hbm_gbs": d["hbm_gbs"], "source": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


def conv_traffic_per_launch(batch, precision):
    """DRAM bytes per conv launch from the committed ncu capture (only valid for the configuration it was taken at)."""
    p = os.path.join(ROOT, "profiles", "r01_conv_traffic_final.json")
    if batch != 128 or not os.path.exists(p):
        return None
    d = json.load(open(p))
    return d["dram_bytes_total"] / d["conv_launches_per_step"]


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 7:
                    self.samples.append(f)
            except Exception:
                pass
            self._halt.wait(0.1)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


def host_cores():
    """Cores this process may really use: affinity mask capped by the cgroup CPU quota."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = max(1, min(n, int(int(quota) / int(period))))
    except Exception:
        pass
    return max(1, n)


_ORACLE_W = None


def oracle_images_per_s(images, threads, warm=True):
    """CPU restatement of the reference path on `images` preprocessed images; returns (img/s, seconds)."""
    global _ORACLE_W
    import torch
    from oracle import inception_oracle as O
    from tensorflow_web_deploy_b200 import weights
    torch.set_num_threads(threads)
    if _ORACLE_W is None:
        _ORACLE_W = weights.synthetic_weights(NUM_CLASSES)
    W = _ORACLE_W
    x = np.random.default_rng(1).uniform(-1, 1, (images, 299, 299, 3)).astype(np.float32)
    if warm:
        O.Tower(W).forward(x[:1], NUM_CLASSES, with_aux=False)      # warm-up (oneDNN primitive creation)
    t0 = time.perf_counter()
    for i in range(images):                                         # batch 1, as the reference serves
        pred = O.Tower(W).forward(x[i:i + 1], NUM_CLASSES, with_aux=False)[1]
        O.top_k(pred, 5)
    dt = time.perf_counter() - t0
    return images / dt, dt


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path.  The TF graph cannot run here
    (SURVEY 8(c)), so this is the oracle port on all host threads; rank 0 only."""
    if rank != 0:
        return
    cores = min(host_cores(), 32)
    per_step = 2
    for _ in range(args.warmup):
        oracle_images_per_s(1, cores, warm=False)
    t0 = time.perf_counter()
    tot_img = 0
    for _ in range(args.steps):
        _, dt = oracle_images_per_s(per_step, cores, warm=False)
        tot_img += per_step
        if time.perf_counter() - t0 > 150:      # bounded sample: never more than a few minutes
            break
    wall = time.perf_counter() - t0
    v = tot_img / wall
    done_steps = tot_img // per_step
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "images/s", "n_gpus": args.gpus,
            "steps": done_steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(done_steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "classify path, batch 1 x %d images per step, CPU oracle port of the reference "
                                   "TF graph (reference itself not runnable: no TensorFlow/Python2/checkpoint)" % per_step,
                       "num_classes": NUM_CLASSES},
            "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": "%d steps x %d images, batch 1, torch fp32 CPU" % (args.steps, per_step)},
            "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp16"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="images per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import __graft_entry__ as g
    if rank == 0:
        g.build()
    if world > 1:
        dist.barrier()
    from tensorflow_web_deploy_b200 import engine, weights

    B, C = args.batch, NUM_CLASSES
    wpath = os.path.join(tempfile.gettempdir(), "icp_bench_c%d_rank%d.icpw" % (C, rank))
    weights.save_blob(wpath, weights.synthetic_weights(C))
    eng = engine.Engine(wpath, C, device=local_rank, max_batch=B, precision=args.precision, use_cuda_graph=True)
    dev = torch.device("cuda", local_rank)
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    x_dev = torch.empty((B, 299, 299, 3), device=dev, dtype=torch.float32).uniform_(-1, 1, generator=gen)
    probs = torch.empty((B, C), device=dev, dtype=torch.float32)
    gathered = torch.empty((B * world, C), device=dev, dtype=torch.float32)
    stream = torch.cuda.Stream(device=dev)          # a real (non-NULL) stream: the engine enqueues on the handle we pass
    torch.cuda.set_stream(stream)

    def step():
        eng.forward_device(x_dev.data_ptr(), B, probs.data_ptr(), stream.cuda_stream)
        if world > 1:
            dist.all_gather_into_tensor(gathered, probs)      # the path's only collective: [B,C] probabilities

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    sync_all()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    launches = eng.launches_per_forward * args.steps

    # ---------------- end-to-end through the public API: pinned host u8 -> H2D -> tower -> D2H ----------------
    # Two alternating pinned host batches through Engine.submit_u8 / collect (two slots): the PCIe copy of batch
    # i+1 overlaps the tower of batch i; every step's input crosses PCIe and its probabilities/top-5 come back.
    x_hosts = [torch.randint(0, 256, (B, 299, 299, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
    x_nps = [t.numpy() for t in x_hosts]
    outs = [(torch.empty((B, C), dtype=torch.float32).pin_memory().numpy(),
             torch.empty((B, 5), dtype=torch.float32).pin_memory().numpy(),
             torch.empty((B, 5), dtype=torch.int32).pin_memory().numpy()) for _ in range(2)]

    def e2e_loop(steps):
        eng.submit_u8(x_nps[0], 0, 5, outs[0])
        for i in range(1, steps):
            eng.submit_u8(x_nps[i & 1], i & 1, 5, outs[i & 1])
            p, v, ix = eng.collect((i - 1) & 1)
            if world > 1:
                probs.copy_(torch.from_numpy(p), non_blocking=False)
                dist.all_gather_into_tensor(gathered, probs)
        p, v, ix = eng.collect((steps - 1) & 1)
        if world > 1:
            probs.copy_(torch.from_numpy(p), non_blocking=False)
            dist.all_gather_into_tensor(gathered, probs)
            torch.cuda.synchronize()

    e2e_loop(args.warmup)
    sync_all()
    t0 = time.perf_counter()
    e2e_loop(args.steps)
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_ips = B * world * args.steps / float(e2e_s.item())
    # synchronous single-call variant (no overlap), for reference
    t0 = time.perf_counter()
    for _ in range(max(3, args.steps // 4)):
        eng.forward_u8(x_nps[0], 5)
    e2e_sync_ips = B * max(3, args.steps // 4) / (time.perf_counter() - t0)
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        peaks = measured_peaks()
        value = B * world * args.steps / (ms_total * 1e-3)
        # ---------------- roofline of the dominant kernel (conv_igemm_kernel), live ----------------
        eng.forward(x_dev[:B].cpu().numpy())                    # resident input for the profiled passes
        per_op = eng.profile_forward(B, reps=3)
        table = eng.op_table()
        conv_ms = sum(float(t) for (nm, kind, fl, by), t in zip(table, per_op) if kind == 1)
        conv_fl = sum(fl for (nm, kind, fl, by) in table if kind == 1) * B
        n_conv = sum(1 for t in table if t[1] == 1)
        achieved = conv_fl / (conv_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": "conv_igemm_persistent_kernel (63 launches) / conv_halo_kernel (2): the "
                                                 "tcgen05 implicit-GEMM family, all conv launches but conv0",
                    "achieved": achieved,
                    "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_tflops"],
                    "peak_source": peaks["source"] + " burst bf16 (MEASURED_PEAKS.json)",
                    "launches_per_step": n_conv, "avg_launch_us": 1e3 * conv_ms / n_conv,
                    "share_of_step": conv_ms / float(per_op.sum()), "traffic": conv_traffic_per_launch(B, args.precision),
                    "traffic_note": "dram__bytes_read+write per launch, mean over the 65 conv launches of one step, from "
                                    "the committed ncu capture profiles/r01_conv_traffic_final.json (batch 128 bf16); "
                                    "null for other configurations",
                    "whole_tower_frac": value / world * FLOP_PER_IMAGE / 1e12 / peaks["bf16_tflops"]}
        # ---------------- batch-1 latency (BASELINE configs[1]) ----------------
        lat = engine.Engine(wpath, C, device=local_rank, max_batch=1, precision=args.precision, use_cuda_graph=True)
        x1 = x_dev[:1].cpu().numpy()
        dms = []
        for i in range(203):
            lat.forward(x1)
            if i >= 3:
                dms.append(lat.last_forward_ms)
        rgb = np.random.default_rng(2).integers(0, 256, (375, 500, 3), dtype=np.uint8)
        wall = []
        for i in range(103):
            t0 = time.perf_counter()
            lat.classify_rgb8(rgb, 5)
            if i >= 3:
                wall.append(1e3 * (time.perf_counter() - t0))
        latency = {"batch1_tower_device_ms_p50": float(np.median(dms)), "batch1_tower_device_ms_p99": float(np.percentile(dms, 99)),
                   "classify_rgb8_wall_ms_p50": float(np.median(wall)), "classify_rgb8_wall_ms_p99": float(np.percentile(wall, 99))}
        lat.close()
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = min(host_cores(), 32)
            ips, secs = oracle_images_per_s(12, cores)
            cpu = {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                   "sample": "12 images, batch 1 each (as the reference serves), torch fp32 CPU oracle, %.1f s" % secs,
                   "p50_latency_s": secs / 12}
        line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
                "config": {"workload": "batch=%d/GPU %s synthetic 299x299x3, full Inception-v3 tower + softmax + top-5 "
                                       "(BASELINE configs[2]; configs[3] at 8 GPUs)" % (B, args.precision),
                           "global_batch": B * world, "num_classes": C, "parallelism": "dp%d" % world,
                           "collective": "all_gather probs [%d,%d] f32" % (B, C) if world > 1 else "none",
                           "l2": "inputs larger than L2 (%.0f MB fp32 batch, activations ~8 GB)" % (B * 299 * 299 * 3 * 4 / 1e6),
                           "cuda_graph": True},
                "e2e": {"value": e2e_ips, "unit": "images/s", "h2d_bytes_per_step": B * 299 * 299 * 3,
                        "d2h_bytes_per_step": B * C * 4 + B * 5 * 8,
                        "api": "Engine.submit_u8/collect (icp_submit_u8, two slots: H2D of batch i+1 overlaps tower i)",
                        "sync_call_value": e2e_sync_ips},
                "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
                "latency": latency,
                "reference_published": {"classify_image_latency_s": 0.47, "source": "reference README.md:22, TF CPU, hardware unstated"}}
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
