#!/usr/bin/env python
"""BASELINE configs[4]: N-worker HTTP load test with concurrent POSTs of data/sample.jpg, micro-batched on the GPU(s).

Starts `python -m tensorflow_web_deploy_b200.serving.server -w <workers>` (the reference's `gunicorn -w N`, run.sh:1-3),
checks the documented web-service call (README.md:30: curl --request POST --data-binary @sample.jpg http://host:port),
then drives it from `--procs` client processes x `--threads` keep-alive connections each and prints ONE JSON line:
requests/s, p50 / p99 latency, the batch-size histogram of the GPU owners, and whether every response carried the
labels / ClassificationAccuracy of the single-request answer.
"""
import argparse
import http.client
import json
import multiprocessing as mp
import os
import shutil
import signal
import socket
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def post(conn, body):
    conn.request("POST", "/", body=body, headers={"Content-Type": "application/octet-stream"})
    r = conn.getresponse()
    data = r.read()
    return r.status, data.decode(), r.getheader("ClassificationAccuracy")


def client_proc(port, body, threads, requests_per_thread, out_q):
    import threading
    lat, answers, errors = [], set(), [0]
    lock = threading.Lock()

    def work():
        conn = http.client.HTTPConnection("127.0.0.1", port, timeout=60)
        mine, mans = [], set()
        for _ in range(requests_per_thread):
            t0 = time.perf_counter()
            try:
                status, text, acc = post(conn, body)
            except Exception:
                conn.close()
                conn = http.client.HTTPConnection("127.0.0.1", port, timeout=60)
                with lock:
                    errors[0] += 1
                continue
            mine.append(time.perf_counter() - t0)
            if status != 200:
                with lock:
                    errors[0] += 1
            else:
                mans.add((text, round(float(acc), 3)))
        with lock:
            lat.extend(mine)
            answers.update(mans)

    ts = [threading.Thread(target=work) for _ in range(threads)]
    t0 = time.perf_counter()
    [t.start() for t in ts]
    [t.join() for t in ts]
    out_q.put((lat, list(answers), errors[0], time.perf_counter() - t0))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workers", type=int, default=8)
    ap.add_argument("--gpus", default=None, help="comma list; default: all visible")
    ap.add_argument("--procs", type=int, default=8)
    ap.add_argument("--threads", type=int, default=4)
    ap.add_argument("--requests", type=int, default=4000, help="total requests of the timed run")
    ap.add_argument("--max-batch", type=int, default=32)
    ap.add_argument("--window-ms", type=float, default=2.0)
    ap.add_argument("--precision", default="fp16")
    ap.add_argument("--decode", default="worker", choices=["worker", "owner"])
    args = ap.parse_args()
    import torch
    from tensorflow_web_deploy_b200 import weights
    ngpu = torch.cuda.device_count()
    if ngpu == 0:
        raise SystemExit("http_bench: no CUDA device — the B200 path has no CPU fallback")
    gpus = args.gpus or ",".join(str(i) for i in range(ngpu))
    import __graft_entry__ as g
    g.build()
    tmp = tempfile.mkdtemp(prefix="icp_http_")
    model = os.path.join(tmp, "model")
    weights.save_blob(model, weights.synthetic_weights(6))
    labels = os.path.join(ROOT, "tests", "golden", "reference_synset.txt")
    sample = os.path.join(ROOT, "tests", "golden", "reference_sample.jpg")
    body = open(sample, "rb").read()
    port = free_port()
    cmd = [sys.executable, "-m", "tensorflow_web_deploy_b200.serving.server", "-w", str(args.workers), "-b", "127.0.0.1:%d" % port,
           "--model", model, "--labels", labels, "--gpus", gpus, "--max-batch", str(args.max_batch), "--window-ms",
           str(args.window_ms), "--precision", args.precision, "--decode", args.decode, "--sample", sample, "-p", os.path.join(tmp, "server.pid")]
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), TMPDIR=tmp)
    srv = subprocess.Popen(cmd, cwd=ROOT, env=env, stdout=open(os.path.join(tmp, "server.log"), "w"), stderr=subprocess.STDOUT,
                           start_new_session=True)
    try:
        t0 = time.time()
        while True:
            if srv.poll() is not None:
                raise SystemExit("server exited during start-up:\n" + open(os.path.join(tmp, "server.log")).read()[-3000:])
            try:
                c = http.client.HTTPConnection("127.0.0.1", port, timeout=5)
                c.request("GET", "/")
                if c.getresponse().status == 200:
                    break
            except OSError:
                pass
            if time.time() - t0 > 600:
                raise SystemExit("server did not come up")
            time.sleep(0.2)
        startup_s = time.time() - t0
        conn = http.client.HTTPConnection("127.0.0.1", port, timeout=60)
        status, single_text, single_acc = post(conn, body)
        assert status == 200, status
        curl_ok = None
        if shutil.which("curl"):                      # the reference's documented call, verbatim
            r = subprocess.run(["curl", "-s", "-i", "--request", "POST", "--data-binary", "@" + sample,
                                "http://127.0.0.1:%d" % port], capture_output=True, text=True, timeout=60)
            curl_ok = r.returncode == 0 and single_text in r.stdout and "ClassificationAccuracy" in r.stdout
        seq = []
        for _ in range(50):                           # sequential (one request in flight): the unbatched latency
            t1 = time.perf_counter()
            post(conn, body)
            seq.append(time.perf_counter() - t1)
        per_thread = max(1, args.requests // (args.procs * args.threads))
        q = mp.Queue()
        for phase in ("warm", "timed"):
            n = max(4, per_thread // 8) if phase == "warm" else per_thread
            ps = [mp.Process(target=client_proc, args=(port, body, args.threads, n, q)) for _ in range(args.procs)]
            t1 = time.perf_counter()
            [p.start() for p in ps]
            res = [q.get() for _ in ps]
            [p.join() for p in ps]
            wall = time.perf_counter() - t1
        lat = np.array([x for r in res for x in r[0]])
        answers = {tuple(a) for r in res for a in r[1]}
        errors = sum(r[2] for r in res)
        time.sleep(0.7)
        hist = {}
        run_dir = [d for d in os.listdir(tmp) if d.startswith("icp_serving_")]
        for d in run_dir:
            for f in os.listdir(os.path.join(tmp, d)):
                if f.endswith(".stats"):
                    st = json.load(open(os.path.join(tmp, d, f)))
                    for k, v in st["batch_hist"].items():
                        hist[int(k)] = hist.get(int(k), 0) + v
        same = answers == {(single_text, round(float(single_acc), 3))}
        tot_b = sum(hist.values())
        line = {"metric": "HTTP requests/s (POST / with data/sample.jpg, %d workers)" % args.workers, "value": len(lat) / wall,
                "unit": "requests/s", "n_gpus": len(gpus.split(",")), "workers": args.workers,
                "clients": "%d processes x %d keep-alive connections" % (args.procs, args.threads),
                "requests": int(len(lat)), "errors": int(errors), "latency_ms_p50": float(np.percentile(lat, 50) * 1e3),
                "latency_ms_p99": float(np.percentile(lat, 99) * 1e3),
                "sequential_latency_ms_p50": float(np.median(seq) * 1e3),
                "jpeg_decode": "libjpeg (PIL) in the HTTP workers" if args.decode == "worker" else "nvJPEG in the GPU owner",
                "micro_batching": {"window_ms": args.window_ms, "max_batch": args.max_batch,
                                   "mean_batch": (sum(k * v for k, v in hist.items()) / tot_b) if tot_b else None,
                                   "batch_histogram": {str(k): hist[k] for k in sorted(hist)}},
                "answer": single_text, "ClassificationAccuracy": float(single_acc),
                "all_responses_equal_single_request": bool(same), "curl_contract_ok": curl_ok,
                "server_startup_s": startup_s, "higher_is_better": True,
                "config": {"workload": "BASELINE configs[4]: concurrent POSTs of the reference's 134,914-byte sample.jpg to the "
                                       "py3/werkzeug port of its front-end; nvJPEG decode + tower in one GPU-owner process per GPU"}}
        print(json.dumps(line), flush=True)
        return 0 if same and errors == 0 else 1
    finally:
        try:
            os.killpg(srv.pid, signal.SIGTERM)
        except ProcessLookupError:
            pass
        try:
            srv.wait(timeout=10)
        except subprocess.TimeoutExpired:
            os.killpg(srv.pid, signal.SIGKILL)
        if os.environ.get("ICP_KEEP_TMP") is None:
            log_tail = open(os.path.join(tmp, "server.log")).read()[-1500:]
            sys.stderr.write("--- server.log tail ---\n" + log_tail + "\n")
            shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    sys.exit(main())
