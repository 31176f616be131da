"""Pre-fork HTTP server + GPU owners: the reference's `gunicorn -w N -b ADDR PyClassification:app` (run.sh:1-3).

    python -m tensorflow_web_deploy_b200.serving.server -w 8 -b 0.0.0.0:5001 --model data/model --labels data/synset.txt \
           --gpus 0,1,2,3,4,5,6,7 --max-batch 32 --window-ms 2 [-p gunicorn.pid]

The parent binds the listening socket, starts one GPU-owner process per GPU (batcher.owner_main: engine + nvJPEG +
micro-batching loop), waits until each has captured its graphs, then forks `-w` workers that all accept() on the
shared socket; each worker runs a threaded werkzeug WSGI server over app.ClassifierApp with a BatcherClient as
`app.clf`.  Workers hold no CUDA context and no weights.  SIGTERM / SIGINT to the parent (or `kill $(cat pidfile)`,
the reference's stop.sh) stops everything.
"""
from __future__ import annotations

import argparse
import logging
import multiprocessing as mp
import os
import signal
import socket
import sys
import tempfile
import time

log = logging.getLogger("icp.server")


def read_labels(path):
    names = ["none"]                                       # PyClassification.py:175-178
    with open(path, "rb") as f:
        names += f.read().decode("utf-8").splitlines()
    return names


def worker_main(listen_sock, socket_paths, label_names, sample_path, decode="worker"):
    from werkzeug.serving import make_server
    from .app import ClassifierApp, setup_app
    from .batcher import BatcherClient
    logging.basicConfig(level=logging.INFO)
    logging.getLogger("werkzeug").setLevel(logging.WARNING)
    app = setup_app(ClassifierApp(), BatcherClient(socket_paths, label_names, decode=decode), sample_path)
    srv = make_server("0.0.0.0", 0, app, threaded=True, fd=listen_sock.fileno())
    signal.signal(signal.SIGTERM, lambda *a: os._exit(0))
    srv.serve_forever()


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("-w", "--workers", type=int, default=2)                 # run.sh: NUM_WORKER=2
    ap.add_argument("-b", "--bind", default="0.0.0.0:5001")                 # run.sh: BIND_ADDR
    ap.add_argument("-p", "--pid", default=None)
    ap.add_argument("--model", required=True, help="ICPW weight blob (ImagenetClassifier's model_def_file)")
    ap.add_argument("--labels", required=True, help="synset file (class_labels_file)")
    ap.add_argument("--gpus", default="0")
    ap.add_argument("--max-batch", type=int, default=32)
    ap.add_argument("--window-ms", type=float, default=2.0)
    ap.add_argument("--precision", default=os.environ.get("ICP_PRECISION", "fp16"))
    ap.add_argument("--decode", default=os.environ.get("ICP_SERVE_DECODE", "worker"), choices=["worker", "owner"],
                    help="worker: libjpeg in the HTTP workers (parallel, reference-exact pixels); owner: nvJPEG on the GPU")
    ap.add_argument("--sample", default=None, help="warm-up picture (the reference uses data/sample.jpg)")
    args = ap.parse_args(argv)
    logging.basicConfig(level=logging.INFO)
    for key, val in (("model_def_file", args.model), ("class_labels_file", args.labels)):
        if not os.path.exists(val):                         # the reference raises at import, :140-143
            raise SystemExit("File for {} is missing. Should be at: {}".format(key, val))
    label_names = read_labels(args.labels)
    host, port = args.bind.rsplit(":", 1)
    lsock = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
    lsock.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
    lsock.bind((host, int(port)))
    lsock.listen(1024)
    lsock.set_inheritable(True)

    ctx = mp.get_context("spawn")                           # owners create CUDA contexts: never fork them
    run_dir = tempfile.mkdtemp(prefix="icp_serving_")
    gpus = [int(g) for g in args.gpus.split(",") if g != ""]
    owners, paths = [], []
    from .batcher import owner_main

    def start_owner(g):
        path = os.path.join(run_dir, "gpu%d.sock" % g)
        ready = path + ".ready"
        for stale in (ready,):
            if os.path.exists(stale):
                os.unlink(stale)
        p = ctx.Process(target=owner_main, args=(path, args.model, len(label_names), g, args.max_batch, args.window_ms,
                                                 args.precision, ready), daemon=True)
        p.start()
        return p, ready, path

    for g in gpus:
        p, ready, path = start_owner(g)
        owners.append((p, ready))
        paths.append(path)
    t0 = time.time()
    for p, ready in owners:
        while not os.path.exists(ready):
            if not p.is_alive():
                raise SystemExit("a GPU owner died during start-up (exit code %s)" % p.exitcode)
            if time.time() - t0 > 300:
                raise SystemExit("GPU owners did not come up within 300 s")
            time.sleep(0.05)
    log.info("%d GPU owner(s) ready after %.1f s", len(owners), time.time() - t0)

    workers = []
    for _ in range(args.workers):
        pid = os.fork()
        if pid == 0:
            try:
                worker_main(lsock, paths, label_names, args.sample, args.decode)
            finally:
                os._exit(0)
        workers.append(pid)
    if args.pid:
        open(args.pid, "w").write(str(os.getpid()))
    log.info("listening on %s with %d workers (pids %s)", args.bind, args.workers, workers)

    def shutdown(*_):
        for pid in workers:
            try:
                os.kill(pid, signal.SIGTERM)
            except ProcessLookupError:
                pass
        for p, _ in owners:
            p.terminate()
        sys.exit(0)

    signal.signal(signal.SIGTERM, shutdown)
    signal.signal(signal.SIGINT, shutdown)
    while True:
        try:
            pid, status = os.wait()
        except ChildProcessError:
            break
        dead = [i for i, (p, _) in enumerate(owners) if p.pid == pid]
        if dead:                                           # a GPU owner exited (exit code 3 = CUDA error, batcher.py): its CUDA
            i = dead[0]                                    # context is gone with it; a fresh process on the same socket
            log.error("GPU owner %d (gpu %d) exited with status %d; starting a replacement", pid, gpus[i], status)
            p, ready, _ = start_owner(gpus[i])             # takes over (workers' pooled connections reconnect on error)
            owners[i] = (p, ready)
            continue
        if pid in workers:                                 # gunicorn re-forks dead workers; so do we
            workers.remove(pid)
            log.warning("worker %d exited (status %d); re-forking", pid, status)
            npid = os.fork()
            if npid == 0:
                try:
                    worker_main(lsock, paths, label_names, None, args.decode)
                finally:
                    os._exit(0)
            workers.append(npid)


if __name__ == "__main__":
    main()
