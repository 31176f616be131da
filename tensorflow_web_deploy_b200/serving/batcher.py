"""Cross-process micro-batcher: many HTTP workers, ONE GPU-owner process per GPU.

The reference runs N gunicorn workers, each with a private TF session and a private copy of the weights, each
serving batch 1 (/root/reference/run.sh:1-3, PyClassification.py:248-252).  On a B200 a batch-1 tower uses ~1 % of
the chip, so here the workers only parse HTTP; every request is handed to the process that owns the GPU, which
merges whatever arrived within a short window into one tower launch (north_star: "requests are micro-batched"):

    worker thread --(payload in a shared-memory segment, 28-byte header on an AF_UNIX socket)--> GpuOwner
    GpuOwner: collect up to `max_batch` requests, or until `window_ms` after the first one  ->
              ONE tower launch at batch n (engine.classify_batch_rgb8 / classify_batch_jpeg)
    GpuOwner --(status, top-5 values, top-5 indices)--> worker thread

Where the JPEG is decoded is the client's choice (`decode=`):
  "worker" (default)  the HTTP worker decodes with libjpeg (PIL) — the decoder family TF's DecodeJpeg wraps, so the
                      tower sees the pixels the reference would, and N workers decode N requests in parallel — and
                      hands over the uint8 [h,w,3] pixels; the owner does H2D + crop/resize kernel + tower;
  "owner"             the worker hands over the JPEG bytes and the owner decodes them on the GPU with nvJPEG
                      (icp_classify_batch_jpeg), falling back to libjpeg per request for streams nvJPEG refuses.

`BatcherClient.classify_image(bytes)` has the return contract of ImagenetClassifier.classify_image
(PyClassification.py:205-218): (names[3], probs[:3], seconds, sum of top-5) or None on any error.

Wire format (little-endian): request  = u32 magic | u32 length | u64 request id | u32 kind | u32 h | u32 w
                                        (payload: shm[0:length]; kind 0 = JPEG bytes, 1 = decoded RGB8 [h,w,3])
                             response = u64 request id | i32 status | i32 k | k f32 values | k i32 indices
"""
from __future__ import annotations

import ctypes
import logging
import os
import queue
import socket
import struct
import threading
import time
from multiprocessing import shared_memory
from typing import Callable, List, Optional, Sequence

import numpy as np

MAGIC = 0x49435042            # "ICPB"
REQ = struct.Struct("<IIQIII")
KIND_JPEG, KIND_RGB8 = 0, 1
RSP_HEAD = struct.Struct("<Qii")
HELLO = struct.Struct("<I64s")     # shm size, shm name
TOP_K = 5
SHM_BYTES = 16 << 20          # largest request body a worker thread can hand over

log = logging.getLogger("icp.batcher")


def _recv_exact(sock: socket.socket, n: int) -> Optional[bytes]:
    buf = bytearray()
    while len(buf) < n:
        chunk = sock.recv(n - len(buf))
        if not chunk:
            return None
        buf += chunk
    return bytes(buf)


class _Pending:
    __slots__ = ("conn", "lock", "req_id", "data", "kind", "h", "w", "t_arrival")

    def __init__(self, conn, lock, req_id, data, kind=KIND_JPEG, h=0, w=0):
        self.conn, self.lock, self.req_id, self.data = conn, lock, req_id, data
        self.kind, self.h, self.w = kind, h, w
        self.t_arrival = time.perf_counter()


class GpuOwner:
    """Owns one engine; serves every connected worker thread.  `engine` needs classify_batch_jpeg(list[bytes], k)
    -> (vals[n,k], idx[n,k]) and classify_rgb8(rgb, k); tests inject a fake."""

    def __init__(self, engine, socket_path: str, max_batch: int = 32, window_ms: float = 2.0,
                 decode_fallback: Optional[Callable[[bytes], np.ndarray]] = None):
        self.engine, self.path, self.max_batch, self.window = engine, socket_path, int(max_batch), window_ms * 1e-3
        self.decode_fallback = decode_fallback
        self.q: "queue.Queue[_Pending]" = queue.Queue()
        self.stop_ev = threading.Event()
        self.stats = {"requests": 0, "batches": 0, "max_batch_seen": 0, "errors": 0, "batch_hist": {}}
        # set when the engine reports a CUDA error: a trapped kernel or a lost device poisons the process's CUDA context,
        # every later call would fail too, and only a new process gets a new context — owner_main exits on it and
        # serving/server.py starts a replacement owner on the same socket
        self.fatal = threading.Event()
        self._threads: List[threading.Thread] = []
        self._listener: Optional[socket.socket] = None

    # ------------------------------------------------------------------ lifecycle
    def start(self):
        if os.path.exists(self.path):
            os.unlink(self.path)
        s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        s.bind(self.path)
        s.listen(256)
        s.settimeout(0.2)
        self._listener = s
        for target in (self._accept_loop, self._batch_loop):
            t = threading.Thread(target=target, daemon=True)
            t.start()
            self._threads.append(t)
        return self

    def stop(self):
        self.stop_ev.set()
        for t in self._threads:
            t.join(timeout=2)
        if self._listener is not None:
            self._listener.close()
        if os.path.exists(self.path):
            os.unlink(self.path)

    # ------------------------------------------------------------------ connections
    def _accept_loop(self):
        while not self.stop_ev.is_set():
            try:
                conn, _ = self._listener.accept()
            except socket.timeout:
                continue
            except OSError:
                return
            threading.Thread(target=self._serve_conn, args=(conn,), daemon=True).start()

    def _serve_conn(self, conn: socket.socket):
        shm = None
        pinned = None
        try:
            hello = _recv_exact(conn, HELLO.size)
            if hello is None:
                return
            size, name = HELLO.unpack(hello)
            shm = shared_memory.SharedMemory(name=name.rstrip(b"\0").decode())
            lock = threading.Lock()
            # ICP_SERVE_PIN=1: page-lock the worker's segment once and hand the engine views into it instead of copies, so
            # the per-image host->device copies of a batch are asynchronous DMA.  MEASURED (16 workers, 64 requests in
            # flight, one GPU): p50 14.7 -> 9.0 ms but p99 19 -> 56 ms (cudaHostRegister of segments that appear while the
            # pool grows stalls the batch thread) and 4.2k -> 4.0k requests/s — the closed-loop load generator is
            # latency-bound, not the owner — so it is opt-in.
            zero_copy = os.environ.get("ICP_SERVE_PIN", "0") == "1"
            reg = getattr(self.engine, "host_register", None) if zero_copy else None
            if reg is not None:
                try:
                    anchor = ctypes.c_char.from_buffer(shm.buf)
                    addr = ctypes.addressof(anchor)
                    del anchor
                    if reg(addr, shm.size):
                        pinned = addr
                except Exception as err:                      # registration is an optimisation only
                    log.info("could not page-lock a worker segment: %s", err)
            while not self.stop_ev.is_set():
                head = _recv_exact(conn, REQ.size)
                if head is None:
                    return
                magic, length, req_id, kind, h, w = REQ.unpack(head)
                if magic != MAGIC or length > size or (kind == KIND_RGB8 and length != h * w * 3):
                    log.warning("bad request header from a worker; closing the connection")
                    return
                # Default: the bytes are copied out of the segment (the worker may reuse it as soon as it has its answer).
                # ICP_SERVE_PIN=1: decoded pixels are NOT copied — the worker waits for its answer before it touches the
                # segment again, so a view into it stays valid until the engine call has returned.
                data = shm.buf[:length] if (zero_copy and kind == KIND_RGB8) else bytes(shm.buf[:length])
                self.q.put(_Pending(conn, lock, req_id, data, kind, h, w))
        except OSError:
            pass
        finally:
            if shm is not None:
                if pinned is not None:
                    try:
                        self.engine.host_unregister(pinned)
                    except Exception:
                        pass
                try:
                    shm.close()
                except BufferError:                           # a queued request still holds a view: the mapping goes with it
                    pass
            try:
                conn.close()
            except OSError:
                pass

    # ------------------------------------------------------------------ batching
    def _batch_loop(self):
        while not self.stop_ev.is_set():
            try:
                first = self.q.get(timeout=0.1)
            except queue.Empty:
                continue
            batch = [first]
            deadline = time.perf_counter() + self.window
            while len(batch) < self.max_batch:
                left = deadline - time.perf_counter()
                try:
                    batch.append(self.q.get(timeout=max(left, 0)) if left > 0 else self.q.get_nowait())
                except queue.Empty:
                    break
            self._run(batch)

    def _run(self, batch: Sequence[_Pending]):
        n = len(batch)
        st = self.stats
        st["requests"] += n
        st["batches"] += 1
        st["max_batch_seen"] = max(st["max_batch_seen"], n)
        st["batch_hist"][n] = st["batch_hist"].get(n, 0) + 1
        results = [None] * n
        rgb = [i for i, p in enumerate(batch) if p.kind == KIND_RGB8]
        jpg = [i for i, p in enumerate(batch) if p.kind != KIND_RGB8]
        if rgb:                                         # decoded by the workers: one launch for all of them
            imgs = None
            try:
                imgs = [np.frombuffer(batch[i].data, np.uint8).reshape(batch[i].h, batch[i].w, 3) for i in rgb]
                vals, idx = self.engine.classify_batch_rgb8(imgs, TOP_K)
                for j, i in enumerate(rgb):
                    results[i] = (vals[j], idx[j])
            except Exception as err:
                log.info("Classification error: %s", err)
                self._note(err)
            imgs = None
            for i in rgb:                               # give the views into the workers' segments back before answering
                if isinstance(batch[i].data, memoryview):
                    try:
                        batch[i].data.release()
                    except Exception:
                        pass
                batch[i].data = None
        if jpg:
            try:
                vals, idx = self.engine.classify_batch_jpeg([batch[i].data for i in jpg], TOP_K)
                for j, i in enumerate(jpg):
                    results[i] = (vals[j], idx[j])
            except Exception as err:                    # one undecodable stream fails the batched call: go one by one
                self._note(err)
                if not self.fatal.is_set():
                    log.info("batched call failed (%s); retrying the %d requests individually", err, len(jpg))
                    for i in jpg:
                        results[i] = self._single(batch[i].data)
        for p, r in zip(batch, results):
            if r is None:
                st["errors"] += 1
                payload = RSP_HEAD.pack(p.req_id, -1, 0)
            else:
                v, ix = r
                k = len(v)
                payload = RSP_HEAD.pack(p.req_id, 0, k) + np.asarray(v, "<f4").tobytes() + np.asarray(ix, "<i4").tobytes()
            try:
                with p.lock:
                    p.conn.sendall(payload)
            except OSError:
                pass

    def _note(self, err: Exception):
        if getattr(err, "code", 0) == -3:               # ICP_E_CUDA (include/inception_b200.h)
            if not self.fatal.is_set():
                log.error("CUDA error in the GPU owner (%s): the context is lost, this process will be replaced", err)
            self.fatal.set()

    def _single(self, data: bytes):
        try:
            v, ix = self.engine.classify_batch_jpeg([data], TOP_K)
            return v[0], ix[0]
        except Exception:
            pass
        if self.decode_fallback is not None:
            try:
                v, ix = self.engine.classify_rgb8(self.decode_fallback(data), TOP_K)
                return v, ix
            except Exception as err:
                log.info("Classification error: %s", err)
        return None


class BatcherClient:
    """What a worker holds instead of an ImagenetClassifier: same `classify_image` contract, the work happens in the
    GPU-owner process(es).  A pool of connections (socket + shared-memory segment) per owner, grown on demand to the
    number of requests in flight; owners are used round-robin."""

    def __init__(self, socket_paths: Sequence[str], label_names: Sequence[str], timeout_s: float = 30.0,
                 decode: str = "worker"):
        self.paths, self.label_names, self.timeout = list(socket_paths), list(label_names), timeout_s
        if decode not in ("worker", "owner"):
            raise ValueError("decode must be 'worker' or 'owner'")
        self.decode = decode
        self._pools = {}
        self._pool_lock = threading.Lock()
        self._rr = 0
        self._rr_lock = threading.Lock()
        self._all = []
        self._all_lock = threading.Lock()

    # Connections (socket + shared-memory segment) are POOLED per owner, not per thread: a threaded WSGI server runs
    # every request on a fresh thread, so thread-local connections meant one new 16 MiB segment, one new socket and one new
    # owner-side thread per request (measured: 9 requests/s with time-outs, 6908 leaked segments).
    def _acquire(self, owner: int):
        with self._pool_lock:
            pool = self._pools.setdefault(owner, [])
            if pool:
                return pool.pop()
        s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        s.settimeout(self.timeout)
        s.connect(self.paths[owner])
        shm = shared_memory.SharedMemory(create=True, size=SHM_BYTES)
        s.sendall(HELLO.pack(SHM_BYTES, shm.name.encode()))
        c = {"sock": s, "shm": shm, "next_id": 1}
        with self._all_lock:
            self._all.append(c)
        return c

    def _release(self, owner: int, c):
        with self._pool_lock:
            self._pools.setdefault(owner, []).append(c)

    def _discard(self, c):
        try:
            c["sock"].close()
        except OSError:
            pass
        try:
            c["shm"].close()
            c["shm"].unlink()
        except Exception:
            pass
        with self._all_lock:
            if c in self._all:
                self._all.remove(c)

    def topk(self, image: bytes):
        """JPEG bytes -> (values[k], indices[k]) or raises."""
        kind, h, w, payload = KIND_JPEG, 0, 0, image
        if self.decode == "worker":
            from ..PyClassification import decode_jpeg_rgb
            rgb = decode_jpeg_rgb(image)                 # raises for anything that is not a JPEG (the graph only has DecodeJpeg)
            if rgb.nbytes <= SHM_BYTES:
                kind, (h, w), payload = KIND_RGB8, rgb.shape[:2], rgb.reshape(-1)
        if len(payload) > SHM_BYTES:
            raise ValueError("request of %d bytes exceeds the %d-byte hand-over segment" % (len(payload), SHM_BYTES))
        with self._rr_lock:
            owner = self._rr % len(self.paths)
            self._rr += 1
        for attempt in range(3):                        # stale pooled connections: the owner may have been restarted
            c = self._acquire(owner)
            try:
                c["shm"].buf[:len(payload)] = payload
                rid = c["next_id"]
                c["next_id"] += 1
                c["sock"].sendall(REQ.pack(MAGIC, len(payload), rid, kind, h, w))
                head = _recv_exact(c["sock"], RSP_HEAD.size)
                if head is None:
                    raise OSError("GPU owner closed the connection")
                got_id, status, k = RSP_HEAD.unpack(head)
                if got_id != rid:
                    raise OSError("response id mismatch")
                body = _recv_exact(c["sock"], 8 * k) if status == 0 else b""
                if body is None:
                    raise OSError("truncated response")
            except OSError:
                self._discard(c)
                if attempt == 2:
                    raise
                continue
            self._release(owner, c)
            if status != 0:
                raise RuntimeError("classification failed in the GPU owner")
            return np.frombuffer(body[:4 * k], "<f4").copy(), np.frombuffer(body[4 * k:], "<i4").copy()
        raise OSError("unreachable")

    def classify_image(self, image: bytes):
        try:
            t0 = time.time()
            probs, labels = self.topk(image)
            dt = time.time() - t0
            names = [self.label_names[labels[0]], self.label_names[labels[1]], self.label_names[labels[2]]]
            return names, probs[:3], dt, sum(probs)
        except Exception as err:
            logging.info("Classification error: %s", err)
            return None

    def close(self):
        with self._all_lock:
            for c in self._all:
                try:
                    c["sock"].close()
                except OSError:
                    pass
                try:
                    c["shm"].close()
                    c["shm"].unlink()
                except Exception:
                    pass
            self._all.clear()


def owner_main(socket_path: str, weights_path: str, num_classes: int, device: int, max_batch: int, window_ms: float,
               precision: str, ready_path: Optional[str] = None):
    """Entry point of a GPU-owner process."""
    from .. import engine as _engine
    from ..PyClassification import decode_jpeg_rgb
    logging.basicConfig(level=logging.INFO)
    eng = _engine.Engine(weights_path, num_classes, device=device, max_batch=max_batch, precision=precision,
                         use_cuda_graph=True)
    # warm-up: capture the graphs of the common batch sizes before traffic arrives (PyClassification.py:221-226 warms
    # the reference up with one request)
    import io
    from PIL import Image
    buf = io.BytesIO()
    Image.fromarray(np.zeros((64, 64, 3), np.uint8)).save(buf, "JPEG")
    for n in sorted({1, 2, 4, 8, 16, max_batch}):
        if n <= max_batch:
            eng.classify_batch_jpeg([buf.getvalue()] * n, TOP_K)
    owner = GpuOwner(eng, socket_path, max_batch, window_ms, decode_fallback=decode_jpeg_rgb).start()
    if ready_path:
        open(ready_path, "w").write(str(os.getpid()))
    log.info("GPU owner pid %d: device %d, max_batch %d, window %.1f ms, socket %s", os.getpid(), device, max_batch,
             window_ms, socket_path)
    import json
    try:
        while True:
            if owner.fatal.wait(0.5):      # sticky CUDA error: answer nothing more from this context
                log.error("GPU owner pid %d on device %d exits after a CUDA error (exit code 3)", os.getpid(), device)
                os._exit(3)                # no destructors: they would call into the dead context
            if os.getppid() == 1:          # parent gone
                break
            try:                           # batch statistics for the load test (tools/http_bench.py)
                with open(socket_path + ".stats.tmp", "w") as f:
                    json.dump(owner.stats, f)
                os.replace(socket_path + ".stats.tmp", socket_path + ".stats")
            except OSError:
                pass
    except KeyboardInterrupt:
        pass
    finally:
        owner.stop()
        if os.environ.get("ICP_OWNER_STATS"):
            log.info("owner stats: %s", owner.stats)
        eng.close()
