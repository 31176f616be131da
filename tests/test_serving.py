"""Front-end shim + micro-batcher without a GPU: the three routes of the reference
(/root/reference/PyClassification.py:29-113) against a fake classifier, and the cross-process batching protocol
against a fake engine."""
import io
import threading
import time

import numpy as np
import pytest
from werkzeug.test import Client

from tensorflow_web_deploy_b200.serving import app as A
from tensorflow_web_deploy_b200.serving import batcher as B


def _jpeg(h=40, w=50, fmt="JPEG"):
    from PIL import Image
    buf = io.BytesIO()
    Image.fromarray(np.random.default_rng(1).integers(0, 256, (h, w, 3), dtype=np.uint8)).save(buf, fmt)
    return buf.getvalue()


class FakeClf:
    def __init__(self):
        self.seen = []

    def classify_image(self, data):
        self.seen.append(data)
        if data.startswith(b"BAD"):
            return None
        return ["roses", "sunflowers", "dandelion"], np.float32([0.7, 0.2, 0.05]), 0.0123, np.float32(0.99)


@pytest.fixture
def client():
    clf = FakeClf()
    return Client(A.ClassifierApp(clf)), clf


def test_post_raw_body_is_the_webservice_call(client):
    """README.md:30: curl --request POST --data-binary @sample.jpg http://host:port/ (PyClassification.py:47-55)."""
    c, clf = client
    body = _jpeg()
    r = c.post("/", data=body)
    assert r.status_code == 200
    assert r.get_data(as_text=True) == "roses,sunflowers,dandelion"
    assert float(r.headers["ClassificationAccuracy"]) == pytest.approx(0.99)
    assert clf.seen == [body]
    assert c.post("/", data=b"").status_code == 400                 # :49-52
    assert c.post("/", data=b"BAD bytes").status_code == 500        # classify_image -> None -> unpack error -> 500


def test_get_index_file_and_page(client, tmp_path):
    c, clf = client
    r = c.get("/")
    assert r.status_code == 200 and "<form" in r.get_data(as_text=True) and not clf.seen      # has_result=False page
    p = tmp_path / "x.jpg"
    p.write_bytes(_jpeg())
    r = c.get("/", query_string={"file": str(p)})
    assert r.get_data(as_text=True) == "roses,sunflowers,dandelion" and clf.seen == [p.read_bytes()]
    assert c.get("/nope").status_code == 404
    assert c.post("/classify_url").status_code == 405


def test_classify_upload_renders_result_and_converts_png(client):
    c, clf = client
    r = c.post("/classify_upload", data={"imagefile": (io.BytesIO(_jpeg()), "flower.jpg")}, content_type="multipart/form-data")
    page = r.get_data(as_text=True)
    assert r.status_code == 200 and "roses" in page and "0.012 seconds" in page and "data:image/png;base64," in page
    r = c.post("/classify_upload", data={"imagefile": (io.BytesIO(_jpeg(fmt="PNG")), "flower.png")},
               content_type="multipart/form-data")
    assert r.status_code == 200 and clf.seen[-1][:2] == b"\xff\xd8"          # the model only ever sees JPEG (:93-96)
    r = c.post("/classify_upload", data={"imagefile": (io.BytesIO(b"not an image"), "x.jpg")}, content_type="multipart/form-data")
    assert "Cannot open uploaded image." in r.get_data(as_text=True)
    r = c.get("/classify_url", query_string={"imageurl": "file:///definitely/not/here.jpg"})
    assert "Cannot open image from URL." in r.get_data(as_text=True)


def test_classify_url_with_a_file_url(client, tmp_path):
    c, clf = client
    p = tmp_path / "y.jpg"
    p.write_bytes(_jpeg())
    r = c.get("/classify_url", query_string={"imageurl": "file://" + str(p)})
    page = r.get_data(as_text=True)
    assert r.status_code == 200 and "sunflowers" in page and clf.seen[-1] == p.read_bytes()


# ------------------------------------------------------------------------------------------------ batcher
class FakeEngine:
    """classify_batch_jpeg answers with a value derived from each payload so that routing errors are visible."""

    def __init__(self, delay=0.0):
        self.batches, self.delay, self.lock = [], delay, threading.Lock()

    def classify_batch_jpeg(self, jpegs, k):
        with self.lock:
            self.batches.append(len(jpegs))
        if any(j.startswith(b"BAD") for j in jpegs):
            raise RuntimeError("undecodable stream in the batch")
        time.sleep(self.delay)
        vals = np.stack([np.float32([len(j) % 97, 4, 3, 2, 1]) / 100 for j in jpegs])
        idx = np.stack([np.int32([len(j) % 6, 1, 2, 3, 4]) for j in jpegs])
        return vals, idx

    def classify_rgb8(self, rgb, k):
        return np.float32([0.5, 0.2, 0.1, 0.1, 0.1]), np.int32([5, 4, 3, 2, 1])


LABELS = ["none", "daisy", "dandelion", "roses", "sunflowers", "tulips"]


def test_concurrent_requests_are_merged_and_routed_back(tmp_path):
    eng = FakeEngine(delay=0.02)
    owner = B.GpuOwner(eng, str(tmp_path / "o.sock"), max_batch=8, window_ms=30.0).start()
    cli = B.BatcherClient([owner.path], LABELS, decode="owner")
    payloads = [bytes([65 + i]) * (100 + 13 * i) for i in range(24)]
    out = [None] * len(payloads)

    def work(i):
        out[i] = cli.classify_image(payloads[i])

    ts = [threading.Thread(target=work, args=(i,)) for i in range(len(payloads))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for i, p in enumerate(payloads):
        names, probs, secs, acc = out[i]
        assert names == [LABELS[len(p) % 6], "daisy", "dandelion"]                       # the answer to THIS payload
        assert probs[0] == pytest.approx((len(p) % 97) / 100) and len(probs) == 3 and secs > 0
        assert acc == pytest.approx((len(p) % 97 + 10) / 100)
    assert sum(eng.batches) == len(payloads) and max(eng.batches) > 1 and max(eng.batches) <= 8
    assert len(eng.batches) < len(payloads)                                              # requests really were merged
    assert owner.stats["requests"] == 24 and owner.stats["max_batch_seen"] == max(eng.batches)
    cli.close()
    owner.stop()


def test_bad_stream_does_not_poison_its_batch_mates(tmp_path):
    from tensorflow_web_deploy_b200.PyClassification import decode_jpeg_rgb
    eng = FakeEngine()
    owner = B.GpuOwner(eng, str(tmp_path / "o.sock"), max_batch=4, window_ms=50.0, decode_fallback=decode_jpeg_rgb).start()
    cli = B.BatcherClient([owner.path], LABELS, decode="owner")
    payloads = [b"BAD not a jpeg", b"x" * 300, b"y" * 301]
    out = [None] * 3
    ts = [threading.Thread(target=lambda i=i: out.__setitem__(i, cli.classify_image(payloads[i]))) for i in range(3)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert out[0] is None                                   # the reference's None-on-error (:216-218)
    assert out[1][0][0] == LABELS[300 % 6] and out[2][0][0] == LABELS[301 % 6]
    assert owner.stats["errors"] == 1
    cli.close()
    owner.stop()


def test_round_robin_over_owners_and_reconnect(tmp_path):
    engs = [FakeEngine(), FakeEngine()]
    owners = [B.GpuOwner(e, str(tmp_path / ("o%d.sock" % i)), max_batch=4, window_ms=1.0).start() for i, e in enumerate(engs)]
    cli = B.BatcherClient([o.path for o in owners], LABELS, decode="owner")
    for i in range(10):
        assert cli.classify_image(b"z" * (50 + i)) is not None
    assert sum(engs[0].batches) == 5 and sum(engs[1].batches) == 5        # one GPU owner per GPU, used alternately
    owners[0].stop()                                                       # owner restart: the client reconnects
    owners[0] = B.GpuOwner(engs[0], owners[0].path, max_batch=4, window_ms=1.0).start()
    time.sleep(0.05)
    res = [cli.classify_image(b"w" * 70) for _ in range(4)]
    assert all(r is not None for r in res)
    cli.close()
    [o.stop() for o in owners]


@pytest.mark.parametrize("pin", ["0", "1"])
def test_worker_side_decode_hands_over_pixels(tmp_path, monkeypatch, pin):
    """decode="worker" (default): libjpeg in the worker, uint8 [h,w,3] pixels through the segment, ONE
    classify_batch_rgb8 call for the merged requests; a non-JPEG body never reaches the owner.  pin=1 (ICP_SERVE_PIN): the
    segments are page-locked through the engine and the engine sees views into them instead of copies."""
    from PIL import Image
    monkeypatch.setenv("ICP_SERVE_PIN", pin)
    registered = []

    class RgbEngine(FakeEngine):
        def __init__(self):
            super().__init__(delay=0.02)
            self.shapes = []
            self.owned = []

        def host_register(self, address, nbytes):
            registered.append((address, nbytes))
            return True

        def host_unregister(self, address):
            registered.append((address, -1))

        def classify_batch_rgb8(self, imgs, k):
            with self.lock:
                self.batches.append(len(imgs))
                self.shapes += [im.shape for im in imgs]
                self.owned += [bool(im.flags["OWNDATA"]) or im.base is None or isinstance(im.base, bytes) or
                               (isinstance(im.base, np.ndarray) and isinstance(im.base.base, bytes)) for im in imgs]
            time.sleep(self.delay)
            vals = np.stack([np.float32([int(im[0, 0, 0]) % 50, 4, 3, 2, 1]) / 100 for im in imgs])
            idx = np.stack([np.int32([im.shape[0] % 6, 1, 2, 3, 4]) for im in imgs])
            return vals, idx

    eng = RgbEngine()
    owner = B.GpuOwner(eng, str(tmp_path / "o.sock"), max_batch=8, window_ms=30.0).start()
    cli = B.BatcherClient([owner.path], LABELS)
    assert cli.decode == "worker"
    jpegs = [_jpeg(h=40 + i, w=50 + 2 * i) for i in range(6)]
    want = [np.asarray(Image.open(io.BytesIO(j)).convert("RGB")) for j in jpegs]
    out = [None] * len(jpegs)
    ts = [threading.Thread(target=lambda i=i: out.__setitem__(i, cli.classify_image(jpegs[i]))) for i in range(len(jpegs))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for i, o in enumerate(out):
        assert o[0][0] == LABELS[(40 + i) % 6]
        assert o[1][0] == pytest.approx((int(want[i][0, 0, 0]) % 50) / 100)       # the owner saw libjpeg's pixels
    assert sorted(eng.shapes) == sorted(w.shape for w in want) and max(eng.batches) > 1
    assert cli.classify_image(b"definitely not a jpeg") is None and owner.stats["requests"] == 6
    cli.close()
    owner.stop()
    if pin == "1":
        assert any(n > 0 for _, n in registered) and not any(eng.owned)      # segments registered, views (not copies) handed over
    else:
        assert not registered and all(eng.owned)


def test_cuda_error_marks_the_owner_fatal(tmp_path):
    """An engine error with code ICP_E_CUDA (-3) = the owner's CUDA context is lost: the waiting requests get the error
    answer (-> None, like PyClassification.py:216-218), nothing is retried on the dead context, and `fatal` is set so that
    owner_main exits and serving/server.py starts a replacement."""
    class CudaErr(Exception):
        code = -3

    class DeadEngine(FakeEngine):
        def classify_batch_jpeg(self, jpegs, k):
            with self.lock:
                self.batches.append(len(jpegs))
            raise CudaErr("unspecified launch failure")

    eng = DeadEngine()
    owner = B.GpuOwner(eng, str(tmp_path / "o.sock"), max_batch=4, window_ms=1.0).start()
    cli = B.BatcherClient([owner.path], LABELS, decode="owner")
    assert cli.classify_image(b"x" * 50) is None
    assert owner.fatal.wait(2.0)
    assert eng.batches == [1]                              # no one-by-one retry on a dead context
    cli.close()
    owner.stop()


def test_connections_are_pooled_not_per_thread(tmp_path):
    """A threaded WSGI server runs every request on a fresh thread: the client must reuse its connections (one shared-memory
    segment + socket each), growing the pool only to the number of requests in flight."""
    eng = FakeEngine()
    owner = B.GpuOwner(eng, str(tmp_path / "o.sock"), max_batch=8, window_ms=1.0).start()
    cli = B.BatcherClient([owner.path], LABELS, decode="owner")
    for i in range(40):                                       # 40 sequential requests, each on its own thread
        t = threading.Thread(target=lambda: cli.classify_image(b"q" * 64))
        t.start()
        t.join()
    assert len(cli._all) == 1
    ts = [threading.Thread(target=lambda: cli.classify_image(b"r" * 65)) for _ in range(6)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert 1 <= len(cli._all) <= 6 and owner.stats["requests"] == 46
    cli.close()
    owner.stop()
