"""Host-side mirror of PyClassification.ImagenetClassifier (reference :133-218) with the engine faked:
the drop-in contract — return tuple, label table, None-on-error — does not need a GPU."""
import io

import numpy as np
import pytest

from tensorflow_web_deploy_b200 import PyClassification as P


class FakeEngine:
    def __init__(self, *a, **k):
        self.args, self.kw = a, k
        self.calls = []

    def classify_rgb8(self, rgb, k):
        self.calls.append(rgb.shape)
        return np.float32([0.5, 0.25, 0.125, 0.0625, 0.03125])[:k], np.int32([3, 0, 5, 1, 2])[:k]

    def preprocess_rgb8(self, rgb):
        return np.zeros((299, 299, 3), np.float32)


def _jpeg(h=60, w=80, mode="RGB"):
    from PIL import Image
    arr = np.random.default_rng(0).integers(0, 256, (h, w, 3), dtype=np.uint8)
    im = Image.fromarray(arr).convert(mode)
    buf = io.BytesIO()
    im.save(buf, "JPEG")
    return buf.getvalue()


@pytest.fixture
def clf(monkeypatch, tmp_path):
    monkeypatch.setattr(P._engine, "Engine", FakeEngine)
    monkeypatch.delenv("ICP_JPEG", raising=False)     # default: host libjpeg decode -> classify_rgb8
    labels = tmp_path / "synset.txt"
    labels.write_text("daisy\ndandelion\nroses\nsunflowers\ntulips\n")
    model = tmp_path / "model"
    model.write_bytes(b"x")
    return P.ImagenetClassifier(str(model), str(labels))


def test_label_table_and_engine_config(clf):
    assert clf.label_names == ["none", "daisy", "dandelion", "roses", "sunflowers", "tulips"]   # :175-178
    assert clf.engine.kw["num_classes"] == 6                                                    # NUM_CLASSES + 1
    assert P.NUM_CLASSES == 5 and P.NUM_TOP_CLASSES == 5
    assert set(P.ImagenetClassifier.default_args) == {"model_def_file", "class_labels_file"}


def test_classify_image_return_contract(clf):
    out = clf.classify_image(_jpeg())
    names, probs, secs, acc = out                     # callers unpack 4 values (:54)
    assert names == ["roses", "none", "tulips"]       # label_names[idx[0..2]]
    assert isinstance(probs, np.ndarray) and probs.dtype == np.float32 and probs.tolist() == [0.5, 0.25, 0.125]
    assert isinstance(secs, float) and secs >= 0
    assert acc == pytest.approx(0.96875)              # sum of the top-5 probabilities (:215)
    assert clf.engine.calls == [(60, 80, 3)]
    assert list(zip(names, probs))                    # handlers zip them (:78,111)


def test_grayscale_jpeg_is_expanded_to_rgb(clf):
    assert clf.classify_image(_jpeg(mode="L")) is not None
    assert clf.engine.calls[-1] == (60, 80, 3)        # decode_jpeg(channels=3)


def test_errors_return_none(clf, monkeypatch):
    assert clf.classify_image(b"definitely not a jpeg") is None        # :216-218
    from PIL import Image
    buf = io.BytesIO()
    Image.new("RGB", (8, 8)).save(buf, "PNG")
    assert clf.classify_image(buf.getvalue()) is None                   # graph only has DecodeJpeg
    monkeypatch.setattr(clf.engine, "classify_rgb8", lambda *a: (_ for _ in ()).throw(RuntimeError("gpu")))
    assert clf.classify_image(_jpeg()) is None


def test_cuda_error_rebuilds_the_engine_once(clf, monkeypatch):
    """ICP_E_CUDA (-3) is sticky for the context: the classifier answers None, rebuilds its engine (which works when the
    error was not fatal), and the next request is served by the new engine."""
    first = clf.engine
    first.close = lambda: first.calls.append("closed")

    def boom(*a):
        raise P._engine.EngineError(-3, "an illegal memory access was encountered")

    monkeypatch.setattr(first, "classify_rgb8", boom)
    assert clf.classify_image(_jpeg()) is None
    assert clf.engine is not first and first.calls[-1] == "closed"
    assert clf.engine.kw["num_classes"] == 6 and clf.engine.kw["max_batch"] == 1
    assert clf.classify_image(_jpeg()) is not None


def test_missing_files_raise(monkeypatch, tmp_path):
    monkeypatch.setattr(P._engine, "Engine", FakeEngine)
    with pytest.raises(Exception, match="missing"):
        P.ImagenetClassifier(str(tmp_path / "nope"), str(tmp_path / "nope2"))     # :140-143


def test_nvjpeg_route_and_fallback_to_libjpeg(monkeypatch, tmp_path):
    """ICP_JPEG=nvjpeg: the JPEG bytes go to the engine untouched (icp_classify_jpeg); a stream nvJPEG refuses
    (ICP_E_INVALID) is retried through PIL -> classify_rgb8; any other engine error is an error (None)."""
    class JpegEngine(FakeEngine):
        mode = "ok"

        def classify_jpeg(self, data, k):
            self.calls.append(("jpeg", len(data)))
            if self.mode == "refuse":
                raise P._engine.EngineError(-1, "unsupported stream")
            if self.mode == "broken":
                raise P._engine.EngineError(-3, "cuda")
            return np.float32([0.4, 0.3, 0.2, 0.05, 0.05])[:k], np.int32([1, 2, 3, 4, 5])[:k]

    monkeypatch.setattr(P._engine, "Engine", JpegEngine)
    monkeypatch.setenv("ICP_JPEG", "nvjpeg")
    labels = tmp_path / "synset.txt"
    labels.write_text("daisy\ndandelion\nroses\nsunflowers\ntulips\n")
    model = tmp_path / "model"
    model.write_bytes(b"x")
    clf = P.ImagenetClassifier(str(model), str(labels))
    assert clf.decode_backend == "nvjpeg"
    data = _jpeg()
    out = clf.classify_image(data)
    assert out[0] == ["daisy", "dandelion", "roses"] and clf.engine.calls == [("jpeg", len(data))]
    clf.engine.mode = "refuse"
    out = clf.classify_image(data)
    assert out[0] == ["roses", "none", "tulips"] and clf.engine.calls[-1] == (60, 80, 3)
    clf.engine.mode = "broken"
    assert clf.classify_image(data) is None
