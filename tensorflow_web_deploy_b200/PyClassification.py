"""Drop-in mirror of the reference's classifier class on the B200 engine.

Mirrors /root/reference/PyClassification.py:133-218 — same class name, constructor
arguments, attributes and ``classify_image`` contract — so the reference's Flask
handlers (PyClassification.py:54,76,108-109) and warm-up (:221-226) keep working when
they import ``ImagenetClassifier`` from here instead:

    classify_image(image: bytes) ->
        [names[0..2]], probs[:3] (float32 ndarray), seconds (float), sum(top-5 probs)
        or None on any error (the reference logs and swallows, :216-218).

What changed underneath: the single ``sess.run`` (:208-209) is now one C-ABI call into
libinception_b200.so (crop/resize/normalise + Inception-v3 tower + softmax + top-k on the
GPU); ``model_def_file`` names an ICPW weight blob (weights.py) instead of a TF checkpoint.
JPEG entropy decoding stays on the host (PIL/libjpeg), as TF's DecodeJpeg did.
"""
from __future__ import annotations

import io
import logging
import os
import threading
import time

import numpy as np

from . import engine as _engine

REPO_DIRNAME = os.path.abspath(os.path.dirname(os.path.abspath(__file__)) + "/../data")
NUM_CLASSES = 5        # PyClassification.py:20 (flowers-5; +1 background inside the model)
NUM_TOP_CLASSES = 5    # PyClassification.py:21

logger = logging.getLogger("PyClassification")


def decode_jpeg_rgb(image: bytes) -> np.ndarray:
    """tf.image.decode_jpeg(buffer, channels=3) (PyClassification.py:151): uint8 [H,W,3]."""
    from PIL import Image
    with Image.open(io.BytesIO(image)) as im:
        if im.format not in ("JPEG", "MPO"):
            # the reference graph only has DecodeJpeg; PNG uploads are re-encoded upstream (:93-96)
            raise ValueError("not a JPEG image (%s)" % im.format)
        return np.asarray(im.convert("RGB"), dtype=np.uint8)


class ImagenetClassifier(object):
    default_args = {
        "model_def_file": "{}/model".format(REPO_DIRNAME),
        "class_labels_file": "{}/synset.txt".format(REPO_DIRNAME),
    }

    def __init__(self, model_def_file, class_labels_file, device=None, precision=None, use_cuda_graph=True):
        logging.info("Loading net and associated files...")
        for key, val in (("model_def_file", model_def_file), ("class_labels_file", class_labels_file)):
            if not os.path.exists(val):   # the reference raises at import time, :140-143
                raise Exception("File for {} is missing. Should be at: {}".format(key, val))
        self.label_names = ["none"]                       # :175-178
        with open(class_labels_file, "rb") as f:
            for line in f.read().decode("utf-8").splitlines():
                self.label_names.append(line)
        # model_def_file: an ICPW blob, or the TF V1 checkpoint the reference restores (`data/model`, keys
        # <var>/ExponentialMovingAverage, :166-172) — converted once into a blob next to a temp dir
        self._converted = None
        from . import tf_checkpoint
        if tf_checkpoint.is_tf_checkpoint(model_def_file):
            import tempfile
            self._converted = tempfile.NamedTemporaryFile(prefix="icp_model_", suffix=".icpw", delete=False).name
            tf_checkpoint.convert(model_def_file, self._converted, len(self.label_names))
            model_def_file = self._converted
        device = int(os.environ.get("ICP_DEVICE", "0")) if device is None else device
        precision = os.environ.get("ICP_PRECISION", "fp16") if precision is None else precision
        self._engine_args = (model_def_file,)
        self._engine_kwargs = dict(num_classes=len(self.label_names), device=device, max_batch=1, precision=precision,
                                   use_cuda_graph=use_cuda_graph)
        self.engine = _engine.Engine(*self._engine_args, **self._engine_kwargs)
        if self._converted:                     # kept for a rebuild after a CUDA error (_cuda_lost); removed at exit
            import atexit
            atexit.register(lambda p=self._converted: os.path.exists(p) and os.unlink(p))
        self.top_k = min(NUM_TOP_CLASSES, len(self.label_names))
        # JPEG decode (ICP_JPEG): "pil" (default) = host libjpeg(-turbo) — the decoder family TF's DecodeJpeg wraps, so
        # the tower sees the pixels the reference's graph would — then icp_classify_rgb8; "nvjpeg" = on the GPU inside
        # the engine (icp_classify_jpeg: JPEG string in, top-k out, no host pixel buffer).  MEASURED on B200 (nvJPEG
        # GPU-hybrid backend, profiles/r02_nvjpeg_vs_libjpeg.md): decoded pixels differ from libjpeg's by up to 4 LSB
        # (mean 0.5-0.7; different IDCT / chroma upsampling), which moves the scores of data/sample.jpg by 1.2e-3 —
        # outside the 1e-3 bar, hence not the default.  Anything nvJPEG refuses (CMYK, arithmetic coding) falls back to PIL.
        self.decode_backend = os.environ.get("ICP_JPEG", "pil").lower()
        # The reference's tf.Session.run is thread-safe and Flask's development server / gthread workers are threaded;
        # the engine keeps one set of staging buffers, so requests of one process are serialised here (the C ABI
        # also locks per engine).  Cost: nothing next to the tower; throughput comes from the micro-batcher (serving/).
        self._lock = threading.Lock()

    def eval_image(self, image, height, width, scope=None):
        """Compat shim for PyClassification.py:180-203: decoded uint8 HWC -> the float32 [S,S,3]
        tensor the tower sees (crop 87.5 %, legacy bilinear resize, [-1,1]), computed on the GPU."""
        return self.engine.preprocess_rgb8(np.asarray(image, np.uint8))

    def classify_image(self, image):
        try:
            start_time = time.time()
            probs = None
            if self.decode_backend == "nvjpeg":
                try:
                    with self._lock:
                        probs, labels = self.engine.classify_jpeg(image, self.top_k)
                except _engine.EngineError as err:
                    if err.code != -1:              # ICP_E_INVALID: nvJPEG could not take this stream; try libjpeg
                        raise
            if probs is None:
                rgb = decode_jpeg_rgb(image)
                with self._lock:
                    probs, labels = self.engine.classify_rgb8(rgb, self.top_k)
            end_time = time.time()
            logger.info("classify_image cost %.2f secs", end_time - start_time)
            names = [self.label_names[labels[0]], self.label_names[labels[1]], self.label_names[labels[2]]]
            return names, probs[:3], end_time - start_time, sum(probs)
        except Exception as err:
            logging.info("Classification error: %s", err)
            if getattr(err, "code", 0) == -3:       # ICP_E_CUDA: sticky — see _cuda_lost
                self._cuda_lost(err)
            return None

    def _cuda_lost(self, err):
        """A CUDA error (a trapped kernel, a lost device) poisons this process's CUDA context: every later call fails too,
        and the reference's own recovery is gunicorn re-forking a worker that died.  One rebuild of the engine is tried (it
        succeeds for errors that were not sticky); if that fails and ICP_EXIT_ON_CUDA_ERROR=1 the process exits with status 3
        so that the pre-fork server replaces it, otherwise the failure is logged once more and the calls keep returning None."""
        with self._lock:
            try:
                self.engine.close()
            except Exception:
                pass
            try:
                self.engine = _engine.Engine(*self._engine_args, **self._engine_kwargs)
                logger.warning("engine rebuilt after a CUDA error (%s)", err)
                return
            except Exception as err2:
                logger.error("CUDA context lost (%s; rebuild failed: %s): this worker process must be restarted", err, err2)
        if os.environ.get("ICP_EXIT_ON_CUDA_ERROR", "0") == "1":
            os._exit(3)
