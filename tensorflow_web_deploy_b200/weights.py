"""Weight blob format + seeded synthetic weights.

The reference restores a TF V1 checkpoint at ``data/model`` keyed by
``<variable>/ExponentialMovingAverage`` (PyClassification.py:166-172).  That
checkpoint was never published (reference README.md:14), so this build defines
its own container, keyed by the *reference's variable names* and holding the
reference's layouts (conv HWIO fp32, fc [in,out] fp32):

    "ICPWv001" | u32 count | count x { u16 len | name | u8 ndim | u32 dims[ndim] | f32 data }

(all little-endian).  csrc/engine.cu (read_blob) reads the same format; tf_checkpoint.py converts the TF V1
checkpoint the reference restores into it.
"""
from __future__ import annotations

import struct
from typing import Dict

import numpy as np

from . import tower_spec

MAGIC = b"ICPWv001"


def save_blob(path: str, tensors: Dict[str, np.ndarray]) -> None:
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<I", len(tensors)))
        for name, a in tensors.items():
            a = np.ascontiguousarray(a, dtype="<f4")
            nb = name.encode()
            f.write(struct.pack("<H", len(nb)))
            f.write(nb)
            f.write(struct.pack("<B", a.ndim))
            f.write(struct.pack("<%dI" % a.ndim, *a.shape))
            f.write(a.tobytes())


def load_blob(path: str) -> Dict[str, np.ndarray]:
    out: Dict[str, np.ndarray] = {}
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError("%s: not an ICPW weight blob" % path)
        (count,) = struct.unpack("<I", f.read(4))
        for _ in range(count):
            (ln,) = struct.unpack("<H", f.read(2))
            name = f.read(ln).decode()
            (nd,) = struct.unpack("<B", f.read(1))
            dims = struct.unpack("<%dI" % nd, f.read(4 * nd))
            n = int(np.prod(dims)) if nd else 1
            out[name] = np.frombuffer(f.read(4 * n), dtype="<f4").reshape(dims).copy()
    return out


def synthetic_weights(num_classes: int, seed: int = 1234, height: int = 299, width: int = 299) -> Dict[str, np.ndarray]:
    """Seeded stand-in for the missing checkpoint.

    conv W ~ N(0, 2/(kh*kw*Cin)) keeps activations O(1) through 47 ReLU layers
    (the reference's own init, truncated_normal(stddev=0.1) at
    inception/inception_model.py:78, explodes without training); BN statistics
    are randomised so the scale/shift fold is exercised; the FC is scaled so the
    softmax is neither uniform nor saturated.
    """
    rng = np.random.default_rng(seed)
    t: Dict[str, np.ndarray] = {}
    for name, shape in tower_spec.variables(num_classes, height, width):
        leaf = name.rsplit("/", 1)[1]
        if leaf == "weights" and len(shape) == 4:
            kh, kw, cin, _ = shape
            a = rng.standard_normal(shape, dtype=np.float32) * np.float32(np.sqrt(2.0 / (kh * kw * cin)))
        elif leaf == "weights":
            a = rng.standard_normal(shape, dtype=np.float32) * np.float32(0.5 / np.sqrt(shape[0]))
        elif leaf == "biases" or leaf == "beta" or leaf == "moving_mean":
            a = rng.standard_normal(shape, dtype=np.float32) * np.float32(0.1)
        elif leaf == "moving_variance":
            a = rng.uniform(0.5, 1.5, shape).astype(np.float32)
        else:
            raise AssertionError(name)
        t[name] = a
    return t
