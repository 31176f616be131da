"""ctypes binding of libinception_b200.so (include/inception_b200.h).

There is deliberately no fallback: if the shared library is missing or no B200
is visible, constructing an ``Engine`` raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import numpy as np

PRECISION_FP16 = 0
PRECISION_BF16 = 1
_PREC = {"fp16": PRECISION_FP16, "bf16": PRECISION_BF16, 0: 0, 1: 1}

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libinception_b200.so")


class EngineError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("inception_b200 error %d: %s" % (code, msg))
        self.code = code


class _Config(C.Structure):
    _fields_ = [("weights_path", C.c_char_p), ("num_classes", C.c_int), ("device", C.c_int),
                ("max_batch", C.c_int), ("precision", C.c_int), ("image_size", C.c_int),
                ("enable_aux", C.c_int), ("use_cuda_graph", C.c_int)]


_lib = None
_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)
_u8p = C.POINTER(C.c_uint8)

# name -> (restype, argtypes); also the list tests/test_abi.py checks against the header
SIGNATURES = {
    "icp_version": (C.c_char_p, []),
    "icp_crc32c": (C.c_uint32, [C.c_char_p, C.c_size_t]),
    "icp_last_error": (C.c_char_p, [C.c_void_p]),
    "icp_create": (C.c_int, [C.POINTER(_Config), C.POINTER(C.c_void_p)]),
    "icp_destroy": (None, [C.c_void_p]),
    "icp_forward": (C.c_int, [C.c_void_p, _f32p, C.c_int, _f32p]),
    "icp_forward_topk": (C.c_int, [C.c_void_p, _f32p, C.c_int, C.c_int, _f32p, _f32p, _i32p]),
    "icp_forward_u8": (C.c_int, [C.c_void_p, _u8p, C.c_int, C.c_int, _f32p, _f32p, _i32p]),
    "icp_submit_u8": (C.c_int, [C.c_void_p, _u8p, C.c_int, C.c_int, _f32p, _f32p, _i32p, C.c_int]),
    "icp_collect": (C.c_int, [C.c_void_p, C.c_int]),
    "icp_submit_u8_device_out": (C.c_int, [C.c_void_p, _u8p, C.c_int, C.c_void_p, C.c_int]),
    "icp_num_ops": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "icp_op_info": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double),
                              C.POINTER(C.c_double)]),
    "icp_profile_forward": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _f32p]),
    "icp_forward_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "icp_forward_u8_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "icp_host_register": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "icp_host_unregister": (C.c_int, [C.c_void_p, C.c_void_p]),
    "icp_classify_rgb8": (C.c_int, [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, _f32p, _i32p]),
    "icp_preprocess_rgb8": (C.c_int, [C.c_void_p, _u8p, C.c_int, C.c_int, _f32p]),
    "icp_classify_jpeg": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t, C.c_int, _f32p, _i32p]),
    "icp_decode_jpeg": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t, _u8p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "icp_jpeg_backend": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "icp_classify_batch_rgb8": (C.c_int, [C.c_void_p, C.POINTER(_u8p), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_int,
                                          _f32p, _i32p]),
    "icp_classify_batch_jpeg": (C.c_int, [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_int, C.c_int, _f32p, _i32p]),
    "icp_endpoint_shape": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_int * 4)]),
    "icp_endpoint": (C.c_int, [C.c_void_p, C.c_char_p, _f32p]),
    "icp_last_forward_ms": (C.c_int, [C.c_void_p, _f32p]),
    "icp_launches_per_forward": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "icp_stream": (C.c_void_p, [C.c_void_p]),
    "icp_conv2d_bn_relu": (C.c_int, [C.c_int, C.c_int, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, C.c_int,
                                     C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _f32p, C.c_int, _f32p, C.c_int,
                                     _f32p]),
    "icp_pool2d": (C.c_int, [C.c_int, C.c_int, C.c_int, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                             C.c_int, _f32p]),
}


def load_library():
    """dlopen the in-tree library; raises (no fallback) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(-3, "%s not built — run `python -c 'import __graft_entry__ as g; g.build()'`; "
                              "there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _f32(a: np.ndarray):
    return a.ctypes.data_as(_f32p)


def _check(rc: int, handle=None):
    if rc != 0:
        msg = load_library().icp_last_error(handle)
        raise EngineError(rc, msg.decode(errors="replace") if msg else "")


class Engine:
    """One Inception-v3 tower on one B200."""

    def __init__(self, weights_path: str, num_classes: int, device: int = 0, max_batch: int = 1,
                 precision="fp16", image_size: int = 299, enable_aux: bool = False, use_cuda_graph: bool = False):
        self._lib = load_library()
        self._h = C.c_void_p()
        self.num_classes = int(num_classes)
        self.max_batch = int(max_batch)
        self.image_size = int(image_size)
        self.precision = _PREC[precision]
        cfg = _Config(os.fsencode(weights_path), self.num_classes, int(device), self.max_batch, self.precision,
                      self.image_size, int(bool(enable_aux)), int(bool(use_cuda_graph)))
        _check(self._lib.icp_create(C.byref(cfg), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.icp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ tower
    def forward(self, x: np.ndarray) -> np.ndarray:
        """x float32 [n,S,S,3] in [-1,1] -> softmax probabilities [n,C]."""
        x = self._images(x)
        probs = np.empty((x.shape[0], self.num_classes), np.float32)
        _check(self._lib.icp_forward(self._h, _f32(x), x.shape[0], _f32(probs)), self._h)
        return probs

    def forward_topk(self, x: np.ndarray, k: int = 5) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        x = self._images(x)
        n = x.shape[0]
        probs = np.empty((n, self.num_classes), np.float32)
        vals = np.empty((n, k), np.float32)
        idx = np.empty((n, k), np.int32)
        _check(self._lib.icp_forward_topk(self._h, _f32(x), n, k, _f32(probs), _f32(vals),
                                          idx.ctypes.data_as(_i32p)), self._h)
        return probs, vals, idx

    def forward_u8(self, x: np.ndarray, k: int = 5) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """x uint8 [n,S,S,3] (0..255, already at network size) -> (probs [n,C], vals [n,k], idx [n,k])."""
        x = np.ascontiguousarray(x, np.uint8)
        s = self.image_size
        if x.ndim != 4 or x.shape[1:] != (s, s, 3):
            raise ValueError("expected uint8 [n,%d,%d,3], got %s" % (s, s, x.shape))
        n = x.shape[0]
        probs = np.empty((n, self.num_classes), np.float32)
        vals = np.empty((n, k), np.float32)
        idx = np.empty((n, k), np.int32)
        _check(self._lib.icp_forward_u8(self._h, x.ctypes.data_as(_u8p), n, k, _f32(probs), _f32(vals),
                                        idx.ctypes.data_as(_i32p)), self._h)
        return probs, vals, idx

    def submit_u8(self, x: np.ndarray, slot: int, k: int = 5, out=None):
        """Asynchronous forward_u8 into `slot` (0/1).  `x` (and `out`, a (probs, vals, idx) triple of preallocated
        arrays — pinned for real overlap) must stay alive until collect(slot), which returns the triple."""
        if x.dtype != np.uint8 or not x.flags["C_CONTIGUOUS"]:
            raise ValueError("submit_u8 needs a C-contiguous uint8 array (no hidden copy: the transfer is asynchronous)")
        s = self.image_size
        if x.ndim != 4 or x.shape[1:] != (s, s, 3):
            raise ValueError("expected uint8 [n,%d,%d,3], got %s" % (s, s, x.shape))
        n = x.shape[0]
        if out is None:
            out = (np.empty((n, self.num_classes), np.float32), np.empty((n, k), np.float32), np.empty((n, k), np.int32))
        probs, vals, idx = out
        _check(self._lib.icp_submit_u8(self._h, x.ctypes.data_as(_u8p), n, k, _f32(probs), _f32(vals),
                                       idx.ctypes.data_as(_i32p), slot), self._h)
        if not hasattr(self, "_inflight"):
            self._inflight = {}
        self._inflight[slot] = (x, out)
        return out

    def collect(self, slot: int):
        _check(self._lib.icp_collect(self._h, slot), self._h)
        x, out = self._inflight.pop(slot)
        return out

    def submit_u8_device_out(self, x: np.ndarray, slot: int, probs_ptr: int) -> None:
        """submit_u8 whose probabilities stay on the device (float32 [n,C] at `probs_ptr`), for callers that feed a
        collective; ordered on Engine.stream; collect(slot) waits for the tower."""
        if x.dtype != np.uint8 or not x.flags["C_CONTIGUOUS"]:
            raise ValueError("submit_u8_device_out needs a C-contiguous uint8 array")
        _check(self._lib.icp_submit_u8_device_out(self._h, x.ctypes.data_as(_u8p), x.shape[0], C.c_void_p(probs_ptr), slot),
               self._h)
        if not hasattr(self, "_inflight"):
            self._inflight = {}
        self._inflight[slot] = (x, None)

    def op_table(self):
        """[(name, kind, flops_per_image, bytes_per_image)] for every launch of one forward pass."""
        n = C.c_int()
        _check(self._lib.icp_num_ops(self._h, C.byref(n)), self._h)
        out = []
        for i in range(n.value):
            buf = C.create_string_buffer(160)
            kind, fl, by = C.c_int(), C.c_double(), C.c_double()
            _check(self._lib.icp_op_info(self._h, i, buf, 160, C.byref(kind), C.byref(fl), C.byref(by)), self._h)
            out.append((buf.value.decode(), kind.value, fl.value, by.value))
        return out

    def profile_forward(self, n: int, reps: int = 5) -> np.ndarray:
        """Mean device milliseconds of each launch (CUDA events between launches) at batch n, on the
        input of the previous forward call."""
        nops = C.c_int()
        _check(self._lib.icp_num_ops(self._h, C.byref(nops)), self._h)
        ms = np.zeros(nops.value, np.float32)
        _check(self._lib.icp_profile_forward(self._h, n, reps, _f32(ms)), self._h)
        return ms

    def forward_device(self, x_ptr: int, n: int, probs_ptr: int = 0, stream: int = 0, u8: bool = False) -> None:
        """Enqueue (no sync) on device pointers; see icp_forward_device / icp_forward_u8_device (u8=True: the input is
        uint8 [n,S,S,3] 0..255)."""
        fn = self._lib.icp_forward_u8_device if u8 else self._lib.icp_forward_device
        _check(fn(self._h, C.c_void_p(x_ptr), n, C.c_void_p(probs_ptr or None), C.c_void_p(stream or None)), self._h)

    def classify_rgb8(self, rgb: np.ndarray, k: int = 5) -> Tuple[np.ndarray, np.ndarray]:
        rgb = np.ascontiguousarray(rgb, np.uint8)
        if rgb.ndim != 3 or rgb.shape[2] != 3:
            raise ValueError("expected uint8 [H,W,3]")
        vals = np.empty(k, np.float32)
        idx = np.empty(k, np.int32)
        _check(self._lib.icp_classify_rgb8(self._h, rgb.ctypes.data_as(_u8p), rgb.shape[0], rgb.shape[1], k,
                                           _f32(vals), idx.ctypes.data_as(_i32p)), self._h)
        return vals, idx

    # ------------------------------------------------------------------ JPEG in (nvJPEG on the GPU)
    JPEG_BACKENDS = {-1: "none", 0: "default", 1: "hybrid (CPU Huffman + GPU IDCT)", 2: "gpu_hybrid (GPU Huffman)",
                     3: "hardware engine", 4: "gpu_hybrid_device", 5: "hardware_device"}

    def classify_jpeg(self, jpeg: bytes, k: int = 5) -> Tuple[np.ndarray, np.ndarray]:
        """The reference's request: JPEG bytes -> (top-k probabilities, indices); decode, crop/resize, tower on the GPU."""
        vals = np.empty(k, np.float32)
        idx = np.empty(k, np.int32)
        _check(self._lib.icp_classify_jpeg(self._h, jpeg, len(jpeg), k, _f32(vals), idx.ctypes.data_as(_i32p)), self._h)
        return vals, idx

    def decode_jpeg(self, jpeg: bytes) -> np.ndarray:
        h, w = C.c_int(), C.c_int()
        _check(self._lib.icp_decode_jpeg(self._h, jpeg, len(jpeg), None, 0, C.byref(h), C.byref(w)), self._h)
        out = np.empty((h.value, w.value, 3), np.uint8)
        _check(self._lib.icp_decode_jpeg(self._h, jpeg, len(jpeg), out.ctypes.data_as(_u8p), out.nbytes, C.byref(h),
                                         C.byref(w)), self._h)
        return out

    @property
    def jpeg_backend(self) -> str:
        v = C.c_int()
        _check(self._lib.icp_jpeg_backend(self._h, C.byref(v)), self._h)
        return self.JPEG_BACKENDS.get(v.value, str(v.value))

    # ------------------------------------------------------------------ micro-batched requests
    def host_register(self, address: int, nbytes: int) -> bool:
        """Page-lock a host range that will be handed to the batch calls repeatedly (icp_host_register); False if the
        driver refuses (the range stays usable, copies are then staged)."""
        return self._lib.icp_host_register(self._h, C.c_void_p(address), C.c_size_t(nbytes)) == 0

    def host_unregister(self, address: int) -> None:
        self._lib.icp_host_unregister(self._h, C.c_void_p(address))

    def classify_batch_rgb8(self, images, k: int = 5) -> Tuple[np.ndarray, np.ndarray]:
        """images: sequence of uint8 [h,w,3] arrays of any sizes -> (vals [n,k], idx [n,k]) from ONE tower launch."""
        imgs = [np.ascontiguousarray(im, np.uint8) for im in images]
        n = len(imgs)
        for im in imgs:
            if im.ndim != 3 or im.shape[2] != 3:
                raise ValueError("expected uint8 [H,W,3]")
        ptrs = (_u8p * n)(*[im.ctypes.data_as(_u8p) for im in imgs])
        hs = (C.c_int * n)(*[im.shape[0] for im in imgs])
        ws = (C.c_int * n)(*[im.shape[1] for im in imgs])
        vals = np.empty((n, k), np.float32)
        idx = np.empty((n, k), np.int32)
        _check(self._lib.icp_classify_batch_rgb8(self._h, ptrs, hs, ws, n, k, _f32(vals), idx.ctypes.data_as(_i32p)), self._h)
        return vals, idx

    def classify_batch_jpeg(self, jpegs, k: int = 5) -> Tuple[np.ndarray, np.ndarray]:
        n = len(jpegs)
        ptrs = (C.c_char_p * n)(*jpegs)
        lens = (C.c_size_t * n)(*[len(j) for j in jpegs])
        vals = np.empty((n, k), np.float32)
        idx = np.empty((n, k), np.int32)
        _check(self._lib.icp_classify_batch_jpeg(self._h, ptrs, lens, n, k, _f32(vals), idx.ctypes.data_as(_i32p)), self._h)
        return vals, idx

    def preprocess_rgb8(self, rgb: np.ndarray) -> np.ndarray:
        rgb = np.ascontiguousarray(rgb, np.uint8)
        out = np.empty((self.image_size, self.image_size, 3), np.float32)
        _check(self._lib.icp_preprocess_rgb8(self._h, rgb.ctypes.data_as(_u8p), rgb.shape[0], rgb.shape[1],
                                             _f32(out)), self._h)
        return out

    def endpoint(self, name: str) -> np.ndarray:
        dims = (C.c_int * 4)()
        _check(self._lib.icp_endpoint_shape(self._h, name.encode(), C.byref(dims)), self._h)
        n, h, w, c = list(dims)
        out = np.empty((n, h, w, c), np.float32)
        _check(self._lib.icp_endpoint(self._h, name.encode(), _f32(out)), self._h)
        return out.reshape(n, c) if h == 1 and w == 1 and name in ("logits", "predictions", "aux_logits") else out

    @property
    def last_forward_ms(self) -> float:
        v = C.c_float()
        _check(self._lib.icp_last_forward_ms(self._h, C.byref(v)), self._h)
        return float(v.value)

    @property
    def launches_per_forward(self) -> int:
        v = C.c_int()
        _check(self._lib.icp_launches_per_forward(self._h, C.byref(v)), self._h)
        return int(v.value)

    @property
    def stream(self) -> int:
        return int(self._lib.icp_stream(self._h) or 0)

    def _images(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, np.float32)
        s = self.image_size
        if x.ndim != 4 or x.shape[1:] != (s, s, 3):
            raise ValueError("expected float32 [n,%d,%d,3], got %s" % (s, s, x.shape))
        return x


def conv2d_bn_relu(x: np.ndarray, w_hwio: np.ndarray, scale: np.ndarray, shift: np.ndarray, stride: int = 1,
                   same: bool = True, relu: bool = True, precision="fp16", device: int = 0,
                   reps: int = 0) -> Tuple[np.ndarray, Optional[float]]:
    """One ops.conv2d layer through the tcgen05 kernel (host arrays in/out)."""
    lib = load_library()
    x = np.ascontiguousarray(x, np.float32)
    w_hwio = np.ascontiguousarray(w_hwio, np.float32)
    scale = np.ascontiguousarray(scale, np.float32)
    shift = np.ascontiguousarray(shift, np.float32)
    n, h, w, cin = x.shape
    kh, kw, cin2, cout = w_hwio.shape
    assert cin == cin2
    ph, pw = ((kh - 1) // 2, (kw - 1) // 2) if same else (0, 0)
    oh, ow = (h + 2 * ph - kh) // stride + 1, (w + 2 * pw - kw) // stride + 1
    y = np.empty((n, oh, ow, cout), np.float32)
    ms = C.c_float(0)
    _check(lib.icp_conv2d_bn_relu(device, _PREC[precision], _f32(x), n, h, w, cin, _f32(w_hwio), kh, kw, cout, stride,
                                  int(same), _f32(scale), _f32(shift), int(relu), _f32(y), reps,
                                  C.byref(ms) if reps else None))
    return y, (float(ms.value) if reps else None)


def pool2d(x: np.ndarray, kind: str, k: int = 3, stride: int = 1, same: bool = True, precision="fp16",
           device: int = 0) -> np.ndarray:
    lib = load_library()
    x = np.ascontiguousarray(x, np.float32)
    n, h, w, c = x.shape
    if kind == "max":
        k, stride, same = 3, 2, False
    pad = (k - 1) // 2 if same else 0
    oh, ow = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    y = np.empty((n, oh, ow, c), np.float32)
    _check(lib.icp_pool2d(device, _PREC[precision], 0 if kind == "max" else 1, _f32(x), n, h, w, c, k, stride,
                          int(same), _f32(y)))
    return y
