"""TensorFlow V1 ("tensor slice") checkpoint reader / writer / converter  (SURVEY 8(f)3).

The reference restores `data/model` with
    variable_averages = tf.train.ExponentialMovingAverage(0.9999)
    saver = tf.train.Saver(variable_averages.variables_to_restore()); saver.restore(sess, model_def_file)
(/root/reference/PyClassification.py:166-172): a TF r0.9 Saver file in which every trainable and every BatchNorm
moving statistic is stored under `<variable name>/ExponentialMovingAverage`.  This module reads such a file
without TensorFlow and turns it into the ICPW blob the engine loads (weights.py), so a user's
`model.ckpt-5000` (reference README.md:16) drops in; `ImagenetClassifier` does the conversion on the fly when
`model_def_file` is not an ICPW blob.

Format (TF <= 0.11 `Saver`, tensorflow/core/util/tensor_slice_writer.cc + tensorflow/core/lib/io/table*):
  * the file is an SSTable (TF's port of LevelDB's table): data blocks of prefix-compressed (key, value) entries
    with a restart array, each followed by a 5-byte trailer (compression type, masked CRC32C); an index block;
    an (empty) metaindex block; a 48-byte footer ending in the magic 0xdb4775248b80fb57.
  * key ""  ->  SavedTensorSlices{ meta: SavedTensorSliceMeta{ tensor: [SavedSliceMeta{name, shape, type, slice}] } }
  * key OrderedCode(0, name, ndims, (start, length)*)  ->  SavedTensorSlices{ data: SavedSlice{name, slice, data: TensorProto} }
    (float tensors travel in TensorProto.float_val; full-tensor slices have start 0 / length -1.)
No real checkpoint of this model exists here (reference README.md:14), so the writer below produces files in the
same format from a name->array dict and the tests round-trip through it; field numbers and encodings follow the
published .proto / source files named above.
"""
from __future__ import annotations

import struct
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
EMA_SUFFIX = "/ExponentialMovingAverage"
DT_FLOAT, DT_DOUBLE, DT_INT32, DT_INT64 = 1, 2, 3, 9
_NP_OF_DT = {DT_FLOAT: np.float32, DT_DOUBLE: np.float64, DT_INT32: np.int32, DT_INT64: np.int64}
BLOCK_SIZE = 4096
RESTART_INTERVAL = 16


# ---------------------------------------------------------------------------------------------- crc32c
_CRC_TABLE = None


def _crc_table():
    global _CRC_TABLE
    if _CRC_TABLE is None:
        t = np.zeros(256, np.uint32)
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ (0x82F63B78 if c & 1 else 0)
            t[i] = c
        _CRC_TABLE = t
    return _CRC_TABLE


def crc32c(data: bytes) -> int:
    """CRC-32C (Castagnoli).  Uses the engine library's implementation when it is loadable (1 GB/s), else Python."""
    try:
        from . import engine as _engine
        lib = _engine.load_library()
        return int(lib.icp_crc32c(data, len(data)))
    except Exception:
        pass
    t = _crc_table()
    c = 0xFFFFFFFF
    for b in data:
        c = int(t[(c ^ b) & 0xFF]) ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def _mask(crc: int) -> int:
    return (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ---------------------------------------------------------------------------------------------- varints / protobuf
def _put_varint(n: int) -> bytes:
    n &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _get_varint(b, i: int) -> Tuple[int, int]:
    v = s = 0
    while True:
        c = b[i]
        i += 1
        v |= (c & 0x7F) << s
        if c < 0x80:
            return v, i
        s += 7


def _fields(b) -> Iterable[Tuple[int, int, object]]:
    i, n = 0, len(b)
    while i < n:
        key, i = _get_varint(b, i)
        f, wt = key >> 3, key & 7
        if wt == 0:
            v, i = _get_varint(b, i)
        elif wt == 2:
            ln, i = _get_varint(b, i)
            v = b[i:i + ln]
            i += ln
        elif wt == 5:
            v = b[i:i + 4]
            i += 4
        elif wt == 1:
            v = b[i:i + 8]
            i += 8
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield f, wt, v


def _pb(field: int, wt: int, payload) -> bytes:
    key = _put_varint((field << 3) | wt)
    if wt == 0:
        return key + _put_varint(payload)
    return key + _put_varint(len(payload)) + payload


def _signed(v: int) -> int:
    return v if v < (1 << 63) else v - (1 << 64)


# ---------------------------------------------------------------------------------------------- OrderedCode (keys)
def _oc_num_increasing(n: int) -> bytes:
    body = b"" if n == 0 else n.to_bytes((n.bit_length() + 7) // 8, "big")
    return bytes([len(body)]) + body


def _oc_string(s: bytes) -> bytes:
    """OrderedCode::WriteString: \x00 -> \x00\xff, \xff -> \xff\x00, terminator \x00\x01."""
    out = bytearray()
    for c in s:
        if c == 0:
            out += b"\x00\xff"
        elif c == 0xFF:
            out += b"\xff\x00"
        else:
            out.append(c)
    return bytes(out) + b"\x00\x01"


_OC_HEADER = [(0, 0), (0x80, 0), (0xC0, 0), (0xE0, 0), (0xF0, 0), (0xF8, 0), (0xFC, 0), (0xFE, 0), (0xFF, 0), (0xFF, 0x80),
              (0xFF, 0xC0)]


def _oc_signed_num_increasing(v: int) -> bytes:
    """OrderedCode::WriteSignedNumIncreasing (tensorflow/core/lib/strings/ordered_code.cc): the value sign-extended
    to 10 big-endian bytes, the last `len` of them kept (len = bits/7 + 1), XORed with a unary length header."""
    x = ~v if v < 0 else v
    if x < 64:
        return bytes([(0x80 ^ v) & 0xFF])
    length = x.bit_length() // 7 + 1
    buf = bytearray((v & ((1 << 80) - 1)).to_bytes(10, "big"))
    begin = 10 - length
    buf[begin] ^= _OC_HEADER[length][0]
    buf[begin + 1] ^= _OC_HEADER[length][1]
    return bytes(buf[begin:])


def encode_tensor_name_slice(name: str, ndims: int) -> bytes:
    """checkpoint::EncodeTensorNameSlice for a full-tensor slice (start 0, length -1 in every dimension)."""
    out = _oc_num_increasing(0) + _oc_string(name.encode()) + _oc_num_increasing(ndims)
    for _ in range(ndims):
        out += _oc_signed_num_increasing(0) + _oc_signed_num_increasing(-1)
    return out


# ---------------------------------------------------------------------------------------------- snappy (read side)
def _snappy_uncompress(src: bytes) -> bytes:
    n, i = _get_varint(src, 0)
    out = bytearray()
    while i < len(src):
        tag = src[i]
        i += 1
        t = tag & 3
        if t == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(src[i:i + nb], "little")
                i += nb
            ln += 1
            out += src[i:i + ln]
            i += ln
            continue
        if t == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | src[i]
            i += 1
        elif t == 2:
            ln = (tag >> 2) + 1
            off = src[i] | (src[i + 1] << 8)
            i += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(src[i:i + 4], "little")
            i += 4
        if off == 0 or off > len(out):
            raise ValueError("corrupt snappy block")
        for _ in range(ln):
            out.append(out[-off])
    if len(out) != n:
        raise ValueError("corrupt snappy block (length)")
    return bytes(out)


# ---------------------------------------------------------------------------------------------- table reader
def _read_block(data: bytes, offset: int, size: int, verify: bool) -> bytes:
    raw = data[offset:offset + size]
    if len(raw) != size or offset + size + 5 > len(data):
        raise ValueError("truncated table block")
    ctype = data[offset + size]
    if verify:
        want = struct.unpack("<I", data[offset + size + 1:offset + size + 5])[0]
        if _mask(crc32c(data[offset:offset + size + 1])) != want:
            raise ValueError("table block checksum mismatch at offset %d" % offset)
    if ctype == 0:
        return raw
    if ctype == 1:
        return _snappy_uncompress(raw)
    raise ValueError("unknown block compression %d" % ctype)


def _block_entries(block: bytes) -> List[Tuple[bytes, bytes]]:
    if len(block) < 4:
        raise ValueError("bad table block")
    num_restarts = struct.unpack("<I", block[-4:])[0]
    limit = len(block) - 4 - 4 * num_restarts
    out, i, key = [], 0, b""
    while i < limit:
        shared, i = _get_varint(block, i)
        non_shared, i = _get_varint(block, i)
        vlen, i = _get_varint(block, i)
        key = key[:shared] + block[i:i + non_shared]
        i += non_shared
        out.append((key, block[i:i + vlen]))
        i += vlen
    return out


def read_table(path: str, verify_checksums: bool = True) -> List[Tuple[bytes, bytes]]:
    data = open(path, "rb").read()
    if len(data) < 48 or struct.unpack("<Q", data[-8:])[0] != TABLE_MAGIC:
        raise ValueError("%s: not a TensorFlow V1 checkpoint (no SSTable footer)" % path)
    footer = data[-48:]
    _, i = _get_varint(footer, 0)          # metaindex handle (offset, size): unused
    _, i = _get_varint(footer, i)
    idx_off, i = _get_varint(footer, i)
    idx_size, i = _get_varint(footer, i)
    entries = []
    for _, handle in _block_entries(_read_block(data, idx_off, idx_size, verify_checksums)):
        off, j = _get_varint(handle, 0)
        size, j = _get_varint(handle, j)
        entries += _block_entries(_read_block(data, off, size, verify_checksums))
    return entries


# ---------------------------------------------------------------------------------------------- checkpoint reader
def _decode_shape(b) -> Tuple[int, ...]:
    dims = []
    for f, _, v in _fields(b):
        if f == 2:
            size = 0
            for f2, _, v2 in _fields(v):
                if f2 == 1:
                    size = _signed(v2)
            dims.append(size)
    return tuple(dims)


def _decode_tensor(b, shape: Tuple[int, ...], dtype: Optional[int]) -> np.ndarray:
    content, floats, doubles, ints = None, [], [], []
    for f, wt, v in _fields(b):
        if f == 1:
            dtype = v
        elif f == 2:
            s = _decode_shape(v)
            if s:
                shape = s
        elif f == 4:
            content = bytes(v)
        elif f == 5:
            floats.append(np.frombuffer(bytes(v), "<f4") if wt == 2 else np.frombuffer(bytes(v), "<f4"))
        elif f == 6:
            doubles.append(np.frombuffer(bytes(v), "<f8"))
        elif f in (7, 10):
            if wt == 2:
                j, vals = 0, []
                while j < len(v):
                    x, j = _get_varint(v, j)
                    vals.append(_signed(x))
                ints.append(np.array(vals, np.int64))
            else:
                ints.append(np.array([_signed(v)], np.int64))
    np_t = _NP_OF_DT.get(dtype)
    if np_t is None:
        raise ValueError("unsupported tensor dtype %s" % dtype)
    if content is not None:
        a = np.frombuffer(content, np.dtype(np_t).newbyteorder("<"))
    elif floats:
        a = np.concatenate(floats)
    elif doubles:
        a = np.concatenate(doubles)
    elif ints:
        a = np.concatenate(ints)
    else:
        a = np.zeros(0, np_t)
    n = int(np.prod(shape)) if shape else 1
    if a.size == 1 and n > 1:
        a = np.repeat(a, n)
    if a.size != n:
        raise ValueError("tensor holds %d values, shape %s needs %d" % (a.size, shape, n))
    return a.astype(np_t).reshape(shape)


def read_checkpoint(path: str, verify_checksums: bool = True) -> Dict[str, np.ndarray]:
    """name -> array for every tensor stored as ONE full slice (what tf.train.Saver writes for unpartitioned
    variables).  Partitioned variables (several slices per name) are rejected."""
    meta: Dict[str, Tuple[Tuple[int, ...], int]] = {}
    out: Dict[str, np.ndarray] = {}
    pending = []
    for key, value in read_table(path, verify_checksums):
        for f, _, v in _fields(value):
            if f == 1:                                   # SavedTensorSliceMeta
                for f2, _, v2 in _fields(v):
                    if f2 == 1:                          # SavedSliceMeta
                        name, shape, dtype, nslices = None, (), None, 0
                        for f3, _, v3 in _fields(v2):
                            if f3 == 1:
                                name = bytes(v3).decode()
                            elif f3 == 2:
                                shape = _decode_shape(v3)
                            elif f3 == 3:
                                dtype = v3
                            elif f3 == 4:
                                nslices += 1
                        if nslices > 1:
                            raise ValueError("tensor '%s' is stored in %d slices (partitioned variable): not supported" % (name, nslices))
                        meta[name] = (shape, dtype)
            elif f == 2:                                 # SavedSlice
                pending.append(v)
    for v in pending:
        name, tensor = None, None
        for f, _, v2 in _fields(v):
            if f == 1:
                name = bytes(v2).decode()
            elif f == 3:
                tensor = v2
        if name is None or tensor is None:
            raise ValueError("malformed SavedSlice entry")
        shape, dtype = meta.get(name, ((), None))
        out[name] = _decode_tensor(tensor, shape, dtype)
    missing = set(meta) - set(out)
    if missing:
        raise ValueError("checkpoint metadata lists tensors without data: %s" % sorted(missing)[:3])
    return out


# ---------------------------------------------------------------------------------------------- writer
class _BlockBuilder:
    def __init__(self):
        self.buf, self.restarts, self.count, self.last = bytearray(), [0], 0, b""

    def add(self, key: bytes, value: bytes):
        shared = 0
        if self.count < RESTART_INTERVAL:
            m = min(len(self.last), len(key))
            while shared < m and self.last[shared] == key[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.count = 0
        self.buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value)) + key[shared:] + value
        self.last = key
        self.count += 1

    def finish(self) -> bytes:
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))

    def empty(self) -> bool:
        return not self.buf


def write_table(path: str, entries: List[Tuple[bytes, bytes]]):
    """entries must be sorted by key (bytewise), as TableBuilder requires."""
    out = bytearray()
    index = _BlockBuilder()

    def emit(block: bytes) -> bytes:
        off = len(out)
        out.extend(block)
        out.append(0)                                              # kNoCompression, as tensor_slice_writer.cc sets
        out.extend(struct.pack("<I", _mask(crc32c(block + b"\x00"))))
        return _put_varint(off) + _put_varint(len(block))

    cur = _BlockBuilder()
    for key, value in entries:
        cur.add(key, value)
        if len(cur.buf) >= BLOCK_SIZE:
            index.add(key, emit(cur.finish()))
            cur = _BlockBuilder()
    if not cur.empty():
        index.add(cur.last, emit(cur.finish()))
    meta_handle = emit(_BlockBuilder().finish())
    index_handle = emit(index.finish())
    footer = meta_handle + index_handle
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    out.extend(footer)
    with open(path, "wb") as f:
        f.write(out)


def _shape_pb(shape) -> bytes:
    return b"".join(_pb(2, 2, _pb(1, 0, int(d))) for d in shape)


def write_checkpoint(path: str, tensors: Dict[str, np.ndarray]):
    """Writes `tensors` the way tf.train.Saver (TF <= 0.11) does: one full slice per tensor, float32 in float_val."""
    metas, entries = [], []
    for name in sorted(tensors):
        a = np.ascontiguousarray(tensors[name], "<f4")
        full_slice = b"".join(_pb(1, 2, b"") for _ in range(a.ndim))          # TensorSliceProto: one empty Extent per dim
        metas.append(_pb(1, 2, _pb(1, 2, name.encode()) + _pb(2, 2, _shape_pb(a.shape)) + _pb(3, 0, DT_FLOAT) + _pb(4, 2, full_slice)))
        tensor = _pb(1, 0, DT_FLOAT) + _pb(5, 2, a.tobytes())
        saved = _pb(1, 2, name.encode()) + _pb(2, 2, full_slice) + _pb(3, 2, tensor)
        entries.append((encode_tensor_name_slice(name, a.ndim), _pb(2, 2, saved)))
    versions = _pb(1, 0, 10)                                                 # VersionDef.producer (GraphDef producer 10 in model.meta)
    meta_value = _pb(1, 2, b"".join(metas) + _pb(2, 2, versions))
    entries.append((b"", meta_value))
    entries.sort(key=lambda kv: kv[0])
    write_table(path, entries)


# ---------------------------------------------------------------------------------------------- converter
def is_tf_checkpoint(path: str) -> bool:
    try:
        with open(path, "rb") as f:
            f.seek(-8, 2)
            return struct.unpack("<Q", f.read(8))[0] == TABLE_MAGIC
    except (OSError, struct.error):
        return False


def checkpoint_to_weights(path: str, num_classes: int, height: int = 299, width: int = 299) -> Dict[str, np.ndarray]:
    """The restore of PyClassification.py:166-172: every variable of the tower (tower_spec.variables) is looked up as
    `<name>/ExponentialMovingAverage` (variable_averages.variables_to_restore()), then as `<name>`; shapes are checked
    against the tower; missing tensors are an error (TF would silently keep the random initial value, :169-170)."""
    from . import tower_spec
    ck = read_checkpoint(path)
    out, missing = {}, []
    for name, shape in tower_spec.variables(num_classes, height, width):
        a = ck.get(name + EMA_SUFFIX)
        if a is None:
            a = ck.get(name)
        if a is None:
            missing.append(name)
            continue
        if tuple(a.shape) != tuple(shape):
            raise ValueError("checkpoint tensor '%s' has shape %s, the %d-class tower needs %s" % (name, a.shape, num_classes, tuple(shape)))
        out[name] = a.astype(np.float32)
    if missing:
        raise ValueError("checkpoint lacks %d of the tower's variables, e.g. %s" % (len(missing), missing[:3]))
    return out


def convert(ckpt_path: str, blob_path: str, num_classes: int) -> None:
    from . import weights
    weights.save_blob(blob_path, checkpoint_to_weights(ckpt_path, num_classes))


if __name__ == "__main__":
    import sys
    if len(sys.argv) != 4:
        raise SystemExit("usage: python -m tensorflow_web_deploy_b200.tf_checkpoint <model.ckpt> <out.icpw> <num_classes incl. background>")
    convert(sys.argv[1], sys.argv[2], int(sys.argv[3]))
