"""Property tests (hypothesis) of the byte-level codecs on the host side: the TF checkpoint's varints / OrderedCode keys /
table round trip, snappy literals, and the micro-batcher's wire headers.  Pure CPU."""
import struct

import numpy as np
from hypothesis import given, settings, strategies as st

from tensorflow_web_deploy_b200 import tf_checkpoint as T
from tensorflow_web_deploy_b200.serving import batcher as B


@given(st.integers(min_value=0, max_value=(1 << 64) - 1))
def test_varint_round_trip(n):
    enc = T._put_varint(n)
    val, end = T._get_varint(enc + b"\xff", 0)
    assert (val, end) == (n, len(enc)) and len(enc) == max(1, (n.bit_length() + 6) // 7)


@given(st.lists(st.integers(min_value=-(1 << 62), max_value=(1 << 62)), min_size=2, max_size=20))
def test_ordered_code_signed_numbers_sort_like_integers(vals):
    enc = [T._oc_signed_num_increasing(v) for v in vals]
    assert [v for _, v in sorted(zip(enc, vals))] == sorted(vals)


@given(st.lists(st.binary(max_size=12), min_size=2, max_size=12))
def test_ordered_code_strings_sort_like_bytes_and_are_prefix_free(strs):
    enc = [T._oc_string(s) for s in strs]
    assert [s for _, s in sorted(zip(enc, strs))] == sorted(strs)
    for a in enc:
        for b in enc:
            assert a == b or not b.startswith(a)           # a terminator no payload byte sequence can contain


@settings(max_examples=25, deadline=None)
@given(st.dictionaries(st.text(alphabet="abcdefghij/_0123456789", min_size=1, max_size=24),
                       st.tuples(st.integers(1, 5), st.integers(1, 40), st.integers(0, 2 ** 31 - 1)), min_size=1, max_size=8))
def test_checkpoint_round_trip_of_arbitrary_tensors(tmp_path_factory, spec):
    tensors = {name: np.random.default_rng(seed).standard_normal((r, c)).astype(np.float32) for name, (r, c, seed) in spec.items()}
    p = str(tmp_path_factory.mktemp("ckpt") / "model.ckpt")
    T.write_checkpoint(p, tensors)
    got = T.read_checkpoint(p)
    assert set(got) == set(tensors)
    for k, v in tensors.items():
        np.testing.assert_array_equal(got[k], v)
    keys = [k for k, _ in T.read_table(p)]
    assert keys == sorted(keys)


@given(st.binary(min_size=1, max_size=200))
def test_snappy_literal_only_stream(data):
    n = len(data) - 1
    tag = bytes([n << 2]) if n < 60 else bytes([60 << 2, n])      # literal with a 1-byte length for 61..256 bytes
    assert T._snappy_uncompress(T._put_varint(len(data)) + tag + data) == data


@given(st.integers(0, 2 ** 32 - 1), st.integers(0, 2 ** 32 - 1), st.integers(0, 2 ** 63 - 1), st.integers(0, 2), st.integers(0, 4096),
       st.integers(0, 4096))
def test_request_header_round_trip(req_id, nbytes, off, kind, h, w):
    raw = B.REQ.pack(req_id, nbytes, off, kind, h, w)
    assert len(raw) == B.REQ.size == struct.calcsize("<IIQIII") and B.REQ.unpack(raw) == (req_id, nbytes, off, kind, h, w)
