"""TF V1 checkpoint reader / writer / converter (the restore of /root/reference/PyClassification.py:166-172).
No real checkpoint exists (reference README.md:14): files are produced by the writer, which follows the published
format (SSTable + SavedTensorSlices protos), and read back; the format's fixed points are asserted byte for byte."""
import struct

import numpy as np
import pytest

from tensorflow_web_deploy_b200 import tf_checkpoint as T
from tensorflow_web_deploy_b200 import weights


def test_crc32c_and_mask_known_answers(built_lib):
    assert T.crc32c(b"123456789") == 0xE3069283                     # the standard CRC-32C check value
    assert T.crc32c(b"") == 0
    assert T.crc32c(bytes(32)) == 0x8A9136AA                         # RFC 3720 B.4: 32 bytes of zeros
    assert T._mask(0) == 0xA282EAD8                                  # leveldb crc32c::Mask


def test_ordered_code_key_encoding():
    # OrderedCode: numbers are length-prefixed big-endian, strings end in \x00\x01, signed numbers XOR 0x80
    assert T._oc_num_increasing(0) == b"\x00" and T._oc_num_increasing(4) == b"\x01\x04" and T._oc_num_increasing(300) == b"\x02\x01\x2c"
    assert T._oc_string(b"a\x00b\xff") == b"a\x00\xffb\xff\x00\x00\x01"
    assert T._oc_signed_num_increasing(0) == b"\x80" and T._oc_signed_num_increasing(-1) == b"\x7f"
    assert T._oc_signed_num_increasing(63) == b"\xbf" and T._oc_signed_num_increasing(-64) == b"\x40"
    assert T._oc_signed_num_increasing(64) == b"\xc0\x40" and T._oc_signed_num_increasing(-65) == b"\x3f\xbf"
    vals = [-(1 << 40), -70000, -65, -64, -1, 0, 1, 63, 64, 8191, 8192, 1 << 33]
    enc = [T._oc_signed_num_increasing(v) for v in vals]
    assert enc == sorted(enc)                                        # the point of the code: bytewise order == numeric order
    key = T.encode_tensor_name_slice("conv0/weights", 4)
    assert key == b"\x00conv0/weights\x00\x01\x01\x04" + b"\x80\x7f" * 4
    assert b"" < key                                                 # the metadata entry (key "") sorts first


def test_table_layout_and_round_trip(tmp_path, built_lib):
    rng = np.random.default_rng(0)
    tensors = {"a/weights" + T.EMA_SUFFIX: rng.standard_normal((3, 3, 8, 16)).astype(np.float32),
               "a/BatchNorm/beta" + T.EMA_SUFFIX: rng.standard_normal(16).astype(np.float32),
               "global_step": np.float32([5000]),
               "big/weights": rng.standard_normal((64, 1000)).astype(np.float32)}       # spans many 4 KiB blocks
    p = str(tmp_path / "model.ckpt-5000")
    T.write_checkpoint(p, tensors)
    raw = open(p, "rb").read()
    assert struct.unpack("<Q", raw[-8:])[0] == 0xDB4775248B80FB57 and T.is_tf_checkpoint(p)
    entries = T.read_table(p)
    assert [k for k, _ in entries] == sorted(k for k, _ in entries) and entries[0][0] == b""
    got = T.read_checkpoint(p)
    assert set(got) == set(tensors)
    for k in tensors:
        np.testing.assert_array_equal(got[k], tensors[k])
        assert got[k].dtype == np.float32
    bad = bytearray(raw)
    bad[100] ^= 0x40                                                 # flip a bit inside the first data block
    (tmp_path / "corrupt").write_bytes(bytes(bad))
    with pytest.raises(ValueError, match="checksum"):
        T.read_checkpoint(str(tmp_path / "corrupt"))
    (tmp_path / "notckpt").write_bytes(b"ICPWv001" + bytes(100))
    assert not T.is_tf_checkpoint(str(tmp_path / "notckpt"))
    with pytest.raises(ValueError):
        T.read_checkpoint(str(tmp_path / "notckpt"))


def test_snappy_blocks_are_readable():
    # literal "abcd" + copy(offset 4, length 8) -> "abcdabcdabcd"
    comp = bytes([12]) + bytes([(4 - 1) << 2]) + b"abcd" + bytes([((8 - 4) << 2) | 1, 4])
    assert T._snappy_uncompress(comp) == b"abcdabcdabcd"


def test_checkpoint_to_blob_follows_the_ema_restore(tmp_path, built_lib):
    """variables_to_restore(): trainables and BN moving statistics under <name>/ExponentialMovingAverage; a variable
    without a shadow (here one BN beta) is taken under its own name; shape errors and missing tensors are loud."""
    W = weights.synthetic_weights(6)
    ck = {}
    for i, (k, v) in enumerate(W.items()):
        ck[k if k == "conv1/BatchNorm/beta" else k + T.EMA_SUFFIX] = v
    ck["global_step"] = np.float32([5000])
    p = str(tmp_path / "model")
    T.write_checkpoint(p, ck)
    back = T.checkpoint_to_weights(p, 6)
    assert set(back) == set(W)
    for k in W:
        np.testing.assert_array_equal(back[k], W[k])
    blob = str(tmp_path / "model.icpw")
    T.convert(p, blob, 6)
    rb = weights.load_blob(blob)
    assert all(np.array_equal(rb[k], W[k]) for k in W)
    with pytest.raises(ValueError, match="shape"):
        T.checkpoint_to_weights(p, 1001)                             # a 6-class checkpoint is not a 1001-class tower
    del ck["conv4/weights" + T.EMA_SUFFIX]
    T.write_checkpoint(p, ck)
    with pytest.raises(ValueError, match="lacks"):
        T.checkpoint_to_weights(p, 6)
