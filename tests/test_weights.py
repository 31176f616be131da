import numpy as np
import pytest

from tensorflow_web_deploy_b200 import tower_spec, weights


def test_blob_roundtrip(tmp_path):
    t = {"a/weights": np.arange(24, dtype=np.float32).reshape(1, 2, 3, 4), "b": np.float32([1.5, -2.0])}
    p = str(tmp_path / "t.icpw")
    weights.save_blob(p, t)
    r = weights.load_blob(p)
    assert list(r) == list(t)
    for k in t:
        np.testing.assert_array_equal(r[k], t[k])
        assert r[k].dtype == np.float32
    with pytest.raises(ValueError):
        (tmp_path / "x").write_bytes(b"12345678abcd")
        weights.load_blob(str(tmp_path / "x"))


def test_synthetic_weights_are_seeded_and_complete():
    a, b = weights.synthetic_weights(6), weights.synthetic_weights(6)
    spec = tower_spec.variables(6)
    assert list(a) == [n for n, _ in spec]
    for n, shp in spec:
        assert a[n].shape == tuple(shp)
        np.testing.assert_array_equal(a[n], b[n])
    assert (a["conv1/BatchNorm/moving_variance"] > 0).all()
    assert not np.array_equal(a["conv1/weights"], weights.synthetic_weights(6, seed=7)["conv1/weights"])
