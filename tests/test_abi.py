"""The C-ABI library loads on a CPU-only box, exports every symbol include/inception_b200.h declares,
and fails loudly (no CPU fallback) when asked to compute without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "inception_b200.h")


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(icp_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(built_lib):
    from tensorflow_web_deploy_b200 import engine
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(built_lib, n), "library does not export %s" % n
    assert sorted(engine.SIGNATURES) == names, "engine.py binding and header disagree"
    assert b"sm_100a" in built_lib.icp_version()


def test_error_codes_without_gpu(built_lib, synth6, tmp_path):
    import torch
    from tensorflow_web_deploy_b200 import engine
    with pytest.raises(engine.EngineError) as ei:
        engine.Engine(str(tmp_path / "missing.icpw"), 6)
    assert ei.value.code == -2 and "cannot open" in str(ei.value)
    bad = tmp_path / "bad.icpw"
    bad.write_bytes(b"not a blob at all")
    with pytest.raises(engine.EngineError) as ei:
        engine.Engine(str(bad), 6)
    assert ei.value.code == -2
    with pytest.raises(engine.EngineError) as ei:
        engine.Engine(synth6[0], 6, precision="fp16", max_batch=0)
    assert ei.value.code == -1
    if not torch.cuda.is_available():
        with pytest.raises(engine.EngineError) as ei:       # no silent CPU path
            engine.Engine(synth6[0], 6)
        assert ei.value.code == -3 and "no CPU fallback" in str(ei.value)
        with pytest.raises(engine.EngineError) as ei:
            engine.conv2d_bn_relu(np.zeros((1, 4, 4, 8), np.float32), np.zeros((1, 1, 8, 16), np.float32),
                                  np.ones(16, np.float32), np.zeros(16, np.float32))
        assert ei.value.code == -3


def test_null_handle_is_rejected(built_lib):
    assert built_lib.icp_forward(None, None, 1, None) == -1
    assert built_lib.icp_last_forward_ms(None, None) == -1
    built_lib.icp_destroy(None)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tensorflow_web_deploy_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "inception_oracle" not in src, f
