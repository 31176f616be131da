import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built_lib():
    """Builds (if stale) and loads the C-ABI library; works without a GPU."""
    import __graft_entry__ as g
    g.build()
    from tensorflow_web_deploy_b200 import engine
    return engine.load_library()


@pytest.fixture(scope="session")
def synth6(tmp_path_factory):
    """(path, dict) of the seeded synthetic C=6 weights."""
    from tensorflow_web_deploy_b200 import weights
    W = weights.synthetic_weights(6)
    path = str(tmp_path_factory.mktemp("w") / "synthetic_c6.icpw")
    weights.save_blob(path, W)
    return path, W
