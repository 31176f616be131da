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


def calibrated_weights(num_classes, seed=4321, images=2):
    """Synthetic weights with TRAINED-LIKE BatchNorm statistics (test helper; uses the oracle, so it lives in tests/).

    The default synthetic blob draws moving_mean ~ N(0, 0.1) and variance in [0.5, 1.5] — tame.  A trained net's
    statistics are whatever the conv outputs really are: a mean of several sigma and variances spread over orders
    of magnitude.  Here every conv's output channels are rescaled by 10^U(-1,1) and moving_mean / moving_variance
    are then SET to the measured per-channel statistics of the oracle's activations on a few random images, layer by
    layer, so |mean| / sigma is O(1..5) and the BN fold has real work to do (the pool-branch reorder)."""
    import numpy as np
    import torch
    from oracle import inception_oracle as O
    from tensorflow_web_deploy_b200 import weights
    rng = np.random.default_rng(seed)
    W = weights.synthetic_weights(num_classes, seed=seed)
    for name in list(W):
        if name.endswith("/weights") and W[name].ndim == 4:
            W[name] = W[name] * (10.0 ** rng.uniform(-1, 1, W[name].shape[3])).astype(np.float32)

    class Calibrating(O.Tower):
        def bn_relu(self, y, scope):
            c = y.shape[1]
            self.w[scope + "/BatchNorm/moving_mean"] = y.mean(dim=(0, 2, 3)).numpy().astype(np.float32)
            self.w[scope + "/BatchNorm/moving_variance"] = y.var(dim=(0, 2, 3), unbiased=False).numpy().astype(np.float32) + np.float32(1e-4)
            return super().bn_relu(y, scope)

    x = rng.uniform(-1, 1, (images, 299, 299, 3)).astype(np.float32)
    Calibrating(W).forward(x, num_classes, with_aux=True)
    return W
