"""pytest configuration: `gpu` marker + shared fixtures.

`-m "not gpu"` runs everywhere (oracle vs golden vectors, host logic, C-ABI symbol export,
gloo world_size-2 sharding).  `-m gpu` needs a B200 and calls the CUDA path through the C-ABI.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


try:        # property tests: same examples on every run, no example database written into the tree
    from hypothesis import settings as _hyp_settings
    _hyp_settings.register_profile("repo", derandomize=True, database=None, deadline=None)
    _hyp_settings.load_profile("repo")
except ImportError:  # pragma: no cover
    pass


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference mounted (build container only)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    skip_gpu = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords and not has_gpu:
            item.add_marker(skip_gpu)


@pytest.fixture(scope="session", autouse=True)
def _fresh_library():
    """Rebuild librenv_b200.so when it is older than its sources (nvcc cross-compiles without a GPU).  Only a
    convenience for the test run: the product path never builds or falls back, it raises when the library is missing."""
    import shutil
    from random_envs_b200 import build as lib_build
    if lib_build.is_stale() and (shutil.which("nvcc") or os.path.isfile("/usr/local/cuda/bin/nvcc")):
        lib_build.build()
    yield


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def golden_traj():
    return dict(np.load(os.path.join(GOLDEN, "cartpole_traj.npz")))
