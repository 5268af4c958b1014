"""pytest configuration: the `gpu` marker (tests that need a B200) and import paths.

`-m "not gpu"` : oracle vs golden vectors, host logic, C-ABI symbol checks, gloo sharding logic - runs on CPU.
`-m gpu`       : parity tests proper - the CUDA path through the C ABI against the oracle / goldens.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (sm_100a); run on the GPU box with -m gpu")


def pytest_collection_modifyitems(config, items):
    # GPU tests must never silently pass on a CPU box: they are skipped loudly when no device is visible.
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def cuda_lib():
    """The built library; building is part of the fixture so a stale .so can never be tested."""
    from cmh_b200 import _cabi
    _cabi.build()
    return _cabi.lib()
