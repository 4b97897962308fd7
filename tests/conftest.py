"""pytest configuration: path setup and the ``gpu`` marker.

``-m "not gpu"`` runs on CPU (oracle vs golden vectors, host logic, C-ABI symbol
checks, gloo world_size-2 sharding); ``-m gpu`` runs the parity tests proper through
the C-ABI on a B200.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "torch-geometric-pool_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import torch

    path = os.path.join(ROOT, "tests", "golden", "ref_vectors.pt")
    return torch.load(path, weights_only=False)
