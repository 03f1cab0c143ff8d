import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (runs on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """Without a CUDA device the `gpu` tests are SKIPPED (the product has no CPU fallback, so they could
    only fail with ClvError): a plain `pytest tests` on a CPU box then gates on the CPU suite alone."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device: the CL-VAE/CL-VRNN hot path has no CPU fallback")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
