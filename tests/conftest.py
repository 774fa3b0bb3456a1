import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU tests fail (not skip) without a device when selected with -m gpu; with no
    selection on a CPU-only host they are skipped so a bare `pytest tests` is green."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu or "gpu" in (config.getoption("-m") or ""):
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def golden_cases():
    """Per-scene traces of the reference (trace_io format); track0_export.npz holds the dataset-builder blocks."""
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and f != "track0_export.npz")
