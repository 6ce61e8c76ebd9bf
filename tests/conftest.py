"""pytest configuration: the `gpu` marker and shared fixtures."""

import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: CPU test that takes more than ~20 s")


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
def golden_cases():
    with open(os.path.join(GOLDEN_DIR, "cases.json")) as f:
        data = json.load(f)
    return {r["name"]: r for r in data["cases"]}


@pytest.fixture(scope="session")
def golden_sweep():
    with open(os.path.join(GOLDEN_DIR, "sweep.json")) as f:
        return json.load(f)["sweep"]


def parse_float(v):
    """cases.json stores inf/nan as their repr strings."""
    return float(v) if isinstance(v, str) else v
