"""Shared fixtures.  `-m "not gpu"` runs here (no GPU): oracle vs golden vectors,
host logic, C-ABI symbol checks, gloo sharding.  `-m gpu` runs on a B200: the
parity tests proper, through the C-ABI."""
import gzip
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "high-order-entropy-compressed-suffix-array_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)
os.environ.setdefault("HKCSA_QUIET", "1")      # drop-in modules: no import-time demo prints


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


class Golden:
    def __init__(self):
        g = os.path.join(ROOT, "tests", "golden")
        self.arr = np.load(os.path.join(g, "golden_ref.npz"))
        with gzip.open(os.path.join(g, "golden_ref.json.gz"), "rt") as f:
            self.meta = json.load(f)

    def cases(self):
        return [k for k in self.meta if not k.startswith("_")]

    def text(self, name) -> bytes:
        return self.arr[f"{name}/text"].tobytes()

    def get(self, key):
        return self.arr[key]

    def has(self, key):
        return key in self.arr.files


_golden = None


@pytest.fixture(scope="session")
def golden():
    global _golden
    if _golden is None:
        _golden = Golden()
    return _golden


def golden_case_names():
    with gzip.open(os.path.join(ROOT, "tests", "golden", "golden_ref.json.gz"), "rt") as f:
        return [k for k in json.load(f) if not k.startswith("_")]


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)
