import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """Lazy loader of tests/golden/<name>.npz; complex arrays are stored as [..., 2] float32."""

    def __init__(self):
        self._cache = {}

    def __call__(self, name):
        if name not in self._cache:
            self._cache[name] = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
        return self._cache[name]

    @staticmethod
    def weights(d, prefix):
        return {k[len(prefix):]: torch.from_numpy(v) for k, v in d.items() if k.startswith(prefix)}


@pytest.fixture(scope="session")
def golden():
    return Golden()


def rel_l2(a, b):
    a = torch.as_tensor(a).detach().cpu()
    b = torch.as_tensor(b).detach().cpu()
    if a.is_complex():
        a = torch.view_as_real(a)
    if b.is_complex():
        b = torch.view_as_real(b)
    a, b = a.double(), b.double()
    den = b.norm().item()
    return (a - b).norm().item() / (den if den > 0 else 1.0)


NORMS = ["backward", "ortho", "forward", "none"]
MASK_DTYPES = {0: torch.uint8, 1: torch.float32, 2: torch.bool}


def mask_from_golden(arr, code):
    t = torch.from_numpy(arr)
    return t.to(MASK_DTYPES[int(code)])
