"""pytest configuration: the `gpu` marker and import paths.

`-m "not gpu"` : oracle vs golden vectors, host-side logic, C-ABI symbols, world_size-2 gloo tests (CPU only).
`-m gpu`       : parity of the CUDA path (through the C ABI) against the oracle and the golden fixtures.
Only tests may import `oracle/` (test infrastructure); the product package never does.
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle", "pyg_shim")):
    if p not in sys.path:
        sys.path.insert(0, p)

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


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


def t(a, **kw):
    return torch.from_numpy(np.ascontiguousarray(a)).to(**kw)


@pytest.fixture(scope="session")
def golden():
    return load_golden


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max(1, max|b|) -- the relative measure used for every tolerance in these tests."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    if a.numel() == 0:
        return 0.0
    return float((a - b).abs().max() / max(1.0, float(b.abs().max())))


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b||_2 / ||b||_2 -- the second bound on every bf16 comparison: the max-norm measure above tolerates a large
    relative error on near-zero elements, this one does not let a systematic error hide behind one large element."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    if a.numel() == 0:
        return 0.0
    return float((a - b).norm() / max(float(b.norm()), 1e-30))
