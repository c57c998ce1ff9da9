"""pytest configuration: registers the `gpu` marker and shared fixtures.

`-m "not gpu"` runs here (no GPU): oracle vs golden vectors, host logic, C-ABI symbol check.
`-m gpu` runs on a B200: parity of the CUDA engine (through the C-ABI) against the oracle.
Nothing in here reads /root/reference at run time (the golden fixtures were generated from it
by oracle/make_golden.py and are committed under tests/golden/).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.cpu import Oracle, build
    build()
    return Oracle()


@pytest.fixture(scope="session")
def shipped():
    return dict(np.load(os.path.join(GOLDEN, "shipped_scene.npz")))


@pytest.fixture(scope="session")
def decomp2():
    return dict(np.load(os.path.join(GOLDEN, "decomp2.npz")))
