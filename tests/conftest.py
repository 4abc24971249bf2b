"""Shared pytest plumbing: the ``gpu`` marker, import paths, weight / golden loaders.

``-m "not gpu"`` runs here (no GPU): oracle vs golden vectors, host logic, C-ABI symbols.
``-m gpu`` runs on a B200: CUDA path vs oracle / golden vectors through the C-ABI.
Nothing here reads ``/root/reference`` (it does not exist on the GPU box).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "python-visual-similarity_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
WEIGHTS = os.path.join(ROOT, "python-visual-similarity_b200", "pyvisim_b200", "res", "model_files")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def load_weights(name):
    return dict(np.load(os.path.join(WEIGHTS, name + ".npz")))


def split(desc, offsets):
    return [desc[offsets[i]:offsets[i + 1]] for i in range(len(offsets) - 1)]


def rel_l2(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.fixture(scope="session")
def golden():
    return load_golden


@pytest.fixture(scope="session")
def weights():
    return load_weights
