"""pytest configuration: registers the ``gpu`` marker and shared fixtures."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


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


@pytest.fixture(scope="session")
def golden():
    return load_golden


@pytest.fixture(scope="session")
def rf50mm_state_dict():
    return torch.load(os.path.join(GOLDEN, "rf50mm_PSFNet480x640_ks11.pkl"), map_location="cpu")


@pytest.fixture(scope="session")
def rf50mm_weights(rf50mm_state_dict):
    from oracle.focal_stack_oracle import split_state_dict
    return split_state_dict(rf50mm_state_dict)


def analytic_rgbd(N, H, W):
    """KAT-B inputs, RNG-free."""
    n, c, h, w = np.meshgrid(np.arange(N), np.arange(3), np.arange(H), np.arange(W), indexing="ij")
    img = ((7 * h + 13 * w + 29 * c + 101 * n) % 256) / 255.0
    n, h, w = np.meshgrid(np.arange(N), np.arange(H), np.arange(W), indexing="ij")
    depth_m = 0.5 + 4.5 * (((h * W + w) * 31 + 17 * n) % 1000) / 999.0
    return (torch.tensor(img, dtype=torch.float32),
            torch.tensor(depth_m, dtype=torch.float32).unsqueeze(1))
