import json
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "wavelab_golden.json")) as f:
        return json.load(f)


def wavelet_class(key):
    """golden-vector key ('db4', 'sym8', 'haar', ...) -> WT class object."""
    import wavelets_b200 as wb
    single = {"haar": wb.WT.haar, "beyl": wb.WT.beyl, "vaid": wb.WT.vaid}
    if key in single:
        return single[key]
    m = re.match(r"([a-z]+)(\d+)$", key)
    ctor = {"db": wb.WT.Daubechies, "coif": wb.WT.Coiflet, "sym": wb.WT.Symlet, "batt": wb.WT.Battle}[m.group(1)]
    return ctor(int(m.group(2)))


def rng(seed=42):
    return np.random.default_rng(seed)


@pytest.fixture(autouse=True)
def _seed_torch():
    """every test starts from the same torch seed (CPU and CUDA generators): inputs drawn with torch.randn are the same on
    every run, so a result never depends on an unlucky draw (exact magnitude ties, for example)"""
    import torch
    torch.manual_seed(1234)
    yield
