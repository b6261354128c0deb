"""pytest configuration: path setup, the ``gpu`` marker, golden-fixture loader."""
import json
import os
import sys
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "python-audio-mastering_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

warnings.filterwarnings("ignore", category=DeprecationWarning, message=".*audioop.*")

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and f != "stages.npz")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    d = {k: z[k] for k in z.files}
    d["rate"] = int(d["rate"])
    d["settings"] = json.loads(str(d["settings"]))
    if "loudness" in d:
        d["loudness"] = float(d["loudness"])
    return d


@pytest.fixture(scope="session")
def has_cuda():
    import torch
    return torch.cuda.is_available()
