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
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and f not in ("stages.npz", "exciter_tables.npz"))


def authoring_exciter_table(saturation):
    """The exciter table (ENG:128-134 over the 65536 int16 samples) of the host that WROTE the golden fixtures,
    rebuilt from ``exciter_tables.npz`` (oracle/make_golden.py: ulp differences of its float32 tanh from the
    correctly rounded one) and verified against the stored SHA-256."""
    import hashlib
    z = np.load(os.path.join(GOLDEN_DIR, "exciter_tables.npz"))
    tag = repr(float(saturation))
    s16 = np.arange(65536, dtype=np.uint16).view(np.int16)
    x = s16.astype(np.float32) / (2 ** 15)
    mix = (saturation / 100.0) ** 2
    arg = x * (1 + mix * 4)
    t0 = np.tanh(arg.astype(np.float64)).astype(np.float32)
    t32 = (t0.view(np.int32).astype(np.int64) + z["d_" + tag].astype(np.int64)).astype(np.int32).view(np.float32)
    table = np.ascontiguousarray((1 - mix) * x + mix * t32, dtype=np.float32)
    assert hashlib.sha256(table.tobytes()).hexdigest() == str(z["sha_" + tag]), \
        "could not rebuild the authoring host's exciter table (float64 tanh of this libm differs?)"
    return table


class golden_exciter:
    """Context manager: plans made inside use the authoring host's exciter table for every saturation value in
    ``settings`` (a dict or a list of dicts), so a fixture is compared like with like on any CPU."""

    def __init__(self, settings):
        sets = [settings] if isinstance(settings, dict) else list(settings)
        self.sats = sorted({s.get("saturation", 0) for s in sets} - {0})

    def __enter__(self):
        from b200master import plan
        for s in self.sats:
            plan.install_exciter_table(s, authoring_exciter_table(s))
        return self

    def __exit__(self, *exc):
        from b200master import plan
        for s in self.sats:
            plan.install_exciter_table(s, None)
        return False


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
