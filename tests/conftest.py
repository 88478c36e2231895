"""pytest configuration: registers the ``gpu`` marker and shared golden-fixture helpers."""
import glob
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` run on a machine without a GPU skips the `gpu` tests instead of failing them.  On a machine
    WITH a GPU nothing is skipped: a missing or stale libphet_b200.so must fail loudly there (no CPU fallback)."""
    if glob.glob("/dev/nvidia[0-9]*"):
        return
    skip = pytest.mark.skip(reason="no CUDA device on this machine (the gpu tests run on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden_names():
    return sorted(n for n in (os.path.splitext(os.path.basename(p))[0]
                              for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))) if not n.startswith(("cli_", "siblings_", "multi_", "big_")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    d = {k: z[k] for k in z.files}
    d["params"] = json.loads(str(d.pop("params_json")))
    if d["params"].get("percentiles") is not None:
        d["params"]["percentiles"] = tuple(d["params"]["percentiles"])
    d["dtypes"] = [t for t in ("f64", "f32") if f"scores_{t}" in d]
    d["name"] = name
    return d


@pytest.fixture(scope="session")
def goldens():
    return {n: load_golden(n) for n in golden_names()}
