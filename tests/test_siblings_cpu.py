"""CPU: the sibling-scorer oracle (oracle/siblings_oracle.py) against the committed outputs of the
unmodified reference classes (tests/golden/siblings_*.npz, made by make_golden_siblings.py)."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR

CASES = sorted(os.path.basename(p)[len("siblings_"):-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "siblings_*.npz")))


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, "siblings_%s.npz" % name))
    norm = str(z["normalize"])
    return z, (None if norm == "none" else norm), tuple(int(v) for v in z["iqr_range"])


def test_cases_present():
    assert len(CASES) >= 6


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference(name):
    from oracle import siblings_oracle as so
    z, norm, rng_ = load(name)
    with np.errstate(invalid="ignore", divide="ignore"):
        for pc in (False, True):
            got = so.hiqr(z["X"].copy(), z["y"], per_condition=pc, normalize=norm, iqr_range=rng_)
            assert np.array_equal(got, z["hiqr_pc%d" % pc])
        for dm in (False, True):
            got = so.delta_iqr_mean(z["X"].copy(), z["y"], calculate_deltamean=dm, normalize=norm, iqr_range=rng_)
            assert np.array_equal(got, z["dim_dm%d" % dm])


def test_host_classes_mirror_reference_signatures():
    import inspect
    from phet_b200 import HIQR, DeltaIQRMean
    assert list(inspect.signature(HIQR.__init__).parameters)[1:4] == ["per_condition", "normalize", "iqr_range"]
    assert list(inspect.signature(DeltaIQRMean.__init__).parameters)[1:4] == ["calculate_deltamean", "normalize", "iqr_range"]
    for cls in (HIQR, DeltaIQRMean):
        assert list(inspect.signature(cls.fit_predict).parameters)[1:5] == ["X", "y", "control_class", "case_class"]
    assert DeltaIQRMean().normalize is None and HIQR().normalize == "zscore"   # reference defaults
