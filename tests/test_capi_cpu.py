"""CPU: the C-ABI library builds, loads and exports every symbol include/phet_b200.h declares;
host-evaluated device routines (same source) agree with scipy; the host class validates its
arguments like the reference.  No compute calls that need a GPU."""
import math
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from phet_b200 import _capi
    lib = _capi.load_library()
    header = open(os.path.join(ROOT, "include", "phet_b200.h")).read()
    declared = set(re.findall(r"\b(phet_[a-z0-9_]+)\s*\(", header))
    declared -= {"phet_ctx"}
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), "library does not export %s" % name
    assert set(_capi.EXPORTS) == declared
    assert lib.phet_abi_version() == 1


def test_no_gpu_is_a_loud_error():
    from phet_b200 import _capi
    lib = _capi.load_library()
    if lib.phet_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_capi.PhetError):
        _capi.Context(0)
    from phet_b200 import PHet
    with pytest.raises(_capi.PhetError):
        PHet(num_subsamples=2).fit_predict(np.random.default_rng(0).normal(size=(20, 4)), [0] * 10 + [1] * 10)


def test_device_pvalue_routines_on_host_match_scipy():
    from scipy import special, stats
    from phet_b200 import _capi
    from phet_b200.hostmath import ln_beta_half
    lib = _capi.load_library()
    for df in (2, 4, 10, 28, 56):
        lb = ln_beta_half(df)
        for t in np.concatenate([np.linspace(0, 60, 301), 10.0 ** np.linspace(-8, 3, 60)]):
            p = lib.phet_host_p_two_sided_t(float(t), float(df), lb)
            r = 2 * special.stdtr(df, -abs(t))
            assert abs(p - r) <= 1e-12 * r + 1e-300
    assert math.isnan(lib.phet_host_p_two_sided_t(float("nan"), 56.0, ln_beta_half(56)))
    for z in np.concatenate([np.linspace(0, 37.6, 500), [37.67, 37.68, 38.0, 1e3]]):
        p = lib.phet_host_p_two_sided_z(float(z))
        r = 2 * stats.norm.sf(z)
        assert (r == 0 and p == 0) or abs(p - r) <= 1e-12 * r
    # cephes erfc underflow rule: p == 0 iff z^2/2 > 709.782712893384
    zc = math.sqrt(2 * 709.782712893384)
    for z in (zc * (1 - 1e-6), zc * (1 + 1e-6)):
        assert (lib.phet_host_p_two_sided_z(z) == 0.0) == (2 * stats.norm.sf(z) == 0.0)
    assert lib.phet_host_p_two_sided_z(zc * (1 - 1e-6)) > 0 and lib.phet_host_p_two_sided_z(zc * (1 + 1e-6)) == 0.0


def test_device_permutation_replica_is_a_permutation_and_matches_c():
    from phet_b200 import _capi
    from phet_b200.perm import device_positions
    lib = _capi.load_library()
    for l in (1, 2, 3, 29, 54, 252, 5000, 25000):
        pm = device_positions(12345, 7, 1, l)
        assert sorted(pm.tolist()) == list(range(l))
        dv = [lib.phet_host_perm_apply(12345, 7, 1, l, i) for i in range(min(l, 100))]
        assert dv == pm[:len(dv)].tolist()
    # different genes walk the class differently
    assert not np.array_equal(device_positions(1, 0, 0, 5000, 64), device_positions(1, 1, 0, 5000, 64))


def test_constructor_mirrors_reference():
    from phet_b200 import PHet
    e = PHet()
    assert e.normalize == "zscore" and e.iqr_range == (25, 75) and e.num_subsamples == 1000
    assert np.allclose(e.feature_weight, [0.4, 0.3, 0.2, 0.1])
    assert PHet(iqr_range=(10, 90)).iqr_range == (10, 90)          # the kwarg src/train.py:474 passes
    assert np.isclose(PHet(feature_weight=[2, 1, 1]).feature_weight.sum(), 1.0)
    with pytest.raises(ValueError):
        PHet(feature_weight=[1.0])
    with pytest.raises(Exception, match="valid groups"):
        PHet().fit_predict(np.zeros((4, 3)), np.zeros(4))
    from phet_b200._capi import PhetError
    with pytest.raises(PhetError):                 # multi-class is built, but there is no CPU fallback
        PHet().fit_predict(np.zeros((24, 3)), np.repeat([0, 1, 2], 8))
    with pytest.raises(NotImplementedError):
        PHet(delta_type="hvf").fit_predict(np.zeros((4, 3)), [0, 0, 1, 1])


def test_host_math_matches_oracle():
    from oracle import phet_oracle as po
    from phet_b200 import hostmath
    for n in (1, 2, 8, 29, 45, 70):
        assert np.array_equal(hostmath.ks_exact_table(n), po.ks_table(n))
    for m in (1, 2, 14, 29, 45, 70, 158, 316):
        for pct in ((25, 75), (10, 90), (0, 100), (75, 25)):
            a = hostmath.quantile_positions(m, pct)
            b = po.quantile_positions(m, pct)
            assert a == tuple((x[0], x[1], float(x[2])) for x in b)
    y = np.array([0] * 5000 + [1] * 5000)
    assert hostmath.subsample_size(5000, 5000, "sqrt") == po.subsample_size(y, "sqrt") == 70
    assert hostmath.subsample_size(54, 29, "sqrt") == 29
    assert hostmath.subsample_size(350, 250, "optimum") == po.subsample_size(np.array([0] * 350 + [1] * 250), "optimum")
    assert hostmath.step3_slices(1000, 31) == (33, 8)
