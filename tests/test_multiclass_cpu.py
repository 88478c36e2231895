"""CPU: multi-class branch.  The oracle (oracle/multiclass_oracle.py) is pinned to the
UNMODIFIED reference by tests/golden/make_golden_multiclass.py (max|oracle - reference| == 0 on five cases); here
the committed fixtures are re-derived from the oracle, and the product's host draws (phet_b200/multiclass.py) are
checked against the index sets and permutations the reference consumed."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR

CASES = sorted(os.path.basename(p)[len("multi_"):-len(".npz")] for p in glob.glob(os.path.join(GOLDEN_DIR, "multi_*.npz")))


def test_fixtures_exist():
    assert len(CASES) >= 5


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_scores(name):
    from oracle import multiclass_oracle as mo
    z = np.load(os.path.join(GOLDEN_DIR, "multi_%s.npz" % name))
    np.random.seed(12345)
    o = mo.fit_predict_multiclass_oracle(z["X"], z["y"].astype(np.int64), percentiles=tuple(z["percentiles"]),
                                         num_subsamples=int(z["S"]), strategy=str(z["strategy"]),
                                         aggregation=str(z["aggregation"]), group_test=str(z["group_test"]))
    assert np.array_equal(o["scores"], z["scores_ref"])
    assert np.array_equal(o["bins"], z["bins"]) and np.array_equal(o["H"], z["H"])
    # the formula restatement the device follows agrees with scipy's f_oneway
    np.random.seed(12345)
    p = mo.fit_predict_multiclass_oracle(z["X"], z["y"].astype(np.int64), percentiles=tuple(z["percentiles"]),
                                         num_subsamples=int(z["S"]), strategy=str(z["strategy"]),
                                         aggregation=str(z["aggregation"]), library=False, group_test=str(z["group_test"]))
    tol = 1e-4 if z["X"].dtype == np.float32 else 1e-11
    assert np.allclose(p["P"], o["P"], rtol=tol, atol=1e-300)


@pytest.mark.parametrize("name", CASES)
def test_host_draws_replay_the_reference_stream(name):
    from phet_b200 import PHet
    from phet_b200.multiclass import plan_and_draw
    z = np.load(os.path.join(GOLDEN_DIR, "multi_%s.npz" % name))
    y = z["y"].astype(np.int64)
    K = len(np.unique(y))
    est = PHet(num_subsamples=int(z["S"]), percentiles=tuple(z["percentiles"]),
               multiconditions_strategy=str(z["strategy"]), multiconditions_aggregation=str(z["aggregation"]),
               group_test=str(z["group_test"]))
    np.random.seed(12345)
    plan = plan_and_draw(est, y, K, z["X"].shape[1])
    assert plan["m"] == int(z["m"]) and plan["min_size"] == int(z["min_size"])
    assert plan["n_last"] == z["n_last"].tolist()
    assert np.array_equal(plan["idx_anova"], z["idx_anova"])
    for c in range(len(plan["combos"])):
        assert np.array_equal(plan["idx_combo"][c][:, 0], z["idx_i_%d" % c])
        assert np.array_equal(plan["idx_combo"][c][:, 1], z["idx_j_%d" % c])
        assert np.array_equal(plan["perms"][c][0], z["perm_i_%d" % c])
        assert np.array_equal(plan["perms"][c][1], z["perm_j_%d" % c])
