"""CPU: selection + ranking (phet_b200/select.py) against the reference's own
significant_features / sort_features run on reference scores (tests/golden/make_golden_cli.py)."""
import os

import numpy as np

from conftest import GOLDEN_DIR


def test_selection_and_ranking_match_reference():
    from phet_b200.select import significant_features, sort_features, top_k
    z = np.load(os.path.join(GOLDEN_DIR, "cli_srbct.npz"))
    scores, names = z["scores"], z["names"].tolist()
    df = significant_features(X=scores, features_name=names, alpha=0.05, fit_type="gamma", per=95,
                              file_name="ignored", save_path="ignored")   # kwargs train.py:496-504 passes
    assert df["features"].to_list() == z["sig_features"].tolist()
    assert np.array_equal(df["score"].to_numpy(), z["sig_scores"])
    full = sort_features(X=scores, features_name=names)
    assert full["features"].to_list() == z["rank_features"].tolist()
    assert [names[i] for i in top_k(scores, 100)] == z["rank_features"].tolist()[:100]


def test_cli_filter_and_parser():
    from oracle.ingest_oracle import filter_data
    from phet_b200.main import build_parser
    args = build_parser().parse_args(["--file-name", "srbct", "--methods", "phet", "--iqr-range", "10", "90"])
    assert args.methods == ["phet"] and tuple(args.iqr_range) == (10.0, 90.0) and args.num_subsamples == 1000
    rng = np.random.default_rng(0)
    X = rng.poisson(2.0, (40, 30)).astype(np.float32)
    X[3] = 0                         # empty cell -> dropped
    X[:, 5] = 0                      # silent gene -> dropped
    y = np.array([0] * 20 + [1] * 20)
    Xf, yf, nf = filter_data(X, y, ["g%d" % i for i in range(30)])
    assert Xf.shape[0] == 39 and len(yf) == 39 and "g5" not in nf and Xf.shape[1] == len(nf)


def test_device_order_argument_reproduces_the_sorted_frames():
    """``order=`` (what phet_rank returns: descending, ties in gene order) yields the frames the pandas sort yields."""
    from phet_b200.select import significant_features, sort_features
    z = np.load(os.path.join(GOLDEN_DIR, "cli_srbct.npz"))
    scores, names = z["scores"], z["names"].tolist()
    order = np.argsort(-scores[:, 0], kind="stable").astype(np.int32)
    full = sort_features(X=scores, features_name=names, order=order)
    assert full["features"].to_list() == z["rank_features"].tolist()
    sig = significant_features(X=scores, features_name=names, alpha=0.05, fit_type="gamma", per=95, order=order)
    assert sig["features"].to_list() == z["sig_features"].tolist()
    assert np.array_equal(sig["score"].to_numpy(), z["sig_scores"])
    asc = sort_features(X=scores, features_name=names, order=order, ascending=True)
    assert asc["features"].to_list() == z["rank_features"].tolist()[::-1]
