"""Golden vectors for the CLI path (main.py --methods phet_br on SRBCT) and for the selection /
ranking functions.  Run in the build container only:

    python tests/golden/make_golden_cli.py

Runs the UNMODIFIED reference code: ``PHet.fit_predict`` (via oracle/ref_shim.py) on the float32
matrix exactly as src/train.py:82-118 prepares it, then the reference's own
``significant_features`` / ``sort_features`` -- their source is extracted with ``ast`` from
/root/reference/src/utility/utils.py (the module itself cannot be imported here: it pulls in
umap/seaborn) and executed verbatim.  Output: tests/golden/cli_srbct.npz.
"""
import ast
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.ref_shim import load_reference_phet  # noqa: E402


def reference_utils_functions():
    src = open("/root/reference/src/utility/utils.py").read()
    tree = ast.parse(src)
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("significant_features", "sort_features")]
    mod = ast.Module(body=wanted, type_ignores=[])
    ns = {}
    exec("import numpy as np\nimport pandas as pd\nfrom typing import Literal\n"
         "from scipy.stats import zscore, gamma, expon, lognorm, scoreatpercentile\n", ns)
    exec(compile(mod, "utils.py", "exec"), ns)
    return ns["significant_features"], ns["sort_features"]


def main():
    from scipy.io import mmread
    import pandas as pd
    PHet = load_reference_phet()
    significant_features, sort_features = reference_utils_functions()
    X = np.asarray(mmread("/root/reference/samples/srbct_matrix.mtx").toarray(), dtype=np.float32)
    np.nan_to_num(X, copy=False, nan=0.0, posinf=0.0, neginf=0.0)
    y = pd.read_csv("/root/reference/samples/srbct_classes.csv").iloc[:, 0].to_numpy().astype(np.int64)
    names = pd.read_csv("/root/reference/samples/srbct_feature_names.csv").iloc[:, 0].to_list()
    # train.py:88-118 filter (drops nothing on SRBCT, kept for fidelity)
    sys.path.insert(0, ROOT)
    from oracle.ingest_oracle import filter_data
    Xf, yf, namesf = filter_data(X, y, names)
    assert Xf.shape == X.shape
    S = 200
    np.random.seed(12345)
    scores = PHet(normalize="zscore", percentiles=(25, 75), num_subsamples=S, subsampling_size="sqrt",
                  delta_type="iqr", adjust_pvalue=False, feature_weight=[0.4, 0.3, 0.2, 0.1]).fit_predict(X=Xf, y=yf)
    df_sig = significant_features(X=scores, features_name=namesf, alpha=0.05, fit_type="gamma", per=95,
                                  X_map=None, map_genes=False, ttest=False)
    df_all = sort_features(X=scores, features_name=namesf, X_map=None, map_genes=False, ttest=False)
    out = dict(scores=scores, S=np.int64(S), sig_features=np.array(df_sig["features"].to_list()),
               sig_scores=df_sig["score"].to_numpy(), rank_features=np.array(df_all["features"].to_list()),
               names=np.array(namesf), y=yf.astype(np.int8))
    np.savez_compressed(os.path.join(HERE, "cli_srbct.npz"), **out)
    print("selected", len(df_sig), "top5", df_sig["features"].head(5).to_list())


if __name__ == "__main__":
    main()
