"""CPU: the library's MatrixMarket reader (phet_mtx_parse, csrc/mtx.cu -- SURVEY.md 8(f) item 2) against
scipy.io.mmread(...).astype(float32), which is what sc.read_mtx does at /root/reference/src/train.py:82.
No GPU: only the host parser is called."""
import os

import numpy as np
import pytest
import scipy.io
import scipy.sparse as sp

pytestmark = pytest.mark.filterwarnings("ignore::DeprecationWarning")


def _dense_from(m):
    r, c, v, sym = m.triplets()
    X = np.zeros(m.shape, dtype=np.float32)
    np.add.at(X, (r, c), v)
    if sym:
        off = r != c
        np.add.at(X, (c[off], r[off]), v[off])
    return X


def _scipy_dense(path):
    A = scipy.io.mmread(path)
    A = A.astype(np.float32)
    return np.asarray(A.toarray() if hasattr(A, "toarray") else A)


@pytest.mark.parametrize("threads", [1, 3, 0])
@pytest.mark.parametrize("case", ["real", "real_p17", "integer", "pattern", "symmetric", "big"])
def test_parser_matches_scipy(tmp_path, case, threads):
    from phet_b200 import _capi
    rng = np.random.default_rng(3)
    path = str(tmp_path / "a.mtx")
    if case in ("real", "real_p17", "big"):
        N, G = (2000, 900) if case == "big" else (60, 45)
        X = rng.gamma(2.0, 3.0, (N, G)) * (rng.random((N, G)) < 0.15)
        X[1, 2] = 1.2345678901234567e-7
        X[2, 3] = 9.87654321e12
        X[3, 4] = -0.5
        scipy.io.mmwrite(path, sp.coo_matrix(X), precision=17 if case == "real_p17" else None)
    elif case == "integer":
        X = rng.poisson(0.4, (70, 33)).astype(np.int64)
        scipy.io.mmwrite(path, sp.coo_matrix(X), field="integer")
    elif case == "pattern":
        X = (rng.random((40, 50)) < 0.1).astype(np.float64)
        scipy.io.mmwrite(path, sp.coo_matrix(X), field="pattern")
    else:
        A = rng.normal(size=(30, 30)) * (rng.random((30, 30)) < 0.2)
        scipy.io.mmwrite(path, sp.coo_matrix(A + A.T), symmetry="symmetric")
    m = _capi.MtxFile(path, threads)
    ref = _scipy_dense(path)
    assert m.shape == ref.shape
    assert np.array_equal(_dense_from(m), ref)
    m.close()


def test_parser_layout_quirks(tmp_path):
    """Comments, blank lines, CRLF, tabs, explicit '+', exponents, no trailing newline, duplicates."""
    from phet_b200 import _capi
    path = str(tmp_path / "q.mtx")
    body = ("%%MatrixMarket MATRIX Coordinate Real General\r\n% a comment\n%another\n\n  3 4   6\n"
            "1 1 +1.5\n\n2\t3\t-2.5e-3\r\n3 4 1E2\n   1 2 .25\n3 4 0.5\n2 1 7")
    with open(path, "w", newline="") as fh:
        fh.write(body)
    m = _capi.MtxFile(path, 2)
    assert m.shape == (3, 4) and m.entries == 6
    want = np.zeros((3, 4), np.float32)
    want[0, 0] = 1.5
    want[1, 2] = np.float32(-2.5e-3)
    want[2, 3] = 100.5            # duplicate entries add up (coo -> csr)
    want[0, 1] = 0.25
    want[1, 0] = 7
    assert np.array_equal(_dense_from(m), want)
    with open(path, "w", newline="") as fh:
        fh.write(body.replace("+1.5", "1.5"))     # scipy's reader rejects an explicit '+'
    assert np.array_equal(_dense_from(_capi.MtxFile(path, 2)), _scipy_dense(path))


def test_parser_nonfinite_values(tmp_path):
    from phet_b200 import _capi
    path = str(tmp_path / "n.mtx")
    with open(path, "w") as fh:
        fh.write("%%MatrixMarket matrix coordinate real general\n2 2 3\n1 1 nan\n1 2 inf\n2 2 -inf\n")
    r, c, v, _ = _capi.MtxFile(path, 1).triplets()
    assert np.isnan(v[0]) and v[1] == np.inf and v[2] == -np.inf


@pytest.mark.parametrize("text, what", [
    ("hello\n1 1 1\n1 1 1.0\n", "banner"),
    ("%%MatrixMarket matrix array real general\n2 2\n1\n2\n3\n4\n", "unsupported"),
    ("%%MatrixMarket matrix coordinate complex general\n1 1 1\n1 1 1.0 0.0\n", "unsupported"),
    ("%%MatrixMarket matrix coordinate real general\n2 2 3\n1 1 1.0\n2 2 1.0\n", "declares 3"),
    ("%%MatrixMarket matrix coordinate real general\n2 2 1\n3 1 1.0\n", "malformed"),
    ("%%MatrixMarket matrix coordinate real general\n2 2 1\n1 1 abc\n", "malformed"),
    ("%%MatrixMarket matrix coordinate real general\n2 2 1\n1 1\n", "malformed"),
    ("%%MatrixMarket matrix coordinate real general\n", "size line"),
])
def test_parser_errors_are_loud(tmp_path, text, what):
    from phet_b200 import _capi
    path = str(tmp_path / "bad.mtx")
    with open(path, "w") as fh:
        fh.write(text)
    with pytest.raises(_capi.PhetError, match=what):
        _capi.MtxFile(path, 1)
    with pytest.raises(_capi.PhetError, match="cannot open"):
        _capi.MtxFile(str(tmp_path / "missing.mtx"), 1)


def test_numpy_row_sum_order_is_the_pairwise_recursion():
    """The device filter (k_row_abs_sums) restates NumPy's float32 pairwise summation; pin that restatement
    (leaves <= 128 with eight strided partial sums, halves split at multiples of eight) against np.sum itself
    so a NumPy whose reduction order changed shows up here and not as a silent filter difference."""
    f = np.float32

    def pw(a):
        n = len(a)
        if n < 8:
            r = f(0.0)
            for x in a:
                r = f(r + x)
            return r
        if n <= 128:
            r = [a[k] for k in range(8)]
            i = 8
            while i < n - (n % 8):
                for k in range(8):
                    r[k] = f(r[k] + a[i + k])
                i += 8
            res = f(f(f(r[0] + r[1]) + f(r[2] + r[3])) + f(f(r[4] + r[5]) + f(r[6] + r[7])))
            while i < n:
                res = f(res + a[i])
                i += 1
            return res
        n2 = n // 2
        n2 -= n2 % 8
        return f(pw(a[:n2]) + pw(a[n2:]))

    rng = np.random.default_rng(1)
    for G in (1, 7, 8, 9, 127, 128, 129, 257, 1000, 2308, 4097):
        X = np.abs(rng.gamma(2, 1, (3, G)).astype(f))
        assert np.array_equal(X.sum(1), np.array([pw(list(row)) for row in X], dtype=f)), G
