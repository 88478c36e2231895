"""Import shim for the UNMODIFIED reference ``model.phet.PHet`` (test infrastructure).

TEST INFRASTRUCTURE ONLY.  Nothing under ``phet_b200/`` may import this module.  It is used
by ``tests/golden/make_golden.py`` (run in the build container, where ``/root/reference``
is mounted) to produce the committed golden vectors, and by CPU tests that are skipped when
the reference tree is absent (it does not exist on the GPU box).

Why a shim is needed: ``/root/reference/src/model/phet.py:16-25`` imports
``anndata``, ``decoupler``, ``scanpy`` and ``statsmodels`` at module top.  None of them is
installed in this image.  The first three are touched only on the pseudobulk / hvf branches
(``phet.py:243-250,404-409``), which are out of scope, so empty stub modules suffice.
``statsmodels.stats.weightstats.ztest`` (``phet.py:25,428``) is on the path for m >= 30; it is
restated below from the published statsmodels 0.14 algorithm (``_zstat_generic`` with
``usevar='pooled'``, ``ddof=1``, ``value=0``, two-sided).  statsmodels' source is not in this
container, so that one function is "parity unpinned" by anything but algebra: it is the pooled
two-sample statistic with a normal tail.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
from scipy import stats as _stats

_VENDORED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")   # oracle/make_ref.sh (git-ignored copy)
REFERENCE_SRC = os.environ.get("PHET_REFERENCE_SRC", "/root/reference/src")
if not os.path.isfile(os.path.join(REFERENCE_SRC, "model", "phet.py")) and os.path.isfile(os.path.join(_VENDORED, "model", "phet.py")):
    REFERENCE_SRC = _VENDORED   # the GPU box: the unmodified file as vendored by the build step


def ztest_restated(x1, x2=None, value=0, alternative="two-sided", usevar="pooled", ddof=1.0):
    """statsmodels.stats.weightstats.ztest, two-sample pooled form, along axis 0.

    Operation order follows statsmodels 0.14 ``CompareMeans``-free ``ztest``:
    ``var = n1*x1.var(0) + n2*x2.var(0); var /= n1+n2-2*ddof; var *= 1/n1+1/n2``.
    """
    x1 = np.asarray(x1)
    x2 = np.asarray(x2)
    nobs1 = x1.shape[0]
    nobs2 = x2.shape[0]
    x1_mean = x1.mean(0)
    x2_mean = x2.mean(0)
    x1_var = x1.var(0)
    x2_var = x2.var(0)
    var_pooled = nobs1 * x1_var + nobs2 * x2_var
    var_pooled /= nobs1 + nobs2 - 2 * ddof
    var_pooled *= 1.0 / nobs1 + 1.0 / nobs2
    std_diff = np.sqrt(var_pooled)
    with np.errstate(divide="ignore", invalid="ignore"):
        zstat = (x1_mean - x2_mean - value) / std_diff
    pvalue = _stats.norm.sf(np.abs(zstat)) * 2
    return zstat, pvalue


class _QuietBar:
    """Stand-in for tqdm inside the reference module: same calls, no output."""

    def __init__(self, iterable=None, *args, **kwargs):
        self._it = iterable

    def __iter__(self):
        return iter(self._it if self._it is not None else ())

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def update(self, n=1):
        pass

    def set_description(self, *a, **k):
        pass

    def set_postfix(self, *a, **k):
        pass

    def refresh(self, *a, **k):
        pass

    def close(self):
        pass


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, "model", "phet.py"))


def load_reference_phet():
    """Return the reference's own ``PHet`` class, imported from ``/root/reference/src``."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_SRC)
    os.environ.setdefault("TQDM_DISABLE", "1")   # the class drives a tqdm bar per step: megabytes of progress text otherwise
    for name in ("anndata", "scanpy", "decoupler"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if "statsmodels.stats.weightstats" not in sys.modules:
        sm = types.ModuleType("statsmodels")
        sms = types.ModuleType("statsmodels.stats")
        smw = types.ModuleType("statsmodels.stats.weightstats")
        smw.ztest = ztest_restated
        sm.stats = sms
        sms.weightstats = smw
        sys.modules["statsmodels"] = sm
        sys.modules["statsmodels.stats"] = sms
        sys.modules["statsmodels.stats.weightstats"] = smw
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import logging
    import warnings
    saved_filters = list(warnings.filters)
    from model.phet import PHet  # noqa: E402  (phet.py:7-9 disables logging/warnings globally)
    import model.phet as _mod
    if not os.environ.get("PHET_REFERENCE_PROGRESS"):
        _mod.tqdm = _QuietBar   # the module's progress bar only (its source is untouched): megabytes of text otherwise
    logging.disable(logging.NOTSET)
    warnings.filters[:] = saved_filters
    return PHet
