"""Input path on the device: MatrixMarket file -> filtered dense matrix resident in HBM.

``load_mtx`` replaces ``src/train.py:82-118`` of the reference (``sc.read_mtx`` -> dense float32, ``nan_to_num``,
cell filter, CPM gene filter): the file is parsed by the library's multi-threaded host reader, only the triplets
cross PCIe, and the dense matrix, the NaN clean-up and both filter decisions happen on the B200
(``phet_b200/csrc/mtx.cu``).  The result is a :class:`StagedMatrix` that ``PHet.fit_predict`` accepts in place of
a NumPy array, so the fit starts from HBM instead of from a Python parser and a 4·N·G-byte host copy.
"""
from __future__ import annotations

import os
import time

import numpy as np

from . import _capi


class StagedMatrix:
    """float32 (cells, genes) matrix in HBM, owned by ``ctx``.  ``kept_rows`` / ``kept_cols``: ascending
    row / column numbers of the file that survived the filter (``examples_ids`` / ``feature_ids`` of
    train.py:92,111) -- apply them to ``y`` and to the feature names."""

    dtype = np.dtype(np.float32)
    ndim = 2

    def __init__(self, ctx: _capi.Context, shape, kept_rows, kept_cols, file_shape, timings):
        self.ctx = ctx
        self.shape = (int(shape[0]), int(shape[1]))
        self.kept_rows = kept_rows
        self.kept_cols = kept_cols
        self.file_shape = file_shape
        self.timings = timings

    @property
    def device_ptr(self) -> int:
        return self.ctx.device_ptr(_capi.BUF_STAGED)[0]

    def to_host(self) -> np.ndarray:
        """Copy of the staged matrix (for parity checks)."""
        if self.shape[0] == 0 or self.shape[1] == 0:
            return np.zeros(self.shape, dtype=np.float32)
        return self.ctx.read(_capi.BUF_STAGED, np.float32, self.shape)

    def close(self):
        self.ctx.close()


def load_mtx(path: str, do_filter: bool = True, minimum_samples: int = 5, device: int | None = None,
             threads: int = 0) -> StagedMatrix:
    """Parse ``path`` (MatrixMarket coordinate, cells x genes) and stage it on ``device``.

    ``do_filter`` / ``minimum_samples``: train.py:88-118 (``--no-filtering`` turns it off; 5 is train.py's local
    default).  ``threads``: parser threads (0 = one per hardware thread)."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    t0 = time.perf_counter()
    mtx = _capi.MtxFile(path, threads)
    t1 = time.perf_counter()
    ctx = _capi.Context(device)
    try:
        n, g, rows, cols = ctx.stage_mtx(mtx, do_filter, minimum_samples)
    except Exception:
        ctx.close()
        raise
    finally:
        file_shape, entries = mtx.shape, mtx.entries
        mtx.close()
    t2 = time.perf_counter()
    return StagedMatrix(ctx, (n, g), rows, cols, file_shape,
                        {"parse_s": t1 - t0, "stage_s": t2 - t1, "entries": entries})
