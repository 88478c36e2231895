"""Host replica of the device-mode Step-3 permutation (csrc/step3_impl.cuh ``make_affine`` /
``affine_at`` and csrc/api.cu ``coprime_list``), so ``permutation="device"`` runs can be replayed
on the CPU: tests feed the replayed permutations to the oracle.

Device mode = a fixed pseudo-random layout of each class's cells in HBM (chosen by
``phet_set_classes``; readable as ``PHET_BUF_ROWS``) composed with a per-gene affine bijection of
the class-relative positions, ``t -> (a_g * t + b_g) mod n``.
"""
from __future__ import annotations

import math

import numpy as np

_M64 = (1 << 64) - 1
COPRIME_COUNT = 1024


def _splitmix64(z: int) -> int:
    z = (z + 0x9E3779B97F4A7C15) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


_COPRIME_CACHE: dict[int, list[int]] = {}


def coprime_list(n: int) -> list[int]:
    if n in _COPRIME_CACHE:
        return _COPRIME_CACHE[n]
    out, j = [], 0
    while len(out) < COPRIME_COUNT:
        z = _splitmix64((0xC0FFEE123456789 ^ ((n * 0x9E3779B97F4A7C15) & _M64) ^ j) & _M64)
        j += 1
        cand = 1 + z % (n - 1) if n > 2 else 1
        if math.gcd(cand, n) == 1:
            out.append(cand)
    _COPRIME_CACHE[n] = out
    return out


def affine_params(seed: int, gene: int, cls: int, n: int):
    z = _splitmix64((seed ^ ((0xD1B54A32D192ED03 * (2 * gene + cls + 1)) & _M64)) & _M64)
    a = coprime_list(n)[z & (COPRIME_COUNT - 1)]
    b = (z >> 32) % n
    return a, b


def device_positions(seed: int, gene: int, cls: int, n: int, count: int | None = None) -> np.ndarray:
    """Class-relative grouped positions visited by gene ``gene`` for t = 0..count-1."""
    count = n if count is None else count
    a, b = affine_params(seed, gene, cls, n)
    t = np.arange(count, dtype=np.uint64)
    return ((np.uint64(a) * t + np.uint64(b)) % np.uint64(n)).astype(np.int64)


def device_permutation(seed: int, gene: int, cls: int, class_rows: np.ndarray, rows_grouped: np.ndarray,
                       offset: int, count: int | None = None) -> np.ndarray:
    """The permutation in the reference's terms: for t = 0..count-1 the CLASS-LOCAL index (into the
    ascending ``class_rows``) of the cell gene ``gene`` puts at permuted position t.
    ``rows_grouped`` is the device layout (PHET_BUF_ROWS); ``offset`` = 0 for controls, n0 for cases."""
    n = len(class_rows)
    pos = device_positions(seed, gene, cls, n, count)
    rows = np.asarray(rows_grouped)[offset + pos]
    return np.searchsorted(class_rows, rows)
