"""Host replica of the device's keyed permutation (csrc/step3_impl.cuh ``perm_apply``), so that
``perm_mode='device'`` runs can be replayed on the CPU (tests feed these to the oracle)."""
from __future__ import annotations

import numpy as np

_M64 = (1 << 64) - 1


def _splitmix64(z: int) -> int:
    z = (z + 0x9E3779B97F4A7C15) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def device_permutation(seed: int, gene: int, cls: int, l: int, count: int | None = None) -> np.ndarray:
    """First ``count`` entries of the class-local permutation the device uses for (gene, cls)."""
    count = l if count is None else count
    a = _splitmix64((seed ^ ((0xD1B54A32D192ED03 * (2 * gene + cls + 1)) & _M64)) & _M64)
    b = _splitmix64(a)
    k = [a & 0xFFFFFFFF, a >> 32, b & 0xFFFFFFFF, b >> 32]
    w = max(l - 1, 0)
    for s in (1, 2, 4, 8, 16):
        w |= w >> s
    bits = int(w).bit_length()
    s1 = max(bits // 2, 1)
    s2 = max(bits // 3, 1)
    mul = [0x9E3779B1, 0x85EBCA77, 0xC2B2AE3D, 0x27D4EB2F]
    sh = [s1, s2, s1, s2]
    x = np.arange(count, dtype=np.uint64)
    todo = np.ones(count, dtype=bool)
    while todo.any():
        v = x[todo]
        for r in range(4):
            v = (v + np.uint64(k[r])) & np.uint64(w)
            v = (v * np.uint64(mul[r])) & np.uint64(w)
            v ^= v >> np.uint64(sh[r])
        x[todo] = v
        todo[todo] = v >= l
    return x.astype(np.uint32)
