"""Host speed of the native np.random replay (no GPU): seconds for the draws of BASELINE config 4."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phet_b200._capi import MtStream  # noqa: E402

np.random.seed(12345)
mt = MtStream.from_numpy()
n, rounds = 25000, 1000
mt.shuffle_draws([n, n], 50)
t = time.perf_counter()
mt.shuffle_draws([n, n], rounds)
dt = time.perf_counter() - t
print("simd=%d store: %.2f us/permutation -> config 4 (62,000 permutations of 25,000): %.3f s" % (mt.simd(), dt / (2 * rounds) * 1e6, dt / (2 * rounds) * 62000))
t = time.perf_counter()
mt.shuffle_draws([n, n], rounds, store=False)
dt = time.perf_counter() - t
print("skip:  %.2f us/permutation -> %.3f s" % (dt / (2 * rounds) * 1e6, dt / (2 * rounds) * 62000))
t = time.perf_counter()
for _ in range(100):
    np.random.permutation(n)
dt = time.perf_counter() - t
print("np.random.permutation(25000): %.1f us" % (dt / 100 * 1e6))
for workers in (1, 2, 4, 6):
    out = None
    t = time.perf_counter()
    mt.shuffle_draws([n, n], rounds, workers=workers, chunks=8)
    dt = time.perf_counter() - t
    print("parallel replay, %d store workers: %.2f us/permutation -> config 4: %.3f s" % (workers, dt / (2 * rounds) * 1e6, dt / (2 * rounds) * 62000))
