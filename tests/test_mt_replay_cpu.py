"""CPU: the native replay of NumPy's legacy global stream (csrc/mt_replay.cpp) against ``np.random`` itself --
permutations, the index-set draws of phet.py:401-402, bounded draws, and the state the stream is left in.
No GPU: the swaps are applied by the library's host routine here (the device routine is checked in the GPU tests)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _mt():
    from phet_b200._capi import MtStream
    return MtStream.from_numpy()


def _same_state(mt):
    key, pos = mt.state()
    st = np.random.get_state()
    return np.array_equal(key, st[1]) and pos == st[2]


@pytest.mark.parametrize("n", [1, 2, 3, 29, 45, 252, 1000, 25000, 65535, 65536, 65537, 100000])
@pytest.mark.parametrize("burn", [0, 7, 623])
def test_permutation_bit_identical_to_numpy(n, burn):
    np.random.seed(12345 + n)
    if burn:
        np.random.randint(0, 2 ** 31 - 1, size=burn)      # start anywhere inside a 624-word block
    mt = _mt()
    for _ in range(3):
        got = mt.permutation(n)
        ref = np.random.permutation(n)
        assert np.array_equal(got, ref)
        assert _same_state(mt), "state advance differs from NumPy's"


def test_choice_without_replacement_is_a_permutation_prefix():
    """phet.py:401-402: np.random.choice(rows, m, replace=False) == rows[np.random.permutation(len(rows))[:m]]."""
    rows = np.arange(1000, 1252)
    np.random.seed(5)
    mt = _mt()
    for m in (45, 1, 252):
        ref = np.random.choice(a=rows, size=m, replace=False)
        got = rows[mt.permutation(len(rows), take=m)]
        assert np.array_equal(ref, got) and _same_state(mt)


def test_choice_with_replacement_and_randint():
    """replace=True (integer subsampling_size above a class size): rows[randint(0, n, m)]; the device-mode seed draw."""
    rows = np.arange(7, 22)
    np.random.seed(11)
    mt = _mt()
    for m in (20, 33):
        ref = np.random.choice(a=rows, size=m, replace=True)
        got = rows[mt.bounded(len(rows) - 1, m)]
        assert np.array_equal(ref, got) and _same_state(mt)
    ref = int(np.random.randint(0, 2 ** 31 - 1))
    assert int(mt.bounded(2 ** 31 - 2, 1)[0]) == ref and _same_state(mt)
    assert mt.bounded(0, 3).tolist() == [0, 0, 0] and _same_state(mt)   # rk_interval(0) does not touch the stream


def test_rounds_follow_the_reference_order_and_skipping_only_advances():
    """{n0, n1} x rounds is the order of phet.py:397-402 / 484-487; store=False must leave the same state."""
    n0, n1, rounds = 300, 45, 17
    np.random.seed(99)
    a, b = _mt(), _mt()
    d0, d1 = a.shuffle_draws([n0, n1], rounds)
    b.shuffle_draws([n0, n1], rounds, store=False)
    assert a.state()[1] == b.state()[1] and np.array_equal(a.state()[0], b.state()[0])
    from phet_b200 import _capi
    import ctypes as C
    lib = _capi.load_library()
    for r in range(rounds):
        for draws, n in ((d0, n0), (d1, n1)):
            out = np.empty(n, dtype=np.uint32)
            assert lib.phet_host_fisher_yates(draws[r].ctypes.data_as(C.c_void_p), draws.dtype.itemsize, n, n,
                                              out.ctypes.data_as(C.c_void_p)) == 0
            assert np.array_equal(out, np.random.permutation(n))
    assert _same_state(a)


def test_state_round_trip_into_numpy():
    np.random.seed(3)
    np.random.standard_normal(3)                 # leaves a cached gaussian in the legacy state
    mt = _mt()
    mt.permutation(5000)
    st_before = np.random.get_state()
    np.random.permutation(5000)
    want_next = np.random.random(4)
    np.random.set_state(st_before)
    mt.to_numpy()                                 # native state back into np.random (gaussian cache untouched)
    np.random.permutation(1)                      # n = 1 consumes nothing
    # the native stream already made the draw: np.random must now continue exactly behind it
    assert np.array_equal(np.random.random(4), want_next)


def test_scalar_and_avx512_scans_agree():
    """The vectorised scan (64 outputs per trip) against the one-output-at-a-time scan, in a fresh process
    (PHET_MT_SCALAR is read once)."""
    code = ("import numpy as np, sys; sys.path.insert(0, %r)\n"
            "from phet_b200._capi import MtStream\n"
            "np.random.seed(2024); mt = MtStream.from_numpy()\n"
            "d = mt.shuffle_draws([25000, 70000, 513, 64], 3)\n"
            "import hashlib; h = hashlib.sha256()\n"
            "[h.update(x.tobytes()) for x in d]; k, p = mt.state(); h.update(k.tobytes()); print(mt.simd(), p, h.hexdigest())\n" % ROOT)
    outs = []
    for env in ({}, {"PHET_MT_SCALAR": "1"}):
        e = dict(os.environ)
        e.update(env)
        outs.append(subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=e, check=True).stdout.split())
    assert outs[1][0] == "0"
    assert outs[0][1:] == outs[1][1:]


@pytest.mark.parametrize("sizes,rounds,chunks", [([25000, 25000], 40, 5), ([300, 45], 500, 7), ([100000, 70000], 6, 2),
                                                 ([1, 2, 3], 50, 3), ([65536, 65537], 4, 1)])
def test_parallel_replay_equals_sequential(sizes, rounds, chunks):
    """Generator / scanner / store-worker threads against the one-thread scan: same targets, same final state --
    which the tests above tie to np.random itself."""
    for burn in (0, 311):
        np.random.seed(4242)
        if burn:
            np.random.randint(0, 2 ** 31 - 1, size=burn)
        a, b = _mt(), _mt()
        da = a.shuffle_draws(sizes, rounds)
        db = b.shuffle_draws(sizes, rounds, workers=3, chunks=chunks)
        for x, y in zip(da, db):
            assert np.array_equal(x, y)
        ka, pa = a.state()
        kb, pb = b.state()
        assert np.array_equal(ka, kb) and pa == pb
        # and the stream goes on identically
        assert np.array_equal(a.permutation(1000), b.permutation(1000))


def test_parallel_replay_skip_blocks_only_advance():
    np.random.seed(8)
    a, b = _mt(), _mt()
    a.shuffle_draws([5000, 4000], 30, store=False)
    b.shuffle_draws([5000, 4000], 30, store=False, workers=2, chunks=3)
    assert np.array_equal(a.state()[0], b.state()[0]) and a.state()[1] == b.state()[1]


@pytest.mark.parametrize("sizes,rounds,env", [
    ([25000, 25000], 150, {}),                                                   # 16-bit ring, default threads
    ([25000, 25000], 150, {"PHET_REPLAY_RING_MB": "1", "PHET_REPLAY_GENERATORS": "2"}),   # a ring that wraps every few permutations
    ([30000, 20000], 120, {"PHET_REPLAY_GENERATORS": "3"}),
    ([100000, 70000], 24, {"PHET_REPLAY_GENERATORS": "2"}),                      # 32-bit ring
    ([5000, 300], 900, {"PHET_REPLAY_RING_MB": "1"}),
    ([25000, 25000], 60, {"PHET_REPLAY_GENERATORS": "0"}),                       # no ring: scanner generates for itself
    ([25000, 25000], 60, {"PHET_MT_SCALAR": "1", "PHET_REPLAY_GENERATORS": "2"}),         # no AVX-512
])
def test_ring_replay_equals_sequential(sizes, rounds, env):
    """Generator threads -> ring -> scanner -> store workers against the one-thread scan: same targets, same final
    state, for 16- and 32-bit rings, rings that wrap often, and the fallbacks (fresh process: the knobs are read once)."""
    code = ("import numpy as np, sys; sys.path.insert(0, %r)\n"
            "from phet_b200._capi import MtStream\n"
            "np.random.seed(777); np.random.randint(0, 2**31 - 1, size=100)\n"
            "a, b = MtStream.from_numpy(), MtStream.from_numpy()\n"
            "da = a.shuffle_draws(%r, %d)\n"
            "db = b.shuffle_draws(%r, %d, workers=3, chunks=6)\n"
            "ok = all(np.array_equal(x, y) for x, y in zip(da, db))\n"
            "ka, pa = a.state(); kb, pb = b.state()\n"
            "ok = ok and np.array_equal(ka, kb) and pa == pb and np.array_equal(a.permutation(999), b.permutation(999))\n"
            "print(int(ok), *b.par_threads)\n" % (ROOT, sizes, rounds, sizes, rounds))
    e = dict(os.environ)
    e.update(env)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=e, check=True, timeout=600).stdout.split()
    assert out[0] == "1", out
    if env.get("PHET_REPLAY_GENERATORS") == "0":
        assert out[1] == "0"


@pytest.mark.parametrize("procs", ["1", "2", "8"])
def test_thread_budget_per_local_process(procs):
    """Several ranks replay one stream side by side on one host: PHET_REPLAY_PROCS (torchrun's LOCAL_WORLD_SIZE) divides the
    cores between them and phet_mt_par_begin picks generator / worker counts from the share -- down to no ring at all
    (the scanner generating for itself) on a one-core share.  Whatever it picks, the draws and the final state are those
    of the one-thread scan."""
    code = ("import numpy as np, sys; sys.path.insert(0, %r)\n"
            "from phet_b200._capi import MtStream\n"
            "np.random.seed(31337); np.random.randint(0, 2**31 - 1, size=17)\n"
            "a, b, c = MtStream.from_numpy(), MtStream.from_numpy(), MtStream.from_numpy()\n"
            "da = a.shuffle_draws([25000, 24000], 40)\n"
            "db = b.shuffle_draws([25000, 24000], 40, workers=0, chunks=4)\n"
            "c.shuffle_draws([25000, 24000], 40, store=False, workers=0, chunks=4)\n"
            "ok = all(np.array_equal(x, y) for x, y in zip(da, db))\n"
            "for m in (b, c):\n"
            "    ok = ok and np.array_equal(a.state()[0], m.state()[0]) and a.state()[1] == m.state()[1]\n"
            "print(int(ok), *b.par_threads)\n" % ROOT)
    e = dict(os.environ)
    e["PHET_REPLAY_PROCS"] = procs
    e.pop("PHET_REPLAY_GENERATORS", None)
    e.pop("PHET_REPLAY_WORKERS", None)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=e, check=True, timeout=600).stdout.split()
    assert out[0] == "1", out
    gens, scanners, workers = (int(v) for v in out[1:4])
    cores = os.cpu_count() or 1
    share = max(cores // int(procs), 1)
    assert scanners == 1 and workers >= 1
    assert gens + scanners + workers <= max(share, 2) or gens == 0, (gens, workers, share)
