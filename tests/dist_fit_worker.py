"""torchrun worker: PHet.fit_predict sharded over the ranks of a NCCL group, checked against a
golden fixture (outputs of the unmodified reference).  Used by tests/test_gpu_distributed.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from conftest import load_golden
    from phet_b200 import PHet
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank()
    for name in [a for a in sys.argv[1:] if a.startswith("multi_")]:
        # multi-class branch (K >= 3), gene-sharded: fixtures of tests/golden/make_golden_multiclass.py
        from conftest import GOLDEN_DIR
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        for capture in (True, False):
            np.random.seed(12345 if rank == 0 else 999 + rank)
            est = PHet(num_subsamples=int(z["S"]), percentiles=tuple(z["percentiles"]),
                       multiconditions_strategy=str(z["strategy"]), multiconditions_aggregation=str(z["aggregation"]),
                       group_test=str(z["group_test"]), capture=capture, distributed=True)
            scores = est.fit_predict(z["X"], z["y"].astype(np.int64))
            f32 = z["X"].dtype == np.float32
            assert np.array_equal(est.o_, z["o"]), "o differs after the all-reduce"
            assert np.array_equal(est.bin_counts_, z["bin_counts"])
            if capture:
                assert np.array_equal(est.index_sets_anova_, z["idx_anova"]), "broadcast draws differ"
                assert np.array_equal(est.o_combo_, z["o_combo"])
                for c in range(z["H"].shape[1]):
                    assert np.array_equal(est.H_combo_[c], z["H"][:, c, :])
                assert np.allclose(est.fsum_, z["fsum"], rtol=1e-3 if f32 else 1e-5, atol=1e-9)
            assert np.allclose(scores, z["scores_ref"], rtol=1e-3 if f32 else 1e-5, atol=1e-12)
        if rank == 0:
            print("rank0 ok", name, "world", dist.get_world_size(), flush=True)
    for name in [a for a in sys.argv[1:] if not a.startswith("multi_")]:
        g = load_golden(name)
        X = np.ascontiguousarray(g["X"].astype(np.float64))
        ref = g["scores_f64"]
        for shard in ("subsamples", "genes"):
            for capture in (True, False):   # False: Step 3 takes the bound-and-prune kernel on each rank's gene slice
                np.random.seed(int(g["seed"]) if rank == 0 else 999 + rank)   # only rank 0's stream matters
                est = PHet(**g["params"], capture=capture, distributed=True, shard=shard)
                scores = est.fit_predict(X, g["y"].astype(np.int64))
                assert np.array_equal(est.index_sets_[:, 0], g["idx_i"]), "broadcast index sets differ"
                assert np.array_equal(est.o_, g["o_f64"]), "o differs after all-reduce"
                assert np.array_equal(est.bin_counts_, g["bin_counts_f64"])
                if capture:
                    assert np.allclose(est.fsum_, g["fsum_f64"], rtol=1e-5, atol=1e-9)
                assert np.allclose(scores, ref, rtol=1e-5, atol=1e-12)
                k = min(100, len(ref) // 4)
                assert np.array_equal(np.argsort(-scores[:, 0], kind="stable")[:k], np.argsort(-ref[:, 0], kind="stable")[:k])
        # the stream: every rank stops replaying after its own genes' draws and takes the final state from the last rank --
        # np.random must end, on EVERY rank, where a one-process fit (= the reference) leaves it
        np.random.seed(int(g["seed"]))
        PHet(**g["params"], distributed=False).fit_predict(X, g["y"].astype(np.int64))
        want = np.random.get_state()
        np.random.seed(int(g["seed"]) if rank == 0 else 999 + rank)
        PHet(**g["params"], distributed=True).fit_predict(X, g["y"].astype(np.int64))
        got = np.random.get_state()
        assert np.array_equal(want[1], got[1]) and want[2] == got[2], "np.random does not end in the reference's state"
        if rank == 0:
            print("rank0 ok", name, "world", dist.get_world_size(), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
