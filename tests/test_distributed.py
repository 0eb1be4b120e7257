"""N > 1 path on CPU: world-size-2 gloo process group exercising shard arithmetic and the
end-of-batch result collation (the solve itself has no collective)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from forces_resilient_planner_b200 import distributed as D, workloads as W  # noqa: E402


def test_shard_ranges_partition_the_batch():
    for B in (0, 1, 7, 4096, 4097, 262144):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, B, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O          # CPU stand-in for the local solver (test infrastructure)
    batch = W.config2(B)

    def local_solver(sh):
        r = O.solve_batch(sh, nthreads=1)
        return r["z"], r["flag"], r["it"]

    z, flag, it = D.solve_sharded(batch, local_solver, torch.device("cpu"))
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), z=z, flag=flag, it=it)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world_size_2_gloo_sharded_solve_matches_single_process(tmp_path):
    B, world = 37, 2          # odd batch: shards of 19 and 18
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, B, str(tmp_path)), nprocs=world, join=True)
    from oracle import oracle as O
    ref = O.solve_batch(W.config2(B), nthreads=2)
    for r in range(world):
        got = np.load(tmp_path / f"rank{r}.npz")
        assert got["z"].shape == (B, 20, 17)
        assert np.array_equal(got["flag"], ref["flag"]) and np.array_equal(got["it"], ref["it"])
        assert np.array_equal(got["z"], ref["z"])
