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
    """Equal blocks of per = ceil(B / world) rounded up to even (16-byte aligned slices); only the tail is short, so the
    global index of row i of rank r's slice is r * per + i and the gathered buffer needs no re-packing."""
    for B in (1, 7, 37, 4096, 4097, 262144):
        for world in (1, 2, 3, 8):
            per = D.per_rank(B, world)
            assert per % 2 == 0 and per * world >= B and (per - 2) * world < B
            spans = [D.shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert all(lo == min(r * per, B) for r, (lo, hi) in enumerate(spans))
            b = W.config2(min(B, 64))
            for r in range(world):
                assert D.shard(b, r, world, pad=True).B == D.per_rank(b.B, world)


def _worker(rank, world, port, B, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O          # CPU stand-in for the local solver (test infrastructure)
    batch = W.config2(B)

    def local_solver(sh, z_out, info_out):        # writes into its slice of the collation buffers
        r = O.solve_batch(sh, nthreads=1)
        z_out.copy_(torch.from_numpy(r["z"]))
        info_out[:, 0] = torch.from_numpy(r["flag"]); info_out[:, 1] = torch.from_numpy(r["it"])

    z, flag, it = D.solve_sharded(batch, local_solver, torch.device("cpu"), D.TorchCollator())
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), z=z.numpy(), flag=flag.numpy(), it=it.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world_size_2_gloo_sharded_solve_matches_single_process(tmp_path):
    B, world = 37, 2          # odd batch: blocks of 20, the second rank owns 17 problems + 3 padding rows
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, B, str(tmp_path)), nprocs=world, join=True)
    from oracle import oracle as O
    ref = O.solve_batch(W.config2(B), nthreads=2)
    for r in range(world):
        got = np.load(tmp_path / f"rank{r}.npz")
        assert got["z"].shape == (B, 20, 17)
        assert np.array_equal(got["flag"], ref["flag"]) and np.array_equal(got["it"], ref["it"])
        assert np.array_equal(got["z"], ref["z"])
