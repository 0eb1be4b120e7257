"""One rank of the peer-store collation test (tests/test_gpu_peers.py): `python p2p_worker.py RANK WORLD PORT`.

Ranks share the visible GPUs round-robin (two ranks on ONE GPU still exercise the CUDA IPC mappings, the peer stores and
the barrier kernels; the contexts then time-slice).  Rendezvous over gloo on 127.0.0.1.  Every rank solves every rank's
shard locally as well and checks that the collated buffers are bit-identical to that, for the fp64, the mixed and the
warp-group kernel, three batches in a row (epochs of the barrier)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))


def main():
    rank, world, port = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    import torch
    import torch.distributed as dist
    from forces_resilient_planner_b200 import _lib, distributed as D, solver as S, workloads as W

    dev = torch.device("cuda", rank % torch.cuda.device_count())
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    per, N = 24, 20
    opts = _lib.default_opts()
    for np_dt, modes in ((np.float64, (0, 1, 2)), (np.float32, (1,))):
        col = D.PeerCollator(rank, world, dev, per, N, dtype=np_dt)
        for mode in modes:
            for step in range(3):
                shards = [W.config2(per, N, seed=100 + 10 * step + r) for r in range(world)]
                db = S.DeviceBatch(shards[rank], np_dt, dev)
                col.solve_sharded(db, opts, mode=mode)
                torch.cuda.synchronize(dev)
                col.check()
                z_all, info_all = col.z_all.cpu().numpy(), col.info_all.cpu().numpy()
                for r in range(world):
                    ref = S.DeviceBatch(shards[r], np_dt, dev)
                    if np_dt == np.float32:
                        S.solve_device(ref, opts)
                    elif mode == 0:
                        S.solve_device(ref, opts)
                    else:
                        S.solve_device(ref, opts, mixed=True, lowlatency=(mode == 2))
                    res = ref.result()
                    assert np.array_equal(z_all[r * per:(r + 1) * per], res.z), (np_dt, mode, step, r, "z")
                    assert np.array_equal(info_all[r * per:(r + 1) * per, 0], res.flag), (np_dt, mode, step, r, "flag")
                    assert np.array_equal(info_all[r * per:(r + 1) * per, 1], res.it), (np_dt, mode, step, r, "it")
                    assert (res.flag == 1).all()
        dist.barrier()
        col.close()
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank}: OK")


if __name__ == "__main__":
    main()
