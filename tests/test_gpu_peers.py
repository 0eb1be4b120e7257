"""Peer-store collation (nmpc_peers_* / nmpc_solve_batch_sharded_p2p_*): two OS processes, CUDA IPC, real kernels.

Runs on a single GPU as well (both ranks on cuda:0): what is under test is the cross-process mapping, the peer stores of
the three kernels' epilogues and the barrier kernels, not the NVLink bandwidth."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [1, 2, 3])
def test_peer_store_collation_matches_local_solves(world):
    port = _free_port()
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "tools", "p2p_worker.py"), str(r), str(world), str(port)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = []
    try:
        for p in procs:
            out, _ = p.communicate(timeout=240)
            outs.append(out)
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"rank {r}: OK" in out, f"rank {r} failed:\n{out[-3000:]}"
