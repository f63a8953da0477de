"""Multi-process host-side logic without GPUs: tests/multi_cpu_worker.py under torchrun (gloo, 2 and 3 ranks) -- rank 0's
Keystone over RPC, one worker process per rank, ring placement into ANOTHER process's DRAM pool through the one-sided
shared-memory path (and, with it disabled, through the TCP data servers), replication on distinct ranks, fan-out."""
import os
import socket
import subprocess
import sys

import pytest

from test_multi_gpu import _json_objects

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,shm", [(2, True), (3, False)])
def test_ring_replication_fanout_across_processes_on_cpu(world, shm):
    env = dict(os.environ, BB_TEST_SHM="1" if shm else "0", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multi_cpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = _json_objects(r.stdout)
    assert len(res) == world and all(x["ring"] == "ok" and x["fanout"] == "ok" and x["one_sided_shm"] == shm for x in res), res
