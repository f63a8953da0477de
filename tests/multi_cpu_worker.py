"""Rank program of tests/test_multi_cpu.py: torchrun + gloo, no GPU.  Every rank runs a worker with a memfd-backed
DRAM pool; the device batch API (through the loopback transport) puts to the ring neighbour's pool -- a different
PROCESS -- and gets back; replication over distinct ranks; rank 0 puts, everybody gets (fan-out); host get of a
neighbour-resident object; per-rank JSON result."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get("BB_PKG_ROOT"):  # sanitizer builds of the package (tests/conftest.py): the worker processes must load the same one
    sys.path.insert(0, os.path.abspath(os.environ["BB_PKG_ROOT"]))
from blackbird_b200 import _bb  # noqa: E402
from blackbird_b200.parallel import CpuRankCluster  # noqa: E402

OK = _bb.ErrorCode.OK
MiB = 1 << 20


def main():
    shm = os.environ.get("BB_TEST_SHM", "1") == "1"
    cl = CpuRankCluster(dram_bytes=96 * MiB, cluster_id="t-cpu-multi", shared_memory=shm)
    rank, world = cl.rank, cl.world
    nxt = (rank + 1) % world
    res = {"rank": rank, "world": world, "shm": shm}
    n, size = 6, 2 * MiB + 77
    rng = np.random.default_rng(100 + rank)
    src = rng.integers(0, 256, n * size, dtype=np.uint8)
    out = np.zeros_like(src)
    for algo in (_bb.ChecksumAlgo.BBH64, _bb.ChecksumAlgo.CRC32C):
        keys = [f"ring/{rank}/{int(algo)}/{i}" for i in range(n)]
        cfg = _bb.WorkerConfig(replication_factor=1, max_workers_per_copy=1, preferred_node=f"cpu{nxt}", ttl_ms=0, checksum=algo)
        assert cl.client.batch_put_device(keys, [src.ctypes.data + i * size for i in range(n)], [size] * n, cfg, 0) == [OK] * n
        sh = cl.client.get_workers(keys[2])[0].shards[0]
        assert sh.worker_id == f"worker-cpu{nxt}"  # the bytes live in another process
        ref = src[2 * size:3 * size]
        assert sh.checksum == (_bb.bbh64_reference(ref) if algo == _bb.ChecksumAlgo.BBH64 else _bb.crc32c(ref))
        out[:] = 0
        ecs, sizes = cl.client.batch_get_device(keys, [out.ctypes.data + i * size for i in range(n)], [size] * n, 0)
        assert ecs == [OK] * n and sizes == [size] * n and np.array_equal(src, out)
        assert cl.client.get(keys[0]) == bytes(src[:size])  # plain host get of the same object
        cl.client.batch_remove(keys)
    res["ring"] = "ok"
    m = cl.client.metrics_text() + cl.io_client.metrics_text()
    res["one_sided_shm"] = "bb_client_shm_put_bytes_total" in m
    assert res["one_sided_shm"] == (shm and world > 1 or shm)
    cl.barrier()
    # replication on distinct ranks
    if world >= 2:
        key = [f"rep/{rank}"]
        cfg = _bb.WorkerConfig(replication_factor=2, max_workers_per_copy=1, ttl_ms=0)
        assert cl.client.batch_put_device(key, [src.ctypes.data], [size], cfg, 0) == [OK]
        copies = cl.client.get_workers(key[0])
        assert len(copies) == 2 and copies[0].shards[0].worker_id != copies[1].shards[0].worker_id
        res["replicas"] = sorted(c.shards[0].worker_id for c in copies)
    cl.barrier()
    # fan-out: rank 0 puts (same seed everywhere), everybody gets
    shared = np.random.default_rng(7).integers(0, 256, 4 * MiB, dtype=np.uint8)
    fkeys = [f"feat/{j}" for j in range(4)]
    if rank == 0:
        cfg = _bb.WorkerConfig(replication_factor=1, max_workers_per_copy=1, ttl_ms=0, enable_locality_awareness=False)
        assert cl.client.batch_put_device(fkeys, [shared.ctypes.data + j * MiB for j in range(4)], [MiB] * 4, cfg, 0) == [OK] * 4
    cl.barrier()
    got = np.zeros_like(shared)
    ecs, _ = cl.client.batch_get_device(fkeys, [got.ctypes.data + j * MiB for j in range(4)], [MiB] * 4, 0)
    assert ecs == [OK] * 4 and np.array_equal(got, shared)
    res["fanout"] = "ok"
    cl.barrier()
    cl.stop()
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
