"""Rank program of tests/test_multi_gpu.py (launched with torchrun, one rank per GPU).

Checks, on N >= 2 GPUs: ring-neighbour put/get over NVLink with both checksums, bit-exact data and CPU-model digests;
replication = 2 single-read fan-out to distinct GPUs with fail-over to the surviving replica after corruption;
1 -> N read fan-out (BASELINE config 5); a DRAM pool owned by ANOTHER process reached by the fused kernel through its
memfd mapping.  Prints one JSON line per rank; any assertion failure makes torchrun exit non-zero."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get("BB_PKG_ROOT"):  # sanitizer builds of the package (tests/conftest.py): the worker processes must load the same one
    sys.path.insert(0, os.path.abspath(os.environ["BB_PKG_ROOT"]))
from blackbird_b200 import _bb  # noqa: E402
from blackbird_b200.models.workloads import feature_store_fanout, replicated_put_verify  # noqa: E402
from blackbird_b200.parallel import GpuRankCluster  # noqa: E402

OK = _bb.ErrorCode.OK


def main():
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    cl = GpuRankCluster(slab_bytes=1 << 30, cluster_id="t-multi", dram_bytes=256 << 20, nvls_arena_bytes=128 << 20,
                        nvls_group_size=min(3, world_env))
    dev = torch.device("cuda", cl.local_rank)
    s = torch.cuda.current_stream().cuda_stream
    rank, world = cl.rank, cl.world
    nxt = (rank + 1) % world
    res = {"rank": rank, "world": world}

    # ---- ring neighbour, both checksums, odd sizes
    n, size = 12, (3 << 20) + 112
    stride = (size + 255) // 256 * 256
    src = torch.randint(0, 256, (n * stride,), dtype=torch.uint8, device=dev)
    for algo in (_bb.ChecksumAlgo.BBH64, _bb.ChecksumAlgo.CRC32C):
        out = torch.zeros_like(src)
        keys = [f"ring/{rank}/{int(algo)}/{i}" for i in range(n)]
        cfg = _bb.WorkerConfig(replication_factor=1, max_workers_per_copy=1, preferred_node=f"gpu{nxt}", ttl_ms=0, checksum=algo,
                               preferred_classes=[_bb.StorageClass.RAM_GPU])
        assert all(e == OK for e in cl.client.batch_put_device(keys, [src.data_ptr() + i * stride for i in range(n)], [size] * n, cfg, s))
        sh = cl.client.get_workers(keys[1])[0].shards[0]
        assert sh.worker_id == f"worker-gpu{nxt}", sh.worker_id  # the bytes live on the neighbour's HBM
        ref = src[stride:stride + size].cpu().numpy()
        assert sh.checksum == (_bb.bbh64(ref) if algo == _bb.ChecksumAlgo.BBH64 else _bb.crc32c(ref))
        ecs, sizes = cl.client.batch_get_device(keys, [out.data_ptr() + i * stride for i in range(n)], [stride] * n, s)
        assert all(e == OK for e in ecs) and sizes == [size] * n
        torch.cuda.synchronize()
        for i in range(n):
            assert torch.equal(src[i * stride:i * stride + size], out[i * stride:i * stride + size])
        cl.client.batch_remove(keys)
    res["ring"] = "ok"
    assert cl.fabric.path_bytes(True, 1) >= 2 * n * size and cl.fabric.path_bytes(False, 1) >= 2 * n * size  # all of it crossed NVLink
    assert 'path="nvlink"' in cl.client.metrics_text()
    cl.barrier()

    # ---- replication 2: fan-out, then corrupt one replica and read through the other
    r = replicated_put_verify(cl, replication=2, nobj=8, size=4 << 20, iters=1)
    res["replication2_put_payload_GBps"] = round(r["put_payload_GBps"], 1)
    key = [f"fo/{rank}"]
    cfg = _bb.WorkerConfig(replication_factor=2, max_workers_per_copy=1, ttl_ms=0, preferred_classes=[_bb.StorageClass.RAM_GPU])
    blob = torch.randint(0, 256, (1 << 20,), dtype=torch.uint8, device=dev)
    assert cl.client.batch_put_device(key, [blob.data_ptr()], [1 << 20], cfg, s) == [OK]
    copies = cl.client.get_workers(key[0])
    assert len(copies) == 2 and copies[0].shards[0].worker_id != copies[1].shards[0].worker_id
    cl.barrier()
    mine = [c.shards[0] for c in copies if c.shards[0].worker_id == f"worker-gpu{rank}"]
    if mine:  # corrupt the replica held by this rank's own worker; the get must fail over to the remote one
        cl.worker.backend(mine[0].pool_id).write(mine[0].offset + 64, b"\x5a" * 32)
    back = torch.zeros_like(blob)
    ecs, _ = cl.client.batch_get_device(key, [back.data_ptr()], [1 << 20], s)
    torch.cuda.synchronize()
    assert ecs == [OK] and torch.equal(back, blob)
    res["failover"] = "corrupted local replica -> remote replica" if mine else "no local replica"
    cl.barrier()

    # ---- NVLS: symmetric replicas are written with ONE multimem.st stream per object (the switch replicates); every rank
    # reads the objects back (its own replica when it holds one, a peer's otherwise) and compares
    if cl.arena is not None:
        R = min(3, world)
        before = cl.fabric.multicast_puts
        m = replicated_put_verify(cl, replication=R, nobj=6, size=(2 << 20) + 4096, iters=2, symmetric=True)
        assert cl.fabric.multicast_puts > before, "symmetric replicas did not take the multicast path"
        assert cl.fabric.path_bytes(True, 3) > 0
        res["nvls_multicast_put_payload_GBps"] = round(m["put_payload_GBps"], 1)
        res["nvls_groups"] = cl.arena.num_groups()
    else:
        res["nvls_multicast_put_payload_GBps"] = None
    cl.barrier()

    # ---- config 5: rank 0 puts, everybody gets
    f = feature_store_fanout(cl, nshards=32, size=1 << 20, iters=2)
    res["fanout_aggregate_get_GBps"] = round(f["aggregate_get_GBps"], 1)
    cl.barrier()

    # ---- DRAM pool of another process, through the fused kernel (memfd mapping + cudaHostRegister in this process)
    n2, sz2 = 4, 2 << 20
    d_src = torch.randint(0, 256, (n2 * sz2,), dtype=torch.uint8, device=dev)
    d_out = torch.zeros_like(d_src)
    keys = [f"dram/{rank}/{i}" for i in range(n2)]
    cfg = _bb.WorkerConfig(replication_factor=1, max_workers_per_copy=1, preferred_node=f"gpu{nxt}", ttl_ms=0,
                           preferred_classes=[_bb.StorageClass.RAM_CPU])
    l0 = cl.fabric.launches
    assert all(e == OK for e in cl.client.batch_put_device(keys, [d_src.data_ptr() + i * sz2 for i in range(n2)], [sz2] * n2, cfg, s))
    sh = cl.client.get_workers(keys[0])[0].shards[0]
    assert sh.storage_class == _bb.StorageClass.RAM_CPU and sh.worker_id == f"worker-gpu{nxt}"
    ecs, _ = cl.client.batch_get_device(keys, [d_out.data_ptr() + i * sz2 for i in range(n2)], [sz2] * n2, s)
    torch.cuda.synchronize()
    assert all(e == OK for e in ecs) and torch.equal(d_src, d_out)
    assert cl.fabric.launches == l0 + 2 and cl.fabric.mapped_host_pools() >= 1  # fused kernel both ways, no TCP staging
    res["remote_dram_pool"] = "fused kernel over PCIe (mapped from the neighbour's worker process)"
    cl.barrier()
    cl.stop()
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
