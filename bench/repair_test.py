#!/usr/bin/env python3
"""Worker loss + re-replication on the GPU tier, >= 3 ranks under torchrun.
Rank 0 puts objects with two replicas spread over the GPUs, one GPU-tier worker is removed, the Keystone re-replicates
every degraded object onto a surviving GPU: the destination worker pulls the shard out of the surviving replica's slab
itself (D_PULL: CUDA IPC mapping + one fused-kernel launch over NVLink, digest checked against the recorded one)."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blackbird_b200 import _bb  # noqa: E402
from blackbird_b200.parallel import GpuRankCluster  # noqa: E402


def main():
    cl = GpuRankCluster(slab_bytes=2 << 30, cluster_id="repair")
    assert cl.world >= 3, "needs >= 3 GPUs"
    dev = torch.device("cuda", cl.local_rank)
    s = torch.cuda.current_stream().cuda_stream
    nobj, size = 16, 32 << 20
    victim = cl.world - 1
    res = {"rank": cl.rank}
    if cl.rank == victim:
        cl.worker.inject_fault("drop_heartbeat")  # stays silent after it is removed (a crashed worker would)
    cl.barrier()
    if cl.rank == 0:
        src = torch.empty(nobj * size, dtype=torch.uint8, device=dev)
        _bb.random_fill(src.data_ptr(), nobj * size, 99, s)
        keys = [f"rep/{j}" for j in range(nobj)]
        cfg = _bb.WorkerConfig(replication_factor=2, max_workers_per_copy=1, ttl_ms=0, preferred_classes=[_bb.StorageClass.RAM_GPU],
                               enable_locality_awareness=False)
        assert all(e == _bb.ErrorCode.OK for e in cl.client.batch_put_device(keys, [src.data_ptr() + j * size for j in range(nobj)], [size] * nobj, cfg, s))
        before = {k: sorted(c.shards[0].worker_id for c in cl.client.get_workers(k)) for k in keys}
        hit = [k for k, w in before.items() if f"worker-gpu{victim}" in w]
        assert cl.keystone.remove_worker(f"worker-gpu{victim}") == _bb.ErrorCode.OK
        degraded = [k for k in keys if len(cl.client.get_workers(k)) == 1]
        t0 = time.perf_counter()
        repaired = cl.keystone.run_repair_once()
        repair_s = time.perf_counter() - t0
        after = {k: sorted(c.shards[0].worker_id for c in cl.client.get_workers(k)) for k in keys}
        ok = all(len(w) == 2 and f"worker-gpu{victim}" not in w for w in after.values())
        digests_ok = all(len({c.shards[0].checksum for c in cl.client.get_workers(k)}) == 1 for k in keys)
        out = torch.zeros_like(src)
        ecs, _ = cl.client.batch_get_device(keys, [out.data_ptr() + j * size for j in range(nobj)], [size] * nobj, s)
        torch.cuda.synchronize()
        res.update({"objects": nobj, "on_victim": len(hit), "degraded_after_removal": len(degraded), "repaired": repaired,
                    "all_objects_back_to_2_replicas": ok, "replica_digests_agree": digests_ok,
                    "read_back_ok": all(e == _bb.ErrorCode.OK for e in ecs) and bool(torch.equal(src, out)),
                    "repair_GBps": round(repaired * size / repair_s / 1e9, 1) if repaired else 0.0, "repair_s": round(repair_s, 4)})
        assert cl.client.put("repair/done", b"1", _bb.WorkerConfig(replication_factor=1, max_workers_per_copy=1)) == _bb.ErrorCode.OK
    else:
        # do NOT park in an NCCL barrier while the repair runs: the pull kernels run on these ranks' GPUs and a
        # spinning collective kernel plus a device-synchronising driver call is a deadlock recipe
        while cl.client.object_exists("repair/done") is not True:
            time.sleep(0.05)
    cl.barrier()
    allp = cl.rdv.gather_int(cl.worker.backend(f"hbm{cl.rank}").device_copies)
    if cl.rank == 0:
        res["fused_pulls_per_rank"] = allp
        print(json.dumps(res))
    cl.barrier()
    cl.stop()


if __name__ == "__main__":
    main()
