#!/usr/bin/env python3
"""Per-tile pipeline trace of the fused transfer kernel (globaltimer-stamped trace ring).

Runs one batched put and one batched get with tile tracing on and reports, per direction: load latency
(issue -> landed in shared memory), time in shared memory before the store was issued, store drain time
(store issued -> slot released), and the achieved issue rate.  `torchrun --nproc-per-node 2` traces the NVLink
path (objects placed on the ring neighbour); a single process traces HBM -> HBM.  Output: JSON on stdout, copy it
into profiles/."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blackbird_b200 import _bb  # noqa: E402
from blackbird_b200.parallel import GpuRankCluster  # noqa: E402


def summarize(tr, tile_bytes=16384):
    t = np.asarray(tr, dtype=np.uint64).reshape(-1, 4).astype(np.int64)
    ok = (t > 0).all(axis=1)
    t = t[ok]
    if len(t) == 0:
        return {"tiles": 0}
    span = (t[:, 3].max() - t[:, 0].min()) / 1e3

    def pct(x):
        return {"p50_us": round(float(np.percentile(x, 50)) / 1e3, 2), "p99_us": round(float(np.percentile(x, 99)) / 1e3, 2), "mean_us": round(float(x.mean()) / 1e3, 2)}

    return {"tiles": int(len(t)), "span_us": round(float(span), 1), "GBps": round(len(t) * tile_bytes / span / 1e3, 1),
            "load_issue_to_landed": pct(t[:, 1] - t[:, 0]), "landed_to_store_issued": pct(t[:, 2] - t[:, 1]),
            "store_issued_to_slot_released": pct(t[:, 3] - t[:, 2]), "tile_lifetime": pct(t[:, 3] - t[:, 0])}


def main():
    cl = GpuRankCluster(slab_bytes=3 << 30, cluster_id="trace")
    dev = torch.device("cuda", cl.local_rank)
    stream = torch.cuda.current_stream().cuda_stream
    nobj, size = 16, 64 << 20
    src = torch.empty(nobj * size, dtype=torch.uint8, device=dev)
    _bb.random_fill(src.data_ptr(), nobj * size, 7 + cl.rank, stream)
    out = torch.zeros_like(src)
    target = f"gpu{(cl.rank + 1) % cl.world}"
    cfg = _bb.WorkerConfig(replication_factor=1, max_workers_per_copy=1, preferred_node=target, ttl_ms=0, preferred_classes=[_bb.StorageClass.RAM_GPU])
    cl.client.set_device_pipeline_chunks(1)
    res = {"rank": cl.rank, "world": cl.world, "placement": "ring neighbour over NVLink" if cl.world > 1 else "local HBM"}
    for it in range(3):  # last iteration is the traced one
        keys = [f"tr/{cl.rank}/{it}/{j}" for j in range(nobj)]
        cl.fabric.set_tile_trace(it == 2)
        cl.barrier()
        assert all(e == _bb.ErrorCode.OK for e in cl.client.batch_put_device(keys, [src.data_ptr() + j * size for j in range(nobj)], [size] * nobj, cfg, stream))
        if it == 2:
            res["put"] = summarize(cl.fabric.tile_trace())
        cl.barrier()
        ecs, _ = cl.client.batch_get_device(keys, [out.data_ptr() + j * size for j in range(nobj)], [size] * nobj, stream)
        assert all(e == _bb.ErrorCode.OK for e in ecs)
        if it == 2:
            res["get"] = summarize(cl.fabric.tile_trace())
        cl.client.batch_remove(keys)
    cl.fabric.set_tile_trace(False)
    print(json.dumps(res))
    cl.stop()


if __name__ == "__main__":
    main()
