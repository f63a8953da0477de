#!/usr/bin/env python3
"""Why are 256-byte gets over NVLink slower than 4 KiB ones?  Single-object latency per size in both orders, with the engine
path (mailbox vs launch) that served each call.   torchrun --nproc-per-node 2 bench/lat256.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blackbird_b200 import _bb  # noqa: E402
from blackbird_b200.models import latency_sweep  # noqa: E402
from blackbird_b200.parallel import GpuRankCluster  # noqa: E402

cl = GpuRankCluster(slab_bytes=1 << 30, cluster_id="lat256")
target = f"gpu{(cl.rank + 1) % cl.world}"
rows = latency_sweep(cl, [4096, 256, 512, 1024, 2048, 4096, 256], target, iters=200, algo=_bb.ChecksumAlgo.XXH3)
local = latency_sweep(cl, [256, 4096], f"gpu{cl.rank}", iters=200, algo=_bb.ChecksumAlgo.XXH3)
for r in rows + local:
    r["rank"] = cl.rank
out = {"rank": cl.rank, "remote": rows, "local": local}
for i in range(cl.world):
    cl.host_barrier()
    if i == cl.rank:
        print(json.dumps(out), flush=True)
cl.stop()
