#!/usr/bin/env python3
"""DRAM tier through the device API: objects placed in a pinned / shared DRAM pool are written and read by the fused
kernel over PCIe (TMA to / from registered host memory, digest on the tensor cores).  Compared with the host-staged
path a client without the mapping takes (worker data server over loopback TCP + CPU digest + cudaMemcpy).
Device-timed (CUDA events around the public client calls, best of 5)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blackbird_b200 import _bb  # noqa: E402
from blackbird_b200.parallel import GpuRankCluster  # noqa: E402


def timed(fn, iters=5):
    best = 1e9
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    nobj, size = 16, 64 << 20
    cl = GpuRankCluster(slab_bytes=256 << 20, cluster_id="dram", dram_bytes=3 * nobj * size)
    dev = torch.device("cuda", cl.local_rank)
    s = torch.cuda.current_stream().cuda_stream
    src = torch.empty(nobj * size, dtype=torch.uint8, device=dev)
    _bb.random_fill(src.data_ptr(), nobj * size, 7, s)
    out = torch.empty_like(src)
    cfg = _bb.WorkerConfig(replication_factor=1, max_workers_per_copy=1, ttl_ms=0, preferred_classes=[_bb.StorageClass.RAM_CPU])
    OK = _bb.ErrorCode.OK
    st = {"it": 0, "k": None}
    sp = [src.data_ptr() + i * size for i in range(nobj)]
    op = [out.data_ptr() + i * size for i in range(nobj)]

    def put():
        if st["k"]:
            cl.client.batch_remove(st["k"])
        st["it"] += 1
        st["k"] = [f"dram/{st['it']}/{j}" for j in range(nobj)]
        assert all(e == OK for e in cl.client.batch_put_device(st["k"], sp, [size] * nobj, cfg, s))

    def get():
        ecs, _ = cl.client.batch_get_device(st["k"], op, [size] * nobj, s)
        assert all(e == OK for e in ecs)

    put_ms = timed(put)
    get_ms = timed(get)
    torch.cuda.synchronize()
    assert torch.equal(src, out)
    res = {"objects": nobj, "object_bytes": size, "fused_put_GBps": nobj * size / put_ms / 1e6, "fused_get_GBps": nobj * size / get_ms / 1e6,
           "mapped_host_pools": cl.fabric.mapped_host_pools()}
    # host-staged comparator: the same objects through a client that has no device transport for the DRAM pool
    plain = _bb.BlackbirdClient(cl.client_api, _bb.BlackbirdClientOptions(node_id="plain", io_parallelism=8))
    assert plain.connect() == OK
    host = torch.empty(size, dtype=torch.uint8).pin_memory()

    def staged_get():
        for j in range(nobj):
            data = plain.get(st["k"][j])  # data server -> host bytes (CRC / BBH64 verified on the CPU)
            host.copy_(torch.frombuffer(data, dtype=torch.uint8))
            out[j * size:(j + 1) * size].copy_(host, non_blocking=True)
        torch.cuda.synchronize()

    import time
    t0 = time.perf_counter()
    staged_get()
    res["host_staged_get_GBps"] = nobj * size / (time.perf_counter() - t0) / 1e9
    cl.stop()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
