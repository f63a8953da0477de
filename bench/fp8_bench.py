#!/usr/bin/env python3
"""MXFP8 tensors through the store: fused (pack inside the put kernel, unpack inside the get kernel) vs unfused
(mxfp8_pack kernel -> put of the packed bytes; get -> mxfp8_unpack kernel).  Reports bf16-side GB/s (the bytes the
application hands over), device time only (CUDA events around the kernels involved).  Single process: local HBM;
`torchrun --nproc-per-node 2`: objects live on the ring neighbour (packed bytes cross NVLink)."""
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
    cl = GpuRankCluster(slab_bytes=4 << 30, cluster_id="fp8")
    dev = torch.device("cuda", cl.local_rank)
    s = torch.cuda.current_stream().cuda_stream
    nobj, n = 16, 32 << 20  # 16 tensors x 32 Mi bf16 elements = 1 GiB of bf16
    xs = [(torch.randn(n, device=dev) * 3).to(torch.bfloat16) for _ in range(nobj)]
    outs = [torch.empty_like(x) for x in xs]
    packed = [torch.empty(_bb.mxfp8_packed_bytes(n), dtype=torch.uint8, device=dev) for _ in range(nobj)]
    target = f"gpu{(cl.rank + 1) % cl.world}"
    cfg = _bb.WorkerConfig(replication_factor=1, max_workers_per_copy=1, preferred_node=target, ttl_ms=0, preferred_classes=[_bb.StorageClass.RAM_GPU])
    OK = _bb.ErrorCode.OK
    state = {"it": 0}

    def keys(tag):
        state["it"] += 1
        return [f"{tag}/{cl.rank}/{state['it']}/{j}" for j in range(nobj)]

    res = {"rank": cl.rank, "world": cl.world, "placement": "ring neighbour over NVLink" if cl.world > 1 else "local HBM",
           "bf16_bytes": nobj * n * 2}
    last = {}

    def fused_put():
        last["k"] = keys("f")
        assert all(e == OK for e in cl.client.batch_put_device_fp8(last["k"], [x.data_ptr() for x in xs], [n] * nobj, cfg, s))

    def fused_get():
        assert all(e == OK for e in cl.client.batch_get_device_fp8(last["k"], [o.data_ptr() for o in outs], [n] * nobj, s))

    def unfused_put():
        last["k"] = keys("u")
        for x, p in zip(xs, packed):
            _bb.mxfp8_pack(x.data_ptr(), n, p.data_ptr(), s)
        assert all(e == OK for e in cl.client.batch_put_device(last["k"], [p.data_ptr() for p in packed], [p.numel() for p in packed], cfg, s))

    def unfused_get():
        ecs, _ = cl.client.batch_get_device(last["k"], [p.data_ptr() for p in packed], [p.numel() for p in packed], s)
        assert all(e == OK for e in ecs)
        for o, p in zip(outs, packed):
            _bb.mxfp8_unpack(p.data_ptr(), n, o.data_ptr(), s)

    def plain_put():
        last["k"] = keys("p")
        assert all(e == OK for e in cl.client.batch_put_device(last["k"], [x.data_ptr() for x in xs], [n * 2] * nobj, cfg, s))

    def plain_get():
        ecs, _ = cl.client.batch_get_device(last["k"], [o.data_ptr() for o in outs], [n * 2] * nobj, s)
        assert all(e == OK for e in ecs)

    for name, put, get in (("fused_mxfp8", fused_put, fused_get), ("unfused_mxfp8", unfused_put, unfused_get), ("plain_bf16", plain_put, plain_get)):
        cl.barrier()
        put_ms = 1e9
        get_ms = 1e9
        for _ in range(4):
            put_ms = min(put_ms, timed(put, 1))
            get_ms = min(get_ms, timed(get, 1))
            cl.client.batch_remove(last["k"])
        res[name] = {"put_ms": round(put_ms, 3), "get_ms": round(get_ms, 3), "put_bf16_GBps": round(nobj * n * 2 / put_ms / 1e6, 1),
                     "get_bf16_GBps": round(nobj * n * 2 / get_ms / 1e6, 1)}
    # numerics: fused get of a fused put equals the unfused round trip bit for bit
    fused_put()
    fused_get()
    torch.cuda.synchronize()
    ref = torch.empty_like(xs[0])
    _bb.mxfp8_pack(xs[0].data_ptr(), n, packed[0].data_ptr(), s)
    _bb.mxfp8_unpack(packed[0].data_ptr(), n, ref.data_ptr(), s)
    torch.cuda.synchronize()
    res["bit_exact_vs_unfused"] = bool(torch.equal(outs[0].view(torch.int16), ref.view(torch.int16)))
    print(json.dumps(res))
    cl.stop()


if __name__ == "__main__":
    main()
