#!/usr/bin/env python3
"""Single-process, two-GPU driver of the fused transfer kernel over NVLink (for `ncu`, which must not wrap a
multi-rank command).  GPU 0 runs the kernels: a *put* pushes local HBM -> GPU 1's memory (peer stores), a *get* pulls
GPU 1's memory -> local HBM (peer loads).  Prints device-timed GB/s per algorithm and direction.

  python bench/nvlink_single_proc.py --mib 1024 --objects 16 --iters 5 [--algo crc32c] [--ctas N]
  ncu --set full --section Nvlink_Tables --section Nvlink_Topology -k regex:bb_xfer -c 4 ... python bench/nvlink_single_proc.py --iters 1
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=1024)
    ap.add_argument("--objects", type=int, default=16)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--algo", default="all")
    ap.add_argument("--ctas", type=int, default=0)
    args = ap.parse_args()

    import torch

    from blackbird_b200 import _bb

    if torch.cuda.device_count() < 2:
        print(json.dumps({"error": "needs 2 GPUs"}))
        return 0
    n = args.mib << 20
    osz = n // args.objects
    d0, d1 = torch.device("cuda", 0), torch.device("cuda", 1)
    local_src = torch.empty(n, dtype=torch.uint8, device=d0)
    local_dst = torch.zeros(n, dtype=torch.uint8, device=d0)
    peer_buf = torch.zeros(n, dtype=torch.uint8, device=d1)
    torch.cuda.set_device(0)
    stream = torch.cuda.current_stream().cuda_stream
    _bb.random_fill(local_src.data_ptr(), n, 0xB200, stream)
    # make torch enable peer access in both directions (cudaDeviceEnablePeerAccess under the hood)
    peer_buf[:4096].copy_(local_src[:4096])
    local_dst[:4096].copy_(peer_buf[:4096])
    torch.cuda.synchronize(0)
    torch.cuda.synchronize(1)
    eng = _bb.XferEngine(0)
    if args.ctas:
        eng.set_max_ctas(args.ctas)
    algos = {"none": _bb.ChecksumAlgo.NONE, "bbh64": _bb.ChecksumAlgo.BBH64, "crc32c": _bb.ChecksumAlgo.CRC32C}
    if hasattr(_bb.ChecksumAlgo, "XXH3"):
        algos["xxh3"] = _bb.ChecksumAlgo.XXH3
    pick = list(algos) if args.algo == "all" else [args.algo]
    for name in pick:
        algo = algos[name]
        put = [(local_src.data_ptr() + i * osz, [peer_buf.data_ptr() + i * osz], osz) for i in range(args.objects)]
        get = [(peer_buf.data_ptr() + i * osz, [local_dst.data_ptr() + i * osz], osz) for i in range(args.objects)]
        res = {}
        for tag, items in (("put", put), ("get", get)):
            eng.run(items, algo, stream)  # warm-up
            best = 1e30
            for _ in range(args.iters):
                _, status, ms = eng.run(items, algo, stream)
                best = min(best, ms)
                assert not any(status)
            res[tag + "_GBps"] = round(n / best / 1e6, 1)
            res[tag + "_ms"] = round(best, 4)
        torch.cuda.synchronize(0)
        ok = bool(torch.equal(local_src, local_dst))
        print(json.dumps({"algo": name, "mib": args.mib, "objects": args.objects, "roundtrip_ok": ok, **res}), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
