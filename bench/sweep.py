#!/usr/bin/env python3
"""BASELINE metric sweep: batched put/get GB/s and single-object p50/p99 latency vs object size
(256 B ... 256 MB) on N GPU workers, plus configs 3 (replication = 3) and 5 (feature-store fan-out).
Launch like bench.py (torchrun for N > 1); rank 0 writes profiles/sweep_n<N>.json."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    import torch

    from blackbird_b200 import _bb
    from blackbird_b200.models import feature_store_fanout, latency_sweep, replicated_put_verify, throughput_sweep
    from blackbird_b200.parallel import GpuRankCluster

    cl = GpuRankCluster(slab_bytes=6 << 30, cluster_id="sweep")
    target = f"gpu{(cl.rank + 1) % cl.world}"
    sizes = [256, 1024, 4096, 16384, 65536, 262144, 1 << 20, 4 << 20, 16 << 20, 64 << 20, 256 << 20]
    if args.quick:
        sizes = [256, 4096, 65536, 1 << 20, 16 << 20]
    res = {"n_gpus": cl.world, "placement": "ring neighbour" if cl.world > 1 else "local slab", "gpu": torch.cuda.get_device_name(0)}
    res["throughput_bbh64"] = throughput_sweep(cl, sizes, target)
    res["throughput_crc32c"] = throughput_sweep(cl, [s for s in sizes if s >= 65536], target, algo=_bb.ChecksumAlgo.CRC32C)
    res["latency_bbh64"] = latency_sweep(cl, sizes, target, iters=100 if args.quick else 300)
    if cl.world >= 3:
        res["replication3"] = replicated_put_verify(cl, 3)
    res["feature_store_fanout"] = feature_store_fanout(cl)
    if cl.world >= 3:
        res["feature_store_fanout_repl3"] = feature_store_fanout(cl, replication=3)
    cl.stop()
    if cl.rank == 0:
        out = args.out or os.path.join("profiles", f"sweep_n{cl.world}.json")
        os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
        json.dump(res, open(out, "w"), indent=1)
        print(json.dumps({"written": out, "sizes": len(sizes)}))
        for r in res["throughput_bbh64"]:
            print("  %10d B x %5d: put %8.1f get %8.1f GB/s (kernel) | put %8.1f get %8.1f GB/s (client)" % (
                r["size"], r["batch"], r["put_GBps_kernel"], r["get_GBps_kernel"], r["put_GBps_client"], r["get_GBps_client"]))
        for r in res["latency_bbh64"]:
            print("  %10d B: put p50 %7.1f p99 %7.1f us | get p50 %7.1f p99 %7.1f us" % (r["size"], r["put_p50_us"], r["put_p99_us"], r["get_p50_us"], r["get_p99_us"]))


if __name__ == "__main__":
    main()
