#!/usr/bin/env python3
"""Single-object latency of the data plane alone (XferEngine.run = descriptor in kernel params + one launch + event
sync + results in pinned memory) and of the whole client call, with the warp-per-object path on and off.
N = 1 (local slab) or under torchrun (ring neighbour)."""
import json
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pct(v, q):
    v = sorted(v)
    return v[min(len(v) - 1, int(q * len(v)))]


def main():
    import torch

    from blackbird_b200 import _bb
    from blackbird_b200.models import latency_sweep
    from blackbird_b200.parallel import GpuRankCluster

    cl = GpuRankCluster(slab_bytes=1 << 30, cluster_id="lat")
    stream = torch.cuda.current_stream().cuda_stream
    dev = torch.device("cuda", cl.local_rank)
    sizes = [256, 1024, 4096, 16384, 65536]
    out = {"n_gpus": cl.world, "engine": [], "client": {}}
    eng = _bb.XferEngine(cl.local_rank, 64, 2)
    src = torch.randint(0, 256, (1 << 16,), dtype=torch.uint8, device=dev)
    dst = torch.zeros_like(src)
    for algo_name in ("XXH3", "BBH64", "CRC32C"):
        algo = getattr(_bb.ChecksumAlgo, algo_name)
        for small, flag, mail in ((True, True, True), (True, True, False), (True, False, False), (False, False, False)):
            eng.set_small_path(small)
            eng.set_flag_completion(flag)
            eng.set_mailbox(mail)
            for size in sizes:
                ts, ks = [], []
                for it in range(305):
                    t0 = time.perf_counter()
                    _, _, ms = eng.run([(src.data_ptr(), dst.data_ptr(), size)], algo, stream)
                    t1 = time.perf_counter()
                    if it >= 5:
                        ts.append((t1 - t0) * 1e6), ks.append(ms * 1e3)
                out["engine"].append({"algo": algo_name, "small_path": small, "flag_completion": flag, "mailbox": mail, "size": size, "call_p50_us": round(pct(ts, 0.5), 2),
                                      "call_p99_us": round(pct(ts, 0.99), 2), "kernel_p50_us": round(pct(ks, 0.5), 2)})
    target = f"gpu{(cl.rank + 1) % cl.world}"
    for small in (True, False):
        cl.fabric.set_small_path(small)
        cl.fabric.set_mailbox(small)
        out["client"]["small_path_on" if small else "small_path_off"] = latency_sweep(cl, sizes, target, iters=300, algo=_bb.ChecksumAlgo.XXH3)
    cl.stop()
    if cl.rank == 0:
        print(json.dumps(out))
        for r in out["engine"]:
            print("  engine %-6s small=%-5s flag=%-5s mailbox=%-5s %6d B: call p50 %6.2f p99 %6.2f us, kernel p50 %5.2f us" % (
                r["algo"], r["small_path"], r["flag_completion"], r["mailbox"], r["size"], r["call_p50_us"], r["call_p99_us"], r["kernel_p50_us"]), file=sys.stderr)
        for k, rows in out["client"].items():
            for r in rows:
                print("  client %-14s %6d B: put p50 %6.1f p99 %6.1f | get p50 %6.1f p99 %6.1f us" % (
                    k, r["size"], r["put_p50_us"], r["put_p99_us"], r["get_p50_us"], r["get_p99_us"]), file=sys.stderr)


if __name__ == "__main__":
    main()
