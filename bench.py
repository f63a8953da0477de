#!/usr/bin/env python3
"""Headline benchmark: batched put + get throughput of the object store on N B200 workers.

Metric (BASELINE.json): batched put/get GB/s at 1/2/4/8 GPU-tier workers, synthetic random-byte
objects, device-timed, max over ranks.  One *step* = one `batch_put_device` of B objects of S
bytes (fused kernel: copy + BBH64 digest, placements from the Keystone) followed by one
`batch_get_device` of the same objects (fused kernel: copy + digest verify), then a
`batch_remove` so that the slab is recycled.  `value` = payload bytes moved per second by the
whole job (put bytes + get bytes, all ranks).

Topology: one process per GPU (torchrun); rank 0 hosts the Keystone (RPC over loopback for the
other ranks); every rank runs a GPU-tier worker whose HBM slab is exported to all peers (CUDA
IPC).  With N >= 2 every object is placed on the writer's ring neighbour, so every payload byte
crosses NVLink once per put and once per get; with N = 1 the slab is local (HBM-bound).

  python bench.py --gpus 1 --steps 20 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
      --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BASELINE_GBPS = 0.233  # the only throughput figure in the reference tree (configs/worker.yaml:19, unsourced)


def reference_arm() -> int:
    """The reference cannot be installed or built offline (see DESIGN.md §Reference arm)."""
    print(json.dumps({
        "impl": "reference",
        "unavailable": "blackbird-io/blackbird is a CMake C++ project with no Python package; pip install fails (no setup.py/"
                       "pyproject), configure needs network (yalantinglibs/gtest FetchContent), UCX/etcd/glog/yaml-cpp/liburing "
                       "are absent, and src/worker/storage/cxl_memory_backend.cpp does not compile at this commit",
    }))
    return 0


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--ctas", type=int, default=0, help="persistent CTAs of the fused kernel (0 = default)")
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--objects", type=int, default=64, help="objects per batch per rank")
    ap.add_argument("--object-mib", type=float, default=64.0)
    ap.add_argument("--algo", default="bbh64", choices=["bbh64", "crc32c", "none"])
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--no-comparators", action="store_true")
    ap.add_argument("--sync", choices=["none", "step", "phase"], default="phase", help="N>1: ranks rendezvous per step / per put-get phase (in the timed region)")
    ap.add_argument("--idle-odd", action="store_true", help="diagnostic: odd ranks idle (unidirectional NVLink traffic)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm()

    import torch
    import torch.distributed as dist

    from blackbird_b200 import _bb
    from blackbird_b200.parallel import GpuRankCluster
    from blackbird_b200.utils import ClockSampler

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if args.gpus > 1:
            print(f"bench.py: --gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})", file=sys.stderr)
            return 2
    nobj = args.objects
    osz = int(args.object_mib * (1 << 20))
    step_bytes = nobj * osz
    algo = {"bbh64": _bb.ChecksumAlgo.BBH64, "crc32c": _bb.ChecksumAlgo.CRC32C, "none": _bb.ChecksumAlgo.NONE}[args.algo]

    cl = GpuRankCluster(slab_bytes=3 * step_bytes + (64 << 20))
    rank, dev = cl.rank, torch.device("cuda", cl.local_rank)
    stream = torch.cuda.current_stream().cuda_stream
    target_node = f"gpu{(rank + 1) % world}"  # ring neighbour (== self when N = 1)
    cfg = _bb.WorkerConfig(replication_factor=1, max_workers_per_copy=1, preferred_node=target_node, ttl_ms=0,
                           checksum=algo, preferred_classes=[_bb.StorageClass.RAM_GPU])

    # synthetic random-byte objects: payload (1 GiB by default) is larger than L2 (126 MB)
    src = torch.empty(step_bytes, dtype=torch.uint8, device=dev)
    _bb.random_fill(src.data_ptr(), step_bytes, 0xB200 + rank, stream)
    out = torch.empty(step_bytes, dtype=torch.uint8, device=dev)
    src_ptrs = [src.data_ptr() + i * osz for i in range(nobj)]
    out_ptrs = [out.data_ptr() + i * osz for i in range(nobj)]
    sizes = [osz] * nobj
    OK = _bb.ErrorCode.OK

    def step(tag: str, i: int, keys=None):
        if args.idle_odd and rank % 2 == 1:
            return
        keys = keys or [f"r{rank}/{tag}{i}/o{j}" for j in range(nobj)]
        if args.sync != "none":
            rendezvous()  # bulk-synchronous step (checkpoint / KV hand-off pattern): all ranks put, then all ranks get
        ecs = cl.client.batch_put_device(keys, src_ptrs, sizes, cfg, stream)
        assert all(e == OK for e in ecs), f"put failed: {[str(e) for e in ecs if e != OK][:3]}"
        if args.sync == "phase":
            rendezvous()
        ecs, got = cl.client.batch_get_device(keys, out_ptrs, sizes, stream)
        assert all(e == OK for e in ecs), f"get failed: {[str(e) for e in ecs if e != OK][:3]}"
        ecs = cl.client.batch_remove(keys)
        assert all(e == OK for e in ecs)

    sync_tok = torch.zeros(1, dtype=torch.int32, device=dev)

    def rendezvous():
        # inside the timed region: its cost is part of the reported number
        if world > 1:
            dist.all_reduce(sync_tok)
            torch.cuda.current_stream().synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    if args.ctas:
        cl.fabric.set_max_ctas(args.ctas)
    # ---------------------------------------------------------------- warm-up + correctness
    sampler = ClockSampler(cl.local_rank, 100).start() if rank == 0 else None
    for i in range(max(3, args.warmup)):
        step("w", i)
    torch.cuda.synchronize()
    assert torch.equal(src, out), "payload mismatch after put+get"

    # ---------------------------------------------------------------- device-resident timed region
    launches0 = cl.fabric.launches
    all_keys = [[f"r{rank}/t{i}/o{j}" for j in range(nobj)] for i in range(args.steps)]  # fresh keys every step
    cl.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step("t", i, all_keys[i])
    e1.record()
    torch.cuda.synchronize()
    cl.barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = int(sum_over_ranks(cl.fabric.launches - launches0))
    if ms < 1500:  # give nvidia-smi time to take samples under the same load (same count on every rank: steps rendezvous)
        for i in range(int(1200.0 / max(ms / args.steps, 0.05)) + 1):
            step("c", i)
    clocks = sampler.stop() if sampler else None
    phases = {k: round(v[1] / max(v[0], 1), 1) for k, v in cl.client.phase_summary().items() if k.startswith("phase_")}
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, phases)
        phases = {f"rank{r}": g for r, g in enumerate(gathered)}
    total_bytes = 2.0 * step_bytes * args.steps * world
    value = total_bytes / (ms * 1e-3) / 1e9

    # kernel-only view of one put and one get (explains the headline)
    keys = [f"r{rank}/k/o{j}" for j in range(nobj)]
    m0 = cl.fabric.total_device_ms
    assert all(e == OK for e in cl.client.batch_put_device(keys, src_ptrs, sizes, cfg, stream))
    m1 = cl.fabric.total_device_ms
    ecs, _ = cl.client.batch_get_device(keys, out_ptrs, sizes, stream)
    m2 = cl.fabric.total_device_ms
    put_ms, get_ms = m1 - m0, m2 - m1  # sum of the (pipelined) chunk kernels
    cl.client.batch_remove(keys)
    put_ms, get_ms = max_over_ranks(put_ms), max_over_ranks(get_ms)

    # ---------------------------------------------------------------- end-to-end (host -> device -> store -> device -> host)
    h_src = torch.empty(step_bytes, dtype=torch.uint8).pin_memory()
    h_src.copy_(src.cpu())
    sample = 4096
    h_res = torch.empty(nobj * sample, dtype=torch.uint8).pin_memory()
    d_res = torch.empty(nobj * sample, dtype=torch.uint8, device=dev)
    idx = (torch.arange(nobj, device=dev).repeat_interleave(sample) * osz + torch.arange(sample, device=dev).repeat(nobj))

    h_ptrs = [h_src.data_ptr() + i * osz for i in range(nobj)]

    def e2e_step(i: int, zero_copy: bool):
        if zero_copy:
            # the put kernel reads the pinned host buffer itself (TMA over PCIe): the H2D copy IS the put
            keys = [f"r{rank}/z{i}/o{j}" for j in range(nobj)]
            if args.sync != "none":
                rendezvous()
            ecs = cl.client.batch_put_device(keys, h_ptrs, sizes, cfg, stream)
            assert all(e == OK for e in ecs)
            if args.sync == "phase":
                rendezvous()
            ecs, _ = cl.client.batch_get_device(keys, out_ptrs, sizes, stream)
            assert all(e == OK for e in ecs)
            assert all(e == OK for e in cl.client.batch_remove(keys))
        else:
            src.copy_(h_src, non_blocking=True)      # H2D of this step's inputs from pinned memory
            step("e", i)                               # public API: batch_put_device / batch_get_device
        torch.index_select(out, 0, idx, out=d_res)    # result read-back: 4 KiB of every object
        h_res.copy_(d_res, non_blocking=True)
        torch.cuda.synchronize()

    def e2e_run(zero_copy: bool) -> float:
        out.zero_()
        e2e_step(-1, zero_copy)
        cl.barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        for i in range(args.e2e_steps):
            e2e_step(i, zero_copy)
        e3.record()
        torch.cuda.synchronize()
        cl.barrier()
        assert torch.equal(h_res.view(nobj, sample), h_src.view(nobj, osz)[:, :sample]), "end-to-end sample mismatch"
        return max_over_ranks(e2.elapsed_time(e3))

    e2e_staged_ms = e2e_run(False)
    e2e_zc_ms = e2e_run(True)
    e2e_mode = "zero_copy" if e2e_zc_ms < e2e_staged_ms else "staged"
    e2e_ms = min(e2e_zc_ms, e2e_staged_ms)
    e2e_value = 2.0 * step_bytes * args.e2e_steps * world / (e2e_ms * 1e-3) / 1e9
    h2d = step_bytes + nobj * 64 * 2 + (nobj + 1) * 4 * 2  # payload + put/get descriptor tables
    d2h = nobj * sample + nobj * 12 * 2                    # sampled result + digests/status of put and get

    # ---------------------------------------------------------------- comparators (same buffers, outside the timed regions)
    comparators = {}
    if not args.no_comparators:
        def timed(fn, iters=5):
            fn()
            torch.cuda.synchronize()
            cl.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(iters):
                fn()
            b.record()
            torch.cuda.synchronize()
            return max_over_ranks(a.elapsed_time(b)) / iters

        if world == 1:
            t = timed(lambda: [out[i * osz:(i + 1) * osz].copy_(src[i * osz:(i + 1) * osz]) for i in range(nobj)])
            comparators["per_object_cudaMemcpyAsync_GBps"] = step_bytes / t / 1e6
            scratch = torch.empty(step_bytes // 512 + 2, dtype=torch.int32, device=dev)
            crc = torch.zeros(1, dtype=torch.int32, device=dev)

            def unfused():
                _bb.copy_simt(out.data_ptr(), src.data_ptr(), step_bytes, stream)
                _bb.crc32c_device(out.data_ptr(), step_bytes, crc.data_ptr(), scratch.data_ptr(), stream)

            t = timed(unfused, 2)
            comparators["unfused_copy_plus_crc32c_GBps"] = step_bytes / t / 1e6
        else:
            nxt, prv = (rank + 1) % world, (rank - 1) % world

            def nccl_ring():
                ops = []
                for i in range(nobj):
                    ops.append(dist.P2POp(dist.isend, src[i * osz:(i + 1) * osz], nxt))
                    ops.append(dist.P2POp(dist.irecv, out[i * osz:(i + 1) * osz], prv))
                for r in dist.batch_isend_irecv(ops):
                    r.wait()

            t = timed(nccl_ring, 3)
            comparators["nccl_grouped_send_recv_GBps_per_rank"] = step_bytes / t / 1e6

    cl.stop()
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        per_gpu = value / world
        roof = (peaks.get("hbm_gbs", 6650.0) / 2.0) if world == 1 else 770.0
        line = {
            "metric": "batched put+get payload throughput (GB/s), %g MiB random-byte objects, GPU tier, checksum fused" % args.object_mib,
            "value": round(value, 2),
            "unit": "GB/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": max(3, args.warmup),
            "ms_per_step": round(ms / args.steps, 4),
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": round(value / BASELINE_GBPS, 1),
            "dtype": "uint8 payload (byte objects; no reduced precision)",
            "data": "synthetic random bytes, device-resident; fresh keys every step; slab recycled by batch_remove",
            "config": {
                "model": "object-store sweep point: batch of %d x %.0f MiB objects per rank" % (nobj, args.object_mib),
                "global_batch": nobj * world,
                "seq_len": osz,
                "parallelism": "ring%d" % world if world > 1 else "local1",
                "placement": "writer's ring neighbour (all payload crosses NVLink)" if world > 1 else "local HBM slab",
                "checksum": args.algo,
                "rank_sync": args.sync if world > 1 else "n/a",
                "l2_policy": "per-step payload %.2f GiB per rank >> 126 MB L2 (inputs larger than L2)" % (step_bytes / 2**30),
                "control_plane_in_timed_region": "batch_put_start + batch_put_complete + batch_get_workers + batch_remove RPCs every step",
            },
            "kernel_only": {"put_ms": round(put_ms, 4), "get_ms": round(get_ms, 4),
                            "put_GBps_per_gpu": round(step_bytes / put_ms / 1e6, 1), "get_GBps_per_gpu": round(step_bytes / get_ms / 1e6, 1)},
            "roofline": {"per_gpu_GBps": round(per_gpu, 1), "denominator_GBps": roof,
                         "fraction": round(per_gpu / roof, 3),
                         "note": "N=1: measured HBM copy peak / 2 (payload read+written); N>=2: every payload byte crosses NVLink once; a rank's egress carries its own puts and its neighbour's gets, so payload/GPU is bound by the measured 770 GB/s per direction"},
            "e2e": {"value": round(e2e_value, 2), "unit": "GB/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": args.e2e_steps, "ms_per_step": round(e2e_ms / args.e2e_steps, 3),
                    "mode": e2e_mode,
                    "staged_ms_per_step": round(e2e_staged_ms / args.e2e_steps, 3), "zero_copy_ms_per_step": round(e2e_zc_ms / args.e2e_steps, 3),
                    "note": "every step: the payload comes from pinned host memory -- staged: cudaMemcpyAsync to HBM, then batch_put_device; zero_copy: "
                            "batch_put_device is handed the pinned host pointers and the fused kernel pulls them over PCIe itself -- then "
                            "batch_get_device through the public client API, D2H of 4 KiB of every returned object (verified on the host) plus "
                            "digests/status; value = the faster of the two modes (both reported)"},
            "gpu_launches": launches,
            "host_phase_mean_us": phases,
            "clocks": clocks,
            "comparators": comparators,
            "baseline_note": "vs_baseline divides by the reference's only throughput figure, an unsourced '~233 MB/sec' config comment (BASELINE.md section 1)",
        }
        print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
