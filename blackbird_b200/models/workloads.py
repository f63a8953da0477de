"""The five workload configurations of BASELINE.json, as reusable functions.

1. plumbing_cpu            in-process Keystone + 2 workers over loopback TCP, 1 KB put/get/exists/remove
2. throughput_sweep /      GPU tier, batched put/get vs object size 256 B - 256 MB (GB/s, p50/p99)
   latency_sweep
3. replicated_put_verify   replication = 3 batched put (single-read fan-out), digest verified on get from every replica
4. tier_spill              GPU -> DRAM -> NVMe demotion under memory pressure with TTL + soft-pin
5. feature_store_fanout    rank 0 puts 128 x 1 MB activation shards, every rank batch-gets all of them
"""
from __future__ import annotations

import os
import statistics
import time

from .. import _bb

OK = None


def _ok(ecs):
    return all(e == _bb.ErrorCode.OK for e in ecs)


def plumbing_cpu(n_objects: int = 1000, size: int = 1024) -> dict:
    from ..parallel import LocalCluster

    with LocalCluster("plumbing", n_workers=2, pool_bytes=64 << 20) as c:
        cl = c.client()
        cfg = _bb.WorkerConfig(replication_factor=1, max_workers_per_copy=1)
        blobs = [os.urandom(size) for _ in range(16)]
        t0 = time.perf_counter()
        for i in range(n_objects):
            assert cl.put(f"k{i}", blobs[i % 16], cfg) == _bb.ErrorCode.OK
        t1 = time.perf_counter()
        for i in range(n_objects):
            assert cl.get(f"k{i}") == blobs[i % 16]
        t2 = time.perf_counter()
        assert all(cl.object_exists(f"k{i}") for i in range(0, n_objects, 10))
        t3 = time.perf_counter()
        for i in range(n_objects):
            assert cl.remove(f"k{i}") == _bb.ErrorCode.OK
        t4 = time.perf_counter()
        return {"objects": n_objects, "size": size, "put_per_s": n_objects / (t1 - t0), "get_per_s": n_objects / (t2 - t1),
                "exists_per_s": (n_objects // 10) / (t3 - t2), "remove_per_s": n_objects / (t4 - t3)}


def _pct(v, q):
    v = sorted(v)
    return v[min(len(v) - 1, int(q * len(v)))]


def throughput_sweep(cluster, sizes, target_node: str, batch_bytes: int = 1 << 30, max_batch: int = 4096, iters: int = 5,
                     algo=None) -> list:
    """Batched put + get GB/s per object size on the GPU tier (device time of the fused kernels and
    wall time of the whole client call including the Keystone round trips)."""
    import torch

    algo = algo or _bb.ChecksumAlgo.BBH64
    dev = torch.device("cuda", cluster.local_rank)
    stream = torch.cuda.current_stream().cuda_stream
    cfg = _bb.WorkerConfig(replication_factor=1, max_workers_per_copy=1, preferred_node=target_node, ttl_ms=0, checksum=algo,
                           preferred_classes=[_bb.StorageClass.RAM_GPU])
    rows = []
    for size in sizes:
        stride = (size + 255) // 256 * 256
        nobj = max(1, min(max_batch, batch_bytes // max(stride, 1)))
        src = torch.empty(nobj * stride, dtype=torch.uint8, device=dev)
        _bb.random_fill(src.data_ptr(), nobj * stride, size, stream)
        out = torch.zeros_like(src)
        sp = [src.data_ptr() + i * stride for i in range(nobj)]
        op = [out.data_ptr() + i * stride for i in range(nobj)]
        put_dev, get_dev, put_wall, get_wall = [], [], [], []
        ph0 = None
        for it in range(iters + 1):
            if it == 1:
                ph0 = cluster.client.phase_summary()
            keys = [f"sw/{cluster.rank}/{size}/{it}/{j}" for j in range(nobj)]
            t0 = time.perf_counter()
            m0 = cluster.fabric.total_device_ms
            assert _ok(cluster.client.batch_put_device(keys, sp, [size] * nobj, cfg, stream))
            t1 = time.perf_counter()
            m1 = cluster.fabric.total_device_ms
            ecs, _ = cluster.client.batch_get_device(keys, op, [stride] * nobj, stream)
            t2 = time.perf_counter()
            assert _ok(ecs)
            pd, gd = m1 - m0, cluster.fabric.total_device_ms - m1
            cluster.client.batch_remove(keys)
            if it:  # first iteration is warm-up
                put_dev.append(pd), get_dev.append(gd), put_wall.append((t1 - t0) * 1e3), get_wall.append((t2 - t1) * 1e3)
        torch.cuda.synchronize()
        assert torch.equal(src.view(nobj, stride)[:, :size], out.view(nobj, stride)[:, :size])
        total = nobj * size
        # where the client call spends its time: Keystone round trips vs descriptor build + launch vs waiting for the kernel
        ph1 = cluster.client.phase_summary()
        phases = {}
        for name, v in ph1.items():
            if name.startswith("phase_"):
                d = v[1] - (ph0.get(name, [0, 0])[1] if ph0 else 0)
                phases[name[len("phase_"):-len("_us")] + "_ms_per_batch"] = round(d / 1e3 / max(1, iters), 4)
        rows.append({"size": size, "batch": nobj, "put_GBps_kernel": total / min(put_dev) / 1e6, "get_GBps_kernel": total / min(get_dev) / 1e6,
                     "put_GBps_client": total / statistics.median(put_wall) / 1e6, "get_GBps_client": total / statistics.median(get_wall) / 1e6,
                     "put_batch_ms_p50": statistics.median(put_wall), "get_batch_ms_p50": statistics.median(get_wall), "phases": phases})
        del src, out
    return rows


def latency_sweep(cluster, sizes, target_node: str, iters: int = 200, algo=None) -> list:
    """Single-object put / get latency (p50 / p99, microseconds) per object size: the whole client call
    (Keystone round trips + descriptor upload + one fused launch + digest read-back)."""
    import torch

    algo = algo or _bb.ChecksumAlgo.BBH64
    dev = torch.device("cuda", cluster.local_rank)
    stream = torch.cuda.current_stream().cuda_stream
    cfg = _bb.WorkerConfig(replication_factor=1, max_workers_per_copy=1, preferred_node=target_node, ttl_ms=0, checksum=algo,
                           preferred_classes=[_bb.StorageClass.RAM_GPU])
    rows = []
    for size in sizes:
        n_it = iters if size <= (1 << 20) else max(20, iters // 10)
        stride = (size + 255) // 256 * 256
        src = torch.empty(stride, dtype=torch.uint8, device=dev)
        _bb.random_fill(src.data_ptr(), stride, size + 1, stream)
        out = torch.zeros_like(src)
        put_us, get_us, put_dev, get_dev = [], [], [], []
        fab = cluster.fabric
        paths = {"put_mailbox": 0, "put_launches": 0, "get_mailbox": 0, "get_launches": 0}  # which engine path served the calls
        for it in range(n_it + 5):
            key = [f"lat/{cluster.rank}/{size}/{it}"]
            c0 = (fab.mailbox_requests, fab.launches)
            t0 = time.perf_counter()
            ecs = cluster.client.batch_put_device(key, [src.data_ptr()], [size], cfg, stream)
            t1 = time.perf_counter()
            c1 = (fab.mailbox_requests, fab.launches)
            pd = cluster.fabric.last_device_ms
            ecs2, _ = cluster.client.batch_get_device(key, [out.data_ptr()], [stride], stream)
            t2 = time.perf_counter()
            c2 = (fab.mailbox_requests, fab.launches)
            gd = cluster.fabric.last_device_ms
            if it >= 5:
                paths["put_mailbox"] += c1[0] - c0[0]
                paths["put_launches"] += c1[1] - c0[1]
                paths["get_mailbox"] += c2[0] - c1[0]
                paths["get_launches"] += c2[1] - c1[1]
            cluster.client.batch_remove(key)
            assert _ok(ecs) and _ok(ecs2)
            if it >= 5:
                put_us.append((t1 - t0) * 1e6), get_us.append((t2 - t1) * 1e6), put_dev.append(pd * 1e3), get_dev.append(gd * 1e3)
        rows.append({"size": size, "put_p50_us": _pct(put_us, 0.5), "put_p99_us": _pct(put_us, 0.99), "get_p50_us": _pct(get_us, 0.5),
                     "get_p99_us": _pct(get_us, 0.99), "put_kernel_p50_us": _pct(put_dev, 0.5), "get_kernel_p50_us": _pct(get_dev, 0.5),
                     "engine_paths": paths})
    return rows


def replicated_put_verify(cluster, replication: int = 3, nobj: int = 32, size: int = 16 << 20, iters: int = 3, symmetric: bool = False) -> dict:
    """Config 3: every object gets `replication` copies on distinct GPUs; the put kernel reads the source once
    and fans out (TMA stores to every replica); the get verifies the digest against every replica in turn."""
    import torch

    dev = torch.device("cuda", cluster.local_rank)
    stream = torch.cuda.current_stream().cuda_stream
    # symmetric = the replicas share one offset in an NVLS replica arena: the put is ONE multimem.st stream per object
    # (the switch replicates); otherwise the kernel issues one TMA store per replica (single HBM read either way)
    cfg = _bb.WorkerConfig(replication_factor=replication, max_workers_per_copy=1, ttl_ms=0, preferred_classes=[_bb.StorageClass.RAM_GPU],
                           checksum=_bb.ChecksumAlgo.CRC32C, symmetric_replicas=symmetric)
    src = torch.empty(nobj * size, dtype=torch.uint8, device=dev)
    _bb.random_fill(src.data_ptr(), nobj * size, 33 + cluster.rank, stream)
    out = torch.zeros_like(src)
    sp = [src.data_ptr() + i * size for i in range(nobj)]
    op = [out.data_ptr() + i * size for i in range(nobj)]
    put_ms, get_ms = [], []
    copies_seen = set()
    for it in range(iters):
        keys = [f"rep{int(symmetric)}/{cluster.rank}/{it}/{j}" for j in range(nobj)]
        cluster.host_barrier()  # all ranks put at once (every GPU's ingress carries R x payload)
        m0 = cluster.fabric.total_device_ms
        assert _ok(cluster.client.batch_put_device(keys, sp, [size] * nobj, cfg, stream))
        put_ms.append(cluster.fabric.total_device_ms - m0)
        placed = cluster.client.get_workers(keys[0])
        assert len(placed) == replication and len({c.shards[0].worker_id for c in placed}) == replication
        assert len({c.shards[0].checksum for c in placed}) == 1
        copies_seen |= {c.shards[0].worker_id for c in placed}
        out.zero_()
        m0 = cluster.fabric.total_device_ms
        ecs, _ = cluster.client.batch_get_device(keys, op, [size] * nobj, stream)
        assert _ok(ecs)
        get_ms.append(cluster.fabric.total_device_ms - m0)
        torch.cuda.synchronize()
        assert torch.equal(src, out)
        cluster.client.batch_remove(keys)
    total = nobj * size
    return {"replication": replication, "objects": nobj, "size": size, "put_payload_GBps": total / min(put_ms) / 1e6,
            "put_wire_GBps": total * replication / min(put_ms) / 1e6, "get_GBps": total / min(get_ms) / 1e6, "replica_workers": sorted(copies_seen)}


def tier_spill(cluster, nobj: int = 10, size: int = 6 << 20) -> dict:
    """Config 4: fill the GPU tier beyond the watermark, let the Keystone demote LRU objects down the ladder
    (GPU -> DRAM -> NVMe) through the workers' D_COPY path, read everything back through the device API and
    compare bytes; a soft-pinned object must stay in HBM.  `cluster` is a single-rank GpuRankCluster created with
    dram_bytes / nvme_bytes / high_watermark < 1 (tests/test_gpu_stack.py)."""
    import torch

    dev = torch.device("cuda", cluster.local_rank)
    stream = torch.cuda.current_stream().cuda_stream
    gpu = dict(replication_factor=1, max_workers_per_copy=1, ttl_ms=0, preferred_classes=[_bb.StorageClass.RAM_GPU])
    src = torch.empty((nobj + 1) * size, dtype=torch.uint8, device=dev)
    _bb.random_fill(src.data_ptr(), src.numel(), 4242, stream)
    keys = [f"spill/{i}" for i in range(nobj)]
    ks = cluster.keystone
    t0 = time.perf_counter()
    assert _ok(cluster.client.batch_put_device(["spill/pinned"], [src.data_ptr() + nobj * size], [size],
                                               _bb.WorkerConfig(enable_soft_pin=True, **gpu), stream))
    demoted_rounds = 0
    for i, k in enumerate(keys):  # one by one: every put may push the tier over the watermark
        ecs = cluster.client.batch_put_device([k], [src.data_ptr() + i * size], [size], _bb.WorkerConfig(**gpu), stream)
        assert _ok(ecs), [str(e) for e in ecs]
        if ks.tier_utilization(_bb.StorageClass.RAM_GPU) > 0.5 or ks.tier_utilization(_bb.StorageClass.RAM_CPU) > 0.5:
            demoted_rounds += 1 if ks.run_eviction_once() else 0
    fill_s = time.perf_counter() - t0
    tiers = {k: cluster.client.get_workers(k)[0].shards[0].storage_class for k in keys + ["spill/pinned"]}
    out = torch.zeros_like(src)
    all_keys = keys + ["spill/pinned"]
    t0 = time.perf_counter()
    ecs, sizes = cluster.client.batch_get_device(all_keys, [out.data_ptr() + i * size for i in range(nobj + 1)], [size] * (nobj + 1), stream)
    torch.cuda.synchronize()
    read_s = time.perf_counter() - t0
    assert _ok(ecs), [str(e) for e in ecs]
    verified = sum(int(torch.equal(src[i * size:(i + 1) * size], out[i * size:(i + 1) * size])) for i in range(nobj + 1))
    text = ks.metrics_text()
    gpu_util_after = ks.tier_utilization(_bb.StorageClass.RAM_GPU)
    # promotion: bring the coldest object back into HBM (DRAM/NVMe -> GPU) and read it through the fused get
    promoted = False
    cold = next((k for k in keys if tiers[k] != _bb.StorageClass.RAM_GPU), None)
    if cold is not None and ks.tier_utilization(_bb.StorageClass.RAM_GPU) + size / max(1, cluster.worker.backend(f"hbm{cluster.rank}").get_total_capacity()) < 1.0:
        i = keys.index(cold)
        assert cluster.client.migrate(cold, _bb.StorageClass.RAM_GPU) == _bb.ErrorCode.OK
        probe = torch.zeros(size, dtype=torch.uint8, device=dev)
        ecs, _ = cluster.client.batch_get_device([cold], [probe.data_ptr()], [size], stream)
        torch.cuda.synchronize()
        promoted = _ok(ecs) and cluster.client.get_workers(cold)[0].shards[0].storage_class == _bb.StorageClass.RAM_GPU and \
            torch.equal(probe, src[i * size:(i + 1) * size])
    hbm = cluster.worker.backend(f"hbm{cluster.rank}")
    return {
        "fused_tier_moves": hbm.device_copies, "promoted_back": promoted,
        "objects": nobj + 1, "size": size, "verified": verified, "eviction_rounds": demoted_rounds,
        "demoted_to_dram": sum(1 for t in tiers.values() if t == _bb.StorageClass.RAM_CPU),
        "demoted_to_nvme": sum(1 for t in tiers.values() if t == _bb.StorageClass.NVME),
        "still_in_hbm": sum(1 for t in tiers.values() if t == _bb.StorageClass.RAM_GPU),
        "pinned_tier": str(tiers["spill/pinned"]).split(".")[-1],
        "gpu_util_after": gpu_util_after,
        "dropped": "bb_evictions_total" in text, "fill_s": fill_s, "read_back_s": read_s,
    }


def tier_spill_perf(cluster, nobj: int = 48, size: int = 64 << 20) -> dict:
    """Config 4 with numbers: `nobj` objects of `size` bytes are put into an HBM slab that holds about a third of them; the
    Keystone's watermark eviction demotes LRU objects GPU -> DRAM (one fused launch per shard, digest carried) and, when
    the DRAM tier passes its watermark too, DRAM -> NVMe (io_uring).  Reported: demotion GB/s (bytes moved down / time in
    the movers), the foreground put latency with and without a demotion round in between (the stall a writer sees), the
    final tier census, and a byte-exact read-back of everything through the device API (NVMe / DRAM objects come back
    through their own paths).  A soft-pinned object must still be in HBM at the end."""
    import torch

    dev = torch.device("cuda", cluster.local_rank)
    stream = torch.cuda.current_stream().cuda_stream
    gpu = dict(replication_factor=1, max_workers_per_copy=1, ttl_ms=0, preferred_classes=[_bb.StorageClass.RAM_GPU], checksum=_bb.ChecksumAlgo.CRC32C)
    ks = cluster.keystone
    nsrc = 8  # distinct payloads (the source buffer is reused round-robin; objects are identified by key)
    src = torch.empty(nsrc * size, dtype=torch.uint8, device=dev)
    _bb.random_fill(src.data_ptr(), src.numel(), 4242, stream)
    torch.cuda.synchronize()
    assert _ok(cluster.client.batch_put_device(["spill/pinned"], [src.data_ptr()], [size], _bb.WorkerConfig(enable_soft_pin=True, **gpu), stream))
    put_us, put_us_after_evict, evict_s, moved_bytes = [], [], 0.0, 0
    keys = [f"spill/{i}" for i in range(nobj)]
    classes = (_bb.StorageClass.RAM_GPU, _bb.StorageClass.RAM_CPU)
    t_fill0 = time.perf_counter()
    for i, k in enumerate(keys):
        evicted = False
        if any(ks.tier_utilization(c) > 0.55 for c in classes):
            t0 = time.perf_counter()
            n = ks.run_eviction_once()
            evict_s += time.perf_counter() - t0
            evicted = n > 0
            moved_bytes += n * size
        t0 = time.perf_counter()
        ecs = cluster.client.batch_put_device([k], [src.data_ptr() + (i % nsrc) * size], [size], _bb.WorkerConfig(**gpu), stream)
        dt = (time.perf_counter() - t0) * 1e6
        assert _ok(ecs), [str(e) for e in ecs]
        (put_us_after_evict if evicted else put_us).append(dt)
    fill_s = time.perf_counter() - t_fill0
    tiers = {k: cluster.client.get_workers(k)[0].shards[0].storage_class for k in keys + ["spill/pinned"]}
    # read everything back, HBM or not, and compare
    out = torch.zeros(size, dtype=torch.uint8, device=dev)
    verified, read_s = 0, 0.0
    for i, k in enumerate(keys):
        t0 = time.perf_counter()
        ecs, _ = cluster.client.batch_get_device([k], [out.data_ptr()], [size], stream)
        torch.cuda.synchronize()
        read_s += time.perf_counter() - t0
        verified += int(_ok(ecs) and torch.equal(out, src[(i % nsrc) * size:(i % nsrc + 1) * size]))
    census = {str(c).split(".")[-1]: sum(1 for t in tiers.values() if t == c) for c in set(tiers.values())}
    text = ks.metrics_text()
    demotions = 0
    for ln in text.splitlines():
        if ln.startswith("bb_demotions_total"):
            demotions = int(float(ln.split()[-1]))
    hbm = cluster.worker.backend(f"hbm{cluster.rank}")
    return {"objects": nobj, "size": size, "fill_s": round(fill_s, 3), "time_in_movers_s": round(evict_s, 3), "demotions": demotions,
            "demotion_GBps": round(demotions * size / evict_s / 1e9, 2) if evict_s > 0 else 0.0,
            "put_p50_us_no_demotion": round(_pct(put_us, 0.5), 1) if put_us else None,
            "put_p50_us_after_a_demotion_round": round(_pct(put_us_after_evict, 0.5), 1) if put_us_after_evict else None,
            "foreground_stall_ms_per_demotion_round_p50": None if not put_us_after_evict else round(evict_s / max(1, len(put_us_after_evict)) * 1e3, 2),
            "tier_census": census, "pinned_tier": str(tiers["spill/pinned"]).split(".")[-1], "verified": verified,
            "read_back_GBps": round(nobj * size / read_s / 1e9, 2), "fused_tier_moves": hbm.device_copies,
            "dropped_objects": sum(1 for ln in text.splitlines() if ln.startswith("bb_evictions_total") and float(ln.split()[-1]) > 0)}


def feature_store_fanout(cluster, nshards: int = 128, size: int = 1 << 20, iters: int = 5, replication: int = 1) -> dict:
    """Config 5: rank 0 puts `nshards` activation shards, then every rank batch-gets all of them (1 -> N read
    fan-out).  With replication > 1 the readers spread over the replicas (hash of reader + key)."""
    import torch

    dev = torch.device("cuda", cluster.local_rank)
    stream = torch.cuda.current_stream().cuda_stream
    cfg = _bb.WorkerConfig(replication_factor=replication, max_workers_per_copy=1, ttl_ms=0, preferred_classes=[_bb.StorageClass.RAM_GPU],
                           enable_locality_awareness=False)
    src = torch.empty(nshards * size, dtype=torch.uint8, device=dev)
    _bb.random_fill(src.data_ptr(), nshards * size, 777, stream)  # same seed on every rank: readers can verify locally
    out = torch.zeros_like(src)
    get_ms, put_ms = [], []
    for it in range(iters):
        keys = [f"feat/{it}/{j}" for j in range(nshards)]
        if cluster.rank == 0:
            t0 = time.perf_counter()
            assert _ok(cluster.client.batch_put_device(keys, [src.data_ptr() + j * size for j in range(nshards)], [size] * nshards, cfg, stream))
            put_ms.append((time.perf_counter() - t0) * 1e3)
        cluster.barrier()
        out.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ecs, _ = cluster.client.batch_get_device(keys, [out.data_ptr() + j * size for j in range(nshards)], [size] * nshards, stream)
        e1.record()
        torch.cuda.synchronize()
        assert _ok(ecs) and torch.equal(src, out)
        ms = cluster.max_over_ranks(e0.elapsed_time(e1))
        get_ms.append(ms)
        cluster.barrier()
        if cluster.rank == 0:
            cluster.client.batch_remove(keys)
        cluster.barrier()
    total = nshards * size
    return {"shards": nshards, "size": size, "readers": cluster.world, "replication": replication,
            "aggregate_get_GBps": total * cluster.world / min(get_ms) / 1e6, "per_reader_get_GBps": total / min(get_ms) / 1e6,
            "get_ms_best": min(get_ms), "rank0_put_ms_best": min(put_ms) if put_ms else None}
