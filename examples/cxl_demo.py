#!/usr/bin/env python3
"""CXL tier walk-through (the reference's examples/cxl_example.cpp, which is not even in its CMake): parse
configs/cxl_worker.yaml, show the transport block and how its fall-backs resolve on this machine, bring the worker's
pools up (DAX devices that cannot be opened fall back to an anonymous placeholder mapping) and run the two-phase
shard life cycle on the CXL backend."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blackbird_b200 import _bb  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    cfg = _bb.WorkerServiceConfig.from_yaml(os.path.join(ROOT, "configs", "cxl_worker.yaml"))
    t = cfg.transport
    print(f"worker {cfg.worker_id}: {len(cfg.storage_pools)} pools")
    print(f"transport: {_bb.cxl_interconnect_name(t.interconnect_type)} / {_bb.cxl_protocol_name(t.transport_protocol)}, "
          f"queue_depth={t.queue_depth}, multipath={t.enable_multipath}, fallbacks={t.fallback_transports}")
    print("advertised interconnects here (no CXL device, no GPU):", t.resolve_interconnects(False, False))
    print("advertised interconnects on a CXL + GPU box        :", t.resolve_interconnects(True, True))
    for p in cfg.storage_pools:
        extra = f" dax={p.cxl.dax_device} numa={p.numa_node} interleave={p.cxl.interleave_granularity}" if p.cxl.dax_device else ""
        print(f"  pool {p.pool_id:22s} {str(p.storage_class).split('.')[-1]:18s} {p.size_bytes >> 20:6d} MiB{extra}")
    print("tier policy for a 200 MB object:", _bb.tier_classes_for_size(cfg.preferred_tiers, 200 * 10**6))
    backend = _bb.create_storage_backend(_bb.StorageClass.CXL_MEMORY, 8 << 20, "/dev/dax0.0")
    assert backend.initialize() == _bb.ErrorCode.OK
    tok = backend.reserve_shard(100_000)
    print(f"reserved {tok.size} bytes at {tok.remote_addr:#x} (cache-line aligned), dax={backend.is_dax}")
    assert backend.commit_shard(tok) == _bb.ErrorCode.OK
    off = tok.remote_addr - backend.get_base_address()
    backend.write(off, b"cxl!" * 16)
    assert backend.read(off, 64) == b"cxl!" * 16
    print("stats:", backend.get_stats().used_capacity, "bytes used;", "region id of the shard:", backend.region_id(off))
    assert backend.free_shard(tok.remote_addr, tok.size) == _bb.ErrorCode.OK
    backend.shutdown()
    print("cxl demo OK")


if __name__ == "__main__":
    main()
