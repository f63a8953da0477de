#!/usr/bin/env python3
"""Smallest end-to-end use of the store on one machine (no GPU needed): an in-process Keystone, two DRAM workers
reached over loopback TCP, and a client doing put / get / batch ops with per-shard checksums and replication.
(The multi-process equivalent is scripts/start_cluster.sh + bin/bb-cli.)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blackbird_b200 import _bb  # noqa: E402
from blackbird_b200.parallel import LocalCluster  # noqa: E402


def main():
    with LocalCluster("quickstart", n_workers=2, pool_bytes=64 << 20) as cluster:
        client = cluster.client(node_id="node-0")
        cfg = _bb.WorkerConfig(replication_factor=2, max_workers_per_copy=1, ttl_ms=60_000, checksum=_bb.ChecksumAlgo.CRC32C)
        blob = os.urandom(1 << 20)
        assert client.put("hello", blob, cfg) == _bb.ErrorCode.OK
        assert client.get("hello") == blob
        copies = client.get_workers("hello")
        print(f"'hello': {len(copies)} copies on {[c.shards[0].worker_id for c in copies]}, crc32c={copies[0].shards[0].checksum:#x}")
        keys = [f"batch/{i}" for i in range(8)]
        assert all(e == _bb.ErrorCode.OK for e in client.batch_put(keys, [os.urandom(4096) for _ in keys], cfg))
        print("exists:", client.batch_exists(keys[:3] + ["missing"]))
        stats = client.cluster_stats()
        print(f"cluster: {stats.total_workers} workers, {stats.total_objects} objects, {stats.used_capacity} bytes used")
        assert all(e == _bb.ErrorCode.OK for e in client.batch_remove(keys + ["hello"]))
        print("metrics sample:", [ln for ln in cluster.keystone.metrics_text().splitlines() if ln.startswith("bb_put_start_total")])
    print("quickstart OK")


if __name__ == "__main__":
    main()
