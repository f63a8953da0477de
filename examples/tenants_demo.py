#!/usr/bin/env python3
"""Tenants on one machine (no GPU needed): two applications share a store without sharing keys or capacity.

`trainer` writes checkpoints under ckpt/ within a 4 MiB budget; `serving` may read them and keeps its own kv/ prefix.
Keys outside a tenant's grants are ACCESS_DENIED, a put beyond its budget is QUOTA_EXCEEDED before anything is allocated,
and the budget comes back when objects go.  (Multi-process: `tenants_file:` in the server YAMLs or BB_TENANTS_FILE, and
`bb-cli --tenant NAME --tenant-secret S ...`; see docs/OPERATIONS.md "Tenants".)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blackbird_b200 import _bb  # noqa: E402
from blackbird_b200.parallel import LocalCluster  # noqa: E402

TABLE = """
tenants:
  - {name: trainer, secret: "t-secret", write: ["ckpt/"], quota_bytes: 4MB}
  - {name: serving, secret: "s-secret", read: ["ckpt/"], write: ["kv/"], max_objects: 100}
"""


def main():
    _bb.load_tenants_text(TABLE)
    try:
        with LocalCluster("tenants-demo", n_workers=2, pool_bytes=64 << 20) as cluster:
            cfg = _bb.WorkerConfig(replication_factor=1, max_workers_per_copy=1, checksum=_bb.ChecksumAlgo.CRC32C)
            shard = os.urandom(1 << 20)

            trainer = cluster.client(tenant="trainer", tenant_secret="t-secret")
            for step in range(4):
                assert trainer.put(f"ckpt/step{step}", shard, cfg) == _bb.ErrorCode.OK
            print("trainer: 4 x 1 MiB under ckpt/ ->", [u for u in trainer.keystone().tenant_usage()])  # its own line only
            print("trainer: 5th MiB              ->", trainer.put("ckpt/step4", shard, cfg).name)       # QUOTA_EXCEEDED
            print("trainer: put kv/x             ->", trainer.put("kv/x", b"nope", cfg).name)           # ACCESS_DENIED
            assert trainer.remove("ckpt/step0") == _bb.ErrorCode.OK
            assert trainer.put("ckpt/step4", shard, cfg) == _bb.ErrorCode.OK  # the budget came back with the removal

            serving = cluster.client(tenant="serving", tenant_secret="s-secret")
            assert serving.get("ckpt/step4") == shard  # a read grant
            print("serving: remove ckpt/step4    ->", serving.remove("ckpt/step4").name)                # ACCESS_DENIED
            assert serving.put("kv/session-1", b"cache", cfg) == _bb.ErrorCode.OK
            try:
                trainer_view = cluster.client(tenant="trainer", tenant_secret="t-secret")
                trainer_view.get("kv/session-1")
                raise AssertionError("trainer read serving's prefix")
            except _bb.BlackbirdError as e:
                print("trainer: get kv/session-1     ->", e.code.name)                                  # ACCESS_DENIED

            _bb.set_client_tenant("", "")  # a member (the operator) sees everybody
            print("operator:", cluster.keystone.tenant_usage())
    finally:
        _bb.set_client_tenant("", "")
        _bb.set_tenants([])
    print("tenants demo OK")


if __name__ == "__main__":
    main()
