"""keystone/: object state machine, batch ops, TTL GC, LRU + soft-pin eviction with tier demotion,
dead-worker handling + repair, sessions, metrics, HA fail-over with WAL recovery, and one
regression test per reference defect listed in SURVEY §2.8."""
import time

import pytest


def ks_cfg(bb, **kw):
    c = bb.KeystoneConfig()
    c.cluster_id = kw.pop("cluster_id", "t")
    c.listen_address = "127.0.0.1:0"
    c.http_metrics_port = "0"
    c.gc_interval_sec = 3600
    c.health_check_interval_sec = 3600
    c.service_refresh_interval_sec = 3600
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def mkpool(bb, pid, size, sc=None, node="node-a", worker=None):
    return bb.MemoryPool(pid, size, sc if sc is not None else bb.StorageClass.RAM_CPU, node, worker if worker is not None else "w-" + pid,
                         "127.0.0.1:12345", 0x1000000, "deadbeef")


@pytest.fixture
def ks(bb):
    k = bb.KeystoneService(ks_cfg(bb), None)
    assert k.initialize() == bb.ErrorCode.OK and k.start() == bb.ErrorCode.OK
    for i in range(4):
        assert k.register_memory_pool(mkpool(bb, f"p{i}", 1 << 20)) == bb.ErrorCode.OK
    yield k
    k.stop()


def cfg1(bb, **kw):
    kw.setdefault("replication_factor", 1)
    kw.setdefault("max_workers_per_copy", 1)
    return bb.WorkerConfig(**kw)


def test_put_state_machine_pending_complete(bb, ks):
    v0 = ks.get_view_version()
    copies = ks.put_start("k", 10000, cfg1(bb))
    assert len(copies) == 1 and copies[0].shards[0].length == 10000
    # bug #2: an in-flight put is not visible to readers
    assert ks.object_exists("k") is False
    with pytest.raises(bb.BlackbirdError) as e:
        ks.get_workers("k")
    assert e.value.code == bb.ErrorCode.OBJECT_NOT_READY
    assert ks.get_cluster_stats().pending_objects == 1 and ks.get_cluster_stats().total_objects == 0
    with pytest.raises(bb.BlackbirdError) as e:
        ks.put_start("k", 10, cfg1(bb))
    assert e.value.code == bb.ErrorCode.OBJECT_ALREADY_EXISTS
    assert ks.put_complete("k", [[0xABCDEF]]) == bb.ErrorCode.OK
    assert ks.put_complete("k") == bb.ErrorCode.OK  # idempotent
    got = ks.get_workers("k")
    assert got[0].shards[0].checksum == 0xABCDEF and got[0].shards[0].checksum_algo == bb.ChecksumAlgo.BBH64
    assert ks.object_exists("k") is True and ks.get_view_version() > v0
    assert ks.put_cancel("k") == bb.ErrorCode.INVALID_STATE  # cannot cancel a completed put
    assert ks.remove_object("k") == bb.ErrorCode.OK and ks.remove_object("k") == bb.ErrorCode.OBJECT_NOT_FOUND
    assert ks.get_cluster_stats().used_capacity == 0


def test_argument_validation(bb, ks):
    for key, size, cfg, code in [("", 10, cfg1(bb), bb.ErrorCode.INVALID_KEY),
                                 ("k", 10, cfg1(bb, replication_factor=0), bb.ErrorCode.INVALID_PARAMETERS),
                                 ("k", 10, cfg1(bb, replication_factor=9), bb.ErrorCode.VALUE_OUT_OF_RANGE),
                                 ("k", 64 << 20, cfg1(bb), bb.ErrorCode.INSUFFICIENT_SPACE)]:
        with pytest.raises(bb.BlackbirdError) as e:
            ks.put_start(key, size, cfg)
        assert e.value.code == code
    assert ks.put_complete("nope") == bb.ErrorCode.OBJECT_NOT_FOUND
    assert ks.put_complete("nope", [[1]]) == bb.ErrorCode.OBJECT_NOT_FOUND
    ks.put_start("c", 100, cfg1(bb))
    assert ks.put_complete("c", [[1], [2]]) == bb.ErrorCode.INVALID_PARAMETERS  # shape mismatch


def test_put_cancel_frees_ranges(bb, ks):
    ks.put_start("k", 500000, cfg1(bb))
    assert ks.get_cluster_stats().used_capacity >= 500000
    assert ks.put_cancel("k") == bb.ErrorCode.OK
    assert ks.get_cluster_stats().used_capacity == 0 and ks.put_cancel("k") == bb.ErrorCode.OBJECT_NOT_FOUND


def test_default_client_config_works_wpc1(bb, ks):
    """Bug #6: max_workers_per_copy == 1 returned NOT_IMPLEMENTED in the reference."""
    copies = ks.put_start("single", 4096, bb.WorkerConfig(replication_factor=1, max_workers_per_copy=1))
    assert len(copies[0].shards) == 1


def test_batch_ops_per_item_results(bb, ks):
    res = ks.batch_put_start(["a", "b", "", "a"], [1000, 2000, 10, 5], cfg1(bb))
    assert [r[0] for r in res] == [bb.ErrorCode.OK, bb.ErrorCode.OK, bb.ErrorCode.INVALID_KEY, bb.ErrorCode.OBJECT_ALREADY_EXISTS]
    assert ks.batch_put_complete(["a", "b", "zzz"]) == [bb.ErrorCode.OK, bb.ErrorCode.OK, bb.ErrorCode.OBJECT_NOT_FOUND]
    ex = ks.batch_object_exists(["a", "b", "zzz"])
    assert [(e, v) for e, v in ex] == [(bb.ErrorCode.OK, True), (bb.ErrorCode.OK, True), (bb.ErrorCode.OK, False)]
    gw = ks.batch_get_workers(["a", "zzz"])
    assert gw[0][0] == bb.ErrorCode.OK and len(gw[0][1]) == 1 and gw[1][0] == bb.ErrorCode.OBJECT_NOT_FOUND
    ks.batch_put_start(["c"], [10], cfg1(bb))
    assert ks.batch_put_cancel(["c", "a"]) == [bb.ErrorCode.OK, bb.ErrorCode.INVALID_STATE]
    assert ks.batch_remove_object(["a", "b", "q"]) == [bb.ErrorCode.OK, bb.ErrorCode.OK, bb.ErrorCode.OBJECT_NOT_FOUND]


def test_remove_all_objects_frees_allocator_ranges(bb, ks):
    """Bug #5: the reference cleared the map but leaked every range."""
    for i in range(10):
        ks.put_start(f"o{i}", 100000, cfg1(bb))
        ks.put_complete(f"o{i}")
    assert ks.get_cluster_stats().used_capacity > 0
    assert ks.remove_all_objects() == 10
    st = ks.get_cluster_stats()
    assert st.total_objects == 0 and st.used_capacity == 0 and ks.allocator_stats().total_objects == 0


def test_ttl_expiry_gc_and_inline_reclaim(bb, ks):
    ks.put_start("short", 1000, cfg1(bb, ttl_ms=60))
    ks.put_complete("short")
    ks.put_start("forever", 1000, cfg1(bb, ttl_ms=0))
    ks.put_complete("forever")
    time.sleep(0.15)
    assert ks.object_exists("short") is False
    with pytest.raises(bb.BlackbirdError) as e:
        ks.get_workers("short")
    assert e.value.code == bb.ErrorCode.OBJECT_NOT_FOUND
    # bug #4: an expired-but-unswept key could not be re-put until the GC thread ran
    assert len(ks.put_start("short", 2000, cfg1(bb, ttl_ms=60))) == 1
    ks.put_complete("short")
    time.sleep(0.15)
    assert ks.run_gc_once() == 1
    assert ks.object_exists("forever") is True and ks.allocator_stats().total_objects == 1
    assert "bb_expired_total 2" in ks.metrics_text()


def test_cluster_stats_live_accounting(bb, ks):
    """Bug #3: utilisation came from a registration snapshot that never changed; total_workers was never set."""
    st = ks.get_cluster_stats()
    assert st.total_workers == 4 and st.total_memory_pools == 4 and st.total_capacity == 4 << 20 and st.used_capacity == 0
    ks.put_start("x", 1 << 19, cfg1(bb))
    st = ks.get_cluster_stats()
    assert st.used_capacity == 1 << 19 and abs(st.avg_utilization - 0.125) < 1e-9
    assert sum(p.used for p in ks.get_memory_pools()) == 1 << 19


def test_eviction_lru_soft_pin_and_frees_ranges(bb):
    k = bb.KeystoneService(ks_cfg(bb, high_watermark=0.5, eviction_ratio=0.34), None)
    k.initialize(), k.start()
    k.register_memory_pool(mkpool(bb, "p", 1 << 20))
    for name, pin in [("old", False), ("pinned", True), ("mid", False), ("new", False)]:
        k.put_start(name, 150000, cfg1(bb, enable_soft_pin=pin))
        k.put_complete(name)
        time.sleep(0.01)
    k.get_workers("old")  # touch: "old" becomes most recently used -> LRU victim is "mid"
    assert k.tier_utilization(bb.StorageClass.RAM_CPU) > 0.5
    used0 = k.get_cluster_stats().used_capacity
    assert k.run_eviction_once() == 1
    assert k.object_exists("mid") is False and k.object_exists("pinned") and k.object_exists("old") and k.object_exists("new")
    assert k.get_cluster_stats().used_capacity < used0  # ranges really freed
    assert "bb_evictions_total 1" in k.metrics_text()
    k.stop()


def test_eviction_demotes_to_lower_tier_through_mover(bb):
    k = bb.KeystoneService(ks_cfg(bb, high_watermark=0.5, eviction_ratio=1.0), None)
    k.initialize(), k.start()
    k.register_memory_pool(mkpool(bb, "hbm", 1 << 20, bb.StorageClass.RAM_GPU))
    k.register_memory_pool(mkpool(bb, "dram", 8 << 20, bb.StorageClass.RAM_CPU))
    moved = []

    def mover(key, src, dst, algo):
        moved.append((key, src.shards[0].pool_id, dst.shards[0].pool_id))
        return (bb.ErrorCode.OK, [0x77] * len(dst.shards))

    k.set_copy_mover(mover)
    gpu_only = dict(preferred_classes=[bb.StorageClass.RAM_GPU])
    for n in ("a", "b"):
        k.put_start(n, 300000, cfg1(bb, **gpu_only))
        k.put_complete(n)
    k.put_start("pin", 100000, cfg1(bb, enable_soft_pin=True, **gpu_only))
    k.put_complete("pin")
    assert k.run_eviction_once() == 2
    assert sorted(m[0] for m in moved) == ["a", "b"] and all(m[1] == "hbm" and m[2] == "dram" for m in moved)
    for n in ("a", "b"):
        sh = k.get_workers(n)[0].shards[0]
        assert sh.pool_id == "dram" and sh.storage_class == bb.StorageClass.RAM_CPU and sh.checksum == 0x77
    assert k.get_workers("pin")[0].shards[0].pool_id == "hbm"  # soft-pinned objects stay
    assert k.tier_utilization(bb.StorageClass.RAM_GPU) < 0.2
    assert "bb_demotions_total 2" in k.metrics_text()
    assert k.remove_object("a") == bb.ErrorCode.OK
    assert k.get_cluster_stats().used_capacity == 300032 + 100096  # b (dram) + pin (hbm), 256 B aligned
    k.set_copy_mover(None)
    k.stop()


def test_dead_worker_invalidates_copies_and_repairs(bb):
    k = bb.KeystoneService(ks_cfg(bb), None)
    k.initialize(), k.start()
    for i in range(3):
        k.register_memory_pool(mkpool(bb, f"p{i}", 1 << 20, worker=f"w{i}"))
    k.put_start("r2", 5000, bb.WorkerConfig(replication_factor=2, max_workers_per_copy=1))
    k.put_complete("r2", [[5], [5]])
    k.put_start("r1", 5000, cfg1(bb))
    k.put_complete("r1")
    victim = k.get_workers("r1")[0].shards[0].worker_id
    r2_pools = {c.shards[0].worker_id for c in k.get_workers("r2")}
    k.handle_worker_death(victim)
    st = k.get_cluster_stats()
    assert st.total_workers == 2 and st.total_memory_pools == 2
    # the single-copy object on the dead worker is gone (the reference left a dangling placement)
    with pytest.raises(bb.BlackbirdError) as e:
        k.get_workers("r1")
    assert e.value.code == bb.ErrorCode.OBJECT_NOT_FOUND
    if victim in r2_pools:
        live = k.get_workers("r2")
        assert len(live) == 1 and live[0].shards[0].worker_id != victim  # survivor keeps serving
        k.set_copy_mover(lambda key, src, dst, algo: (bb.ErrorCode.OK, [5]))
        assert k.run_repair_once() == 1
        healed = k.get_workers("r2")
        assert len(healed) == 2 and len({c.shards[0].worker_id for c in healed}) == 2 and victim not in {c.shards[0].worker_id for c in healed}
        k.set_copy_mover(None)
    else:
        assert len(k.get_workers("r2")) == 2
    assert k.remove_worker("ghost") == bb.ErrorCode.INVALID_WORKER
    k.stop()


def test_client_sessions_ttl(bb):
    k = bb.KeystoneService(ks_cfg(bb, client_ttl_sec=1, health_check_interval_sec=1), None)
    k.initialize(), k.start()
    cid = k.client_register("node-x")
    assert cid.startswith("client-") and k.client_ping(cid) == k.get_view_version()
    assert k.get_cluster_stats().active_clients == 1
    time.sleep(2.3)  # health thread expires the silent session
    with pytest.raises(bb.BlackbirdError) as e:
        k.client_ping(cid)
    assert e.value.code == bb.ErrorCode.SESSION_EXPIRED
    k.stop()


def test_metrics_exposition_format(bb, ks):
    ks.put_start("m", 1000, cfg1(bb))
    ks.put_complete("m")
    ks.get_workers("m")
    text = ks.metrics_text()
    for needle in ["# TYPE bb_put_start_total counter", "bb_put_start_total 1", "bb_put_complete_total 1", "bb_get_workers_total 1",
                   "bb_objects 1", "bb_workers 4", "bb_capacity_bytes 4194304", "bb_is_leader 1", 'bb_tier_used_bytes{tier="RAM_CPU"} 1024',
                   "bb_put_start_latency_us_bucket{le=\"+Inf\"} 1", "bb_put_start_latency_us_count 1"]:
        assert needle in text, needle
    sj = ks.stats_json()
    assert sj["cluster"]["total_objects"] == 1 and len(sj["pools"]) == 4 and sj["is_leader"] is True


def test_concurrent_puts_across_shards(bb, ks):
    import threading

    errs = []

    def run(t):
        for i in range(200):
            key = f"t{t}/o{i}"
            try:
                ks.put_start(key, 512, cfg1(bb))
                assert ks.put_complete(key) == bb.ErrorCode.OK
                assert ks.object_exists(key)
                assert ks.remove_object(key) == bb.ErrorCode.OK
            except Exception as ex:  # noqa: BLE001
                errs.append(ex)

    ts = [threading.Thread(target=run, args=(t,)) for t in range(8)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs and ks.get_cluster_stats().used_capacity == 0 and ks.allocator_stats().total_objects == 0


def test_ha_leader_failover_recovers_objects_from_wal(bb):
    """Real leader election + metadata log (the reference has neither: §2.7, §5.4)."""
    store = bb.MemCoord()
    mk = lambda sid: bb.KeystoneService(ks_cfg(bb, cluster_id="ha", enable_ha=True, service_id=sid, service_registration_ttl_sec=4,
                                                service_refresh_interval_sec=1), bb.CoordService(store))
    a, b = mk("ks-a"), mk("ks-b")
    assert a.initialize() == bb.ErrorCode.OK and a.start() == bb.ErrorCode.OK
    assert b.initialize() == bb.ErrorCode.OK and b.start() == bb.ErrorCode.OK
    assert a.is_leader() and not b.is_leader()
    # a worker registers through the coordination store: both keystones learn about it
    pool = mkpool(bb, "p0", 1 << 20, worker="w0")
    store.put("/blackbird/clusters/ha/workers/w0", '{"worker_id":"w0","node_id":"n0"}')
    store.put("/blackbird/clusters/ha/workers/w0/memory_pools/p0", pool.to_json())
    store.flush_events()
    assert a.get_cluster_stats().total_memory_pools == 1 and b.get_cluster_stats().total_memory_pools == 1
    with pytest.raises(bb.BlackbirdError) as e:
        b.put_start("x", 10, cfg1(bb))
    assert e.value.code == bb.ErrorCode.NOT_LEADER  # standby refuses mutations
    placed = a.put_start("obj", 4096, cfg1(bb, ttl_ms=0))
    assert a.put_complete("obj", [[0x1234]]) == bb.ErrorCode.OK
    a.put_start("pending-only", 4096, cfg1(bb))  # never completed: not in the WAL
    # leader crashes (no clean resign): its lease expires, the standby takes over on its next campaign
    store.advance_time_ms(5000)
    deadline = time.time() + 5
    while not b.is_leader() and time.time() < deadline:
        time.sleep(0.05)
    assert b.is_leader()
    got = b.get_workers("obj")
    assert got[0].shards[0].checksum == 0x1234 and got[0].shards[0].offset == placed[0].shards[0].offset
    with pytest.raises(bb.BlackbirdError):
        b.get_workers("pending-only")
    # recovered extents are reserved: a new object must not overlap the recovered one
    fresh = b.put_start("fresh", 4096, cfg1(bb))
    assert fresh[0].shards[0].offset != placed[0].shards[0].offset
    assert store.get("/blackbird/elections/keystone-ha/leader") == b"ks-b"
    a.stop(), b.stop()


def test_rpc_client_with_both_endpoints_follows_the_leader(bb):
    """Two keystones behind their RPC services, one KeystoneRpcClient that knows both: the standby answers NOT_LEADER to
    object-scoped calls (reads too -- it has no object map yet) but serves cluster-scoped ones; the client lands on the
    leader, and after the leader's lease runs out it ends up on the new one with the recovered objects."""
    store = bb.MemCoord()
    cfgs = [ks_cfg(bb, cluster_id="ha3", enable_ha=True, service_id=sid, service_registration_ttl_sec=4, service_refresh_interval_sec=1)
            for sid in ("ks-a", "ks-b")]
    kss = [bb.KeystoneService(c, bb.CoordService(store)) for c in cfgs]
    for k in kss:
        assert k.initialize() == bb.ErrorCode.OK and k.start() == bb.ErrorCode.OK
    a, b = kss
    assert a.is_leader() and not b.is_leader()
    rpcs = [bb.RpcService(k, c) for k, c in zip(kss, cfgs)]
    for r in rpcs:
        assert r.start() == bb.ErrorCode.OK
    ep_a, ep_b = (f"127.0.0.1:{r.rpc_port}" for r in rpcs)
    store.put("/blackbird/clusters/ha3/workers/w0", '{"worker_id":"w0","node_id":"n0"}')
    store.put("/blackbird/clusters/ha3/workers/w0/memory_pools/p0", mkpool(bb, "p0", 1 << 20, worker="w0").to_json())
    store.flush_events()
    try:
        only_b = bb.KeystoneRpcClient()
        assert only_b.connect_any([ep_b]) == bb.ErrorCode.OK
        assert only_b.get_cluster_stats().total_memory_pools == 1  # cluster-scoped: any keystone answers
        for call in (lambda: only_b.object_exists("obj"), lambda: only_b.get_workers("obj"), lambda: only_b.put_start("obj", 10, cfg1(bb))):
            with pytest.raises(bb.BlackbirdError) as e:
                call()
            assert e.value.code == bb.ErrorCode.NOT_LEADER
        assert [e for e, _ in only_b.batch_object_exists(["x", "y"])] == [bb.ErrorCode.NOT_LEADER] * 2 and only_b.failovers() == 0
        assert [r[0] for r in only_b.batch_get_workers(["x"])] == [bb.ErrorCode.NOT_LEADER]

        both = bb.KeystoneRpcClient()
        assert both.connect_any([ep_b, ep_a]) == bb.ErrorCode.OK and both.active_endpoint() == ep_b
        both.put_start("obj", 4096, cfg1(bb, ttl_ms=0))
        assert both.active_endpoint() == ep_a and both.failovers() == 1
        assert both.put_complete("obj", [[0x77]]) == bb.ErrorCode.OK
        # the leader dies without resigning: its RPC port goes away and its lease expires
        rpcs[0].stop()
        a.stop()
        store.advance_time_ms(5000)
        both.set_failover_budget_ms(10000)
        got = both.get_workers("obj")  # transport error on a -> b answers NOT_LEADER until its campaign succeeds -> b serves
        assert got[0].shards[0].checksum == 0x77 and both.active_endpoint() == ep_b and b.is_leader()
        assert [v for _, v in both.batch_object_exists(["obj", "nope"])] == [True, False]
        # nobody left: the budget bounds the wait and the caller gets the transport error
        rpcs[1].stop()
        both.set_failover_budget_ms(300)
        t0 = time.time()
        with pytest.raises(bb.BlackbirdError) as e:
            both.object_exists("obj")
        assert e.value.code in (bb.ErrorCode.RPC_FAILED, bb.ErrorCode.CLIENT_DISCONNECTED) and time.time() - t0 < 5
    finally:
        for r in rpcs:
            r.stop()
        for k in kss:
            k.stop()


def test_size_based_tier_policy_for_puts_without_a_preferred_class(bb):
    """keystone.tier_policy (reference cxl_worker.yaml `allocation.preferred_tiers`, consumed by nothing there)."""
    from blackbird_b200.parallel import LocalCluster

    kc = bb.KeystoneConfig()
    kc.tier_policy = [bb.TierRule("RAM_CPU", 0, 64 << 10), bb.TierRule("CXL_MEMORY", (64 << 10) + 1)]
    with LocalCluster("tierpol", n_workers=0, keystone_cfg=kc) as c:
        c.add_worker("w0", "n0", [("dram", bb.StorageClass.RAM_CPU, 8 << 20, ""), ("cxl", bb.StorageClass.CXL_MEMORY, 8 << 20, "")])
        cl = c.client()
        cfg = bb.WorkerConfig(replication_factor=1, max_workers_per_copy=1)
        assert cl.put("small", b"x" * 1000, cfg) == bb.ErrorCode.OK
        assert cl.put("big", b"y" * (1 << 20), cfg) == bb.ErrorCode.OK
        assert cl.get_workers("small")[0].shards[0].storage_class == bb.StorageClass.RAM_CPU
        assert cl.get_workers("big")[0].shards[0].storage_class == bb.StorageClass.CXL_MEMORY
        # an explicit preference wins over the policy
        assert cl.put("big2", b"z" * (1 << 20), bb.WorkerConfig(replication_factor=1, max_workers_per_copy=1, preferred_classes=[bb.StorageClass.RAM_CPU])) == bb.ErrorCode.OK
        assert cl.get_workers("big2")[0].shards[0].storage_class == bb.StorageClass.RAM_CPU
        assert cl.get("big") == b"y" * (1 << 20)
        # a batch with two sizes = two runs: each follows the policy for its size
        names = [f"s{i}" for i in range(12)] + [f"b{i}" for i in range(12)]
        res = c.keystone.batch_put_start(names, [2000] * 12 + [128 << 10] * 12, cfg)
        assert all(r[0] == bb.ErrorCode.OK for r in res)
        assert {r[1][0].shards[0].storage_class for r in res[:12]} == {bb.StorageClass.RAM_CPU}
        assert {r[1][0].shards[0].storage_class for r in res[12:]} == {bb.StorageClass.CXL_MEMORY}


def test_batch_put_start_places_uniform_runs_together(bb, ks):
    """Runs of >= 8 items with one size and one policy take the run path (one ranking, one allocator call per chunk):
    same per-item results as the object-by-object path, spread over the pools that tie in the ranking, no overlap,
    and every extent comes back on remove."""
    cfg = cfg1(bb, ttl_ms=0)
    keys = [f"run/{i}" for i in range(64)]
    ks.put_start("run/7", 512, cfg)  # a live duplicate inside the run
    names = keys[:20] + [""] + keys[20:] + ["run/3"]  # an invalid key and a duplicate of an item of the same batch
    res = ks.batch_put_start(names, [4096] * len(names), cfg)
    codes = [r[0] for r in res]
    assert codes[20] == bb.ErrorCode.INVALID_KEY and codes[-1] == bb.ErrorCode.OBJECT_ALREADY_EXISTS
    assert codes[7] == bb.ErrorCode.OBJECT_ALREADY_EXISTS
    ok = [(n, r[1]) for n, r in zip(names, res) if r[0] == bb.ErrorCode.OK]
    assert len(ok) == 63
    seen, per_pool = set(), {}
    for name, copies in ok:
        assert len(copies) == 1 and len(copies[0].shards) == 1
        sh = copies[0].shards[0]
        assert sh.length == 4096 and sh.checksum_algo == cfg.checksum
        loc = (sh.pool_id, sh.location["remote_addr"])
        assert loc not in seen
        seen.add(loc)
        per_pool[sh.pool_id] = per_pool.get(sh.pool_id, 0) + 1
    assert len(per_pool) == 4 and min(per_pool.values()) >= 12 and max(per_pool.values()) <= 17  # four equal pools: dealt out evenly
    # extents of one pool do not overlap
    for pid in per_pool:
        addrs = sorted(a for p, a in seen if p == pid)
        assert all(b - a >= 4096 for a, b in zip(addrs, addrs[1:]))
    assert ks.get_cluster_stats().pending_objects == 64
    assert "bb_put_start_run_objects_total 63" in ks.metrics_text()  # the duplicates went through the per-object path
    done = [n for n, _ in ok]
    assert set(ks.batch_put_complete(done)) == {bb.ErrorCode.OK}
    assert ks.get_workers("run/5")[0].shards[0].length == 4096
    assert set(ks.batch_remove_object(done + ["run/7"])) == {bb.ErrorCode.OK}
    assert ks.get_cluster_stats().used_capacity == 0
    # the whole capacity is allocatable again as one extent per pool
    big = ks.batch_put_start([f"big/{i}" for i in range(4)], [1 << 20] * 4, cfg)
    assert [r[0] for r in big] == [bb.ErrorCode.OK] * 4


def test_batch_put_start_run_overflows_down_the_ranking_and_reports_the_rest(bb, ks):
    cfg = cfg1(bb, ttl_ms=0)
    n = 40  # 4 pools x 1 MiB hold 32 objects of 128 KiB
    res = ks.batch_put_start([f"o/{i}" for i in range(n)], [128 << 10] * n, cfg)
    codes = [r[0] for r in res]
    assert codes.count(bb.ErrorCode.OK) == 32 and codes.count(bb.ErrorCode.INSUFFICIENT_SPACE) == 8
    # striped or replicated policies keep the general path (two shards per copy here)
    assert set(ks.batch_remove_object([f"o/{i}" for i in range(n) if codes[i] == bb.ErrorCode.OK])) == {bb.ErrorCode.OK}
    wide = bb.WorkerConfig(replication_factor=1, max_workers_per_copy=2, ttl_ms=0)
    res = ks.batch_put_start([f"s/{i}" for i in range(16)], [64 << 10] * 16, wide)
    assert all(r[0] == bb.ErrorCode.OK and len(r[1][0].shards) == 2 for r in res)


def test_concurrent_symmetric_placements_do_not_starve_each_other(bb, monkeypatch):
    """Symmetric replicas (same offset on every replica, for NVLS multicast) are found by intersecting free lists and then
    claiming the offset pool by pool.  Eight writers doing that at once all aim at the same hole; they used to exhaust the
    retry budget (ALLOCATION_FAILED on the 8-GPU box).  Symmetric placements now go one at a time."""
    import threading

    from blackbird_b200.parallel import LocalCluster

    monkeypatch.setenv("BB_RPC_SHM", "0")  # one epoll pool thread per busy connection: the requests really overlap
    with LocalCluster(cluster_id="sym", n_workers=0) as c:
        G = bb.StorageClass.RAM_GPU
        for i in range(8):
            assert c.keystone.register_memory_pool(bb.MemoryPool(f"g{i}", 64 << 20, G, f"gpu{i}", f"w{i}", "127.0.0.1:1", 0, "00", i)) == bb.ErrorCode.OK
        cfg = bb.WorkerConfig(replication_factor=3, max_workers_per_copy=1, ttl_ms=0, symmetric_replicas=True, preferred_classes=[G])
        errors, lock = [], threading.Lock()

        def writer(t):
            api = bb.KeystoneRpcClient()
            assert api.connect("127.0.0.1", c.rpc.rpc_port, 3000) == bb.ErrorCode.OK
            for rnd in range(6):
                res = api.batch_put_start([f"s/{t}/{rnd}/{j}" for j in range(32)], [64 << 10] * 32, cfg)
                bad = [r[0] for r in res if r[0] != bb.ErrorCode.OK]
                offs = [{cp.shards[0].location["offset"] for cp in r[1]} for r in res if r[0] == bb.ErrorCode.OK]
                with lock:
                    errors.extend(bad)
                    errors.extend("offsets differ" for o in offs if len(o) != 1)

        ts = [threading.Thread(target=writer, args=(t,)) for t in range(8)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        assert errors == []
        assert c.keystone.get_cluster_stats().pending_objects == 8 * 6 * 32


def test_ha_failover_recovers_a_run_placed_batch_extent_by_extent(bb):
    """Objects placed as a run share one pool-allocator extent that was sliced; the metadata log records them one by one.
    After a fail-over the new leader re-reserves every slice (adopt): the survivors keep their offsets, removed ones left
    holes that are reusable, and new objects do not land on any recovered slice."""
    store = bb.MemCoord()
    mk = lambda sid: bb.KeystoneService(ks_cfg(bb, cluster_id="harun", enable_ha=True, service_id=sid, service_registration_ttl_sec=4,
                                                service_refresh_interval_sec=1), bb.CoordService(store))
    a, b = mk("ks-a"), mk("ks-b")
    assert a.initialize() == bb.ErrorCode.OK and a.start() == bb.ErrorCode.OK
    assert b.initialize() == bb.ErrorCode.OK and b.start() == bb.ErrorCode.OK
    assert a.is_leader()
    store.put("/blackbird/clusters/harun/workers/w0", '{"worker_id":"w0","node_id":"n0"}')
    store.put("/blackbird/clusters/harun/workers/w0/memory_pools/p0", mkpool(bb, "p0", 1 << 20, worker="w0").to_json())
    store.flush_events()
    keys = [f"run/{i:02d}" for i in range(32)]
    res = a.batch_put_start(keys, [8192] * 32, cfg1(bb, ttl_ms=0))
    assert all(r[0] == bb.ErrorCode.OK for r in res)
    offs = {k: r[1][0].shards[0].offset for k, r in zip(keys, res)}
    assert len(set(offs.values())) == 32
    assert set(a.batch_put_complete(keys)) == {bb.ErrorCode.OK}
    gone = keys[1::4]
    assert set(a.batch_remove_object(gone)) == {bb.ErrorCode.OK}  # holes inside the sliced extent
    store.advance_time_ms(5000)  # the leader's lease runs out without a resign
    deadline = time.time() + 5
    while not b.is_leader() and time.time() < deadline:
        time.sleep(0.05)
    assert b.is_leader()
    kept = [k for k in keys if k not in gone]
    for k in kept:
        assert b.get_workers(k)[0].shards[0].offset == offs[k]
    for k in gone:
        with pytest.raises(bb.BlackbirdError):
            b.get_workers(k)
    assert b.get_cluster_stats().used_capacity == len(kept) * 8192
    # fill the rest of the pool: every new object avoids the recovered slices, and the holes are usable
    more = b.batch_put_start([f"new/{i:03d}" for i in range(104)], [8192] * 104, cfg1(bb, ttl_ms=0))  # 128 slots - 24 kept
    new_offs = [r[1][0].shards[0].offset for r in more if r[0] == bb.ErrorCode.OK]
    assert len(new_offs) == 104 and not (set(new_offs) & {offs[k] for k in kept})
    assert {offs[k] for k in gone} <= set(new_offs)
    a.stop(), b.stop()


def test_concurrent_run_placements_and_removals_never_overlap(bb, monkeypatch):
    """Several clients place uniform batches (run path: sliced pool extents) and remove parts of them at the same time, over
    the RPC server's thread pool.  Afterwards no two live objects share a byte and the books balance."""
    import threading

    from blackbird_b200.parallel import LocalCluster

    monkeypatch.setenv("BB_RPC_SHM", "0")
    with LocalCluster(cluster_id="runrace", n_workers=0) as c:
        for i in range(3):
            assert c.keystone.register_memory_pool(mkpool(bb, f"p{i}", 8 << 20, worker=f"w{i}")) == bb.ErrorCode.OK
        cfg = bb.WorkerConfig(replication_factor=1, max_workers_per_copy=1, ttl_ms=0)
        live, lock, errors = {}, threading.Lock(), []

        def client(t):
            api = bb.KeystoneRpcClient()
            assert api.connect("127.0.0.1", c.rpc.rpc_port, 3000) == bb.ErrorCode.OK
            mine = {}
            for rnd in range(12):
                size = (1024, 4096, 12288)[(t + rnd) % 3]
                keys = [f"t{t}/{rnd}/{j}" for j in range(48)]
                for key, (ec, copies) in zip(keys, api.batch_put_start(keys, [size] * 48, cfg)):
                    if ec == bb.ErrorCode.OK:
                        sh = copies[0].shards[0]
                        mine[key] = (sh.pool_id, sh.location["remote_addr"], (size + 255) // 256 * 256)
                    elif ec != bb.ErrorCode.INSUFFICIENT_SPACE:
                        errors.append((key, ec))
                done = [k for k in keys if k in mine]
                if done and set(api.batch_put_complete(done)) != {bb.ErrorCode.OK}:
                    errors.append(("complete", t, rnd))
                victims = done[rnd % 2::2]
                if victims and set(api.batch_remove_object(victims)) != {bb.ErrorCode.OK}:
                    errors.append(("remove", t, rnd))
                for k in victims:
                    del mine[k]
            with lock:
                live.update(mine)

        ts = [threading.Thread(target=client, args=(t,)) for t in range(6)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        assert errors == []
        by_pool = {}
        for pool, addr, n in live.values():
            by_pool.setdefault(pool, []).append((addr, n))
        for pool, ext in by_pool.items():
            ext.sort()
            assert all(a + n <= b for (a, n), (b, _) in zip(ext, ext[1:])), f"overlap on {pool}"
        assert c.keystone.get_cluster_stats().used_capacity == sum(n for _, _, n in live.values())
        assert c.keystone.get_cluster_stats().total_objects == len(live)


def test_batch_replies_over_the_rpc_equal_the_in_process_ones_for_every_placement_shape(bb, tmp_path, monkeypatch):
    """Batch replies delta-encode placements (one-shard results that differ from the last fully written one only in position
    and digest travel as 17 bytes).  Whatever the mix -- memory / file / CXL locations, striped and replicated objects in
    between, errors, runs on alternating pools -- the client must rebuild exactly what the Keystone returned."""
    from blackbird_b200.parallel import LocalCluster

    monkeypatch.setenv("BB_RPC_SHM", "0")
    with LocalCluster("delta", n_workers=0) as c:
        c.add_worker("w0", "n0", [("dram0", bb.StorageClass.RAM_CPU, 8 << 20, ""), ("nvme0", bb.StorageClass.NVME, 8 << 20, str(tmp_path / "n0")),
                                  ("cxl0", bb.StorageClass.CXL_MEMORY, 8 << 20, "")])
        c.add_worker("w1", "n1", [("dram1", bb.StorageClass.RAM_CPU, 8 << 20, ""), ("nvme1", bb.StorageClass.NVME, 8 << 20, str(tmp_path / "n1"))])
        ks = c.keystone
        api = bb.KeystoneRpcClient()
        assert api.connect("127.0.0.1", c.rpc.rpc_port, 3000) == bb.ErrorCode.OK
        keys, n = [], 0

        def put(tier, size, repl=1, wpc=1, count=1):
            nonlocal n
            cfg = bb.WorkerConfig(replication_factor=repl, max_workers_per_copy=wpc, ttl_ms=0, preferred_classes=[tier], min_shard_size=256)
            names = [f"o/{n + i:03d}" for i in range(count)]
            n += count
            res = ks.batch_put_start(names, [size] * count, cfg)
            assert all(r[0] == bb.ErrorCode.OK for r in res), [r[0] for r in res]
            assert set(ks.batch_put_complete(names)) == {bb.ErrorCode.OK}
            keys.extend(names)

        R, N, X = bb.StorageClass.RAM_CPU, bb.StorageClass.NVME, bb.StorageClass.CXL_MEMORY
        put(R, 4096, count=12)            # a run: deltas
        put(N, 8192, count=9)             # file locations: deltas on another variant
        put(R, 6000, repl=2)              # replicated: written in full, resets nothing it should not
        put(R, 4096, count=3)
        put(X, 4096, count=8)             # CXL: region id derived from the offset
        put(R, 20000, wpc=2)              # striped
        put(N, 8192, count=2)
        put(R, 100)                       # different length: full again
        asked = keys[:5] + ["missing/1"] + keys[5:] + ["missing/2"]
        direct = ks.batch_get_workers(asked)
        remote = api.batch_get_workers(asked)
        assert len(direct) == len(remote) == len(asked)

        def flat(res):
            out = []
            for ec, copies in res:
                if ec != bb.ErrorCode.OK:
                    out.append((ec,))
                    continue
                out.append((ec, [(cp.copy_index, [(s.pool_id, s.worker_id, s.storage_class, s.length, s.checksum, s.checksum_algo, repr(s.location),
                                                    s.endpoint.ip, s.endpoint.port) for s in cp.shards]) for cp in copies]))
            return out

        assert flat(direct) == flat(remote)
        assert [r[0] for r in remote].count(bb.ErrorCode.OBJECT_NOT_FOUND) == 2
        # and the reply of a put_start batch (run placement on two pools in turn) comes back identical too
        cfg = bb.WorkerConfig(replication_factor=1, max_workers_per_copy=1, ttl_ms=0, preferred_classes=[R])
        names = [f"p/{i:03d}" for i in range(40)]
        placed = api.batch_put_start(names, [2048] * 40, cfg)
        assert all(r[0] == bb.ErrorCode.OK for r in placed)
        assert set(ks.batch_put_complete(names)) == {bb.ErrorCode.OK}
        assert flat(placed) == flat([(ec, [type("C", (), {"copy_index": cp.copy_index, "shards": cp.shards})() for cp in copies]) for ec, copies in ks.batch_get_workers(names)])
