"""coord/: revisioned KV, leases, watches, txn primitives, service registry, leader election,
and the TCP daemon (CoordServer / RemoteCoord) — the etcd replacement."""
import threading
import time

import pytest


def test_kv_prefix_revisions(bb):
    s = bb.MemCoord()
    r0 = s.revision()
    assert s.put("/a/1", "x") == bb.ErrorCode.OK and s.put("/a/2", "y") == bb.ErrorCode.OK and s.put("/b/1", "z") == bb.ErrorCode.OK
    assert s.get("/a/1") == b"x" and s.get("/missing") is None
    assert [k for k, *_ in s.get_with_prefix("/a/")] == ["/a/1", "/a/2"]
    assert s.revision() == r0 + 3
    assert s.delete("/a/1") == bb.ErrorCode.OK and s.delete("/a/1") == bb.ErrorCode.OK  # idempotent
    assert s.del_prefix("/a/") == 1 and s.key_count() == 1


def test_lease_expiry_deletes_keys_and_fires_watch(bb):
    s = bb.MemCoord()
    events = []
    s.watch_prefix("/hb/", lambda t, k, v, rev: events.append((t, k, v)))
    lease = s.grant_lease(10)
    assert s.put("/hb/w1", "alive", lease) == bb.ErrorCode.OK
    assert s.put("/hb/w2", "alive", 424242) == bb.ErrorCode.ETCD_LEASE_ERROR  # unknown lease
    s.advance_time_ms(5000)
    assert s.keep_alive(lease) == bb.ErrorCode.OK  # refreshed at t=5s -> expires at t=15s
    s.advance_time_ms(9000)
    assert s.get("/hb/w1") == b"alive" and 0 < s.lease_remaining_ms(lease) <= 1000
    s.advance_time_ms(2000)
    assert s.get("/hb/w1") is None
    assert s.keep_alive(lease) == bb.ErrorCode.ETCD_LEASE_ERROR
    assert events == [("PUT", "/hb/w1", b"alive"), ("DELETE", "/hb/w1", b"alive")]
    assert s.lease_count() == 0


def test_revoke_and_rebind_lease(bb):
    s = bb.MemCoord()
    l1, l2 = s.grant_lease(5), s.grant_lease(50)
    s.put("/k", "v1", l1)
    s.put("/k", "v2", l2)  # key moves to the second lease
    assert s.revoke_lease(l1) == bb.ErrorCode.OK
    assert s.get("/k") == b"v2"
    assert s.revoke_lease(l2) == bb.ErrorCode.OK and s.get("/k") is None


def test_txn_primitives(bb):
    s = bb.MemCoord()
    assert s.put_if_absent("/lock", "a") is True
    assert s.put_if_absent("/lock", "b") is False and s.get("/lock") == b"a"
    assert s.compare_and_swap("/lock", "zzz", "c") is False
    assert s.compare_and_swap("/lock", "a", "c") is True and s.get("/lock") == b"c"
    assert s.compare_and_delete("/lock", "a") is False
    assert s.compare_and_delete("/lock", "c") is True and s.get("/lock") is None


def test_ttl_put_reuses_one_lease_per_key(bb):
    """Reference bug #12: every put_with_ttl granted a fresh lease and leaked the old one."""
    store = bb.MemCoord()
    svc = bb.CoordService(store)
    for _ in range(20):
        assert svc.put_with_ttl("/hb/w", "t", 10) == bb.ErrorCode.OK
    assert store.lease_count() == 1
    store.advance_time_ms(11000)
    assert svc.get("/hb/w") is None
    assert svc.put_with_ttl("/hb/w", "t", 10) == bb.ErrorCode.OK and store.lease_count() == 1  # re-granted after expiry


def test_service_registry(bb):
    store = bb.MemCoord()
    svc = bb.CoordService(store)
    assert svc.register_service("blackbird-keystone", "ks-1", "10.0.0.1:9090", 60) == bb.ErrorCode.OK
    assert svc.register_service("blackbird-keystone", "ks-2", "10.0.0.2:9090", 60) == bb.ErrorCode.OK
    assert sorted(svc.discover_service("blackbird-keystone")) == ["10.0.0.1:9090", "10.0.0.2:9090"]
    assert store.get("/blackbird/services/blackbird-keystone/ks-1") == b"10.0.0.1:9090"  # reference key schema
    svc.unregister_service("blackbird-keystone", "ks-1")
    assert svc.discover_service("blackbird-keystone") == ["10.0.0.2:9090"]


def test_leader_election_cas_lease_failover(bb):
    """campaign_leader is a stub in the reference (etcd_service.cpp:379-385)."""
    store = bb.MemCoord()
    a, b = bb.CoordService(store), bb.CoordService(store)
    ec, won_a = a.campaign_leader("ks", "A", 5)
    ec, won_b = b.campaign_leader("ks", "B", 5)
    assert won_a and not won_b and a.get_leader("ks") == "A"
    assert store.get("/blackbird/elections/ks/leader") == b"A"
    assert a.refresh_leadership("ks", "A") == bb.ErrorCode.OK
    assert b.refresh_leadership("ks", "B") == bb.ErrorCode.NOT_LEADER
    assert b.resign_leader("ks", "B") == bb.ErrorCode.NOT_LEADER  # cannot depose the leader
    store.advance_time_ms(6000)  # A stops refreshing: lease expires
    assert a.refresh_leadership("ks", "A") == bb.ErrorCode.NOT_LEADER
    ec, won_b = b.campaign_leader("ks", "B", 5)
    assert won_b and b.get_leader("ks") == "B"
    ec, won_a = a.campaign_leader("ks", "A", 5)
    assert not won_a
    assert b.resign_leader("ks", "B") == bb.ErrorCode.OK and a.get_leader("ks") is None


def test_only_one_of_many_concurrent_campaigners_wins(bb):
    store = bb.MemCoord()
    wins = []

    def run(i):
        svc = bb.CoordService(store)
        ec, won = svc.campaign_leader("race", f"c{i}", 30)
        if won:
            wins.append(i)

    ts = [threading.Thread(target=run, args=(i,)) for i in range(16)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert len(wins) == 1


def test_remote_coord_daemon_roundtrip_and_watch_push(bb):
    srv = bb.CoordServer()
    assert srv.start("127.0.0.1", 0) == bb.ErrorCode.OK
    try:
        c1, c2 = bb.RemoteCoord(), bb.RemoteCoord()
        assert c1.connect(f"127.0.0.1:1,127.0.0.1:{srv.port}") == bb.ErrorCode.OK  # tries every endpoint
        assert c2.connect(f"127.0.0.1:{srv.port}") == bb.ErrorCode.OK
        got = []
        done = threading.Event()

        def cb(t, k, v, rev):
            got.append((t, k, v))
            if len(got) >= 3:
                done.set()

        wid = c2.watch_prefix("/w/", cb)
        assert c1.put("/w/a", "1") == bb.ErrorCode.OK
        lease = c1.grant_lease(30)
        assert c1.put("/w/b", b"\x00\x01binary".decode("latin1"), lease) == bb.ErrorCode.OK
        assert c1.put("/other", "x") == bb.ErrorCode.OK
        assert c1.revoke_lease(lease) == bb.ErrorCode.OK
        assert done.wait(5.0)
        assert [g[:2] for g in got] == [("PUT", "/w/a"), ("PUT", "/w/b"), ("DELETE", "/w/b")]
        assert c2.get("/w/a") == b"1" and c2.put_if_absent("/w/a", "2") is False
        assert c2.compare_and_swap("/w/a", "1", "3") is True
        assert [k for k, *_ in c1.get_with_prefix("/w/")] == ["/w/a"]
        assert c2.unwatch(wid) == bb.ErrorCode.OK
        assert c1.revision() == srv.store().revision()
        c1.close()
        c2.close()
    finally:
        srv.stop()
