"""EtcdCoord: CoordStore over etcd's v3 JSON gateway (the reference's substrate, src/etcd/etcd_service.cpp).  No etcd
exists offline, so the adapter talks to tests/fake_etcd.py, an in-memory implementation of the gateway's wire format
(base64 bytes, string integers, omitted zero fields, chunked watch stream).  The last test runs a whole cluster --
Keystone HA election with fenced metadata writes, worker registration and heartbeats, client put/get -- on top of it."""
import os
import time

import pytest

from fake_etcd import FakeEtcd


def wait_for(pred, timeout=8.0, step=0.02):
    deadline = time.time() + timeout
    while time.time() < deadline:
        if pred():
            return True
        time.sleep(step)
    return pred()


@pytest.fixture
def etcd():
    f = FakeEtcd()
    yield f
    f.stop()


def test_kv_leases_and_transactions_over_the_json_gateway(bb, etcd):
    c = bb.EtcdCoord()
    assert c.connect(f"127.0.0.1:{etcd.port}") == bb.ErrorCode.OK
    blob = bytes(range(256)) * 3  # arbitrary bytes survive the base64 round trip
    assert c.put("/a/1", blob) == bb.ErrorCode.OK and c.get("/a/1") == blob and c.get("/nope") is None
    c.put("/a/2", b"two")
    c.put("/b/1", b"other")
    assert [k for k, _, _, _ in c.get_with_prefix("/a/")] == ["/a/1", "/a/2"]
    kv = c.get_kv("/a/2")
    c.put("/a/2", b"two'")
    kv2 = c.get_kv("/a/2")
    assert kv2["create_revision"] == kv["create_revision"] and kv2["mod_revision"] > kv["mod_revision"] and c.revision() >= kv2["mod_revision"]
    # txn mappings
    assert c.put_if_absent("/a/1", b"x") is False and c.put_if_absent("/a/3", b"three") is True
    assert c.compare_and_swap("/a/3", b"wrong", b"y") is False and c.compare_and_swap("/a/3", b"three", b"3") is True and c.get("/a/3") == b"3"
    assert c.compare_and_delete("/a/3", b"nope") is False and c.compare_and_delete("/a/3", b"3") is True and c.get("/a/3") is None
    guard = c.get_kv("/a/1")["create_revision"]
    assert c.guarded_put("/a/1", guard, "/g", b"ok") is True and c.guarded_put("/a/1", guard + 1, "/g", b"bad") is False and c.get("/g") == b"ok"
    assert c.guarded_del("/a/1", guard + 1, "/g") is False and c.guarded_del("/a/1", guard, "/g") is True and c.get("/g") is None
    assert c.del_prefix("/a/") == 2 and c.get_with_prefix("/a/") == []
    # leases
    lease = c.grant_lease(1)
    assert c.put("/l/hb", b"1", lease) == bb.ErrorCode.OK and 0 <= c.lease_remaining_ms(lease) <= 1000
    for _ in range(6):  # refreshed: outlives its TTL
        time.sleep(0.3)
        assert c.keep_alive(lease) == bb.ErrorCode.OK
    assert c.get("/l/hb") == b"1"
    time.sleep(1.4)
    assert c.get("/l/hb") is None and c.keep_alive(lease) == bb.ErrorCode.ETCD_LEASE_ERROR
    l2 = c.grant_lease(30)
    c.put("/l/x", b"1", l2)
    assert c.revoke_lease(l2) == bb.ErrorCode.OK and c.get("/l/x") is None and c.revoke_lease(l2) != bb.ErrorCode.OK
    assert {"/v3/kv/put", "/v3/kv/range", "/v3/kv/deleterange", "/v3/kv/txn", "/v3/lease/grant", "/v3/lease/keepalive", "/v3/lease/revoke",
            "/v3/lease/timetolive"} <= set(etcd.paths())
    c.close()


def test_watch_stream_puts_deletes_lease_expiry_and_unwatch(bb, etcd):
    c = bb.EtcdCoord()
    assert c.connect(etcd.endpoint) == bb.ErrorCode.OK
    events = []
    wid = c.watch_prefix("/w/", lambda t, k, v, rev: events.append((t, k, v)))
    c.put("/w/a", b"1")
    c.put("/elsewhere", b"x")
    c.put("/w/a", b"2")
    c.delete("/w/a")
    lease = c.grant_lease(1)
    c.put("/w/leased", b"hb", lease)
    assert wait_for(lambda: ("DELETE", "/w/leased", b"hb") in events, timeout=5), events  # expiry arrives as a DELETE with the last value
    assert events[:4] == [("PUT", "/w/a", b"1"), ("PUT", "/w/a", b"2"), ("DELETE", "/w/a", b"2"), ("PUT", "/w/leased", b"hb")]
    assert c.unwatch(wid) == bb.ErrorCode.OK
    n = len(events)
    c.put("/w/after", b"z")
    time.sleep(0.3)
    assert len(events) == n  # barrier: nothing is delivered after unwatch returned
    c.close()


@pytest.fixture
def etcd_proc():
    """The gateway in its own process: the C++ services below call etcd from bindings that hold the GIL."""
    import subprocess
    import sys

    p = subprocess.Popen([sys.executable, os.path.join(os.path.dirname(__file__), "fake_etcd.py")], stdout=subprocess.PIPE, text=True)
    port = int(p.stdout.readline().split()[1])

    class E:
        endpoint = f"etcd://127.0.0.1:{port}"

    yield E
    p.kill()
    p.wait()


def test_a_cluster_runs_on_etcd_election_fencing_workers_and_objects(bb, etcd_proc):
    etcd = etcd_proc
    """CoordService("etcd://...") is what the daemons get from `--coord-endpoints etcd://host:2379`: the Keystone pair
    elects through a put-if-absent transaction on the election key, logs object metadata with transactions guarded by that
    key's create revision, the worker registers and heart-beats with leases, the Keystones watch the registration prefix."""
    from blackbird_b200.parallel import LocalCluster
    from test_keystone import ks_cfg

    cfg = ks_cfg(bb, cluster_id="onetcd", enable_ha=True, service_id="ks-a", service_registration_ttl_sec=3, service_refresh_interval_sec=1)
    cl = LocalCluster("onetcd", n_workers=1, pool_bytes=8 << 20, coord=etcd.endpoint, keystone_cfg=cfg, lease_ttl_sec=2, heartbeat_interval_sec=1)
    try:
        assert cl.keystone.is_leader() and cl.keystone.leader_term() > 0
        assert wait_for(lambda: cl.keystone.get_cluster_stats().total_memory_pools == 1)  # learnt through the watch stream
        c = cl.client()
        blob = os.urandom(50_000)
        wc = bb.WorkerConfig(replication_factor=1, max_workers_per_copy=1, ttl_ms=0)
        assert c.put("obj", blob, wc) == bb.ErrorCode.OK and c.get("obj") == blob
        probe = bb.EtcdCoord()
        assert probe.connect(etcd.endpoint) == bb.ErrorCode.OK
        assert probe.get("/blackbird/elections/keystone-onetcd/leader") == b"ks-a"
        assert probe.get("/blackbird/clusters/onetcd/objects/obj") is not None  # fenced txn put
        assert probe.get_kv("/blackbird/elections/keystone-onetcd/leader")["create_revision"] == cl.keystone.leader_term()
        hb = "/blackbird/clusters/onetcd/heartbeat/worker-0"
        assert probe.get(hb) is not None
        time.sleep(2.5)
        assert probe.get(hb) is not None  # the worker keeps its lease alive through /v3/lease/keepalive
        # a standby on the same etcd takes over when the leader goes, and finds the object in the log
        cfg_b = ks_cfg(bb, cluster_id="onetcd", enable_ha=True, service_id="ks-b", service_registration_ttl_sec=3, service_refresh_interval_sec=1)
        b = bb.KeystoneService(cfg_b, bb.CoordService(etcd.endpoint))
        assert b.initialize() == bb.ErrorCode.OK and b.start() == bb.ErrorCode.OK and not b.is_leader()
        cl.keystone.stop()  # resigns
        assert wait_for(lambda: b.is_leader(), timeout=8)
        assert b.get_workers("obj")[0].shards[0].length == len(blob)
        assert b.leader_term() > 0 and probe.get("/blackbird/elections/keystone-onetcd/leader") == b"ks-b"
        b.stop()
        probe.close()
        import json
        import urllib.request

        req = urllib.request.Request("http://" + etcd.endpoint[len("etcd://"):] + "/debug/paths", data=b"{}", method="POST")
        paths = json.load(urllib.request.urlopen(req, timeout=5))["paths"]
        assert {"/v3/watch", "/v3/kv/txn", "/v3/lease/grant", "/v3/lease/keepalive", "/v3/kv/put", "/v3/kv/range"} <= set(paths)
    finally:
        cl.stop()
