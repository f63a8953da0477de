"""Durability, reconnection and fencing (SURVEY §5.4 "snapshot + log"):

* the coordination store persists to a log + snapshots and comes back with keys, revisions and lease ids;
* a RemoteCoord client survives the daemon being replaced: calls reconnect, watches are re-established and re-listed;
* a single Keystone with `wal_path` keeps its object table (and allocator reservations) across a restart;
* a leader whose election key was replaced cannot write the shared metadata log (term fencing) and steps down;
* a leader that was paused past its lease comes back as a non-leader *before* talking to the store.
"""
import os
import signal
import subprocess
import time

import pytest

from test_keystone import cfg1, ks_cfg, mkpool

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.environ.get("BB_BIN_DIR", os.path.join(ROOT, "bin"))


def wait_for(pred, timeout=8.0, step=0.05):
    deadline = time.time() + timeout
    while time.time() < deadline:
        if pred():
            return True
        time.sleep(step)
    return pred()


# ---------------------------------------------------------------- coordination store on disk
def test_mem_coord_persists_keys_revisions_and_leases(bb, tmp_path):
    d = str(tmp_path / "coord")
    a = bb.MemCoord()
    assert a.open_durable(d, fsync=True) == bb.ErrorCode.OK and a.durable
    a.put("/k/plain", b"v1")
    a.put("/k/plain", b"v2")
    lease = a.grant_lease(5)
    a.put("/k/leased", b"hb", lease)
    a.put("/k/gone", b"x")
    a.delete("/k/gone")
    short = a.grant_lease(1)
    a.put("/k/expiring", b"soon", short)
    a.advance_time_ms(1500)  # the short lease expires: logged as a revoke, must not come back
    rev = a.revision()
    kv = a.get_kv("/k/plain")
    del a

    b = bb.MemCoord()
    assert b.open_durable(d) == bb.ErrorCode.OK and b.recovered_records > 0
    assert b.get("/k/plain") == b"v2" and b.get("/k/gone") is None and b.get("/k/expiring") is None
    assert b.get_kv("/k/plain")["create_revision"] == kv["create_revision"] and b.get_kv("/k/plain")["mod_revision"] == kv["mod_revision"]
    assert b.revision() == rev
    # the lease id survived and was re-armed with its full TTL: its holder can keep refreshing it
    assert b.get("/k/leased") == b"hb" and b.keep_alive(lease) == bb.ErrorCode.OK
    assert 4000 < b.lease_remaining_ms(lease) <= 5000
    assert b.grant_lease(5) > lease  # ids are not reused
    b.advance_time_ms(6000)
    assert b.get("/k/leased") is None


def test_mem_coord_snapshots_compact_the_log_and_torn_tails_are_dropped(bb, tmp_path):
    d = tmp_path / "coord"
    a = bb.MemCoord()
    assert a.open_durable(str(d), fsync=False, snapshot_bytes=4096) == bb.ErrorCode.OK
    for i in range(400):
        a.put(f"/bulk/{i % 50}", os.urandom(64))
    want = {k: v for k, v, _, _ in a.get_with_prefix("/bulk/")}
    rev = a.revision()
    del a
    names = sorted(os.listdir(d))
    assert any(n.startswith("coord.snap.") for n in names), names
    assert sum(n.startswith("coord.wal.") for n in names) <= 2  # older generations were removed
    # a crash in the middle of an append leaves a torn record at the end of the newest log: it is dropped
    newest = max((n for n in names if n.startswith("coord.wal.")), key=lambda n: int(n.rsplit(".", 1)[1]))
    with open(d / newest, "ab") as f:
        f.write(b"\x40\x00\x00\x00\x12\x34\x56\x78partial")
    b = bb.MemCoord()
    assert b.open_durable(str(d), fsync=False) == bb.ErrorCode.OK
    assert {k: v for k, v, _, _ in b.get_with_prefix("/bulk/")} == want and b.revision() == rev
    b.put("/after", b"ok")  # and the log is appendable again
    del b
    c = bb.MemCoord()
    assert c.open_durable(str(d), fsync=False) == bb.ErrorCode.OK and c.get("/after") == b"ok"


def test_remote_coord_survives_a_replaced_daemon(bb, tmp_path):
    """In-process stand-in for `kill -9 bb-coord; bb-coord --data-dir same`: the server object is destroyed and a new
    one is started on the same port from the same directory while a client holds a lease and a watch."""
    d = str(tmp_path / "coord")
    store = bb.MemCoord()
    assert store.open_durable(d, fsync=False) == bb.ErrorCode.OK
    srv = bb.CoordServer(store)
    assert srv.start("127.0.0.1", 0) == bb.ErrorCode.OK
    port = srv.port
    rc = bb.RemoteCoord()
    assert rc.connect(f"127.0.0.1:{port}") == bb.ErrorCode.OK
    events = []
    rc.watch_prefix("/w/", lambda t, k, v, rev: events.append((t, k, v)))
    lease = rc.grant_lease(30)
    rc.put("/w/a", b"1", lease)
    rc.put("/w/b", b"2")
    assert wait_for(lambda: len(events) == 2)

    srv.stop()
    del srv, store
    # while the daemon is away another actor (e.g. an operator with the data dir) cannot reach it either; changes made
    # right after the restart, before the client noticed, are what the re-list has to deliver
    store2 = bb.MemCoord()
    assert store2.open_durable(d, fsync=False) == bb.ErrorCode.OK
    store2.delete("/w/b")
    store2.put("/w/c", b"3")
    srv2 = bb.CoordServer(store2)
    assert srv2.start("127.0.0.1", port) == bb.ErrorCode.OK
    # request/response path: reconnects and retries transparently; the lease id is still valid
    assert rc.get("/w/a") == b"1" and rc.keep_alive(lease) == bb.ErrorCode.OK and rc.reconnects >= 1
    # watch path: re-established, then the difference is replayed (DELETE for b, PUT for c)
    assert wait_for(lambda: ("DELETE", "/w/b", b"") in events and ("PUT", "/w/c", b"3") in events), events
    n = len(events)
    store2.put("/w/live", b"4")  # and live events flow again
    assert wait_for(lambda: ("PUT", "/w/live", b"4") in events[n:])
    rc.close()
    srv2.stop()


# ---------------------------------------------------------------- keystone local metadata log
def test_keystone_wal_path_survives_a_restart(bb, tmp_path):
    wal = str(tmp_path / "ks-wal")
    mk = lambda: bb.KeystoneService(ks_cfg(bb, cluster_id="dur", wal_path=wal, wal_fsync=False, wal_snapshot_mb=1), None)
    a = mk()
    assert a.initialize() == bb.ErrorCode.OK and a.start() == bb.ErrorCode.OK
    pools = [mkpool(bb, f"p{i}", 1 << 20) for i in range(2)]
    for p in pools:
        a.register_memory_pool(p)
    placed = {}
    for i in range(40):
        c = a.put_start(f"o{i}", 8192, cfg1(bb, ttl_ms=0))
        assert a.put_complete(f"o{i}", [[i]]) == bb.ErrorCode.OK
        placed[f"o{i}"] = (c[0].shards[0].pool_id, c[0].shards[0].location["remote_addr"])
    a.put_start("never-completed", 8192, cfg1(bb))
    assert a.remove_object("o7") == bb.ErrorCode.OK
    del placed["o7"]
    assert "bb_wal_write_failed_total" not in a.metrics_text()
    a.stop()
    del a

    b = mk()
    assert b.initialize() == bb.ErrorCode.OK and b.start() == bb.ErrorCode.OK
    st = b.get_cluster_stats()
    assert st.total_objects == 39 and st.pending_objects == 0  # PENDING objects and removed ones are not resurrected
    # the pools register AFTER the recovery (in-process deployments): extents are adopted then
    for p in pools:
        b.register_memory_pool(p)
    for k, (pid, addr) in placed.items():
        got = b.get_workers(k)
        assert (got[0].shards[0].pool_id, got[0].shards[0].location["remote_addr"]) == (pid, addr) and got[0].shards[0].checksum == int(k[1:])
    with pytest.raises(bb.BlackbirdError):
        b.get_workers("o7")
    # recovered extents are reserved: fresh objects do not land on them
    taken = set(placed.values())
    for i in range(40):
        c = b.put_start(f"new{i}", 8192, cfg1(bb))
        assert (c[0].shards[0].pool_id, c[0].shards[0].location["remote_addr"]) not in taken
    assert b.get_cluster_stats().used_capacity == (39 + 40) * 8192
    b.stop()


def test_keystone_wal_snapshot_and_log_replay_agree(bb, tmp_path):
    wal = tmp_path / "ks-wal"
    cfg = ks_cfg(bb, cluster_id="snap", wal_path=str(wal), wal_fsync=False)
    cfg.wal_snapshot_mb = 0  # use the default threshold: force snapshots through stop() instead
    a = bb.KeystoneService(cfg, None)
    assert a.initialize() == bb.ErrorCode.OK and a.start() == bb.ErrorCode.OK
    a.register_memory_pool(mkpool(bb, "p0", 4 << 20))
    for i in range(100):
        a.put_start(f"k{i}", 4096, cfg1(bb, ttl_ms=0))
        a.put_complete(f"k{i}", [[i]])
    a.stop()  # writes a snapshot on the way out
    assert any(n.startswith("keystone-snap.snap.") for n in os.listdir(wal)), os.listdir(wal)
    b = bb.KeystoneService(cfg, None)
    assert b.initialize() == bb.ErrorCode.OK and b.start() == bb.ErrorCode.OK
    b.register_memory_pool(mkpool(bb, "p0", 4 << 20))
    for i in range(50):
        assert b.remove_object(f"k{i}") == bb.ErrorCode.OK  # tombstones land in the log after the snapshot
    del b  # no clean stop: the next start replays snapshot + log
    c = bb.KeystoneService(cfg, None)
    assert c.initialize() == bb.ErrorCode.OK and c.start() == bb.ErrorCode.OK
    assert c.get_cluster_stats().total_objects == 50
    assert c.object_exists("k75") is True and c.object_exists("k10") is False
    c.stop()


# ---------------------------------------------------------------- fencing
def ha_pair(bb, store, cluster, ttl=4):
    mk = lambda sid: bb.KeystoneService(ks_cfg(bb, cluster_id=cluster, enable_ha=True, service_id=sid, service_registration_ttl_sec=ttl,
                                               service_refresh_interval_sec=1), bb.CoordService(store))
    a, b = mk("ks-a"), mk("ks-b")
    for k in (a, b):
        assert k.initialize() == bb.ErrorCode.OK and k.start() == bb.ErrorCode.OK
    store.put(f"/blackbird/clusters/{cluster}/workers/w0", '{"worker_id":"w0","node_id":"n0"}')
    store.put(f"/blackbird/clusters/{cluster}/workers/w0/memory_pools/p0", mkpool(bb, "p0", 1 << 20, worker="w0").to_json())
    store.flush_events()
    return a, b


def test_deposed_leader_cannot_write_the_metadata_log(bb):
    """The election key is replaced behind the leader's back (what a partition + lease expiry + new election looks like
    from the store's side).  The old leader still believes it leads -- its local lease deadline has not passed -- but its
    put_complete is a transaction guarded by the key's create revision (its term): the write is refused, the client gets
    NOT_LEADER instead of an acknowledgement, nothing reaches the log, and the keystone steps down."""
    store = bb.MemCoord()
    a, b = ha_pair(bb, store, "fence", ttl=30)
    assert a.is_leader() and a.leader_term() > 0
    ekey = "/blackbird/elections/keystone-fence/leader"
    term_a = store.get_kv(ekey)["create_revision"]
    assert term_a == a.leader_term()
    a.put_start("before", 4096, cfg1(bb, ttl_ms=0))
    assert a.put_complete("before", [[1]]) == bb.ErrorCode.OK
    assert store.get("/blackbird/clusters/fence/objects/before") is not None
    a.put_start("during", 4096, cfg1(bb, ttl_ms=0))
    # new election behind a's back
    store.delete(ekey)
    lease = store.grant_lease(30)
    assert store.put_if_absent(ekey, b"ks-b", lease)
    assert store.get_kv(ekey)["create_revision"] > term_a  # terms are monotonic
    assert a.is_leader()  # it has not noticed yet
    assert a.put_complete("during", [[2]]) == bb.ErrorCode.NOT_LEADER  # ... but it cannot acknowledge
    assert store.get("/blackbird/clusters/fence/objects/during") is None
    assert not a.is_leader() and a.leader_term() == 0
    with pytest.raises(bb.BlackbirdError) as e:
        a.put_start("after", 10, cfg1(bb))
    assert e.value.code == bb.ErrorCode.NOT_LEADER
    a.stop(), b.stop()


def test_leader_paused_past_its_lease_comes_back_as_a_non_leader(bb):
    """The keep-alive thread of the leader stops running (a stand-in for SIGSTOP / a VM freeze: nothing refreshes the
    lease or the local deadline).  Once TTL - margin of local time has passed, is_leader() is false on its own -- before
    the process has exchanged a single message with the store -- so there is no window in which it double-acks."""
    store = bb.MemCoord()
    cfg = ks_cfg(bb, cluster_id="pause", enable_ha=True, service_id="ks-a", service_registration_ttl_sec=2, service_refresh_interval_sec=3600)
    a = bb.KeystoneService(cfg, bb.CoordService(store))
    assert a.initialize() == bb.ErrorCode.OK
    # start() would spawn the keep-alive thread with a period of TTL/4; skipping it models the pause
    assert a.is_leader()
    time.sleep(2.0)
    assert not a.is_leader()
    with pytest.raises(bb.BlackbirdError) as e:
        a.put_start("x", 10, cfg1(bb))
    assert e.value.code == bb.ErrorCode.NOT_LEADER


def test_standby_rebuilds_from_the_log_and_ex_leader_forgets(bb):
    """a keystone that loses leadership drops its table and allocator; when it wins again it rebuilds
    strictly from the log, so objects the interim leader removed do not come back and the ones it created are reserved."""
    store = bb.MemCoord()
    a, b = ha_pair(bb, store, "flip", ttl=2)
    assert a.is_leader() and not b.is_leader()
    a.put_start("old", 4096, cfg1(bb, ttl_ms=0))
    assert a.put_complete("old", [[1]]) == bb.ErrorCode.OK
    ekey = "/blackbird/elections/keystone-flip/leader"
    store.delete(ekey)  # a loses the election key; b's next campaign wins
    assert wait_for(lambda: b.is_leader() and not a.is_leader(), timeout=6)
    assert wait_for(lambda: a.get_cluster_stats().total_objects == 0, timeout=3)  # a forgot everything
    assert b.get_workers("old")[0].shards[0].checksum == 1
    assert b.remove_object("old") == bb.ErrorCode.OK
    c = b.put_start("new", 4096, cfg1(bb, ttl_ms=0))
    assert b.put_complete("new", [[2]]) == bb.ErrorCode.OK
    b.stop()  # resigns: a takes over again and must see b's world, not its own stale one
    assert wait_for(lambda: a.is_leader(), timeout=6)
    with pytest.raises(bb.BlackbirdError):
        a.get_workers("old")
    assert a.get_workers("new")[0].shards[0].checksum == 2
    fresh = a.put_start("fresh", 4096, cfg1(bb))
    assert fresh[0].shards[0].location["remote_addr"] != c[0].shards[0].location["remote_addr"]
    a.stop()


# ---------------------------------------------------------------- the real daemon, killed with SIGKILL
def test_bb_coord_killed_and_restarted_keeps_the_cluster(tmp_path, bb):
    def free_port():
        import socket
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        p = s.getsockname()[1]
        s.close()
        return p

    def wait_port(port, timeout=8):
        import socket
        deadline = time.time() + timeout
        while time.time() < deadline:
            try:
                socket.create_connection(("127.0.0.1", port), timeout=0.2).close()
                return True
            except OSError:
                time.sleep(0.05)
        return False

    cport, rport = free_port(), free_port()
    data = tmp_path / "coord-data"
    spawned = []

    def spawn(*argv):
        p = subprocess.Popen(list(argv), stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        spawned.append(p)
        return p

    try:
        coord = spawn(os.path.join(BIN, "bb-coord"), "--listen", f"127.0.0.1:{cport}", "--data-dir", str(data))
        assert wait_port(cport)
        kcfg = tmp_path / "ks.yaml"
        kcfg.write_text("keystone:\n  cluster_id: dur2\n  enable_ha: true\n  service_registration_ttl_sec: 6\n  service_refresh_interval_sec: 1\n"
                        "  http_metrics_port: \"0\"\n")
        spawn(os.path.join(BIN, "bb-keystone"), str(kcfg), "--coord-endpoints", f"127.0.0.1:{cport}", "--listen-address", f"127.0.0.1:{rport}",
              "--service-id", "ks-0")
        assert wait_port(rport)
        wcfg = tmp_path / "w.yaml"
        wcfg.write_text('worker:\n  worker_id: "w0"\n  node_id: "n0"\n  lease_ttl_sec: 6\n  heartbeat_interval_sec: 1\n'
                        'storage_pools:\n  - pool_id: "ram-w0"\n    storage_class: "RAM_CPU"\n    size_bytes: 32_MB\n')
        spawn(os.path.join(BIN, "bb-worker"), "--config", str(wcfg), "--coord-endpoints", f"127.0.0.1:{cport}", "--cluster-id", "dur2")
        opts = bb.BlackbirdClientOptions()
        opts.keystone_endpoints = [f"127.0.0.1:{rport}"]
        cl = bb.BlackbirdClient(opts)
        assert cl.connect() == bb.ErrorCode.OK
        assert wait_for(lambda: cl.keystone().get_cluster_stats().total_memory_pools == 1, timeout=10)
        wc = bb.WorkerConfig()
        wc.replication_factor = 1
        wc.max_workers_per_copy = 1
        wc.ttl_ms = 0
        blob = os.urandom(300000)
        assert cl.put("survivor", blob, wc) == bb.ErrorCode.OK

        coord.send_signal(signal.SIGKILL)
        coord.wait()
        time.sleep(0.5)
        spawn(os.path.join(BIN, "bb-coord"), "--listen", f"127.0.0.1:{cport}", "--data-dir", str(data))
        assert wait_port(cport)
        # leader, worker registration and the object log all came back: the same keystone keeps leading (its lease id
        # was preserved), the worker keeps heart-beating, a new put (which needs a fenced log write) succeeds
        time.sleep(2.5)
        assert cl.get("survivor") == blob
        assert cl.put("after-restart", blob, wc) == bb.ErrorCode.OK and cl.get("after-restart") == blob
        st = cl.keystone().get_cluster_stats()
        assert st.total_workers == 1 and st.total_memory_pools == 1 and st.total_objects == 2
        probe = bb.RemoteCoord()
        assert probe.connect(f"127.0.0.1:{cport}") == bb.ErrorCode.OK
        assert probe.get("/blackbird/elections/keystone-dur2/leader") == b"ks-0"
        assert probe.get("/blackbird/clusters/dur2/objects/survivor") is not None
        assert any(k.endswith("/workers/w0") for k, _, _, _ in probe.get_with_prefix("/blackbird/clusters/dur2/workers/"))
        probe.close()
        # worker's heartbeat lease is still being refreshed against the restarted store: it outlives its TTL
        time.sleep(4.0)
        assert cl.keystone().get_cluster_stats().total_workers == 1
    finally:
        for p in spawned:
            if p.poll() is None:
                p.kill()
            p.wait()
