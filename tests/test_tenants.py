"""Tenants (csrc/common/tenant.h): named principals with their own secret, key-prefix ACLs and an admission budget.

The reference has one trust level and lists "Security (mTLS), ACLs" and "admission control" as roadmap / Keystone duties
(README.md:104-108, 146-153).  Here a tenant proves its own secret in the RPC handshake, the Keystone checks every key of
every object call against the tenant's grants, a tenant connection cannot reach the cluster-management methods, and
`quota_bytes` / `max_objects` bound what it may hold (charged at put_start, released wherever the object leaves the table,
rebuilt from the metadata log after a restart)."""
import json
import os
import socket
import subprocess
import time

import pytest

from test_keystone import cfg1, ks_cfg, mkpool

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.environ.get("BB_BIN_DIR", os.path.join(ROOT, "bin"))

TABLE = """
tenants:
  - name: alice
    secret: "alice-secret"
    write: ["alice/"]
    read: ["shared/"]
    quota_bytes: 1MB
    max_objects: 5
  - name: bob
    secret: "bob-secret"
    write: ["bob/"]
    read: ["alice/", "shared/"]
  - name: ops
    secret: "ops-secret"
    admin: true
    write: ["*"]
"""


@pytest.fixture
def tenants(bb):
    bb.load_tenants_text(TABLE)
    yield
    bb.set_tenants([])
    bb.set_client_tenant("", "")


# ---------------------------------------------------------------- the table
def test_table_parses_and_grants_are_prefixes(bb, tenants):
    assert bb.tenant_names() == ["alice", "bob", "ops"]
    assert bb.tenant_may("alice", "alice/ckpt/0", write=True) and bb.tenant_may("alice", "alice/ckpt/0")  # a write grant reads too
    assert not bb.tenant_may("alice", "alic", write=True) and not bb.tenant_may("alice", "bob/x")
    assert bb.tenant_may("alice", "shared/x") and not bb.tenant_may("alice", "shared/x", write=True)
    assert bb.tenant_may("bob", "alice/ckpt/0") and not bb.tenant_may("bob", "alice/ckpt/0", write=True)
    assert bb.tenant_may("ops", "anything/at/all", write=True)
    assert not bb.tenant_may("nobody", "alice/x")


@pytest.mark.parametrize("text,why", [
    ("tenants:\n  - name: a\n", "no secret"),
    ("tenants:\n  - secret: s\n", "name"),
    ("tenants:\n  - {name: a, secret: s}\n  - {name: a, secret: t}\n", "twice"),
    ("tenants:\n  - {name: a, secret: s, write: 7}\n", "list"),
    ("tenants:\n  - {name: a, secret: s, quota_bytes: lots}\n", "size"),
    ("tenants: 3\n", "list"),
    ("tenants:\n  - {name: " + "x" * 65 + ", secret: s}\n", "64"),
])
def test_bad_tables_are_refused_and_change_nothing(bb, tenants, text, why):
    with pytest.raises(ValueError) as e:
        bb.load_tenants_text(text)
    assert why in str(e.value)
    assert bb.tenant_names() == ["alice", "bob", "ops"]


def test_secret_from_the_environment_and_reload_on_change(bb, tmp_path, monkeypatch):
    monkeypatch.setenv("CAROL_SECRET", "from-env")
    f = tmp_path / "tenants.yaml"
    f.write_text("tenants:\n  - {name: carol, secret_env: CAROL_SECRET, write: ['carol/'], quota_bytes: 2GB}\n")
    try:
        bb.load_tenants_file(str(f))
        assert bb.tenant_names() == ["carol"] and bb.tenant_may("carol", "carol/x", write=True)
        assert bb.reload_tenants_if_changed() is False  # untouched
        time.sleep(0.02)
        f.write_text("tenants:\n  - {name: carol, secret_env: CAROL_SECRET, write: ['carol/'], quota_bytes: 2GB}\n"
                     "  - {name: dave, secret: d, read: ['carol/']}\n")
        assert bb.reload_tenants_if_changed() is True and bb.tenant_names() == ["carol", "dave"]
        time.sleep(0.02)
        f.write_text("tenants: [ {name: broken} ]\n")  # a typo must not open or empty the table
        assert bb.reload_tenants_if_changed() is False and bb.tenant_names() == ["carol", "dave"]
        with pytest.raises(ValueError):
            bb.load_tenants_file(str(tmp_path / "missing.yaml"))
    finally:
        bb.set_tenants([])


# ---------------------------------------------------------------- admission control in the Keystone (in process)
@pytest.fixture
def ks(bb, tenants):
    k = bb.KeystoneService(ks_cfg(bb), None)
    assert k.initialize() == bb.ErrorCode.OK and k.start() == bb.ErrorCode.OK
    for i in range(4):
        assert k.register_memory_pool(mkpool(bb, f"p{i}", 4 << 20)) == bb.ErrorCode.OK
    yield k
    k.stop()


def usage(k, name):
    return {u["name"]: u for u in k.tenant_usage()}[name]


def code_of(bb, fn):
    with pytest.raises(bb.BlackbirdError) as e:
        fn()
    return e.value.code


def test_budget_is_charged_at_put_start_and_released_with_the_object(bb, ks):
    assert usage(ks, "alice") == {"name": "alice", "used_bytes": 0, "objects": 0, "quota_bytes": 1 << 20, "max_objects": 5}
    with bb.TenantScope("alice"):
        ks.put_start("alice/a", 300_000, cfg1(bb, ttl_ms=0))
        assert usage(ks, "alice")["used_bytes"] == 300_000  # charged while still PENDING: two racing puts cannot both fit
        ks.put_start("alice/b", 300_000, cfg1(bb, replication_factor=2, ttl_ms=0))
        assert usage(ks, "alice")["used_bytes"] == 900_000 and usage(ks, "alice")["objects"] == 2  # replicas count
        # over the line: refused before anything is allocated, nothing charged
        used = ks.get_cluster_stats().used_capacity
        assert code_of(bb, lambda: ks.put_start("alice/c", 200_000, cfg1(bb))) == bb.ErrorCode.QUOTA_EXCEEDED
        assert ks.get_cluster_stats().used_capacity == used and usage(ks, "alice")["used_bytes"] == 900_000
        # outside the grants: ACCESS_DENIED, also when the budget would allow it
        assert code_of(bb, lambda: ks.put_start("bob/x", 10, cfg1(bb))) == bb.ErrorCode.ACCESS_DENIED
        assert code_of(bb, lambda: ks.put_start("shared/x", 10, cfg1(bb))) == bb.ErrorCode.ACCESS_DENIED  # read grant only
        # a cancelled put gives its share back; so does a removal
        assert ks.put_cancel("alice/a") == bb.ErrorCode.OK
        assert usage(ks, "alice")["used_bytes"] == 600_000
        assert ks.put_complete("alice/b", [[1], [2]]) == bb.ErrorCode.OK
        ks.put_start("alice/c", 200_000, cfg1(bb))  # fits now
    # whoever removes it (a member here), the owner's budget gets it back
    assert ks.remove_object("alice/b") == bb.ErrorCode.OK
    assert usage(ks, "alice")["used_bytes"] == 200_000 and usage(ks, "alice")["objects"] == 1
    # members are not charged and not limited
    ks.put_start("alice/by-a-member", 3 << 20, cfg1(bb))
    assert usage(ks, "alice")["used_bytes"] == 200_000
    text = ks.metrics_text()
    assert "tenant_quota_denials_total 1" in text and "tenant_acl_denials_total 2" in text


def test_object_count_budget_and_expiry(bb, ks):
    with bb.TenantScope("alice"):
        for i in range(5):
            ks.put_start(f"alice/{i}", 100, cfg1(bb, ttl_ms=60 if i < 2 else 0))
            ks.put_complete(f"alice/{i}", [[i]])
        assert code_of(bb, lambda: ks.put_start("alice/5", 100, cfg1(bb))) == bb.ErrorCode.QUOTA_EXCEEDED  # max_objects: 5
        time.sleep(0.12)
        assert ks.run_gc_once() == 2  # TTL expiry releases the share as well
        assert usage(ks, "alice")["objects"] == 3 and usage(ks, "alice")["used_bytes"] == 300
        ks.put_start("alice/5", 100, cfg1(bb))
        # a batch is admitted object by object: what fits is placed, the rest is refused
        res = ks.batch_put_start([f"alice/b{i}" for i in range(3)], [100] * 3, cfg1(bb))
    codes = [ec for ec, _ in res]
    assert codes == [bb.ErrorCode.OK, bb.ErrorCode.QUOTA_EXCEEDED, bb.ErrorCode.QUOTA_EXCEEDED]
    assert usage(ks, "alice")["objects"] == 5
    # an unlimited tenant (no quota) is only counted
    with bb.TenantScope("bob"):
        ks.put_start("bob/big", 2 << 20, cfg1(bb))
    assert usage(ks, "bob") == {"name": "bob", "used_bytes": 2 << 20, "objects": 1, "quota_bytes": 0, "max_objects": 0}


def test_budgets_survive_a_restart_through_the_metadata_log(bb, tenants, tmp_path):
    wal = str(tmp_path / "wal")
    mk = lambda: bb.KeystoneService(ks_cfg(bb, cluster_id="ten", wal_path=wal, wal_fsync=False), None)
    a = mk()
    assert a.initialize() == bb.ErrorCode.OK and a.start() == bb.ErrorCode.OK
    pools = [mkpool(bb, f"p{i}", 4 << 20) for i in range(2)]
    for p in pools:
        a.register_memory_pool(p)
    with bb.TenantScope("alice"):
        for i in range(3):
            a.put_start(f"alice/{i}", 200_000, cfg1(bb, ttl_ms=0))
            assert a.put_complete(f"alice/{i}", [[i]]) == bb.ErrorCode.OK
        a.put_start("alice/pending", 100_000, cfg1(bb))  # never completed: not in the log, not charged after the restart
    a.put_start("plain", 1000, cfg1(bb, ttl_ms=0))
    a.put_complete("plain", [[9]])
    a.stop()
    del a
    b = mk()
    assert b.initialize() == bb.ErrorCode.OK and b.start() == bb.ErrorCode.OK
    for p in pools:
        b.register_memory_pool(p)
    assert usage(b, "alice")["used_bytes"] == 600_000 and usage(b, "alice")["objects"] == 3
    with bb.TenantScope("alice"):
        assert code_of(bb, lambda: b.put_start("alice/more", 500_000, cfg1(bb))) == bb.ErrorCode.QUOTA_EXCEEDED
        assert b.remove_object("alice/0") == bb.ErrorCode.OK
        b.put_start("alice/more", 500_000, cfg1(bb))
    assert usage(b, "alice")["used_bytes"] == 900_000
    b.stop()


# ---------------------------------------------------------------- over the wire (processes, sealed transport)
def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def wait_port(port, timeout=10.0):
    t0 = time.time()
    while time.time() - t0 < timeout:
        try:
            socket.create_connection(("127.0.0.1", port), 0.2).close()
            return True
        except OSError:
            time.sleep(0.05)
    return False


def cli(env, *args, timeout=30):
    return subprocess.run([os.path.join(BIN, "bb-cli"), *args], capture_output=True, text=True, timeout=timeout, env=env)


@pytest.mark.parametrize("sealed", [False, True])
def test_tenants_end_to_end_over_rpc(bb, tmp_path, sealed):
    """Keystone + worker processes with a cluster token and a tenant table; the tools act as tenants (no member token)."""
    TOKEN = "member-token"
    tfile = tmp_path / "tenants.yaml"
    tfile.write_text(TABLE)
    base = {k: v for k, v in os.environ.items() if not k.startswith("BB_")}
    audit = tmp_path / "audit.jsonl"  # keystone, worker and coordinator append to one file (O_APPEND, a line per write)
    srv_env = dict(base, BB_AUTH_TOKEN=TOKEN, BB_TENANTS_FILE=str(tfile), BB_AUDIT_LOG=str(audit))
    if sealed:
        srv_env["BB_ENCRYPT_TRANSPORT"] = "1"
        base["BB_ENCRYPT_TRANSPORT"] = "1"
    cport, rport, hport = free_port(), free_port(), free_port()
    procs = []

    def spawn(*cmd):
        procs.append(subprocess.Popen(list(cmd), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=srv_env))

    try:
        spawn(os.path.join(BIN, "bb-coord"), "--listen", f"127.0.0.1:{cport}")
        assert wait_port(cport)
        kcfg = tmp_path / "k.yaml"
        kcfg.write_text(open(os.path.join(ROOT, "configs", "keystone.yaml")).read().replace("health_check_interval_sec: ", "health_check_interval_sec: 1 #"))
        spawn(os.path.join(BIN, "bb-keystone"), str(kcfg), "--coord-endpoints", f"127.0.0.1:{cport}", "--listen-address", f"127.0.0.1:{rport}",
              "--http-port", str(hport), "--cluster-id", "ten")
        assert wait_port(rport)
        wcfg = tmp_path / "w.yaml"
        wcfg.write_text('worker: {worker_id: "wt", node_id: "node-wt", lease_ttl_sec: 3, heartbeat_interval_sec: 1}\n'
                        'storage_pools:\n  - {pool_id: "ram-wt", storage_class: "RAM_CPU", size_bytes: 64_MB}\n')
        spawn(os.path.join(BIN, "bb-worker"), "--config", str(wcfg), "--coord-endpoints", f"127.0.0.1:{cport}", "--cluster-id", "ten")
        ks = f"127.0.0.1:{rport}"
        member = dict(base, BB_AUTH_TOKEN=TOKEN)
        alice = dict(base, BB_TENANT="alice", BB_TENANT_SECRET="alice-secret")
        bob = dict(base, BB_TENANT="bob", BB_TENANT_SECRET="bob-secret")
        ops = dict(base, BB_TENANT="ops", BB_TENANT_SECRET="ops-secret")
        deadline = time.time() + 15
        while time.time() < deadline:
            st = cli(member, "--keystone", ks, "stats")
            if st.returncode == 0 and json.loads(st.stdout)["total_memory_pools"] == 1:
                break
            time.sleep(0.1)
        else:
            raise AssertionError(st.stdout + st.stderr)
        blob = tmp_path / "blob"
        blob.write_bytes(os.urandom(300_000))
        # alice writes under her prefix (Keystone call + data-server write, both as a tenant) and reads it back
        r = cli(alice, "--keystone", ks, "put", "alice/ckpt", str(blob))
        assert r.returncode == 0, r.stdout + r.stderr
        out = tmp_path / "copy"
        assert cli(alice, "--keystone", ks, "get", "alice/ckpt", str(out)).returncode == 0 and out.read_bytes() == blob.read_bytes()
        # ... nowhere else
        r = cli(alice, "--keystone", ks, "put", "bob/steal", str(blob))
        assert r.returncode != 0 and "ACCESS_DENIED" in r.stdout + r.stderr
        assert cli(alice, "--keystone", ks, "put", "shared/x", str(blob)).returncode != 0
        # bob may read alice's objects, not change them
        out.unlink()
        assert cli(bob, "--keystone", ks, "get", "alice/ckpt", str(out)).returncode == 0 and out.read_bytes() == blob.read_bytes()
        assert "alice/ckpt" in cli(bob, "--keystone", ks, "ls", "alice/").stdout
        for args in (("remove", "alice/ckpt"), ("put", "alice/ckpt2", str(blob))):
            r = cli(bob, "--keystone", ks, *args)
            assert r.returncode != 0 and "ACCESS_DENIED" in r.stdout + r.stderr, (args, r.stdout)
        # alice cannot look into bob's prefix, nor list the whole store
        assert cli(bob, "--keystone", ks, "put", "bob/notes", str(blob)).returncode == 0
        assert cli(alice, "--keystone", ks, "exists", "bob/notes").returncode != 0
        assert cli(alice, "--keystone", ks, "get", "bob/notes", str(out)).returncode != 0
        r = cli(alice, "--keystone", ks, "ls", "")
        assert r.returncode != 0 and "bob/notes" not in r.stdout
        # the budget: 1 MB for alice (300 KB used)
        assert cli(alice, "--keystone", ks, "put", "alice/2", str(blob)).returncode == 0
        assert cli(alice, "--keystone", ks, "put", "alice/3", str(blob)).returncode == 0
        r = cli(alice, "--keystone", ks, "put", "alice/4", str(blob))
        assert r.returncode != 0 and "QUOTA_EXCEEDED" in r.stdout + r.stderr
        assert cli(alice, "--keystone", ks, "remove", "alice/2").returncode == 0
        assert cli(alice, "--keystone", ks, "put", "alice/4", str(blob)).returncode == 0
        # a tenant sees its own usage; members and admins see everybody's
        t = cli(alice, "--keystone", ks, "tenants").stdout
        assert "alice" in t and "900000" in t and "bob" not in t
        t = cli(member, "--keystone", ks, "tenants").stdout
        assert "alice" in t and "bob" in t and "ops" in t
        # cluster management is not for tenants ... unless they are admins
        for args in (("remove-worker", "wt"), ("migrate", "alice/ckpt", "NVME"), ("scrub",), ("workers",)):
            r = cli(alice, "--keystone", ks, *args)
            assert r.returncode != 0, (args, r.stdout)
        assert cli(ops, "--keystone", ks, "workers").returncode == 0
        assert cli(ops, "--keystone", ks, "remove", "bob/notes").returncode == 0
        # identities cannot be guessed or probed: a wrong secret and an unknown name get the same refusal
        wrong = cli(dict(base, BB_TENANT="alice", BB_TENANT_SECRET="guess"), "--keystone", ks, "exists", "alice/ckpt")
        unknown = cli(dict(base, BB_TENANT="mallory", BB_TENANT_SECRET="guess"), "--keystone", ks, "exists", "alice/ckpt")
        assert wrong.returncode != 0 and unknown.returncode != 0
        assert cli(base, "--keystone", ks, "exists", "alice/ckpt").returncode != 0  # and no identity is no entry
        assert cli(base, "--keystone", ks, "--tenant", "alice", "--tenant-secret", "alice-secret", "exists", "alice/ckpt").returncode == 0
        # the coordination store admits no tenants
        bb.set_cluster_token("")
        bb.set_client_tenant("alice", "alice-secret")
        bb.set_transport_encryption(sealed)
        try:
            assert bb.RemoteCoord().connect(f"127.0.0.1:{cport}", 2000) != bb.ErrorCode.OK
        finally:
            bb.set_client_tenant("", "")
            bb.set_transport_encryption(False)
        m = cli(member, "metrics", "--http", f"127.0.0.1:{hport}").stdout
        vals = {ln.split()[0]: float(ln.split()[1]) for ln in m.splitlines() if ln and not ln.startswith("#")}
        assert vals["bb_rpc_tenant_handshakes_total"] >= 10 and vals["bb_rpc_tenant_denials_total"] >= 4
        assert vals['bb_tenant_used_bytes{tenant="alice"}'] == 900_000 and vals['bb_tenant_quota_bytes{tenant="alice"}'] == 1 << 20
        assert vals["bb_tenant_acl_denials_total"] >= 5 and vals["bb_tenant_quota_denials_total"] >= 1
        # revocation: bob leaves the table; the Keystone and the worker pick the file up within a health round
        tfile.write_text(TABLE.replace('  - name: bob\n    secret: "bob-secret"\n    write: ["bob/"]\n    read: ["alice/", "shared/"]\n', ""))
        deadline = time.time() + 10
        while time.time() < deadline and cli(bob, "--keystone", ks, "exists", "alice/ckpt").returncode == 0:
            time.sleep(0.2)
        assert cli(bob, "--keystone", ks, "exists", "alice/ckpt").returncode != 0
        assert cli(alice, "--keystone", ks, "exists", "alice/ckpt").returncode == 0
        # the audit trail (common/audit.h) names what the counters only count
        ev = [json.loads(ln) for ln in audit.read_text().splitlines()]
        assert all({"ts", "event", "who"} <= set(e) for e in ev)
        kinds = lambda k: [e for e in ev if e["event"] == k]
        assert {e["who"] for e in kinds("tenant_admitted")} >= {"alice", "bob", "ops"} and all(e["sealed"] == ("yes" if sealed else "no") for e in kinds("tenant_admitted"))
        assert any(e["who"] == "alice" and e["op"] == "write" and e["key"] == "bob/steal" for e in kinds("acl_denied"))
        assert any(e["who"] == "bob" and e["key"] == "alice/ckpt" and e["op"] == "write" for e in kinds("acl_denied"))
        assert any(e["who"] == "alice" and e["op"] == "list" and e["key"] == "" for e in kinds("acl_denied"))
        assert any(e["who"] == "alice" and e["key"] == "alice/4" and e["bytes"] == "300000" for e in kinds("quota_denied"))
        assert any(e["who"] == "alice" and "peer" in e for e in kinds("method_denied"))  # remove-worker, migrate, scrub, workers
        # wrong secret, unknown name, no identity at all, and bob after he left the table
        # (a client told to encrypt that has nothing to key from gives up before it connects: no "unknown" line when sealed)
        assert {e["who"] for e in kinds("auth_failed")} >= {"tenant:alice", "tenant:mallory", "tenant:bob"} | (set() if sealed else {"unknown"})
        assert any(e["who"] == "keystone" and e["count"] == "2" for e in kinds("tenants_reloaded"))
        m = cli(member, "metrics", "--http", f"127.0.0.1:{hport}").stdout
        assert "bb_audit_events_total" in m
        # a management call by an admin tenant and one by a member are both on record, with their outcome
        assert cli(ops, "--keystone", ks, "migrate", "alice/ckpt", "NVME").returncode != 0  # (no such tier here)
        assert cli(member, "--keystone", ks, "remove-worker", "no-such-worker").returncode != 0
        ev = [json.loads(ln) for ln in audit.read_text().splitlines()]
        adm = [e for e in ev if e["event"] == "admin"]
        assert any(e["who"] == "ops" and e["op"] == "migrate_object" and e["arg"] == "alice/ckpt" and e["result"] != "OK" for e in adm)
        assert any(e["who"] == "member" and e["op"] == "remove_worker" and e["arg"] == "no-such-worker" for e in adm)
    finally:
        for p in reversed(procs):
            p.terminate()
        for p in procs:
            try:
                p.wait(timeout=5)
            except subprocess.TimeoutExpired:
                p.kill()


def test_sample_table_in_configs_parses(bb, monkeypatch):
    for v in ("TRAINER_SECRET", "INFERENCE_SECRET", "OPS_SECRET"):
        monkeypatch.setenv(v, "x-" + v)
    try:
        bb.load_tenants_file(os.path.join(ROOT, "configs", "tenants.yaml"))
        assert bb.tenant_names() == ["inference", "ops", "trainer"]
        assert bb.tenant_may("inference", "ckpt/llama/0") and not bb.tenant_may("inference", "ckpt/llama/0", write=True)
        assert bb.tenant_may("trainer", "ckpt/llama/0", write=True) and not bb.tenant_may("trainer", "kv-cache/x")
    finally:
        bb.set_tenants([])


def test_tenant_hellos_under_noise_never_admit_anything(bb, tenants):
    """Tenant-shaped handshake noise against a server that serves tenants: names of every length (known, unknown, empty, too
    long), proofs that are random / truncated / the secret itself, hellos repeated or mixed with member hellos.  Only a
    handshake reply or the denial marker ever comes back, a known and an unknown name are answered alike (48 bytes), and a
    tenant with the right secret still gets in afterwards."""
    import random
    import struct

    AUTH, DENIED = 0x7FFFFF00, 0x7FFFFFFD
    frame = lambda method, rid, body: struct.pack("<IIQ", len(body), method, rid) + body
    rng = random.Random(0x7E17)
    k = bb.KeystoneService(ks_cfg(bb), None)
    assert k.initialize() == bb.ErrorCode.OK and k.start() == bb.ErrorCode.OK
    rpc = bb.RpcService(k, ks_cfg(bb))
    bb.set_cluster_token("member-token")
    try:
        assert rpc.start() == bb.ErrorCode.OK
        port = rpc.rpc_port
        for _ in range(200):
            s = socket.create_connection(("127.0.0.1", port), 2.0)
            s.settimeout(0.3)
            try:
                for _ in range(rng.randrange(1, 5)):
                    kind = rng.randrange(6)
                    if kind == 0:
                        name = rng.choice([b"alice", b"bob", b"mallory", b"", b"a" * 64, b"a" * 65, rng.randbytes(rng.randrange(1, 80))])
                        body = rng.choice([b"BBT1", b"BBT2"]) + rng.randbytes(16) + name
                    elif kind == 1:
                        body = rng.randbytes(32)
                    elif kind == 2:
                        body = b"alice-secret"
                    elif kind == 3:
                        body = b"BBT1" + rng.randbytes(rng.randrange(0, 16))  # a hello too short to hold a nonce
                    elif kind == 4:
                        body = rng.choice([b"BBA1", b"BBR1"]) + rng.randbytes(16)
                    else:
                        s.sendall(frame(rng.randrange(1, 30), 3, rng.randbytes(rng.randrange(0, 60))))  # a request before being admitted
                        continue
                    s.sendall(frame(AUTH, rng.getrandbits(64), body))
                got = b""
                try:
                    while len(got) < 4096:
                        part = s.recv(4096)
                        if not part:
                            break
                        got += part
                except OSError:
                    pass
                pos = 0
                while len(got) - pos >= 16:
                    n, method, _ = struct.unpack_from("<IIQ", got, pos)
                    assert method in (AUTH, DENIED), hex(method)
                    assert method != AUTH or n == 48, n  # never the empty "admitted" frame
                    pos += 16 + n
            except OSError:
                pass
            finally:
                s.close()
        bb.set_cluster_token("")
        bb.set_client_tenant("alice", "alice-secret")
        c = bb.KeystoneRpcClient()
        assert c.connect("127.0.0.1", port, 2000) == bb.ErrorCode.OK
        assert c.object_exists("alice/x") is False
        with pytest.raises(bb.BlackbirdError) as e:
            c.object_exists("bob/x")
        assert e.value.code == bb.ErrorCode.ACCESS_DENIED
        del c
    finally:
        bb.set_cluster_token("")
        bb.set_client_tenant("", "")
        rpc.stop()
        k.stop()
