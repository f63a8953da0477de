"""Secure mode of the framed RPC protocol (net/tcp.h "BBA2", net/aead.h): with a cluster token and `encrypt_transport` every
frame after the handshake is AES-256-GCM protected.  A man in the middle (a TCP proxy here) sees no key names and no payload
bytes, cannot alter a frame without the connection being dropped, and cannot replay one; clients that do not encrypt are
refused by a server that does.  The reference lists transport security as roadmap only (README.md:146-153)."""
import os
import socket
import struct
import subprocess
import threading
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.environ.get("BB_BIN_DIR", os.path.join(ROOT, "bin"))
TOKEN = "s3cret-cluster-token"


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


class Mitm:
    """TCP proxy that keeps everything it forwards and can corrupt or replay client -> server traffic."""

    def __init__(self, target_port):
        self.target = target_port
        self.up, self.down = bytearray(), bytearray()
        self.flip_after = None   # flip one bit of the first client->server chunk forwarded after this many bytes
        self.replay = False      # send the next client->server chunk (after the handshake) twice
        self.srv = socket.socket()
        self.srv.bind(("127.0.0.1", 0))
        self.srv.listen(8)
        self.port = self.srv.getsockname()[1]
        self.run = True
        threading.Thread(target=self._accept, daemon=True).start()

    def _accept(self):
        self.srv.settimeout(0.2)
        while self.run:
            try:
                c, _ = self.srv.accept()
            except OSError:
                continue
            c.settimeout(None)
            u = socket.create_connection(("127.0.0.1", self.target))
            threading.Thread(target=self._pump, args=(c, u, True), daemon=True).start()
            threading.Thread(target=self._pump, args=(u, c, False), daemon=True).start()

    def _pump(self, a, b, upstream):
        try:
            while True:
                d = a.recv(1 << 16)
                if not d:
                    break
                if upstream:
                    if self.flip_after is not None and len(self.up) >= self.flip_after:
                        d = bytearray(d)
                        d[-1] ^= 0x01  # last byte of the chunk: inside the ciphertext / tag of a sealed frame
                        d = bytes(d)
                        self.flip_after = None
                    self.up += d
                    b.sendall(d)
                    if self.replay and len(self.up) > 200:
                        self.replay = False
                        b.sendall(d)
                else:
                    self.down += d
                    b.sendall(d)
        except OSError:
            pass
        for s in (a, b):
            try:
                s.shutdown(socket.SHUT_RDWR)
            except OSError:
                pass

    def stop(self):
        self.run = False
        self.srv.close()


@pytest.fixture
def secure_cluster(tmp_path):
    """bb-coord + bb-keystone + one bb-worker, all with the token and encrypt_transport from the environment."""
    env = dict(os.environ, BB_AUTH_TOKEN=TOKEN, BB_ENCRYPT_TRANSPORT="1")
    cport, rport, hport = free_port(), free_port(), free_port()
    procs = []

    def spawn(*cmd):
        p = subprocess.Popen(list(cmd), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
        procs.append(p)
        return p

    spawn(os.path.join(BIN, "bb-coord"), "--listen", f"127.0.0.1:{cport}")
    assert wait_port(cport)
    spawn(os.path.join(BIN, "bb-keystone"), os.path.join(ROOT, "configs", "keystone.yaml"), "--coord-endpoints", f"127.0.0.1:{cport}",
          "--listen-address", f"127.0.0.1:{rport}", "--http-port", str(hport), "--cluster-id", "sec")
    assert wait_port(rport)
    cfg = tmp_path / "w.yaml"
    cfg.write_text(f"""
worker:
  worker_id: "ws"
  node_id: "node-ws"
  lease_ttl_sec: 3
  heartbeat_interval_sec: 1
storage_pools:
  - pool_id: "ram-ws"
    storage_class: "RAM_CPU"
    size_bytes: 64_MB
""")
    spawn(os.path.join(BIN, "bb-worker"), "--config", str(cfg), "--coord-endpoints", f"127.0.0.1:{cport}", "--cluster-id", "sec")
    yield {"env": env, "coord": cport, "rpc": rport, "http": hport}
    for p in reversed(procs):
        p.terminate()
    for p in procs:
        try:
            p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            p.kill()


def cli(env, *args, timeout=30):
    return subprocess.run([os.path.join(BIN, "bb-cli"), *args], capture_output=True, text=True, timeout=timeout, env=env)


def wait_pools(env, ks, n):
    import json

    deadline = time.time() + 15
    st = None
    while time.time() < deadline:
        st = cli(env, "--keystone", ks, "stats")
        if st.returncode == 0 and json.loads(st.stdout)["total_memory_pools"] == n:
            return
        time.sleep(0.1)
    raise AssertionError((st.stdout + st.stderr) if st else "no stats")


def test_aead_is_available(bb):
    ok, why = bb.aead_available()
    assert ok, why


def test_encrypted_cluster_end_to_end_and_what_a_listener_sees(secure_cluster, tmp_path):
    env = dict(secure_cluster["env"], BB_RPC_SHM="0")  # keep the control traffic on the wire, through the proxy
    ks_direct = f"127.0.0.1:{secure_cluster['rpc']}"
    wait_pools(env, ks_direct, 1)
    mitm = Mitm(secure_cluster["rpc"])
    try:
        ks = f"127.0.0.1:{mitm.port}"
        secret_key = "very-recognisable-object-name-7f3a"
        blob = tmp_path / "blob"
        blob.write_bytes(b"PLAINTEXT-MARKER-" * 4096)
        r = cli(env, "--keystone", ks, "put", secret_key, str(blob))
        assert r.returncode == 0, r.stdout + r.stderr
        out = tmp_path / "copy"
        r = cli(env, "--keystone", ks, "get", secret_key, str(out))
        assert r.returncode == 0 and out.read_bytes() == blob.read_bytes(), r.stdout + r.stderr
        up, down = bytes(mitm.up), bytes(mitm.down)
        assert len(up) > 300 and len(down) > 300
        # the handshake is visible (magic + nonces + MACs), nothing else is: no key name, no worker / pool ids, no token
        assert b"BBA2" in up[:32]
        for needle in (secret_key.encode(), b"ram-ws", b"node-ws", TOKEN.encode()):
            assert needle not in up and needle not in down
        # every frame after the handshake carries 16 bytes of tag: walk the client's frames
        pos, frames = 0, []
        while pos + 16 <= len(up):
            ln, method, rid = struct.unpack_from("<IIQ", up, pos)
            frames.append((ln, method))
            pos += 16 + ln
        assert pos == len(up) and [m for _, m in frames[:2]] == [0x7FFFFF00, 0x7FFFFF00]
        assert all(ln >= 16 for ln, _ in frames[2:]) and len(frames) > 4
    finally:
        mitm.stop()
    # a client that does not encrypt is refused by this cluster, with or without the token
    plain = {k: v for k, v in env.items() if k != "BB_ENCRYPT_TRANSPORT"}
    assert cli(plain, "--keystone", ks_direct, "stats").returncode != 0
    assert cli(env, "--keystone", ks_direct, "stats").returncode == 0
    # the Keystone's /metrics (clear text, read-only) counts both
    m = cli(env, "metrics", "--http", f"127.0.0.1:{secure_cluster['http']}")
    assert m.returncode == 0, m.stdout + m.stderr
    vals = {ln.split()[0]: float(ln.split()[1]) for ln in m.stdout.splitlines() if ln.startswith("bb_rpc_")}
    assert vals["bb_rpc_secure_handshakes_total"] >= 3 and vals["bb_rpc_auth_failures_total"] >= 1 and vals["bb_rpc_requests_total"] > 0


def test_altered_and_replayed_frames_close_the_connection(secure_cluster, bb):
    env = secure_cluster["env"]
    ks_direct = f"127.0.0.1:{secure_cluster['rpc']}"
    wait_pools(env, ks_direct, 1)
    bb.set_cluster_token(TOKEN)
    bb.set_transport_encryption(True)
    os.environ["BB_RPC_SHM"] = "0"
    try:
        # altered: one bit of a sealed request flipped in transit -> the server drops the connection, the call fails
        mitm = Mitm(secure_cluster["rpc"])
        c = bb.KeystoneRpcClient()
        assert c.connect("127.0.0.1", mitm.port, 3000) == bb.ErrorCode.OK
        assert c.object_exists("nope") is False  # sealed round trip works through the proxy
        mitm.flip_after = len(mitm.up)
        with pytest.raises(bb.BlackbirdError):
            c.object_exists("nope")
        mitm.stop()
        # replayed: the same sealed frame delivered twice -> the second copy fails authentication (the counter moved on)
        mitm = Mitm(secure_cluster["rpc"])
        c = bb.KeystoneRpcClient()
        assert c.connect("127.0.0.1", mitm.port, 3000) == bb.ErrorCode.OK
        assert c.object_exists("nope") is False
        mitm.replay = True
        c.object_exists("nope")  # this request is served (and duplicated on the wire) ...
        time.sleep(0.2)
        with pytest.raises(bb.BlackbirdError):  # ... the duplicate made the server hang up
            for _ in range(3):
                c.object_exists("nope")
        mitm.stop()
        # an altered response is caught by the client
        class FlipDown(Mitm):
            def _pump(self, a, b, upstream):
                if upstream:
                    return super()._pump(a, b, upstream)
                n = 0
                try:
                    while True:
                        d = a.recv(1 << 16)
                        if not d:
                            break
                        n += 1
                        if n == 4:  # hello reply, ack, first sealed response pass; the next one is damaged
                            d = d[:-1] + bytes([d[-1] ^ 0x80])
                        b.sendall(d)
                except OSError:
                    pass

        mitm = FlipDown(secure_cluster["rpc"])
        c = bb.KeystoneRpcClient()
        assert c.connect("127.0.0.1", mitm.port, 3000) == bb.ErrorCode.OK
        assert c.object_exists("nope") is False
        with pytest.raises(bb.BlackbirdError):
            c.object_exists("nope")
        mitm.stop()
    finally:
        os.environ.pop("BB_RPC_SHM", None)
        bb.set_transport_encryption(False)
        bb.set_cluster_token("")  # process-wide: do not leak into the other tests


def test_full_client_with_watches_over_sealed_frames(secure_cluster, bb):
    """The Python client against the encrypted cluster: put / get through the data server (gathered request, scattered
    response), and the coordination client's push channel (watch events) sealed by the server's other threads."""
    env = secure_cluster["env"]
    wait_pools(env, f"127.0.0.1:{secure_cluster['rpc']}", 1)
    bb.set_cluster_token(TOKEN)
    bb.set_transport_encryption(True)
    try:
        opts = bb.BlackbirdClientOptions()
        opts.keystone_host, opts.keystone_port = "127.0.0.1", secure_cluster["rpc"]
        cl = bb.BlackbirdClient(opts)
        assert cl.connect() == bb.ErrorCode.OK
        data = os.urandom(3 << 20)
        cfg = bb.WorkerConfig(replication_factor=1, max_workers_per_copy=1, ttl_ms=0)
        assert cl.put("sealed/obj", data, cfg) == bb.ErrorCode.OK
        assert cl.get("sealed/obj") == data
        assert cl.remove("sealed/obj") == bb.ErrorCode.OK
        # watch channel
        st = bb.RemoteCoord()
        assert st.connect(f"127.0.0.1:{secure_cluster['coord']}", 3000) == bb.ErrorCode.OK
        seen = []
        st.watch_prefix("/sec-test/", lambda t, k, v, rev: seen.append((t, k)))
        st.put("/sec-test/a", "1")
        st.delete("/sec-test/a")
        deadline = time.time() + 5
        while time.time() < deadline and len(seen) < 2:
            time.sleep(0.02)
        assert seen == [("PUT", "/sec-test/a"), ("DELETE", "/sec-test/a")]
        st.close()
    finally:
        bb.set_transport_encryption(False)
        bb.set_cluster_token("")


def test_read_only_token_reads_but_cannot_write(tmp_path):
    """A second secret (`auth_token_ro` / BB_AUTH_TOKEN_RO on the servers) admits read-only members: a client that holds only
    that token gets, lists and inspects, but every mutating call -- put, remove, migrate, a write to a data server, a write
    to the coordination store -- is refused with ACCESS_DENIED while the connection stays usable.  Works sealed as well."""
    import json

    RO = "read-only-secret"
    srv_env = dict(os.environ, BB_AUTH_TOKEN=TOKEN, BB_AUTH_TOKEN_RO=RO, BB_ENCRYPT_TRANSPORT="1")
    cport, rport, hport = free_port(), free_port(), free_port()
    procs = []

    def spawn(*cmd):
        procs.append(subprocess.Popen(list(cmd), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=srv_env))

    try:
        spawn(os.path.join(BIN, "bb-coord"), "--listen", f"127.0.0.1:{cport}")
        assert wait_port(cport)
        spawn(os.path.join(BIN, "bb-keystone"), os.path.join(ROOT, "configs", "keystone.yaml"), "--coord-endpoints", f"127.0.0.1:{cport}",
              "--listen-address", f"127.0.0.1:{rport}", "--http-port", str(hport), "--cluster-id", "roc")
        assert wait_port(rport)
        cfg = tmp_path / "w.yaml"
        cfg.write_text('worker: {worker_id: "wr", node_id: "node-wr", lease_ttl_sec: 3, heartbeat_interval_sec: 1}\n'
                       'storage_pools:\n  - {pool_id: "ram-wr", storage_class: "RAM_CPU", size_bytes: 64_MB}\n')
        spawn(os.path.join(BIN, "bb-worker"), "--config", str(cfg), "--coord-endpoints", f"127.0.0.1:{cport}", "--cluster-id", "roc")
        ks = f"127.0.0.1:{rport}"
        full = dict(os.environ, BB_AUTH_TOKEN=TOKEN, BB_ENCRYPT_TRANSPORT="1")
        ro = {k: v for k, v in os.environ.items() if k != "BB_AUTH_TOKEN"}
        ro.update(BB_AUTH_TOKEN_RO=RO, BB_ENCRYPT_TRANSPORT="1")
        wait_pools(full, ks, 1)
        blob = tmp_path / "blob"
        blob.write_bytes(os.urandom(200_000))
        assert cli(full, "--keystone", ks, "put", "ro/obj", str(blob)).returncode == 0
        # the read-only member reads everything ...
        out = tmp_path / "copy"
        r = cli(ro, "--keystone", ks, "get", "ro/obj", str(out))
        assert r.returncode == 0 and out.read_bytes() == blob.read_bytes(), r.stdout + r.stderr
        assert cli(ro, "--keystone", ks, "exists", "ro/obj").returncode == 0
        st = cli(ro, "--keystone", ks, "stats")
        assert st.returncode == 0 and json.loads(st.stdout)["total_objects"] == 1
        assert "ro/obj" in cli(ro, "--keystone", ks, "ls", "ro/").stdout
        # ... and changes nothing
        for args in (("put", "ro/new", str(blob)), ("remove", "ro/obj"), ("migrate", "ro/obj", "NVME"), ("remove-worker", "wr")):
            bad = cli(ro, "--keystone", ks, *args)
            assert bad.returncode != 0, (args, bad.stdout)
        assert cli(full, "--keystone", ks, "exists", "ro/obj").returncode == 0 and cli(full, "--keystone", ks, "exists", "ro/new").returncode != 0
        # the flag works like the environment variable; a wrong read-only token is no token
        none = {k: v for k, v in os.environ.items() if k not in ("BB_AUTH_TOKEN", "BB_AUTH_TOKEN_RO")}
        none["BB_ENCRYPT_TRANSPORT"] = "1"
        assert cli(none, "--keystone", ks, "--auth-token-ro", RO, "exists", "ro/obj").returncode == 0
        assert cli(none, "--keystone", ks, "--auth-token-ro", "guess", "exists", "ro/obj").returncode != 0
        m = cli(full, "metrics", "--http", f"127.0.0.1:{hport}")
        vals = {ln.split()[0]: float(ln.split()[1]) for ln in m.stdout.splitlines() if ln.startswith("bb_rpc_")}
        assert vals["bb_rpc_read_only_denials_total"] >= 4
        # the coordination store: a read-only member may look and listen, not write, delete or take leases
        from blackbird_b200 import _bb as bb

        bb.set_cluster_token("")
        bb.set_cluster_token_ro(RO)
        bb.set_transport_encryption(True)
        try:
            st = bb.RemoteCoord()
            assert st.connect(f"127.0.0.1:{cport}", 3000) == bb.ErrorCode.OK
            listed = lambda: [kv[0] for kv in st.get_with_prefix("/blackbird/clusters/roc/workers/")]
            assert any(k.endswith("/workers/wr") for k in listed())
            assert st.put("/blackbird/clusters/roc/evil", "x") != bb.ErrorCode.OK
            assert st.delete("/blackbird/clusters/roc/workers/wr") != bb.ErrorCode.OK
            with pytest.raises(bb.BlackbirdError):
                st.grant_lease(5)  # no lease for a read-only member
            assert any(k.endswith("/workers/wr") for k in listed()) and st.get("/blackbird/clusters/roc/evil") is None
            st.close()
        finally:
            bb.set_transport_encryption(False)
            bb.set_cluster_token_ro("")
            bb.set_cluster_token("")
    finally:
        for p in reversed(procs):
            p.terminate()
        for p in procs:
            try:
                p.wait(timeout=5)
            except subprocess.TimeoutExpired:
                p.kill()


def test_http_endpoints_need_the_bearer_token_when_one_is_set(bb, tmp_path):
    """`http_auth_token` / BB_HTTP_TOKEN / --http-token: /metrics and /stats of the Keystone and of a worker answer 401
    without `Authorization: Bearer <token>`, /healthz stays open for liveness probes, the tools send the token."""
    import urllib.error
    import urllib.request

    HTTP = "scrape-secret"
    base = {k: v for k, v in os.environ.items() if not k.startswith("BB_")}
    cport, rport, hport, wport = free_port(), free_port(), free_port(), free_port()
    procs = []
    try:
        procs.append(subprocess.Popen([os.path.join(BIN, "bb-coord"), "--listen", f"127.0.0.1:{cport}"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=base))
        assert wait_port(cport)
        procs.append(subprocess.Popen([os.path.join(BIN, "bb-keystone"), os.path.join(ROOT, "configs", "keystone.yaml"), "--coord-endpoints", f"127.0.0.1:{cport}",
                                       "--listen-address", f"127.0.0.1:{rport}", "--http-port", str(hport), "--cluster-id", "http", "--http-token", HTTP],
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=base))
        assert wait_port(rport) and wait_port(hport)
        cfg = tmp_path / "w.yaml"
        cfg.write_text(f'worker: {{worker_id: "wh", node_id: "node-wh", http_metrics_port: {wport}, http_auth_token: "{HTTP}"}}\n'
                       'storage_pools:\n  - {pool_id: "ram-wh", storage_class: "RAM_CPU", size_bytes: 16_MB}\n')
        procs.append(subprocess.Popen([os.path.join(BIN, "bb-worker"), "--config", str(cfg), "--coord-endpoints", f"127.0.0.1:{cport}", "--cluster-id", "http"],
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=base))
        assert wait_port(wport)

        def get(port, path, token=None, scheme="Bearer"):
            req = urllib.request.Request(f"http://127.0.0.1:{port}{path}")
            if token is not None:
                req.add_header("Authorization", f"{scheme} {token}")
            try:
                with urllib.request.urlopen(req, timeout=5) as r:
                    return r.status, r.read().decode(), dict(r.headers)
            except urllib.error.HTTPError as e:
                return e.code, e.read().decode(), dict(e.headers)

        for port in (hport, wport):
            st, body, hdr = get(port, "/metrics")
            assert st == 401 and "bb_" not in body and hdr.get("WWW-Authenticate") == "Bearer"
            assert get(port, "/stats")[0] == 401
            assert get(port, "/metrics", "guess")[0] == 401 and get(port, "/metrics", HTTP + "x")[0] == 401 and get(port, "/metrics", "")[0] == 401
            assert get(port, "/metrics", HTTP, scheme="Basic")[0] == 401
            st, body, _ = get(port, "/metrics", HTTP)
            assert st == 200 and "bb_" in body
            assert get(port, "/metrics", HTTP, scheme="bearer")[0] == 200  # the scheme is case-insensitive
            assert get(port, "/healthz")[0] == 200  # liveness probes carry no credentials
            assert get(port, "/nope", HTTP)[0] == 404 and get(port, "/nope")[0] == 401  # unknown paths tell a stranger nothing either
        # the tools: flag, environment, and the in-process helper
        assert cli(base, "metrics", "--http", f"127.0.0.1:{hport}").returncode != 0
        r = cli(base, "metrics", "--http", f"127.0.0.1:{hport}", "--http-token", HTTP)
        assert r.returncode == 0 and "bb_put_start_total" in r.stdout
        assert "bb_put_start_total" in cli(dict(base, BB_HTTP_TOKEN=HTTP), "metrics", "--http", f"127.0.0.1:{hport}").stdout
        assert bb.http_get("127.0.0.1", hport, "/metrics")[0] == 401
        bb.set_http_token(HTTP)
        try:
            assert bb.http_get("127.0.0.1", hport, "/metrics")[0] == 200
        finally:
            bb.set_http_token("")
    finally:
        for p in reversed(procs):
            p.terminate()
        for p in procs:
            try:
                p.wait(timeout=5)
            except subprocess.TimeoutExpired:
                p.kill()
