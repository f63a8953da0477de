"""Robustness of every network-facing decoder: the Keystone RPC server, the coordination daemon, the worker data
server and the HTTP endpoint are fed random, truncated, oversized and structure-aware garbage over real sockets; each
server must survive (no crash, no hang, no leak of the connection slot) and keep serving well-formed clients.  Under
`BB_SANITIZE=asan` (build.py) the same test is the ASAN/UBSAN harness for the wire codec."""
import os
import random
import socket
import struct

from blackbird_b200.parallel import LocalCluster


def _frame(method, rid, payload):
    return struct.pack("<IIQ", len(payload), method & 0xFFFFFFFF, rid) + payload


def _rand_payload(rng):
    kind = rng.randrange(5)
    if kind == 0:
        return rng.randbytes(rng.randrange(0, 200))
    if kind == 1:  # length-prefixed strings with lying prefixes
        out = b""
        for _ in range(rng.randrange(1, 6)):
            s = rng.randbytes(rng.randrange(0, 24))
            n = len(s) if rng.random() < 0.6 else rng.choice([0, 1, 0xFFFFFFFF, 0x7FFFFFFF, len(s) + 1, 1 << 20])
            out += struct.pack("<I", n) + s
        return out
    if kind == 2:  # huge element counts
        return struct.pack("<I", rng.choice([0xFFFFFFFF, 1 << 31, 1 << 24])) + rng.randbytes(rng.randrange(0, 64))
    if kind == 3:  # plausible put_start: key, size, then junk config
        key = b"fuzz/" + rng.randbytes(4).hex().encode()
        return struct.pack("<I", len(key)) + key + struct.pack("<Q", rng.choice([0, 1, 4096, 1 << 62, (1 << 64) - 1])) + rng.randbytes(rng.randrange(0, 80))
    return b"\x00" * rng.randrange(0, 64)


def _blast(port, rng, rounds, methods):
    for _ in range(rounds):
        s = socket.create_connection(("127.0.0.1", port), 2.0)
        s.settimeout(0.3)
        try:
            mode = rng.randrange(4)
            if mode == 0:  # raw noise, no framing
                s.sendall(rng.randbytes(rng.randrange(1, 400)))
            elif mode == 1:  # well-framed garbage, several requests per connection
                for _ in range(rng.randrange(1, 8)):
                    s.sendall(_frame(rng.choice(methods), rng.getrandbits(64), _rand_payload(rng)))
                try:
                    s.recv(1 << 16)
                except OSError:
                    pass
            elif mode == 2:  # header promises more than is sent, then the peer goes away
                p = _rand_payload(rng)
                s.sendall(struct.pack("<IIQ", len(p) + rng.randrange(1, 1 << 20), rng.choice(methods), 7) + p)
            else:  # length above the frame limit
                s.sendall(struct.pack("<IIQ", rng.choice([0xFFFFFFFF, (256 << 20) + 1]), rng.choice(methods), 9))
                try:
                    s.recv(16)
                except OSError:
                    pass
        except OSError:
            pass  # the server may close on us at any point
        finally:
            s.close()


def test_keystone_rpc_and_http_survive_garbage(bb):
    rng = random.Random(0xB200)
    with LocalCluster(cluster_id="fuzz", n_workers=2) as c:
        methods = list(range(0, 40)) + [0x7FFFFFFF, 0x80000001, 0xFFFFFFFF]
        _blast(c.rpc.rpc_port, rng, 300, methods)
        cli = c.client()
        wc = bb.WorkerConfig(replication_factor=1, max_workers_per_copy=1)
        data = os.urandom(5000)
        assert cli.put("after-fuzz", data, wc) == bb.ErrorCode.OK
        assert cli.get("after-fuzz") == data
        # garbage puts must not have leaked allocator space: only our object is accounted
        st = cli.cluster_stats()
        live = st.total_objects
        assert live >= 1 and st.used_capacity <= (live + 1) * (1 << 20)
        # HTTP endpoint: malformed request lines, giant headers, binary noise
        for req in (b"GET\r\n\r\n", b"\x00\xff" * 100, b"GET /metrics HTTP/1.1\r\n" + b"X: " + b"a" * 70000 + b"\r\n\r\n",
                    b"POST /metrics HTTP/1.1\r\nContent-Length: 99999999\r\n\r\n", b"GET /" + b"%ff" * 3000 + b" HTTP/1.1\r\n\r\n"):
            s = socket.create_connection(("127.0.0.1", c.rpc.http_port), 2.0)
            s.settimeout(0.5)
            try:
                s.sendall(req)
                s.recv(4096)
            except OSError:
                pass
            s.close()
        status, body = bb.http_get("127.0.0.1", c.rpc.http_port, "/metrics")
        assert status == 200 and "bb_objects" in body


def test_worker_data_server_survives_garbage(bb):
    rng = random.Random(7)
    with LocalCluster(cluster_id="fuzz-data", n_workers=1) as c:
        host, port = c.workers[0].data_endpoint().rsplit(":", 1)
        _blast(int(port), rng, 200, list(range(0, 10)))
        cli = c.client()
        data = os.urandom(70000)
        assert cli.put("k", data, bb.WorkerConfig(replication_factor=1, max_workers_per_copy=1)) == bb.ErrorCode.OK
        assert cli.get("k") == data


def test_coord_daemon_survives_garbage(bb):
    rng = random.Random(99)
    srv = bb.CoordServer()
    assert srv.start("127.0.0.1", 0) == bb.ErrorCode.OK
    try:
        _blast(srv.port, rng, 200, list(range(0, 32)))
        cs = bb.CoordService(f"tcp://127.0.0.1:{srv.port}")
        assert cs.connect() == bb.ErrorCode.OK
        assert cs.put("/fuzz/key", "v") == bb.ErrorCode.OK
        assert cs.get("/fuzz/key") == b"v"
    finally:
        srv.stop()


def test_token_handshake_survives_garbage_and_never_admits_it(bb):
    """A token-gated server under handshake-shaped noise: hellos of every length, proofs without a hello, repeated
    hellos, random and truncated MACs, ordinary requests in between.  Nothing but the denial marker (or a hang-up)
    ever comes back for a request, and a client that holds the token still gets in afterwards."""
    AUTH, DENIED = 0x7FFFFF00, 0x7FFFFFFD
    rng = random.Random(0xA07)
    bb.set_cluster_token("fuzz-token")
    srv = bb.CoordServer()
    try:
        assert srv.start("127.0.0.1", 0) == bb.ErrorCode.OK
        for _ in range(300):
            s = socket.create_connection(("127.0.0.1", srv.port), 2.0)
            s.settimeout(0.3)
            try:
                for _ in range(rng.randrange(1, 5)):
                    kind = rng.randrange(6)
                    if kind == 0:
                        body = b"BBA1" + rng.randbytes(rng.choice([0, 1, 15, 16, 17, 64]))
                    elif kind == 1:
                        body = rng.randbytes(32)  # a proof out of the blue, or a wrong one after a hello
                    elif kind == 2:
                        body = rng.randbytes(rng.randrange(0, 100))
                    elif kind == 3:
                        body = b"fuzz-token"  # the secret itself is not a credential
                    elif kind == 4:
                        body = rng.choice([b"BBA1", b"BBR1", b"BBR2", b"BBA2"]) + rng.randbytes(16)  # (no read-only token here: BBRx is refused)
                    else:
                        s.sendall(_frame(rng.randrange(0, 32), 3, _rand_payload(rng)))  # a request before being admitted
                        continue
                    s.sendall(_frame(AUTH, rng.getrandbits(64), body))
                got = b""
                try:
                    while len(got) < 4096:
                        part = s.recv(4096)
                        if not part:
                            break
                        got += part
                except OSError:
                    pass
                # every complete frame that came back is a handshake reply or the denial marker -- never an RPC response
                pos = 0
                while len(got) - pos >= 16:
                    n, method, _ = struct.unpack_from("<IIQ", got, pos)
                    assert method in (AUTH, DENIED), hex(method)
                    assert method != AUTH or n in (0, 48)
                    if method == AUTH and n == 0:
                        raise AssertionError("admitted without a valid proof")
                    pos += 16 + n
            except OSError:
                pass
            finally:
                s.close()
        cs = bb.CoordService(f"tcp://127.0.0.1:{srv.port}")
        assert cs.connect() == bb.ErrorCode.OK and cs.put("/fuzz/auth", "v") == bb.ErrorCode.OK and cs.get("/fuzz/auth") == b"v"
    finally:
        bb.set_cluster_token("")
        srv.stop()


def _parse(fn, text):
    try:
        return fn(text)
    except ValueError:  # a parse error is reported as ValueError by the bindings
        return None


def test_json_and_yaml_parsers_never_crash_on_mutated_documents(bb):
    """The in-tree JSON / YAML-subset parsers read worker records from the coordination store and operator configs:
    mutated documents must come back as a value or as None (parse error), never crash or hang; deep nesting is bounded."""
    rng = random.Random(1234)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    seeds = [open(os.path.join(root, "configs", f)).read() for f in ("keystone.yaml", "worker.yaml", "cxl_worker.yaml", "tiered_worker.yaml")]
    seeds += ['{"id":"p","size":1048576,"used":0,"storage_class":"RAM_CPU","ucx_rkey_hex":"00ff","nested":{"a":[1,2.5e3,-0,true,null,"\\u00e9\\n"]}}',
              '[[[[[[1]]]]]]', '{"a":{"b":{"c":{"d":{}}}}}']
    alphabet = b'{}[]:,"\\\n\t -#&*!|>\'%@`0123456789eE.+truefalsn\x00\xff'
    for _ in range(3000):
        doc = bytearray(rng.choice(seeds).encode())
        for _ in range(rng.randrange(1, 8)):
            op = rng.randrange(4)
            pos = rng.randrange(len(doc) + 1)
            if op == 0 and doc:
                del doc[pos % len(doc)]
            elif op == 1:
                doc.insert(pos, rng.choice(alphabet))
            elif op == 2 and doc:
                doc[pos % len(doc)] = rng.choice(alphabet)
            else:
                a = rng.randrange(len(doc) + 1)
                doc[pos:pos] = doc[a:a + rng.randrange(0, 40)]
        text = doc.decode("latin-1")
        _parse(bb.parse_json, text)
        _parse(bb.parse_yaml, text)
    # pathological nesting: rejected (or parsed) without exhausting the stack
    for depth in (100, 10_000, 200_000):
        _parse(bb.parse_json, "[" * depth + "]" * depth)
        _parse(bb.parse_json, "[" * depth)
        _parse(bb.parse_json, '{"a":' * depth + "1" + "}" * depth)
        _parse(bb.parse_yaml, "a:\n" + "".join(" " * (i + 1) + "b:\n" for i in range(min(depth, 5000))))
        _parse(bb.parse_yaml, "[" * depth + "]" * depth)


def test_json_parser_agrees_with_python_on_generated_documents(bb):
    """Differential test: documents generated by hypothesis and serialised by Python's json module parse to the same
    value in the in-tree parser, and json_roundtrip (parse + dump) is a fixed point that Python reads back."""
    import json

    from hypothesis import given, settings
    from hypothesis import strategies as st

    leaves = st.one_of(st.none(), st.booleans(), st.integers(-(2 ** 53), 2 ** 53), st.floats(allow_nan=False, allow_infinity=False, width=64),
                       st.text(max_size=20))
    docs = st.recursive(leaves, lambda ch: st.one_of(st.lists(ch, max_size=5), st.dictionaries(st.text(max_size=8), ch, max_size=5)), max_leaves=25)

    def same(a, b):
        if isinstance(a, float) or isinstance(b, float):
            return isinstance(a, (int, float)) and isinstance(b, (int, float)) and (a == b or abs(a - b) <= 1e-9 * max(abs(a), abs(b)))
        if isinstance(a, dict):
            return isinstance(b, dict) and a.keys() == b.keys() and all(same(a[k], b[k]) for k in a)
        if isinstance(a, list):
            return isinstance(b, list) and len(a) == len(b) and all(same(x, y) for x, y in zip(a, b))
        return a == b and type(a) is type(b)

    @settings(max_examples=300, deadline=None)
    @given(docs)
    def check(doc):
        text = json.dumps(doc, ensure_ascii=False)
        assert same(bb.parse_json(text), doc), text
        assert same(bb.parse_json(json.dumps(doc, ensure_ascii=True)), doc)  # \\uXXXX escapes incl. surrogate pairs
        again = bb.json_roundtrip(text)
        assert same(json.loads(again), doc) and bb.json_roundtrip(again) == again

    check()


class _RecordingProxy:
    """TCP proxy that records every framed request a real client sends (structure-aware seeds for the mutator)."""

    def __init__(self, target_port):
        import threading

        self.target = target_port
        self.frames = []
        self.srv = socket.socket()
        self.srv.bind(("127.0.0.1", 0))
        self.srv.listen(8)
        self.port = self.srv.getsockname()[1]
        self.run = True
        self.t = threading.Thread(target=self._accept, daemon=True)
        self.t.start()

    def _accept(self):
        import threading

        self.srv.settimeout(0.2)
        while self.run:
            try:
                c, _ = self.srv.accept()
            except OSError:
                continue
            c.settimeout(None)  # the accepted socket inherits the listener's timeout
            u = socket.create_connection(("127.0.0.1", self.target))
            threading.Thread(target=self._pump, args=(c, u, True), daemon=True).start()
            threading.Thread(target=self._pump, args=(u, c, False), daemon=True).start()

    def _pump(self, a, b, record):
        buf = b""
        try:
            while True:
                d = a.recv(1 << 16)
                if not d:
                    break
                b.sendall(d)
                if record:
                    buf += d
                    while len(buf) >= 16:
                        n = struct.unpack_from("<I", buf)[0]
                        if len(buf) < 16 + n:
                            break
                        self.frames.append(buf[:16 + n])
                        buf = buf[16 + n:]
        except OSError:
            pass
        except Exception as e:  # noqa: BLE001
            print("proxy pump died:", repr(e), flush=True)
        finally:
            for s in (a, b):
                try:
                    s.close()
                except OSError:
                    pass

    def stop(self):
        self.run = False
        self.srv.close()


def _mutate(rng, frame):
    f = bytearray(frame)
    for _ in range(rng.randrange(1, 6)):
        op = rng.randrange(6)
        if op == 0 and len(f) > 16:
            f[rng.randrange(16, len(f))] ^= 1 << rng.randrange(8)                 # bit flip in the payload
        elif op == 1 and len(f) > 20:
            pos = rng.randrange(16, len(f) - 3)
            f[pos:pos + 4] = struct.pack("<I", rng.choice([0, 1, 0xFFFFFFFF, 0x7FFFFFFF, 1 << 24]))  # clobber a length / count
        elif op == 2 and len(f) > 17:
            del f[rng.randrange(16, len(f)):]                                      # truncate the payload
        elif op == 3:
            f += rng.randbytes(rng.randrange(1, 64))                               # trailing garbage
        elif op == 4:
            f[4:8] = struct.pack("<I", rng.randrange(0, 40))                       # same payload, other method
        else:
            a = rng.randrange(16, len(f) + 1)
            f[a:a] = f[16:16 + rng.randrange(0, 32)]                               # duplicate a slice
    struct.pack_into("<I", f, 0, max(0, len(f) - 16) if rng.random() < 0.8 else rng.choice([0, len(f), 1 << 20]))
    return bytes(f)


def test_mutated_real_requests_do_not_break_keystone_or_worker(bb, monkeypatch):
    """Structure-aware fuzzing: the requests of a real client session (put / get / batch / admin calls, D_WRITE / D_READ)
    are recorded through a proxy, mutated (bit flips, clobbered lengths and counts, truncation, method swaps) and fired at
    the Keystone and at the worker data server; both keep serving.  Under the ASAN build this covers the decoders of
    every wire struct with inputs that are almost valid."""
    rng = random.Random(2026)
    monkeypatch.setenv("BB_RPC_SHM", "0")  # the recording proxy only sees TCP: keep the session off the shared-memory channel
    with LocalCluster(cluster_id="fuzz2", n_workers=2) as c:
        kproxy = _RecordingProxy(c.rpc.rpc_port)
        o = bb.BlackbirdClientOptions("127.0.0.1", kproxy.port, 30000, 2, "node-0")
        o.enable_shm = False
        cl = bb.BlackbirdClient(o)
        assert cl.connect() == bb.ErrorCode.OK
        cfg = bb.WorkerConfig(replication_factor=2, max_workers_per_copy=2, ttl_ms=5000, preferred_classes=[bb.StorageClass.RAM_CPU])
        data = os.urandom(30_000)
        assert cl.put("seed/a", data, cfg) == bb.ErrorCode.OK and cl.get("seed/a") == data
        assert cl.batch_put(["seed/b", "seed/c"], [data[:1000], data[:2000]], cfg) == [bb.ErrorCode.OK] * 2
        cl.batch_get(["seed/b", "seed/c", "missing"])
        cl.batch_exists(["seed/a", "nope"])
        cl.object_exists("seed/a")
        cl.cluster_stats()
        api = cl.keystone()
        api.get_memory_pools(), api.get_workers_info(), api.list_objects("seed/", 10, ""), api.get_view_version()
        cl.batch_remove(["seed/b"])
        kframes = list(kproxy.frames)
        kproxy.stop()
        assert len(kframes) >= 12
        # data-server seeds: a D_WRITE and a D_READ built from a real placement
        sh = cl.get_workers("seed/a")[0].shards[0]
        pool = sh.pool_id.encode()
        addr = sh.location["remote_addr"] | (1 << 63)
        dframes = [_frame(1, 1, struct.pack("<I", len(pool)) + pool + struct.pack("<QI", addr, 64) + data[:64]),
                   _frame(2, 2, struct.pack("<I", len(pool)) + pool + struct.pack("<QI", addr, 4096))]
        dport = int(sh.endpoint.port)

        def fire(port, frames, rounds):
            for _ in range(rounds):
                s = socket.create_connection(("127.0.0.1", port), 2.0)
                s.settimeout(0.05)
                try:
                    for _ in range(rng.randrange(1, 5)):
                        s.sendall(_mutate(rng, rng.choice(frames)))
                    try:
                        s.recv(1 << 16)
                    except OSError:
                        pass
                except OSError:
                    pass
                finally:
                    s.close()

        fire(c.rpc.rpc_port, kframes, 250)
        fire(dport, dframes, 150)
        fresh = c.client()
        blob = os.urandom(12345)
        assert fresh.put("after", blob, bb.WorkerConfig(replication_factor=1, max_workers_per_copy=1)) == bb.ErrorCode.OK
        assert fresh.get("after") == blob
        assert bb.http_get("127.0.0.1", c.rpc.http_port, "/healthz")[0] == 200


def test_secure_mode_survives_garbage_after_a_valid_handshake(bb):
    """Secure mode (BBA2): a peer that holds the token completes the handshake by hand and then sends noise instead of
    sealed frames -- random bytes, frames with random tags, truncated frames, frames shorter than a tag, huge length
    fields.  The server must drop every such connection without answering a request, stay up, and keep serving a real
    client (which proves its keys and counters were not disturbed by the other connections)."""
    import hashlib
    import hmac as pyhmac

    AUTH, tok = 0x7FFFFF00, b"fuzz-secure-token"
    rng = random.Random(0x5EC)
    bb.set_cluster_token(tok.decode())
    bb.set_transport_encryption(True)
    srv = bb.CoordServer()
    try:
        assert srv.start("127.0.0.1", 0) == bb.ErrorCode.OK

        def recv_frame(s):
            hdr = b""
            while len(hdr) < 16:
                part = s.recv(16 - len(hdr))
                if not part:
                    return None
                hdr += part
            n, method, rid = struct.unpack("<IIQ", hdr)
            body = b""
            while len(body) < n:
                part = s.recv(n - len(body))
                if not part:
                    return None
                body += part
            return method, body

        for _ in range(120):
            s = socket.create_connection(("127.0.0.1", srv.port), 2.0)
            s.settimeout(0.5)
            try:
                cn = rng.randbytes(16)
                s.sendall(_frame(AUTH, 0, b"BBA2" + cn))
                method, body = recv_frame(s)
                assert method == AUTH and len(body) == 48
                sn, mac = body[:16], body[16:]
                assert pyhmac.compare_digest(mac, pyhmac.new(tok, b"bb-srv" + cn + sn, hashlib.sha256).digest())
                s.sendall(_frame(AUTH, 1, pyhmac.new(tok, b"bb-cli" + cn + sn, hashlib.sha256).digest()))
                method, body = recv_frame(s)
                assert method == AUTH and body == b""  # admitted: from here on everything must be sealed
                kind = rng.randrange(6)
                if kind == 0:
                    junk = rng.randbytes(rng.randrange(1, 400))
                elif kind == 1:
                    junk = _frame(rng.randrange(0, 32), 7, rng.randbytes(rng.randrange(16, 200)))  # random "ciphertext + tag"
                elif kind == 2:
                    junk = _frame(rng.randrange(0, 32), 7, rng.randbytes(rng.randrange(0, 16)))    # shorter than a tag
                elif kind == 3:
                    junk = struct.pack("<IIQ", 0x7FFFFFFF, 3, 9) + rng.randbytes(64)                # absurd length field
                elif kind == 4:
                    junk = _frame(3, 7, rng.randbytes(64))[:40]                                      # truncated, then silence
                else:
                    junk = _frame(3, 7, b"\0" * 48) * 3                                              # several frames, all forged
                s.sendall(junk)
                got = b""
                try:
                    while len(got) < 4096:
                        part = s.recv(4096)
                        if not part:
                            break
                        got += part
                except OSError:
                    pass
                assert got == b"", f"the server answered {len(got)} bytes to a forged frame (kind {kind})"
            finally:
                s.close()
        cs = bb.CoordService(f"tcp://127.0.0.1:{srv.port}")
        assert cs.connect() == bb.ErrorCode.OK and cs.put("/fuzz/secure", "v") == bb.ErrorCode.OK and cs.get("/fuzz/secure") == b"v"
    finally:
        bb.set_transport_encryption(False)
        bb.set_cluster_token("")
        srv.stop()


def test_client_decoders_survive_a_server_that_answers_garbage(bb, monkeypatch):
    """The other direction: a broken (or hostile) server answers the client's batch calls with noise -- huge counts,
    a delta-encoded placement without anything to be a delta of, truncated replies, random bytes.  The client returns
    errors (per item or for the call); it does not crash, hang or allocate by an unchecked count."""
    import threading

    monkeypatch.setenv("BB_RPC_SHM", "0")
    rng = random.Random(0xC11E)
    OK = struct.pack("<i", 0)
    replies = [
        OK + struct.pack("<I", 0xFFFFFFF0),                                  # count far beyond the payload
        OK + struct.pack("<I", 2) + OK + b"\x01" + struct.pack("<QQ", 4096, 7),  # delta placement without a base
        OK + struct.pack("<I", 3) + OK + b"\x07",                            # unknown placement tag
        OK + struct.pack("<I", 1) + OK + b"\x00" + struct.pack("<I", 0xFFFFFF),  # full placement claiming 16 M copies
        OK,                                                                  # truncated right after the status
        b"",
    ] + [rng.randbytes(rng.randrange(1, 300)) for _ in range(40)]
    srv = socket.socket()
    srv.bind(("127.0.0.1", 0))
    srv.listen(4)
    port = srv.getsockname()[1]
    state = {"i": 0, "run": True}

    def serve():
        srv.settimeout(0.2)
        while state["run"]:
            try:
                c, _ = srv.accept()
            except OSError:
                continue
            threading.Thread(target=answer, args=(c,), daemon=True).start()  # a client that gives up on a connection opens the next

    def answer(c):
        if True:
            c.settimeout(2.0)
            try:
                while True:
                    hdr = b""
                    while len(hdr) < 16:
                        part = c.recv(16 - len(hdr))
                        if not part:
                            raise OSError
                        hdr += part
                    n, method, rid = struct.unpack("<IIQ", hdr)
                    left = n
                    while left:
                        part = c.recv(min(left, 1 << 16))
                        if not part:
                            raise OSError
                        left -= len(part)
                    body = replies[state["i"] % len(replies)]
                    state["i"] += 1
                    c.sendall(struct.pack("<IIQ", len(body), method, rid) + body)
            except OSError:
                pass
            finally:
                c.close()

    t = threading.Thread(target=serve, daemon=True)
    t.start()
    try:
        keys = [f"k{i}" for i in range(5)]
        cfg = bb.WorkerConfig(replication_factor=1, max_workers_per_copy=1)
        for _ in range(len(replies) * 2):
            api = bb.KeystoneRpcClient()
            assert api.connect("127.0.0.1", port, 2000) == bb.ErrorCode.OK
            for call in (lambda: api.batch_get_workers(keys), lambda: api.batch_put_start(keys, [4096] * 5, cfg),
                         lambda: api.batch_object_exists(keys), lambda: api.batch_put_complete(keys)):
                try:
                    res = call()
                    assert len(res) == len(keys)  # per-item results, mostly errors
                except bb.BlackbirdError:
                    pass
    finally:
        state["run"] = False
        t.join(timeout=2)
        srv.close()


def test_data_server_checks_a_read_length_before_sizing_anything_from_it(bb):
    """ADVICE r1 class of defect, found again by the mutated-frame fuzzer under TSAN: a D_READ whose 32-bit length is far
    beyond the pool used to zero-fill a reply buffer of that size (4 GiB, 1-2 s of a data-server thread) and only then fail the
    range check.  The length is validated first now: the answer is immediate and carries no payload."""
    import resource
    import time

    with LocalCluster(cluster_id="fuzz3", n_workers=1) as c:
        cl = c.client()
        assert cl.put("seed", os.urandom(4096), bb.WorkerConfig(replication_factor=1, max_workers_per_copy=1, preferred_classes=[bb.StorageClass.RAM_CPU])) == bb.ErrorCode.OK
        sh = cl.get_workers("seed")[0].shards[0]
        pool = sh.pool_id.encode()
        addr = sh.location["remote_addr"] | (1 << 63)
        rss0 = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss
        for n in (0xFFFFFFF0, 0x80000000, (64 << 20) + 1):
            s = socket.create_connection(("127.0.0.1", int(sh.endpoint.port)), 2.0)
            s.settimeout(5.0)
            t0 = time.time()
            s.sendall(_frame(2, 9, struct.pack("<I", len(pool)) + pool + struct.pack("<QI", addr, n)))
            hdr = s.recv(16)
            ln, method, rid = struct.unpack("<IIQ", hdr)
            body = s.recv(64)
            s.close()
            assert time.time() - t0 < 0.5 and ln == 4 and struct.unpack("<i", body[:4])[0] == int(bb.ErrorCode.INVALID_PARAMETERS.value)
        assert resource.getrusage(resource.RUSAGE_SELF).ru_maxrss - rss0 < (256 << 10)  # KiB: nothing like a 4 GiB buffer was touched
