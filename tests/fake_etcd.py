"""A small in-memory etcd v3 *JSON gateway* (the HTTP/1.1 API every etcd >= 3.4 serves next to gRPC): enough of
/v3/kv/{put,range,deleterange,txn}, /v3/lease/{grant,keepalive,revoke,timetolive} and the streaming /v3/watch to run the
EtcdCoord adapter, a Keystone and workers against it.  Wire conventions follow etcd's grpc-gateway: `bytes` are base64,
64-bit integers are decimal strings, zero-valued fields are omitted, a watch is a chunked response with one JSON object
per line.  No etcd binary exists in this environment; this is the stand-in the adapter's tests talk to."""
from __future__ import annotations

import base64
import json
import threading
import time
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer


def b64(b: bytes) -> str:
    return base64.b64encode(b).decode()


def unb64(s) -> bytes:
    return base64.b64decode(s) if s else b""


class Store:
    def __init__(self):
        self.mu = threading.Condition()
        self.rev = 1
        self.kv = {}        # key -> dict(value, create, mod, version, lease)
        self.leases = {}    # id -> dict(ttl, expires, keys)
        self.next_lease = 7000
        self.history = []   # (rev, type, key, kv-dict, prev-kv-dict)
        self.requests = []  # (path, body) log for assertions
        self.stop = False
        threading.Thread(target=self._expire_loop, daemon=True).start()

    # ---- helpers (call with self.mu held)
    def _kv_json(self, key, e):
        j = {"key": b64(key), "create_revision": str(e["create"]), "mod_revision": str(e["mod"]), "version": str(e["version"]), "value": b64(e["value"])}
        if e["lease"]:
            j["lease"] = str(e["lease"])
        return j

    def _put(self, key, value, lease):
        if lease and lease not in self.leases:
            raise KeyError("lease not found")
        self.rev += 1
        prev = self.kv.get(key)
        if prev and prev["lease"] and prev["lease"] != lease and prev["lease"] in self.leases:
            self.leases[prev["lease"]]["keys"].discard(key)
        e = {"value": value, "create": prev["create"] if prev else self.rev, "mod": self.rev, "version": (prev["version"] + 1) if prev else 1, "lease": lease}
        self.kv[key] = e
        if lease:
            self.leases[lease]["keys"].add(key)
        self.history.append((self.rev, "PUT", key, self._kv_json(key, e), self._kv_json(key, prev) if prev else None))
        self.mu.notify_all()

    def _delete(self, key):
        prev = self.kv.pop(key, None)
        if prev is None:
            return 0
        self.rev += 1
        if prev["lease"] in self.leases:
            self.leases[prev["lease"]]["keys"].discard(key)
        self.history.append((self.rev, "DELETE", key, {"key": b64(key), "mod_revision": str(self.rev)}, self._kv_json(key, prev)))
        self.mu.notify_all()
        return 1

    def _range_keys(self, key, end):
        if not end:
            return [key] if key in self.kv else []
        if end == b"\0":
            return sorted(k for k in self.kv if k >= key)
        return sorted(k for k in self.kv if key <= k < end)

    def _revoke(self, lid):
        lease = self.leases.pop(lid, None)
        if lease is None:
            return False
        for k in sorted(lease["keys"]):
            e = self.kv.get(k)
            if e and e["lease"] == lid:
                self._delete(k)
        return True

    def _expire_loop(self):
        while not self.stop:
            time.sleep(0.02)
            with self.mu:
                now = time.monotonic()
                for lid in [i for i, le in self.leases.items() if le["expires"] <= now]:
                    self._revoke(lid)

    def header(self):
        return {"cluster_id": "1", "member_id": "1", "revision": str(self.rev), "raft_term": "2"}


class Handler(BaseHTTPRequestHandler):
    protocol_version = "HTTP/1.1"
    store: Store = None  # set per server

    def log_message(self, *a):  # quiet
        pass

    def _reply(self, obj, status=200):
        body = json.dumps(obj).encode()
        self.send_response(status)
        self.send_header("Content-Type", "application/json")
        self.send_header("Content-Length", str(len(body)))
        self.end_headers()
        self.wfile.write(body)

    def do_POST(self):
        st = self.store
        n = int(self.headers.get("Content-Length", "0"))
        try:
            q = json.loads(self.rfile.read(n) or b"{}")
        except ValueError:
            return self._reply({"error": "bad json", "code": 3}, 400)
        path = self.path
        with st.mu:
            st.requests.append((path, q))
        try:
            if path == "/debug/paths":  # test hook: which endpoints have been called
                with st.mu:
                    return self._reply({"paths": sorted({p for p, _ in st.requests})})
            if path == "/v3/watch":
                return self._watch(q)
            with st.mu:
                if path == "/v3/kv/put":
                    st._put(unb64(q.get("key")), unb64(q.get("value")), int(q.get("lease", 0)))
                    return self._reply({"header": st.header()})
                if path == "/v3/kv/range":
                    keys = st._range_keys(unb64(q.get("key")), unb64(q.get("range_end")))
                    out = {"header": st.header()}
                    if keys and not q.get("count_only"):
                        out["kvs"] = [st._kv_json(k, st.kv[k]) for k in keys]
                    if keys:
                        out["count"] = str(len(keys))
                    return self._reply(out)
                if path == "/v3/kv/deleterange":
                    keys = st._range_keys(unb64(q.get("key")), unb64(q.get("range_end")))
                    d = sum(st._delete(k) for k in keys)
                    out = {"header": st.header()}
                    if d:
                        out["deleted"] = str(d)
                    return self._reply(out)
                if path == "/v3/kv/txn":
                    ok = all(self._compare(c) for c in q.get("compare", []))
                    for op in q.get("success" if ok else "failure", []):
                        if "request_put" in op:
                            p = op["request_put"]
                            st._put(unb64(p.get("key")), unb64(p.get("value")), int(p.get("lease", 0)))
                        elif "request_delete_range" in op:
                            p = op["request_delete_range"]
                            for k in st._range_keys(unb64(p.get("key")), unb64(p.get("range_end"))):
                                st._delete(k)
                    out = {"header": st.header()}
                    if ok:
                        out["succeeded"] = True
                    return self._reply(out)
                if path == "/v3/lease/grant":
                    ttl = int(q.get("TTL", 0))
                    st.next_lease += 1
                    st.leases[st.next_lease] = {"ttl": ttl, "expires": time.monotonic() + ttl, "keys": set()}
                    return self._reply({"header": st.header(), "ID": str(st.next_lease), "TTL": str(ttl)})
                if path == "/v3/lease/keepalive":
                    lid = int(q.get("ID", 0))
                    le = st.leases.get(lid)
                    res = {"header": st.header(), "ID": str(lid)}
                    if le:
                        le["expires"] = time.monotonic() + le["ttl"]
                        res["TTL"] = str(le["ttl"])
                    return self._reply({"result": res})
                if path in ("/v3/lease/revoke", "/v3/kv/lease/revoke"):
                    if not st._revoke(int(q.get("ID", 0))):
                        return self._reply({"error": "etcdserver: requested lease not found", "code": 5}, 404)
                    return self._reply({"header": st.header()})
                if path == "/v3/lease/timetolive":
                    lid = int(q.get("ID", 0))
                    le = st.leases.get(lid)
                    if not le:
                        return self._reply({"header": st.header(), "ID": str(lid), "TTL": "-1"})
                    return self._reply({"header": st.header(), "ID": str(lid), "TTL": str(max(0, int(le["expires"] - time.monotonic()))), "grantedTTL": str(le["ttl"])})
        except KeyError as e:
            return self._reply({"error": f"etcdserver: {e.args[0]}", "code": 5}, 404)
        return self._reply({"error": "not found", "code": 5}, 404)

    def _compare(self, c):
        st = self.store
        e = st.kv.get(unb64(c.get("key")))
        target, result = c.get("target", "VERSION"), c.get("result", "EQUAL")
        if target == "CREATE":
            have, want = (e["create"] if e else 0), int(c.get("create_revision", 0))
        elif target == "MOD":
            have, want = (e["mod"] if e else 0), int(c.get("mod_revision", 0))
        elif target == "VALUE":
            if e is None:
                return False
            have, want = e["value"], unb64(c.get("value"))
        else:
            have, want = (e["version"] if e else 0), int(c.get("version", 0))
        return {"EQUAL": have == want, "NOT_EQUAL": have != want, "GREATER": have > want, "LESS": have < want}[result]

    def _watch(self, q):
        st = self.store
        cr = q.get("create_request", {})
        key, end = unb64(cr.get("key")), unb64(cr.get("range_end"))
        want_prev = bool(cr.get("prev_kv"))
        self.send_response(200)
        self.send_header("Content-Type", "application/json")
        self.send_header("Transfer-Encoding", "chunked")
        self.end_headers()

        def chunk(obj):
            data = (json.dumps(obj) + "\n").encode()
            self.wfile.write(b"%x\r\n%s\r\n" % (len(data), data))
            self.wfile.flush()

        def match(k):
            if not end:
                return k == key
            return k >= key and (end == b"\0" or k < end)

        with st.mu:
            start = int(cr.get("start_revision", 0)) or st.rev + 1
            hdr = st.header()
        try:
            chunk({"result": {"header": hdr, "created": True}})
            sent = start - 1
            while not st.stop:
                with st.mu:
                    evs = [h for h in st.history if h[0] > sent and match(h[2])]
                    if not evs:
                        st.mu.wait(0.2)
                        sent = max(sent, min(st.rev, sent)) if not st.history else sent
                        continue
                    hdr = st.header()
                out = []
                for rev, typ, k, kvj, prev in evs:
                    e = {"kv": kvj}
                    if typ == "DELETE":
                        e["type"] = "DELETE"
                    if want_prev and prev:
                        e["prev_kv"] = prev
                    out.append(e)
                    sent = max(sent, rev)
                chunk({"result": {"header": hdr, "events": out}})
        except (BrokenPipeError, ConnectionResetError, OSError):
            pass
        self.close_connection = True


class FakeEtcd:
    def __init__(self, port: int = 0):
        self.store = Store()
        handler = type("H", (Handler,), {"store": self.store})
        self.httpd = ThreadingHTTPServer(("127.0.0.1", port), handler)
        self.httpd.daemon_threads = True
        self.port = self.httpd.server_address[1]
        self.thread = threading.Thread(target=self.httpd.serve_forever, daemon=True)
        self.thread.start()

    @property
    def endpoint(self):
        return f"etcd://127.0.0.1:{self.port}"

    def paths(self):
        with self.store.mu:
            return [p for p, _ in self.store.requests]

    def stop(self):
        self.store.stop = True
        with self.store.mu:
            self.store.mu.notify_all()
        self.httpd.shutdown()
        self.httpd.server_close()


if __name__ == "__main__":  # stand-alone: `python tests/fake_etcd.py [port]` prints "listening <port>" and serves until killed
    import sys

    srv = FakeEtcd(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
    print(f"listening {srv.port}", flush=True)
    try:
        while True:
            time.sleep(3600)
    except KeyboardInterrupt:
        srv.stop()
