"""Model-based test of tenant admission control (hypothesis): random interleavings of put_start / put_complete / put_cancel /
remove_object / remove_all_objects by two budgeted tenants and a member over a shared key space.  After every step the
Keystone's per-tenant books equal the model's -- size x replicas of every live object, charged to whoever put it -- the
budgets are never exceeded, a refusal charges nothing, and a restart from the metadata log rebuilds the same books."""
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from test_keystone import ks_cfg, mkpool

TABLE = """
tenants:
  - {name: a, secret: sa, write: ["a/", "s/"], quota_bytes: 300000, max_objects: 4}
  - {name: b, secret: sb, write: ["b/", "s/"], quota_bytes: 150000}
"""
WHO = st.sampled_from(["a", "b", ""])  # "" = a member: no grants to check, no budget
KEYS = st.sampled_from(["a/0", "a/1", "b/0", "b/1", "s/0", "s/1", "s/2"])
ops = st.lists(st.one_of(
    st.tuples(st.just("start"), WHO, KEYS, st.integers(1, 120_000), st.integers(1, 2)),
    st.tuples(st.just("complete"), WHO, KEYS),
    st.tuples(st.just("cancel"), WHO, KEYS),
    st.tuples(st.just("remove"), WHO, KEYS),
    st.tuples(st.just("remove_all"), WHO, KEYS),
    st.tuples(st.just("restart"), WHO, KEYS),
), min_size=1, max_size=50)
GRANTS = {"a": ("a/", "s/"), "b": ("b/", "s/")}
QUOTA = {"a": (300_000, 4), "b": (150_000, 0)}


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(ops)
def test_tenant_books_match_a_model(bb, tmp_path_factory, seq):
    import contextlib

    wal = str(tmp_path_factory.mktemp("ten") / "wal")
    bb.load_tenants_text(TABLE)
    pools = [mkpool(bb, f"p{i}", 2 << 20, worker=f"w{i}") for i in range(4)]

    def boot():
        k = bb.KeystoneService(ks_cfg(bb, cluster_id="tp", wal_path=wal, wal_fsync=False), None)
        assert k.initialize() == bb.ErrorCode.OK and k.start() == bb.ErrorCode.OK
        for p in pools:
            assert k.register_memory_pool(p) == bb.ErrorCode.OK
        return k

    k = boot()
    E = bb.ErrorCode
    model = {}  # key -> [state, size, repl, owner]
    try:
        def books():
            out = {}
            for key, (_, size, repl, owner) in model.items():
                if owner:
                    u = out.setdefault(owner, [0, 0])
                    u[0] += size * repl
                    u[1] += 1
            return out

        def check():
            got = {u["name"]: (u["used_bytes"], u["objects"]) for u in k.tenant_usage()}
            want = books()
            for t in ("a", "b"):
                assert got[t] == tuple(want.get(t, [0, 0])), (t, got, want)
                q, n = QUOTA[t]
                assert got[t][0] <= q and (not n or got[t][1] <= n)

        for op, who, key, *rest in seq:
            scope = bb.TenantScope(who) if who else contextlib.nullcontext()
            with scope:
                if op == "start":
                    size, repl = rest
                    try:
                        k.put_start(key, size, bb.WorkerConfig(replication_factor=repl, max_workers_per_copy=1, ttl_ms=0))
                        ec = E.OK
                    except bb.BlackbirdError as e:
                        ec = e.code
                    used, count = books().get(who, [0, 0])
                    if key in model:
                        assert ec == E.OBJECT_ALREADY_EXISTS
                    elif who and not key.startswith(GRANTS[who]):
                        assert ec == E.ACCESS_DENIED
                    elif who and (used + size * repl > QUOTA[who][0] or (QUOTA[who][1] and count >= QUOTA[who][1])):
                        assert ec == E.QUOTA_EXCEEDED
                    else:
                        assert ec == E.OK  # 8 MiB of pools: what the budgets admit always fits
                        model[key] = ["PENDING", size, repl, who]
                elif op == "complete":
                    ec = k.put_complete(key)
                    assert ec == (E.OK if key in model else E.OBJECT_NOT_FOUND)
                    if key in model:
                        model[key][0] = "COMPLETE"
                elif op == "cancel":
                    ec = k.put_cancel(key)
                    if key in model and model[key][0] == "PENDING":
                        assert ec == E.OK
                        del model[key]
                    else:
                        assert ec != E.OK
                elif op == "remove":
                    ec = k.remove_object(key)  # (the in-process API: key ACLs on removals are the RPC layer's)
                    assert ec == (E.OK if key in model else E.OBJECT_NOT_FOUND)
                    model.pop(key, None)
                elif op == "remove_all":
                    k.remove_all_objects()
                    model.clear()
            if op == "restart":
                k.stop()
                del k
                k = boot()
                for key in [key for key, v in model.items() if v[0] == "PENDING"]:
                    del model[key]  # in-flight puts are not in the log: gone, and no longer charged
            check()
    finally:
        k.stop()
        bb.set_tenants([])
