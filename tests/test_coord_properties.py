"""Model-based test of the coordination store (hypothesis): random puts / deletes / CAS / leases with a virtual clock
are replayed against a dictionary model; the key space, lease expiry (keys vanish with their lease, exactly once) and
the ordered watch stream (every mutation delivered once, revisions strictly increasing) must match."""
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

K = st.sampled_from(["/a/1", "/a/2", "/a/3", "/b/1"])
V = st.sampled_from(["x", "y", "z"])
ops = st.lists(st.one_of(
    st.tuples(st.just("put"), K, V, st.integers(0, 3)),          # lease slot 0 = no lease
    st.tuples(st.just("del"), K),
    st.tuples(st.just("pia"), K, V),                              # put-if-absent
    st.tuples(st.just("cas"), K, V, V),
    st.tuples(st.just("cad"), K, V),
    st.tuples(st.just("grant"), st.integers(1, 3), st.integers(1, 5)),   # slot, ttl seconds
    st.tuples(st.just("keepalive"), st.integers(1, 3)),
    st.tuples(st.just("revoke"), st.integers(1, 3)),
    st.tuples(st.just("tick"), st.integers(100, 2500)),           # milliseconds
), min_size=1, max_size=70)


@settings(max_examples=150, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(ops)
def test_mem_coord_matches_a_model_with_leases_and_watches(bb, seq):
    s = bb.MemCoord()
    events = []
    wid = s.watch_prefix("/a/", lambda t, k, v, rev: events.append((t, k, v, rev)))
    kv, key_lease = {}, {}          # model: key -> value, key -> lease id (0 = none)
    leases, slot = {}, {}           # lease id -> [ttl_ms, remaining_ms]; slot -> lease id
    expected = []                   # (type, key, value) under /a/

    def note(t, k, v=""):
        if k.startswith("/a/"):
            expected.append((t, k, v))

    def drop_lease(lid):
        leases.pop(lid, None)
        for k in sorted(k for k, l in key_lease.items() if l == lid):
            note("DELETE", k, kv[k])  # a DELETE event carries the value that was removed
            del kv[k], key_lease[k]

    OK = bb.ErrorCode.OK
    for op in seq:
        if op[0] == "put":
            lid = slot.get(op[3], -1) if op[3] else 0
            ec = s.put(op[1], op[2], max(lid, 0) if lid != -1 else 999999)
            if lid == -1 or (lid and lid not in leases):
                assert ec != OK
            else:
                assert ec == OK
                kv[op[1]], key_lease[op[1]] = op[2], lid
                note("PUT", op[1], op[2])
        elif op[0] == "del":
            assert s.delete(op[1]) == OK  # idempotent, like etcd; only a real deletion is an event
            if op[1] in kv:
                note("DELETE", op[1], kv[op[1]])
                del kv[op[1]], key_lease[op[1]]
        elif op[0] == "pia":
            won = s.put_if_absent(op[1], op[2])
            assert won == (op[1] not in kv)
            if won:
                kv[op[1]], key_lease[op[1]] = op[2], 0
                note("PUT", op[1], op[2])
        elif op[0] == "cas":
            won = s.compare_and_swap(op[1], op[2], op[3])
            assert won == (kv.get(op[1]) == op[2])
            if won:
                kv[op[1]], key_lease[op[1]] = op[3], 0
                note("PUT", op[1], op[3])
        elif op[0] == "cad":
            won = s.compare_and_delete(op[1], op[2])
            assert won == (kv.get(op[1]) == op[2])
            if won:
                note("DELETE", op[1], kv[op[1]])
                del kv[op[1]], key_lease[op[1]]
        elif op[0] == "grant":
            lid = s.grant_lease(op[2])
            assert lid > 0 and lid not in leases
            slot[op[1]] = lid
            leases[lid] = [op[2] * 1000, op[2] * 1000]
        elif op[0] == "keepalive":
            lid = slot.get(op[1])
            ec = s.keep_alive(lid if lid is not None else 424242)
            assert (ec == OK) == (lid in leases)
            if lid in leases:
                leases[lid][1] = leases[lid][0]
        elif op[0] == "revoke":
            lid = slot.get(op[1])
            ec = s.revoke_lease(lid if lid is not None else 424242)
            assert (ec == OK) == (lid in leases)
            if lid in leases:
                drop_lease(lid)
        else:
            s.advance_time_ms(op[1])
            for lid in sorted(leases):
                leases[lid][1] -= op[1]
            for lid in sorted(l for l, (_, rem) in leases.items() if rem <= 0):
                drop_lease(lid)
        # key space after every step
        got = {k: v.decode() for k, v, _, _ in s.get_with_prefix("/")}
        assert got == kv
        assert s.lease_count() == len(leases)
    s.flush_events()
    s.unwatch(wid)
    # every /a/ mutation was delivered exactly once; deletions caused by one lease may come in any order among themselves
    assert sorted((t, k, v.decode()) for t, k, v, _ in events) == sorted(expected)
    revs = [r for _, _, _, r in events]
    assert revs == sorted(revs) and len(set(revs)) == len(revs)
    # per key, the order of events is the order of the mutations
    for key in {k for _, k, _ in expected}:
        assert [(t, v.decode()) for t, k, v, _ in events if k == key] == [(t, v) for t, k, v in expected if k == key]
