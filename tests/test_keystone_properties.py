"""Model-based test of the Keystone object state machine (hypothesis): random interleavings of put_start /
put_complete / put_cancel / get_workers / object_exists / remove_object / remove_all_objects over a small key space are
checked against a dictionary model -- visibility (PENDING objects are invisible), error codes, and exact allocator
accounting (nothing leaks, whatever the order)."""
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from test_keystone import ks_cfg, mkpool

ALIGN = 256
KEYS = st.integers(0, 7)
ops = st.lists(st.one_of(
    st.tuples(st.just("start"), KEYS, st.integers(1, 120_000), st.integers(1, 2), st.integers(1, 3)),
    st.tuples(st.just("complete"), KEYS),
    st.tuples(st.just("cancel"), KEYS),
    st.tuples(st.just("get"), KEYS),
    st.tuples(st.just("exists"), KEYS),
    st.tuples(st.just("remove"), KEYS),
    st.tuples(st.just("remove_all"), KEYS),
), min_size=1, max_size=60)


def code(bb, fn):
    try:
        return bb.ErrorCode.OK, fn()
    except bb.BlackbirdError as e:
        return e.code, None


@settings(max_examples=120, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(ops)
def test_keystone_state_machine_matches_a_dictionary_model(bb, seq):
    k = bb.KeystoneService(ks_cfg(bb), None)
    assert k.initialize() == bb.ErrorCode.OK and k.start() == bb.ErrorCode.OK
    try:
        for i in range(4):
            assert k.register_memory_pool(mkpool(bb, f"p{i}", 1 << 20, worker=f"w{i}")) == bb.ErrorCode.OK
        model = {}  # key -> ["PENDING" | "COMPLETE", size, copies, shards-per-copy lengths]
        E = bb.ErrorCode
        for op in seq:
            key = f"k{op[1]}"
            if op[0] == "start":
                _, _, size, repl, wpc = op
                ec, copies = code(bb, lambda: k.put_start(key, size, bb.WorkerConfig(replication_factor=repl, max_workers_per_copy=wpc, ttl_ms=0, min_shard_size=256)))
                if key in model:
                    assert ec == E.OBJECT_ALREADY_EXISTS
                else:
                    assert ec == E.OK and len(copies) == repl  # 4 MiB of pools: this workload always fits
                    for c in copies:
                        assert sum(s.length for s in c.shards) == size
                    model[key] = ["PENDING", size, copies]
            elif op[0] == "complete":
                ec = k.put_complete(key)
                assert ec == (E.OK if key in model else E.OBJECT_NOT_FOUND)
                if key in model:
                    model[key][0] = "COMPLETE"
            elif op[0] == "cancel":
                ec = k.put_cancel(key)
                if key not in model:
                    assert ec == E.OBJECT_NOT_FOUND
                elif model[key][0] == "PENDING":
                    assert ec == E.OK
                    del model[key]
                else:
                    assert ec == E.INVALID_STATE  # a completed object is removed, not cancelled
            elif op[0] == "get":
                ec, copies = code(bb, lambda: k.get_workers(key))
                if key not in model:
                    assert ec == E.OBJECT_NOT_FOUND
                elif model[key][0] == "PENDING":
                    assert ec == E.OBJECT_NOT_READY
                else:
                    assert ec == E.OK and [[s.length for s in c.shards] for c in copies] == [[s.length for s in c.shards] for c in model[key][2]]
            elif op[0] == "exists":
                assert k.object_exists(key) is (key in model and model[key][0] == "COMPLETE")
            elif op[0] == "remove":
                ec = k.remove_object(key)
                assert ec == (E.OK if key in model else E.OBJECT_NOT_FOUND)
                model.pop(key, None)
            else:
                n = k.remove_all_objects()
                assert n == len(model)
                model.clear()
            # exact accounting after every step
            stt = k.get_cluster_stats()
            assert stt.total_objects == sum(1 for v in model.values() if v[0] == "COMPLETE")
            assert stt.pending_objects == sum(1 for v in model.values() if v[0] == "PENDING")
            want = sum((s.length + ALIGN - 1) // ALIGN * ALIGN for v in model.values() for c in v[2] for s in c.shards)
            assert stt.used_capacity == want
        k.remove_all_objects()
        assert k.get_cluster_stats().used_capacity == 0
    finally:
        k.stop()
