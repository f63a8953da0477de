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


run_ops = st.lists(st.one_of(
    st.tuples(st.just("run"), st.integers(0, 5), st.integers(8, 40), st.sampled_from([1, 256, 1000, 4096, 65536, 70_000])),  # batch id, count, size
    st.tuples(st.just("single"), st.integers(0, 30), st.integers(1, 100_000)),
    st.tuples(st.just("remove_run"), st.integers(0, 5), st.integers(0, 3)),  # remove every (k+2)-th object of a batch: holes
    st.tuples(st.just("remove_single"), st.integers(0, 30)),
), min_size=1, max_size=30)


@settings(max_examples=80, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(run_ops)
def test_run_placement_never_overlaps_and_returns_every_byte(bb, seq):
    """Run placement (one pool-allocator call for a whole chunk of a batch, sliced into per-object extents) interleaved with
    single puts and removals that punch holes into the chunks: whatever the order, no two live objects overlap on a pool,
    the pool accounting equals the sum of the live extents, and after removing everything each pool is one free extent again."""
    k = bb.KeystoneService(ks_cfg(bb), None)
    assert k.initialize() == bb.ErrorCode.OK and k.start() == bb.ErrorCode.OK
    try:
        pool_bytes = 4 << 20
        for i in range(3):
            assert k.register_memory_pool(mkpool(bb, f"p{i}", pool_bytes, worker=f"w{i}")) == bb.ErrorCode.OK
        cfg = bb.WorkerConfig(replication_factor=1, max_workers_per_copy=1, ttl_ms=0)
        live = {}  # key -> (pool, addr, aligned length)
        E = bb.ErrorCode

        def record(key, copies):
            sh = copies[0].shards[0]
            live[key] = (sh.pool_id, sh.location["remote_addr"], (sh.length + ALIGN - 1) // ALIGN * ALIGN)

        gen = 0
        for op in seq:
            if op[0] == "run":
                _, b, count, size = op
                gen += 1
                keys = [f"b{b}/{gen}/{j}" for j in range(count)]
                for key, (ec, copies) in zip(keys, k.batch_put_start(keys, [size] * count, cfg)):
                    assert ec in (E.OK, E.INSUFFICIENT_SPACE)
                    if ec == E.OK:
                        assert len(copies) == 1 and len(copies[0].shards) == 1 and copies[0].shards[0].length == size
                        record(key, copies)
            elif op[0] == "single":
                key = f"s{op[1]}"
                ec, copies = code(bb, lambda: k.put_start(key, op[2], cfg))
                if key in live:
                    assert ec == E.OBJECT_ALREADY_EXISTS
                elif ec == E.OK:
                    record(key, copies)
            elif op[0] == "remove_run":
                victims = [key for key in sorted(live) if key.startswith(f"b{op[1]}/")][:: op[2] + 2]
                if victims:
                    assert set(k.batch_remove_object(victims)) == {E.OK}
                    for key in victims:
                        del live[key]
            else:
                key = f"s{op[1]}"
                assert k.remove_object(key) == (E.OK if key in live else E.OBJECT_NOT_FOUND)
                live.pop(key, None)
            # invariants
            by_pool = {}
            for pool, addr, length in live.values():
                by_pool.setdefault(pool, []).append((addr, length))
            for pool, ext in by_pool.items():
                ext.sort()
                assert all(a + n <= b for (a, n), (b, _) in zip(ext, ext[1:])), f"overlap on {pool}"
            assert k.get_cluster_stats().used_capacity == sum(n for _, _, n in live.values())
        if live:
            assert set(k.batch_remove_object(sorted(live))) == {E.OK}
        assert k.get_cluster_stats().used_capacity == 0
        whole = k.batch_put_start([f"whole{i}" for i in range(3)], [pool_bytes] * 3, cfg)
        assert [r[0] for r in whole] == [E.OK] * 3  # every pool is one free extent again
    finally:
        k.stop()
