"""Property tests of the placement layer (hypothesis): random sequences of allocate / free / allocate_at on a
PoolAllocator are checked against a byte-map model, random put / remove workloads on a RangeAllocator against the
invariants the Keystone relies on (no two live shards overlap, accounting is exact, rollback leaves no trace)."""
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

ALIGN = 256
CAP = 64 * 1024


def mkpool(bb, pid, size, sc=None, node="node-a", worker="", addr=0x1000000):
    return bb.MemoryPool(pid, size, sc if sc is not None else bb.StorageClass.RAM_CPU, node, worker, "127.0.0.1:12345", addr, "deadbeef")


ops = st.lists(st.one_of(
    st.tuples(st.just("alloc"), st.integers(1, 9000), st.booleans()),
    st.tuples(st.just("free"), st.integers(0, 1 << 30), st.booleans()),
    st.tuples(st.just("at"), st.integers(0, CAP // ALIGN), st.integers(1, 6000)),
), min_size=1, max_size=120)


@settings(max_examples=150, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(ops)
def test_pool_allocator_matches_a_byte_map_model(bb, seq):
    pa = bb.PoolAllocator(mkpool(bb, "p", CAP))
    used = bytearray(CAP // ALIGN)  # one cell per alignment unit
    live = []
    for op in seq:
        if op[0] == "alloc":
            need = (op[1] + ALIGN - 1) // ALIGN
            r = pa.allocate(op[1], op[2])
            # feasibility is decided by the model: a hole of `need` cells exists iff the allocator succeeds
            run, best = 0, 0
            for c in used:
                run = run + 1 if not c else 0
                best = max(best, run)
            assert (r is not None) == (best >= need)
            if r is not None:
                assert r.offset % ALIGN == 0 and r.length == need * ALIGN and r.offset + r.length <= CAP
                cells = range(r.offset // ALIGN, (r.offset + r.length) // ALIGN)
                assert not any(used[c] for c in cells)  # never hands out occupied space
                for c in cells:
                    used[c] = 1
                live.append(r)
        elif op[0] == "free" and live:
            r = live.pop(op[1] % len(live))
            pa.free(bb.Range(r.offset, r.length))
            for c in range(r.offset // ALIGN, (r.offset + r.length) // ALIGN):
                used[c] = 0
            if op[2]:
                pa.free(bb.Range(r.offset, r.length))  # double free is rejected (logged) and changes nothing
        elif op[0] == "at":
            off, need = op[1] * ALIGN, (op[2] + ALIGN - 1) // ALIGN
            cells = range(off // ALIGN, off // ALIGN + need)
            fits = off + need * ALIGN <= CAP and not any(used[c] for c in cells)
            ok = pa.allocate_at(off, op[2])
            assert bool(ok) == fits
            if ok:
                for c in cells:
                    used[c] = 1
                live.append(type("R", (), {"offset": off, "length": need * ALIGN})())
        # accounting + coalescing after every step
        free_cells = used.count(0)
        assert pa.total_free() == free_cells * ALIGN
        fr = pa.free_ranges()
        assert sum(x.length for x in fr) == free_cells * ALIGN
        assert all(fr[i].offset + fr[i].length < fr[i + 1].offset for i in range(len(fr) - 1))  # sorted, never adjacent
        assert pa.largest_free_block() == (max((x.length for x in fr), default=0))


def _req(bb, key, size, repl, wpc):
    return bb.AllocationRequest(key, size, repl, wpc, [bb.StorageClass.RAM_CPU], "", enable_striping=wpc > 1, min_shard_size=256)


workload = st.lists(st.one_of(
    st.tuples(st.just("put"), st.integers(0, 40), st.integers(1, 300_000), st.integers(1, 3), st.integers(1, 4)),
    st.tuples(st.just("del"), st.integers(0, 40)),
), min_size=1, max_size=80)


@settings(max_examples=80, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(workload)
def test_range_allocator_invariants_under_random_workloads(bb, seq):
    pools = {f"p{i}": mkpool(bb, f"p{i}", 1 << 20, node=f"node-{i % 3}", worker=f"w{i}", addr=0x1000000 * (i + 1)) for i in range(5)}
    ra = bb.RangeAllocator()
    live = {}
    for op in seq:
        key = f"k{op[1]}"
        if op[0] == "put":
            _, _, size, repl, wpc = op
            before = {p: ra.pool_used_bytes(p) for p in pools}
            try:
                res = ra.allocate(_req(bb, key, size, repl, wpc), pools)
            except bb.BlackbirdError as e:
                assert e.code in (bb.ErrorCode.OBJECT_ALREADY_EXISTS, bb.ErrorCode.INSUFFICIENT_SPACE, bb.ErrorCode.ALLOCATION_FAILED)
                assert (e.code == bb.ErrorCode.OBJECT_ALREADY_EXISTS) == (key in live)
                assert {p: ra.pool_used_bytes(p) for p in pools} == before  # all-or-nothing: a failed put leaves no trace
                continue
            assert key not in live and len(res.copies) == repl
            for c in res.copies:
                assert sum(s.length for s in c.shards) == size and 1 <= len(c.shards) <= wpc
            live[key] = res.copies
        elif key in live:
            assert ra.free(key) == bb.ErrorCode.OK
            del live[key]
        else:
            assert ra.free(key) == bb.ErrorCode.OBJECT_NOT_FOUND
        # no two live shards overlap; per-pool usage equals the aligned sum of the live shards
        per_pool = {}
        for copies in live.values():
            for c in copies:
                for s in c.shards:
                    off = s.location["remote_addr"] - pools[s.pool_id].ucx_remote_addr
                    per_pool.setdefault(s.pool_id, []).append((off, (s.length + ALIGN - 1) // ALIGN * ALIGN))
        for p, extents in per_pool.items():
            extents.sort()
            assert all(extents[i][0] + extents[i][1] <= extents[i + 1][0] for i in range(len(extents) - 1))
            assert extents[-1][0] + extents[-1][1] <= pools[p].size
        for p in pools:
            assert ra.pool_used_bytes(p) == sum(ln for _, ln in per_pool.get(p, []))
    st_ = ra.get_stats()
    assert st_.total_objects == len(live)
