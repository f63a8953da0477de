"""Model-based test of the StorageBackend reservation protocol (hypothesis): random reserve / commit / abort / free
sequences on the DRAM and mmap-disk backends never hand out overlapping extents, keep exact capacity accounting, and
reject stale tokens and wrong frees with the documented codes (reference contract: storage/storage_backend.h:46-126)."""
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

CAP = 1 << 20
ops = st.lists(st.one_of(
    st.tuples(st.just("reserve"), st.integers(1, 200_000)),
    st.tuples(st.just("commit"), st.integers(0, 1 << 20)),
    st.tuples(st.just("abort"), st.integers(0, 1 << 20)),
    st.tuples(st.just("free"), st.integers(0, 1 << 20), st.booleans()),
    st.tuples(st.just("bogus"), st.integers(0, 3)),
), min_size=1, max_size=60)


@settings(max_examples=120, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(ops, st.sampled_from(["RAM_CPU", "SSD"]))
def test_reservation_protocol_matches_a_model(bb, tmp_path_factory, seq, cls):
    sc = getattr(bb.StorageClass, cls)
    b = bb.create_storage_backend(sc, CAP, str(tmp_path_factory.mktemp("be")) if cls != "RAM_CPU" else "", pool_id="prop")
    assert b.initialize() == bb.ErrorCode.OK
    E = bb.ErrorCode
    base = b.get_base_address()
    reserved, committed = [], []  # tokens; (addr, size)
    try:
        for op in seq:
            if op[0] == "reserve":
                try:
                    t = b.reserve_shard(op[1])
                except bb.BlackbirdError as e:
                    assert e.code in (E.INSUFFICIENT_SPACE, E.OUT_OF_MEMORY, E.ALLOCATION_FAILED)
                    continue
                assert t.size >= op[1] and base <= t.remote_addr and t.remote_addr + t.size <= base + CAP
                reserved.append(t)
            elif op[0] == "commit" and reserved:
                t = reserved.pop(op[1] % len(reserved))
                assert b.commit_shard(t) == E.OK
                committed.append((t.remote_addr, t.size))
                assert b.commit_shard(t) != E.OK  # a token is single use
            elif op[0] == "abort" and reserved:
                t = reserved.pop(op[1] % len(reserved))
                assert b.abort_shard(t) == E.OK
                assert b.abort_shard(t) != E.OK and b.commit_shard(t) != E.OK
            elif op[0] == "free" and committed:
                addr, size = committed.pop(op[1] % len(committed))
                if op[2]:
                    unit_ = 4096 if cls == "SSD" else 256
                    assert b.free_shard(addr, size + unit_) != E.OK  # a different extent size is refused, nothing changes
                assert b.free_shard(addr, size) == E.OK
                assert b.free_shard(addr, size) != E.OK
            elif op[0] == "bogus":
                assert b.free_shard(base + CAP + 4096 * op[1], 4096) != E.OK
            # extents handed out (reserved or committed) never overlap; accounting is exact
            unit = 4096 if cls == "SSD" else 256  # O_DIRECT blocks on the io_uring tier, TMA-legal 256 B extents elsewhere
            al = lambda n: (n + unit - 1) // unit * unit
            spans = sorted([(t.remote_addr, t.remote_addr + al(t.size)) for t in reserved] + [(a, a + al(s)) for a, s in committed])
            assert all(spans[i][1] <= spans[i + 1][0] for i in range(len(spans) - 1))
            stt = b.get_stats()
            assert stt.num_reservations == len(reserved) and stt.num_committed_shards == len(committed)
            assert stt.used_capacity == sum(e - s for s, e in spans) and stt.available_capacity == stt.total_capacity - stt.used_capacity
    finally:
        b.shutdown()


from hypothesis import HealthCheck as _HC, given as _given, settings as _settings, strategies as _st  # noqa: E402


@_settings(max_examples=200, deadline=None, suppress_health_check=[_HC.function_scoped_fixture])
@_given(_st.binary(min_size=1, max_size=5000), _st.integers(0, 1 << 40), _st.lists(_st.integers(0, 5000), max_size=6))
def test_offset_cipher_is_position_addressed_and_an_involution(bb, data, base, cuts):
    """AES-256-CTR by byte offset (encryption at rest): encrypting a buffer in one call equals encrypting any partition of it
    piece by piece at the pieces' own offsets (so ranges can be written and read independently), applying it twice gives the
    input back, and the key stream really depends on the position."""
    key, nonce = bytes(range(32)), bytes(range(1, 9))
    whole = bb.offset_cipher_crypt(key, nonce, base, data)
    assert whole is not None and len(whole) == len(data)
    assert bb.offset_cipher_crypt(key, nonce, base, whole) == data
    edges = sorted({0, len(data), *[c for c in cuts if c < len(data)]})
    pieces = b"".join(bb.offset_cipher_crypt(key, nonce, base + a, data[a:b]) for a, b in zip(edges, edges[1:]))
    assert pieces == whole
    if len(data) >= 16:
        assert bb.offset_cipher_crypt(key, nonce, base + 16, data) != whole  # another position, another key stream
