"""worker/storage: reserve/commit/abort/free lifecycle (semantics of the reference's 16
IoUringDiskBackend tests, tests/storage/test_iouring_disk_backend.cpp) on every tier, plus what
the reference lacks: a real data path, real io_uring submissions, manifest recovery."""
import os
import threading
import time

import pytest

MiB = 1 << 20


def make(bb, sc, cap, tmp_path, pool_id="pool0", **kw):
    b = bb.create_storage_backend(sc, cap, str(tmp_path), pool_id=pool_id, **kw)
    assert b is not None, f"factory returned nothing for {sc}"
    assert b.initialize() == bb.ErrorCode.OK
    return b


ALL = ["RAM_CPU", "NVME", "SSD", "HDD", "CXL_MEMORY", "CXL_TYPE2_DEVICE"]


@pytest.mark.parametrize("sc_name", ALL)
def test_factory_builds_every_host_tier_and_accounts_capacity(bb, tmp_path, sc_name):
    """The reference factory returns nullptr for NVME/SSD/HDD (ram_backend.cpp:299-301)."""
    sc = getattr(bb.StorageClass, sc_name)
    b = make(bb, sc, 8 * MiB, tmp_path)
    assert b.get_storage_class() == sc and b.get_total_capacity() == 8 * MiB
    assert b.get_used_capacity() == 0 and b.get_available_capacity() == 8 * MiB
    tok = b.reserve_shard(102400)
    assert tok.size == 102400 and tok.pool_id == "pool0" and tok.remote_addr >= b.get_base_address()
    assert b.get_used_capacity() >= 102400 and b.get_stats().num_reservations == 1
    assert b.commit_shard(tok) == bb.ErrorCode.OK
    st = b.get_stats()
    assert st.num_reservations == 0 and st.num_committed_shards == 1 and 0 < st.utilization < 1
    assert b.free_shard(tok.remote_addr, tok.size) == bb.ErrorCode.OK
    assert b.get_used_capacity() == 0
    b.shutdown()


def test_gpu_tier_needs_the_cuda_factory(bb):
    # on a CPU-only host the RAM_GPU class cannot be built (no silent malloc stand-in)
    if bb.cuda_device_count() == 0:
        assert bb.create_storage_backend(bb.StorageClass.RAM_GPU, MiB) is None


def test_out_of_space_zero_size_unknown_tokens(bb, tmp_path):
    b = make(bb, bb.StorageClass.NVME, 4 * MiB, tmp_path)
    with pytest.raises(bb.BlackbirdError) as e:
        b.reserve_shard(0)
    assert e.value.code == bb.ErrorCode.INVALID_PARAMETERS
    with pytest.raises(bb.BlackbirdError) as e:
        b.reserve_shard(5 * MiB)
    assert e.value.code == bb.ErrorCode.OUT_OF_MEMORY
    tok = b.reserve_shard(MiB)
    assert b.abort_shard(tok) == bb.ErrorCode.OK and b.get_used_capacity() == 0  # abort restores capacity
    assert b.abort_shard(tok) == bb.ErrorCode.INVALID_PARAMETERS
    assert b.commit_shard(tok) == bb.ErrorCode.INVALID_PARAMETERS
    tok = b.reserve_shard(MiB)
    b.commit_shard(tok)
    assert b.free_shard(tok.remote_addr, tok.size + 8192) == bb.ErrorCode.INVALID_PARAMETERS  # size mismatch
    assert b.free_shard(tok.remote_addr + 4096, 4096) == bb.ErrorCode.OBJECT_NOT_FOUND  # unknown address


def test_expired_token_times_out_and_is_reclaimed(bb, tmp_path):
    b = make(bb, bb.StorageClass.RAM_CPU, 4 * MiB, tmp_path)
    b.set_reservation_ttl_ms(30)
    tok = b.reserve_shard(3 * MiB)
    time.sleep(0.08)
    assert b.commit_shard(tok) == bb.ErrorCode.OPERATION_TIMEOUT and b.get_used_capacity() == 0
    tok2 = b.reserve_shard(3 * MiB)  # abandoned reservation...
    time.sleep(0.08)
    tok3 = b.reserve_shard(3 * MiB)  # ...is reclaimed by the next reserve instead of exhausting the tier
    assert tok3.remote_addr == tok2.remote_addr


def test_uncommitted_reservations_never_overlap(bb, tmp_path):
    """Bug #11: RamBackend::find_free_offset ignored uncommitted reservations."""
    b = make(bb, bb.StorageClass.RAM_CPU, 4 * MiB, tmp_path)
    toks = [b.reserve_shard(256 * 1024) for _ in range(8)]
    spans = sorted((t.remote_addr, t.remote_addr + t.size) for t in toks)
    assert all(spans[i][1] <= spans[i + 1][0] for i in range(len(spans) - 1))
    assert b.get_stats().num_reservations == 8  # and get_stats does not self-deadlock (bug #10)


def test_concurrent_reserve_commit(bb, tmp_path):
    b = make(bb, bb.StorageClass.SSD, 64 * MiB, tmp_path)
    toks = []
    lock = threading.Lock()

    def run():
        for _ in range(50):
            t = b.reserve_shard(64 * 1024)
            assert b.commit_shard(t) == bb.ErrorCode.OK
            with lock:
                toks.append(t)

    ts = [threading.Thread(target=run) for _ in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert len({t.remote_addr for t in toks}) == 200 and b.get_stats().num_committed_shards == 200


@pytest.mark.parametrize("sc_name", ALL)
def test_data_path_roundtrip(bb, tmp_path, sc_name):
    """The reference backends have no read/write API at all."""
    b = make(bb, getattr(bb.StorageClass, sc_name), 16 * MiB, tmp_path)
    for off, n in [(0, 1), (4096, 4096), (8192 + 256, 100001), (1 * MiB, 5 * MiB + 13)]:
        data = os.urandom(n)
        assert b.write(off, data) == bb.ErrorCode.OK
        assert b.read(off, n) == data
    assert b.write(16 * MiB - 10, b"x" * 11) == bb.ErrorCode.MEMORY_ACCESS_ERROR  # bounds are enforced
    st = b.get_stats()
    assert st.bytes_written > 5 * MiB and st.bytes_read > 5 * MiB and st.io_errors == 0


@pytest.mark.parametrize("sc_name", ALL)
def test_a_write_never_touches_bytes_past_its_range(bb, tmp_path, sc_name):
    """Keystone packs extents tightly, so what follows a shard is usually another object.  (The O_DIRECT path used to pad
    the last block of a write with zeros and wipe the head of the neighbour: found by `drain-worker` on an NVMe pool.)"""
    import random
    rng = random.Random(7)
    b = make(bb, getattr(bb.StorageClass, sc_name), 8 * MiB, tmp_path)
    image = bytearray(os.urandom(4 * MiB))
    assert b.write(0, bytes(image)) == bb.ErrorCode.OK
    for _ in range(40):
        off = rng.choice([0, 4096, 8192, 256 * rng.randrange(1, 4000), rng.randrange(0, 3 * MiB)])
        n = rng.choice([1, 255, 4095, 4096, 4097, 5000, 3_000_000 % (4 * MiB - off) + 1, rng.randrange(1, MiB)])
        n = min(n, 4 * MiB - off)
        data = os.urandom(n)
        assert b.write(off, data) == bb.ErrorCode.OK
        image[off:off + n] = data
        lo = max(0, off - 8192)
        assert b.read(lo, min(4 * MiB, off + n + 8192) - lo) == bytes(image[lo:min(4 * MiB, off + n + 8192)]), (off, n)
    assert b.read(0, 4 * MiB) == bytes(image)


def test_io_uring_backend_submits_real_sqes_and_persists(bb, tmp_path):
    """The reference opens a ring and never submits an SQE."""
    b = make(bb, bb.StorageClass.NVME, 32 * MiB, tmp_path, queue_depth=32)
    assert os.path.exists(b.file_path) and os.path.getsize(b.file_path) == 32 * MiB
    data = os.urandom(3 * MiB + 777)
    tok = b.reserve_shard(len(data))
    off = tok.remote_addr - b.get_base_address()
    assert b.write(off, data) == bb.ErrorCode.OK
    assert b.commit_shard(tok) == bb.ErrorCode.OK and b.flush() == bb.ErrorCode.OK
    if bb.io_uring_supported():
        assert b.using_uring and b.sqes_submitted >= 12  # 256 KiB pieces through the ring
        assert b.fixed_sqes in (0, b.sqes_submitted)  # registered staging buffer => every SQE is READ_FIXED / WRITE_FIXED (0: memlock limit)
    with open(b.file_path, "rb") as f:  # bytes are really on disk at the advertised offset
        f.seek(off)
        assert f.read(len(data)) == data
    tok2 = b.reserve_shard(4096)
    b.commit_shard(tok2)
    b.free_shard(tok2.remote_addr, 4096)
    b.shutdown()
    # a restarted worker recovers the committed extent (offset, size, crc32c) from the manifest
    b2 = make(bb, bb.StorageClass.NVME, 32 * MiB, tmp_path, queue_depth=32)
    rec = b2.recovered_extents()
    assert rec == [(off, len(data), bb.crc32c(data))]
    assert b2.read(off, len(data)) == data and b2.get_stats().num_committed_shards == 1
    tok3 = b2.reserve_shard(MiB)
    assert not (off < tok3.remote_addr - b2.get_base_address() + MiB and tok3.remote_addr - b2.get_base_address() < off + len(data))


def test_mmap_backend_is_file_backed(bb, tmp_path):
    b = make(bb, bb.StorageClass.HDD, 4 * MiB, tmp_path, pool_id="hdd0")
    assert b.has_direct_ptr() and b.file_path.endswith("hdd0.dat")
    b.write(12345, b"hello-mmap")
    assert b.flush() == bb.ErrorCode.OK
    with open(b.file_path, "rb") as f:
        f.seek(12345)
        assert f.read(10) == b"hello-mmap"


def test_cxl_backend_placeholder_cacheline_and_regions(bb, tmp_path):
    b = make(bb, bb.StorageClass.CXL_MEMORY, 4 * MiB + 17, tmp_path / "no-dax-device")
    assert b.get_total_capacity() % 64 == 0 and not b.is_dax and b.has_direct_ptr()
    tok = b.reserve_shard(100)  # sizes are cache-line granular
    assert tok.size == 128 and b.region_id(0) == 0 and b.region_id(1024) == 4
    bad = bb.create_storage_backend(bb.StorageClass.CXL_MEMORY, 0)
    assert bad.initialize() == bb.ErrorCode.INVALID_ARGUMENT  # one of the codes the reference forgot to declare


def test_unwritable_directory_fails_cleanly(bb):
    b = bb.create_storage_backend(bb.StorageClass.NVME, MiB, "/proc/definitely/not/writable")
    assert b.initialize() == bb.ErrorCode.IO_ERROR


def test_cxl_worker_yaml_is_fully_parsed_and_consumed(bb):
    """configs/cxl_worker.yaml (reference schema): per-pool `config:`, `transport:` and `allocation.preferred_tiers`
    are parsed (the reference parses none of them) and drive the backend options / advertised interconnects."""
    import os

    cfg = bb.WorkerServiceConfig.from_yaml(os.path.join(os.path.dirname(__file__), "..", "configs", "cxl_worker.yaml"))
    assert cfg.worker_id == "worker_cxl_node_1" and len(cfg.storage_pools) == 4
    cxl = next(p for p in cfg.storage_pools if p.pool_id == "cxl_memory_pool")
    assert cxl.storage_class == bb.StorageClass.CXL_MEMORY and cxl.size_bytes == 512 * 10**6 or cxl.size_bytes == 512 << 20
    assert cxl.mount_path == "/dev/dax0.0" and cxl.numa_node == 1 and cxl.cxl.interleave_granularity == 256
    acc = next(p for p in cfg.storage_pools if p.pool_id == "cxl_accelerator_pool")
    assert acc.storage_class == bb.StorageClass.CXL_TYPE2_DEVICE and acc.cxl.interleave_granularity == 4096 and acc.cxl.device_id == "cxl_type2_0"
    t = cfg.transport
    assert cfg.has_transport and t.interconnect_type == bb.CxlInterconnectType.CXL_FABRIC
    assert t.transport_protocol == bb.CxlTransportProtocol.RDMA_OVER_CXL and t.enable_multipath and t.queue_depth == 128
    assert t.fallback_transports == ["ucx", "nvlink", "roce"] and t.cxl_port_id == "cxl_port_0" and t.max_transfer_size >= 4 * 10**9
    assert bb.cxl_protocol_name(t.transport_protocol) == "RDMA over CXL" and bb.cxl_interconnect_name(t.interconnect_type) == "CXL.fabric"
    # no CXL device and no GPU on this box: only tcp survives; with both, the primary protocol then nvlink
    assert t.resolve_interconnects(False, False) == ["tcp"]
    assert t.resolve_interconnects(True, True) == ["rdma_over_cxl", "nvlink", "tcp"]
    rules = cfg.preferred_tiers
    assert [r.storage_class for r in rules] == ["RAM_CPU", "CXL_MEMORY", "CXL_TYPE2_DEVICE", "NVME"]
    assert bb.tier_classes_for_size(rules, 4096) == ["RAM_CPU"]
    assert bb.tier_classes_for_size(rules, 200 * 10**6) == ["CXL_MEMORY", "CXL_TYPE2_DEVICE"]
    assert bb.tier_classes_for_size(rules, 50 * 10**9) == ["NVME"]


def test_shared_ram_pool_is_mappable_through_its_registration_key(bb):
    """DRAM pools created with shared_memory are memfd-backed: the registration key names /proc/<pid>/fd/<n>, which a
    GPU client of another process maps (and registers with CUDA) to reach the tier with the fused kernels."""
    cap = 8 << 20
    b = bb.create_storage_backend(bb.StorageClass.RAM_CPU, cap, pool_id="shared-pool", shared_memory=True)
    assert b.initialize() == bb.ErrorCode.OK
    assert b.shared_path.startswith("/proc/") and os.path.exists(b.shared_path)
    key_hex = b.registration_key_hex()
    assert bytes.fromhex(key_hex).decode() == "file:" + b.shared_path
    payload = os.urandom(70000)
    assert b.write(4096, payload) == bb.ErrorCode.OK
    assert bb.read_shared_pool(key_hex, cap, 4096, len(payload)) == payload  # second mapping of the same pages
    # a private pool keeps the 8-hex-digit rkey and is not mappable
    p = bb.create_storage_backend(bb.StorageClass.RAM_CPU, cap, pool_id="private-pool")
    assert p.initialize() == bb.ErrorCode.OK
    assert p.shared_path == "" and len(p.registration_key_hex()) == 8
    assert bb.read_shared_pool(p.registration_key_hex(), cap, 0, 16) is None
    path = b.shared_path
    b.shutdown()
    assert not os.path.exists(path)


@pytest.mark.parametrize("sc_name", ["NVME", "HDD"])
def test_encryption_at_rest_on_file_backed_tiers(bb, tmp_path, sc_name):
    """`encrypt_at_rest` (reference roadmap v0.5: "encryption-at-rest / in-flight"): what reaches the pool file is AES-256-CTR of
    the pool bytes, addressed by byte offset -- any range reads back independently, nothing recognisable is on disk, a restarted
    worker with the key recovers its extents, one without it (or with another) reads noise that the digests reject."""
    sc = getattr(bb.StorageClass, sc_name)  # NVME: io_uring backend; HDD: mmap backend
    key = "correct horse battery staple"
    b = make(bb, sc, 16 * MiB, tmp_path, pool_id="enc0", at_rest_key=key)
    marker = b"PLAINTEXT-MARKER-0123456789abcdef"
    data = (marker * 40000)[: 1 * MiB + 4321]
    tok = b.reserve_shard(len(data))
    off = tok.remote_addr - b.get_base_address()
    assert b.write(off, data) == bb.ErrorCode.OK and b.commit_shard(tok) == bb.ErrorCode.OK and b.flush() == bb.ErrorCode.OK
    assert not b.has_direct_ptr() or sc_name == "NVME"  # no plain bytes anyone could point at
    on_disk = open(b.file_path, "rb").read()
    assert marker not in on_disk and marker[:8] not in on_disk
    assert on_disk[off:off + len(data)] != data and len(set(on_disk[off:off + 4096])) > 200  # looks like noise
    # every range decrypts on its own: whole, unaligned slices, single bytes
    assert b.read(off, len(data)) == data
    for o, n in ((1, 1), (15, 3), (16, 16), (17, 4096), (65535, 70001), (len(data) - 5, 5)):
        assert b.read(off + o, n) == data[o:o + n]
    # unaligned overwrite in the middle, then read across it
    patch = os.urandom(1000)
    assert b.write(off + 12345, patch) == bb.ErrorCode.OK
    want = data[:12345] + patch + data[12345 + 1000:]
    assert b.read(off, len(want)) == want
    b.flush(), b.shutdown()
    # restart with the key: same bytes
    b2 = make(bb, sc, 16 * MiB, tmp_path, pool_id="enc0", at_rest_key=key)
    assert b2.read(off, len(want)) == want
    if sc_name == "NVME":
        assert b2.recovered_extents()[0][:2] == (off, len(data))
    b2.shutdown()
    # another key, another pool id under the same key, or no key at all: noise
    for kw in (dict(pool_id="enc0", at_rest_key="wrong"), dict(pool_id="enc0"),):
        b3 = make(bb, sc, 16 * MiB, tmp_path, **kw)
        got = b3.read(off, 65536)
        assert got != want[:65536] and marker[:8] not in got
        b3.shutdown()


def test_worker_config_wires_encrypt_at_rest(bb, tmp_path, monkeypatch):
    monkeypatch.setenv("BB_AT_REST_KEY", "from-the-environment")
    cfg = tmp_path / "w.yaml"
    cfg.write_text(f"""
worker: {{worker_id: "we", node_id: "ne"}}
storage_pools:
  - {{pool_id: "plain", storage_class: "NVME", size_bytes: 8_MB, mount_path: "{tmp_path}/p"}}
  - {{pool_id: "sealed", storage_class: "NVME", size_bytes: 8_MB, mount_path: "{tmp_path}/s", encrypt_at_rest: true}}
""")
    w = bb.WorkerService(bb.WorkerServiceConfig.from_yaml(str(cfg)), None, None)
    assert w.create_storage_pools_from_config() == bb.ErrorCode.OK
    secret = b"TOP-SECRET-" * 1000
    for pid in ("plain", "sealed"):
        be = w.backend(pid)
        assert be.initialize() == bb.ErrorCode.OK and be.write(4096, secret) == bb.ErrorCode.OK and be.flush() == bb.ErrorCode.OK
        assert be.read(4096, len(secret)) == secret
    assert b"TOP-SECRET-" in open(w.backend("plain").file_path, "rb").read()
    assert b"TOP-SECRET-" not in open(w.backend("sealed").file_path, "rb").read()
    # without any key the worker refuses to create the pool instead of storing plain text
    monkeypatch.delenv("BB_AT_REST_KEY")
    w2 = bb.WorkerService(bb.WorkerServiceConfig.from_yaml(str(cfg)), None, None)
    assert w2.create_storage_pools_from_config() != bb.ErrorCode.OK
