"""Process-level integration: the real bb-coord / bb-keystone / bb-worker / bb-cli / bb-bench
executables on loopback (what scripts/start_cluster.sh automates), including keystone fail-over
between two server processes."""
import json
import os
import signal
import socket
import subprocess
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.environ.get("BB_BIN_DIR", os.path.join(ROOT, "bin"))


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


class Procs:
    def __init__(self):
        self.procs = []

    def spawn(self, *cmd):
        p = subprocess.Popen(list(cmd), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        self.procs.append(p)
        return p

    def stop(self):
        for p in reversed(self.procs):
            if p.poll() is None:
                p.send_signal(signal.SIGTERM)
        for p in self.procs:
            try:
                p.wait(timeout=5)
            except subprocess.TimeoutExpired:
                p.kill()


@pytest.fixture
def procs():
    p = Procs()
    yield p
    p.stop()


def run_cli(*args, timeout=30):
    return subprocess.run([os.path.join(BIN, "bb-cli"), *args], capture_output=True, text=True, timeout=timeout)


def write_worker_cfg(path, wid, mount):
    path.write_text(f"""
worker:
  worker_id: "{wid}"
  node_id: "node-{wid}"
  lease_ttl_sec: 3
  heartbeat_interval_sec: 1
storage_pools:
  - pool_id: "ram-{wid}"
    storage_class: "RAM_CPU"
    size_bytes: 64_MB
  - pool_id: "nvme-{wid}"
    storage_class: "NVME"
    size_bytes: 64_MB
    mount_path: "{mount}"
""")


def test_cluster_of_real_processes(procs, tmp_path, bb):
    cport, rport, hport = free_port(), free_port(), free_port()
    procs.spawn(os.path.join(BIN, "bb-coord"), "--listen", f"127.0.0.1:{cport}")
    assert wait_port(cport)
    procs.spawn(os.path.join(BIN, "bb-keystone"), os.path.join(ROOT, "configs", "keystone.yaml"), "--coord-endpoints", f"127.0.0.1:{cport}",
                "--listen-address", f"127.0.0.1:{rport}", "--http-port", str(hport), "--cluster-id", "proc")
    assert wait_port(rport) and wait_port(hport)
    workers = []
    for i in range(2):
        cfg = tmp_path / f"w{i}.yaml"
        write_worker_cfg(cfg, f"w{i}", tmp_path / f"nvme{i}")
        workers.append(procs.spawn(os.path.join(BIN, "bb-worker"), "--config", str(cfg), "--coord-endpoints", f"127.0.0.1:{cport}", "--cluster-id", "proc"))
    ks = f"127.0.0.1:{rport}"
    deadline = time.time() + 10
    while time.time() < deadline:
        st = run_cli("--keystone", ks, "stats")
        if st.returncode == 0 and json.loads(st.stdout)["total_memory_pools"] == 4:
            break
        time.sleep(0.1)
    assert json.loads(st.stdout)["total_workers"] == 2
    r = run_cli("--keystone", ks, "smoke", "--size", "1024")
    assert r.returncode == 0 and "verify PASS" in r.stdout, r.stdout + r.stderr
    blob = tmp_path / "blob.bin"
    blob.write_bytes(os.urandom(3 << 20))
    r = run_cli("--keystone", ks, "put", "file-key", str(blob), "--replicas", "2", "--max-workers", "2", "--class", "NVME", "--checksum", "crc32c")
    assert r.returncode == 0, r.stdout + r.stderr
    out = tmp_path / "out.bin"
    assert run_cli("--keystone", ks, "get", "file-key", str(out)).returncode == 0
    assert out.read_bytes() == blob.read_bytes()
    assert run_cli("--keystone", ks, "exists", "file-key").stdout.strip() == "true"
    w = run_cli("--keystone", ks, "where", "file-key")
    assert w.returncode == 0 and w.stdout.count("copy ") == 4 and "tier=NVME" in w.stdout and "crc32c=" in w.stdout  # 2 copies x 2 shards
    pools = json.loads(run_cli("--keystone", ks, "pools").stdout)
    assert len(pools) == 4 and sum(p["used"] for p in pools) >= 2 * (3 << 20)
    m = run_cli("metrics", "--http", f"127.0.0.1:{hport}")
    assert m.returncode == 0 and "bb_objects 1" in m.stdout and 'bb_tier_used_bytes{tier="NVME"}' in m.stdout
    cp = run_cli("--keystone", ks, "compact", "nvme-w0")
    assert cp.returncode == 0 and "moved 0 objects" in cp.stdout, cp.stdout + cp.stderr  # nothing to defragment yet
    assert run_cli("--keystone", ks, "compact", "no-such-pool").returncode != 0
    ls = run_cli("--keystone", ks, "ls", "file-")
    assert ls.returncode == 0 and "file-key" in ls.stdout and "x2" in ls.stdout and "NVME" in ls.stdout
    sc = run_cli("--keystone", ks, "scrub", "file-")  # the workers re-hash what they hold: 1 object, 2 copies, nothing rotted
    assert sc.returncode == 0 and "1 objects, 2 copies hashed, 0 corrupt" in sc.stdout, sc.stdout + sc.stderr
    wk = json.loads(run_cli("--keystone", ks, "workers").stdout)
    assert sorted(w["worker_id"] for w in wk) == ["w0", "w1"] and all(len(w["pools"]) == 2 for w in wk)
    b = subprocess.run([os.path.join(BIN, "bb-bench"), "client", "--keystone", ks, "--size", "65536", "--iterations", "20", "--batch", "4"],
                       capture_output=True, text=True, timeout=60)
    res = json.loads(b.stdout)
    assert b.returncode == 0 and res["failures"] == 0 and res["iterations"] == 20 and res["write_MiBps"] > 0
    # kill -9 one worker: its heartbeat lease expires, the keystone drops it, the replica keeps serving
    workers[0].kill()
    deadline = time.time() + 15
    while time.time() < deadline and json.loads(run_cli("--keystone", ks, "stats").stdout)["total_workers"] != 1:
        time.sleep(0.2)
    assert json.loads(run_cli("--keystone", ks, "stats").stdout)["total_workers"] == 1
    assert run_cli("--keystone", ks, "get", "file-key", str(out)).returncode == 0 and out.read_bytes() == blob.read_bytes()
    rm = run_cli("--keystone", ks, "rm-prefix", "file-")
    assert rm.returncode == 0 and "removed 1 objects" in rm.stdout
    assert run_cli("--keystone", ks, "exists", "file-key").stdout.strip() == "false"


def test_backend_microbenchmark_binary(tmp_path):
    b = subprocess.run([os.path.join(BIN, "bb-bench"), "backend", "--class", "NVME", "--path", str(tmp_path), "--ops", "20", "--size", "4096"],
                       capture_output=True, text=True, timeout=60)
    assert b.returncode == 0 and "lifecycle ops/s" in b.stdout and "20 ops" in b.stdout


def test_keystone_process_failover(procs, tmp_path):
    cport = free_port()
    procs.spawn(os.path.join(BIN, "bb-coord"), "--listen", f"127.0.0.1:{cport}")
    assert wait_port(cport)
    cfg = tmp_path / "ks.yaml"
    cfg.write_text("keystone:\n  cluster_id: ha\n  enable_ha: true\n  service_registration_ttl_sec: 3\n  service_refresh_interval_sec: 1\n  http_metrics_port: \"0\"\n")
    ports, servers = [], []
    for i in range(2):
        rp = free_port()
        ports.append(rp)
        servers.append(procs.spawn(os.path.join(BIN, "bb-keystone"), str(cfg), "--coord-endpoints", f"127.0.0.1:{cport}", "--listen-address",
                                   f"127.0.0.1:{rp}", "--service-id", f"ks-{i}"))
        assert wait_port(rp)
        time.sleep(0.3)
    wcfg = tmp_path / "w.yaml"
    write_worker_cfg(wcfg, "w0", tmp_path / "nvme")
    procs.spawn(os.path.join(BIN, "bb-worker"), "--config", str(wcfg), "--coord-endpoints", f"127.0.0.1:{cport}", "--cluster-id", "ha")
    time.sleep(1.0)
    blob = tmp_path / "b.bin"
    blob.write_bytes(os.urandom(100000))
    r0 = run_cli("--keystone", f"127.0.0.1:{ports[0]}", "put", "ha-key", str(blob))
    assert r0.returncode == 0, r0.stdout + r0.stderr
    r1 = run_cli("--keystone", f"127.0.0.1:{ports[1]}", "put", "other", str(blob))
    assert r1.returncode != 0 and "NOT_LEADER" in r1.stdout + r1.stderr  # the standby refuses mutations
    servers[0].kill()  # leader crashes; the standby wins the next campaign and recovers the object log
    out = tmp_path / "o.bin"
    deadline = time.time() + 20
    ok = False
    while time.time() < deadline and not ok:
        ok = run_cli("--keystone", f"127.0.0.1:{ports[1]}", "get", "ha-key", str(out)).returncode == 0
        time.sleep(0.3)
    assert ok and out.read_bytes() == blob.read_bytes()
    assert run_cli("--keystone", f"127.0.0.1:{ports[1]}", "put", "after-failover", str(blob)).returncode == 0


def test_client_follows_the_keystone_leader(procs, tmp_path):
    """A client that knows both keystones of an HA pair: it is pointed at the standby first and lands on the leader
    (NOT_LEADER -> next endpoint), and when the leader is killed the same client object carries on against the
    newly elected keystone without being reconfigured (reference clients re-resolve the leader through etcd,
    etcd_service.cpp campaign/observe; ours rotate over the configured endpoints)."""
    import blackbird_b200._bb as bb
    cport = free_port()
    procs.spawn(os.path.join(BIN, "bb-coord"), "--listen", f"127.0.0.1:{cport}")
    assert wait_port(cport)
    cfg = tmp_path / "ks.yaml"
    cfg.write_text("keystone:\n  cluster_id: ha2\n  enable_ha: true\n  service_registration_ttl_sec: 3\n  service_refresh_interval_sec: 1\n  http_metrics_port: \"0\"\n")
    ports, servers = [], []
    for i in range(2):
        rp = free_port()
        ports.append(rp)
        servers.append(procs.spawn(os.path.join(BIN, "bb-keystone"), str(cfg), "--coord-endpoints", f"127.0.0.1:{cport}", "--listen-address",
                                   f"127.0.0.1:{rp}", "--service-id", f"ks-{i}"))
        assert wait_port(rp)
        time.sleep(0.3)
    wcfg = tmp_path / "w.yaml"
    write_worker_cfg(wcfg, "w0", tmp_path / "nvme")
    procs.spawn(os.path.join(BIN, "bb-worker"), "--config", str(wcfg), "--coord-endpoints", f"127.0.0.1:{cport}", "--cluster-id", "ha2")
    time.sleep(1.0)
    eps = [f"127.0.0.1:{ports[1]}", f"127.0.0.1:{ports[0]}"]  # standby first
    opts = bb.BlackbirdClientOptions()
    opts.keystone_endpoints = eps
    cl = bb.BlackbirdClient(opts)
    assert cl.connect() == bb.ErrorCode.OK
    api = cl.keystone()
    assert api.active_endpoint() == eps[0]
    wc = bb.WorkerConfig()
    wc.replication_factor = 1
    wc.max_workers_per_copy = 1
    blob = os.urandom(200000)
    assert cl.put("k-before", blob, wc) == bb.ErrorCode.OK  # answered NOT_LEADER by the standby, retried on the leader
    assert api.active_endpoint() == eps[1] and api.failovers() == 1
    assert cl.get("k-before") == blob
    servers[0].kill()
    servers[0].wait()
    api.set_failover_budget_ms(20000)
    assert cl.get("k-before") == blob  # same client: transport failure -> standby -> waits out the election
    assert api.active_endpoint() == eps[0] and api.failovers() >= 2
    assert cl.put("k-after", blob, wc) == bb.ErrorCode.OK and cl.get("k-after") == blob
    # the CLI takes the same list
    out = tmp_path / "o.bin"
    r = run_cli("--keystone", ",".join(reversed(eps)), "get", "k-after", str(out))
    assert r.returncode == 0 and out.read_bytes() == blob, r.stdout + r.stderr


def test_control_plane_benchmark_binary():
    b = subprocess.run([os.path.join(BIN, "bb-bench"), "control", "--threads", "4", "--batch", "256", "--iterations", "4"],
                       capture_output=True, text=True, timeout=120)
    res = json.loads(b.stdout)
    assert b.returncode == 0 and res["objects"] == 4 * 256 * 4 and res["object_lifecycles_per_s"] > 10000, b.stdout + b.stderr


def test_device_client_benchmark_binary():
    """`bb-bench devclient`: batch_put_device / batch_get_device / batch_remove of the real client against a keystone behind
    the real RPC server, with a transport whose transfers cost nothing -- the per-object control cost of small-object batches."""
    b = subprocess.run([os.path.join(BIN, "bb-bench"), "devclient", "--batch", "512", "--iterations", "3"], capture_output=True, text=True, timeout=120)
    res = json.loads(b.stdout)
    assert b.returncode == 0 and res["batch"] == 512 and 0 < res["put_us_per_obj"] < 200 and res["get_objects_per_s"] > 5000, b.stdout + b.stderr


@pytest.mark.gpu
def test_native_gpu_path_without_python(procs, tmp_path):
    """bb-coord + bb-keystone + bb-worker with a RAM_GPU pool + `bb-bench gpu`: the whole device path (CUDA IPC slab
    export, fused put/get kernels, digests in the keystone) with no Python or torch in any of the processes, and
    the client in a different process than the worker (the slab is opened through its IPC handle)."""
    cport, rport, hport = free_port(), free_port(), free_port()
    procs.spawn(os.path.join(BIN, "bb-coord"), "--listen", f"127.0.0.1:{cport}")
    assert wait_port(cport)
    procs.spawn(os.path.join(BIN, "bb-keystone"), os.path.join(ROOT, "configs", "keystone.yaml"), "--coord-endpoints", f"127.0.0.1:{cport}",
                "--listen-address", f"127.0.0.1:{rport}", "--http-port", str(hport), "--cluster-id", "gpuproc")
    assert wait_port(rport)
    cfg = tmp_path / "gpu_worker.yaml"
    cfg.write_text("""
worker:
  worker_id: "worker-gpu0"
  node_id: "gpu0"
  interconnects: ["nvlink", "tcp"]
  fabric_domain: "nvswitch-0"
  lease_ttl_sec: 5
  heartbeat_interval_sec: 1
storage_pools:
  - pool_id: "hbm0"
    storage_class: "RAM_GPU"
    size_bytes: 1_GB
    gpu_device_id: 0
""")
    procs.spawn(os.path.join(BIN, "bb-worker"), "--config", str(cfg), "--coord-endpoints", f"127.0.0.1:{cport}", "--cluster-id", "gpuproc")
    ks = f"127.0.0.1:{rport}"
    deadline = time.time() + 30
    st = None
    while time.time() < deadline:
        st = run_cli("--keystone", ks, "stats")
        if st.returncode == 0 and json.loads(st.stdout)["total_memory_pools"] == 1:
            break
        time.sleep(0.2)
    assert st is not None and json.loads(st.stdout)["total_memory_pools"] == 1, st.stdout if st else ""
    b = subprocess.run([os.path.join(BIN, "bb-bench"), "gpu", "--keystone", ks, "--objects", "16", "--size", str(16 << 20), "--iterations", "5"],
                       capture_output=True, text=True, timeout=120)
    assert b.returncode == 0, b.stdout + b.stderr
    res = json.loads(b.stdout.strip().splitlines()[-1])
    assert res["failures"] == 0 and res["verified"] is True and res["launches"] >= 14 and res["put_GBps"] > 50
    # the same objects' tier is visible to a plain TCP client
    m = run_cli("metrics", "--http", f"127.0.0.1:{hport}")
    assert m.returncode == 0 and "bb_put_start_total" in m.stdout


def test_shared_dram_pool_of_a_worker_process_is_mappable_by_clients(procs, tmp_path, bb):
    """A bb-worker started with `shared_memory: true` backs its DRAM pool with a memfd and advertises
    file:/proc/<pid>/fd/<n>; a client process on the same host maps it (the GPU fabric additionally registers the
    mapping with CUDA) and finds the object's bytes at the placement's offset -- no data-server round trip."""
    cport, rport, hport = free_port(), free_port(), free_port()
    procs.spawn(os.path.join(BIN, "bb-coord"), "--listen", f"127.0.0.1:{cport}")
    assert wait_port(cport)
    procs.spawn(os.path.join(BIN, "bb-keystone"), os.path.join(ROOT, "configs", "keystone.yaml"), "--coord-endpoints", f"127.0.0.1:{cport}",
                "--listen-address", f"127.0.0.1:{rport}", "--http-port", str(hport), "--cluster-id", "shm")
    assert wait_port(rport)
    cfg = tmp_path / "w.yaml"
    cfg.write_text("""
worker:
  worker_id: "ws"
  node_id: "node-ws"
storage_pools:
  - pool_id: "dram-ws"
    storage_class: "RAM_CPU"
    size_bytes: 32_MB
    shared_memory: true
""")
    procs.spawn(os.path.join(BIN, "bb-worker"), "--config", str(cfg), "--coord-endpoints", f"127.0.0.1:{cport}", "--cluster-id", "shm")
    c = bb.BlackbirdClient(bb.BlackbirdClientOptions("127.0.0.1", rport, 30000, 2, "node-ws"))
    assert c.connect() == bb.ErrorCode.OK
    deadline = time.time() + 10
    while time.time() < deadline and c.cluster_stats().total_memory_pools < 1:
        time.sleep(0.1)
    data = os.urandom(300_000)
    wc = bb.WorkerConfig(replication_factor=1, max_workers_per_copy=1, preferred_classes=[bb.StorageClass.RAM_CPU], checksum=bb.ChecksumAlgo.CRC32C)
    assert c.put("shared/obj", data, wc) == bb.ErrorCode.OK
    shard = c.get_workers("shared/obj")[0].shards[0]
    pool = [p for p in c.keystone().get_memory_pools() if p.id == shard.pool_id][0]
    assert bytes.fromhex(pool.ucx_rkey_hex).startswith(b"file:/proc/")
    off = shard.location["remote_addr"] - pool.ucx_remote_addr
    assert bb.read_shared_pool(pool.ucx_rkey_hex, pool.size, off, len(data)) == data


@pytest.mark.gpu
def test_gpu_client_maps_the_dram_pool_of_a_worker_process(procs, tmp_path, bb):
    """Cross-process DRAM tier: the worker process owns a memfd-backed DRAM pool; this process's GPU fabric maps it,
    registers it with CUDA and moves objects to / from it with the fused kernel (PCIe), verified against the worker's
    TCP data path."""
    import torch

    cport, rport, hport = free_port(), free_port(), free_port()
    procs.spawn(os.path.join(BIN, "bb-coord"), "--listen", f"127.0.0.1:{cport}")
    assert wait_port(cport)
    procs.spawn(os.path.join(BIN, "bb-keystone"), os.path.join(ROOT, "configs", "keystone.yaml"), "--coord-endpoints", f"127.0.0.1:{cport}",
                "--listen-address", f"127.0.0.1:{rport}", "--http-port", str(hport), "--cluster-id", "shmgpu")
    assert wait_port(rport)
    cfg = tmp_path / "w.yaml"
    cfg.write_text("""
worker:
  worker_id: "wd"
  node_id: "node-wd"
storage_pools:
  - pool_id: "dram-wd"
    storage_class: "RAM_CPU"
    size_bytes: 128_MB
    shared_memory: true
""")
    worker = procs.spawn(os.path.join(BIN, "bb-worker"), "--config", str(cfg), "--coord-endpoints", f"127.0.0.1:{cport}", "--cluster-id", "shmgpu")
    api = bb.KeystoneRpcClient()
    assert api.connect("127.0.0.1", rport, 10000) == bb.ErrorCode.OK
    c = bb.BlackbirdClient(api, bb.BlackbirdClientOptions(node_id="node-wd"))
    assert c.connect() == bb.ErrorCode.OK
    deadline = time.time() + 15
    while time.time() < deadline and c.cluster_stats().total_memory_pools < 1:
        time.sleep(0.1)
    fabric = bb.GpuFabric(0, api)
    bb.attach_fabric(c, fabric)
    n, size = 6, 2 << 20
    src = torch.randint(0, 256, (n * size,), dtype=torch.uint8, device="cuda")
    out = torch.zeros_like(src)
    keys = [f"x{i}" for i in range(n)]
    wc = bb.WorkerConfig(replication_factor=1, max_workers_per_copy=1, preferred_classes=[bb.StorageClass.RAM_CPU])
    s = torch.cuda.current_stream().cuda_stream
    ecs = c.batch_put_device(keys, [src.data_ptr() + i * size for i in range(n)], [size] * n, wc, s)
    assert all(e == bb.ErrorCode.OK for e in ecs), ecs
    assert fabric.mapped_host_pools() == 1 and fabric.launches == 1
    ecs, _ = c.batch_get_device(keys, [out.data_ptr() + i * size for i in range(n)], [size] * n, s)
    assert all(e == bb.ErrorCode.OK for e in ecs), ecs
    torch.cuda.synchronize()
    assert torch.equal(src, out) and fabric.launches == 2
    # the worker process serves the same bytes over its TCP data server
    host = bb.BlackbirdClient(bb.BlackbirdClientOptions("127.0.0.1", rport, 30000, 2, "elsewhere"))
    assert host.connect() == bb.ErrorCode.OK
    assert host.get(keys[1]) == bytes(src[size:2 * size].cpu().numpy())
    # worker restart with the same pool id: the new incarnation has a new memfd; the fabric notices the changed
    # registration key on the first placement, drops the stale mapping and maps the new pool
    worker.send_signal(signal.SIGTERM)
    worker.wait(timeout=10)
    deadline = time.time() + 15
    while time.time() < deadline and c.cluster_stats().total_memory_pools > 0:
        time.sleep(0.1)
    procs.spawn(os.path.join(BIN, "bb-worker"), "--config", str(cfg), "--coord-endpoints", f"127.0.0.1:{cport}", "--cluster-id", "shmgpu")
    deadline = time.time() + 15
    while time.time() < deadline and c.cluster_stats().total_memory_pools < 1:
        time.sleep(0.1)
    keys2 = [f"y{i}" for i in range(n)]
    out.zero_()
    ecs = c.batch_put_device(keys2, [src.data_ptr() + i * size for i in range(n)], [size] * n, wc, s)
    assert all(e == bb.ErrorCode.OK for e in ecs), ecs
    ecs, _ = c.batch_get_device(keys2, [out.data_ptr() + i * size for i in range(n)], [size] * n, s)
    torch.cuda.synchronize()
    assert all(e == bb.ErrorCode.OK for e in ecs) and torch.equal(src, out)
    assert fabric.remaps == 1 and fabric.mapped_host_pools() == 1 and fabric.launches == 4
    assert host.get(keys2[2]) == bytes(src[2 * size:3 * size].cpu().numpy())  # the NEW worker process holds the bytes


def test_same_host_clients_use_the_shared_memory_fast_path(procs, tmp_path, bb):
    """A CPU client on the worker's host maps the memfd-backed DRAM pool and moves shards with memcpy (one-sided, like
    UCX's shm transports for the reference's intra-node RMA): the worker's data server sees no request, digests are
    still computed / verified, a TCP-only client reads the same bytes, corruption is detected."""
    cport, rport, hport, wport = free_port(), free_port(), free_port(), free_port()
    procs.spawn(os.path.join(BIN, "bb-coord"), "--listen", f"127.0.0.1:{cport}")
    assert wait_port(cport)
    procs.spawn(os.path.join(BIN, "bb-keystone"), os.path.join(ROOT, "configs", "keystone.yaml"), "--coord-endpoints", f"127.0.0.1:{cport}",
                "--listen-address", f"127.0.0.1:{rport}", "--http-port", str(hport), "--cluster-id", "shmcpu")
    assert wait_port(rport)
    cfg = tmp_path / "w.yaml"
    cfg.write_text(f"""
worker:
  worker_id: "wf"
  node_id: "node-wf"
  http_metrics_port: {wport}
storage_pools:
  - pool_id: "dram-wf"
    storage_class: "RAM_CPU"
    size_bytes: 256_MB
    shared_memory: true
""")
    procs.spawn(os.path.join(BIN, "bb-worker"), "--config", str(cfg), "--coord-endpoints", f"127.0.0.1:{cport}", "--cluster-id", "shmcpu")
    assert wait_port(wport)
    c = bb.BlackbirdClient(bb.BlackbirdClientOptions("127.0.0.1", rport, 30000, 4, "node-wf"))
    assert c.connect() == bb.ErrorCode.OK
    deadline = time.time() + 10
    while time.time() < deadline and c.cluster_stats().total_memory_pools < 1:
        time.sleep(0.1)

    def served():
        text = bb.http_get("127.0.0.1", wport, "/metrics")[1]
        return int([ln for ln in text.splitlines() if ln.startswith("bb_worker_data_requests_total")][0].split()[-1])

    data = os.urandom((24 << 20) + 777)
    for algo in (bb.ChecksumAlgo.BBH64, bb.ChecksumAlgo.CRC32C):
        key = f"fast/{int(algo)}"
        before = served()
        assert c.put(key, data, bb.WorkerConfig(replication_factor=1, max_workers_per_copy=1, checksum=algo)) == bb.ErrorCode.OK
        assert c.get(key) == data
        assert served() == before  # not a single data-server request
        sh = c.get_workers(key)[0].shards[0]
        assert sh.checksum == (bb.bbh64_reference(data) if algo == bb.ChecksumAlgo.BBH64 else bb.crc32c(data))
    m = c.metrics_text()
    assert "bb_client_shm_put_bytes_total" in m and "bb_client_shm_get_shards_total 2" in m
    # a client that may not map (other host in real life) takes the TCP path and sees the same object
    o = bb.BlackbirdClientOptions("127.0.0.1", rport, 30000, 2, "elsewhere")
    o.enable_shm = False
    tcp = bb.BlackbirdClient(o)
    assert tcp.connect() == bb.ErrorCode.OK
    before = served()
    assert tcp.get("fast/2") == data and served() > before
    # corruption inside the pool is caught by the one-sided read as well
    sh = c.get_workers("fast/1")[0].shards[0]
    pool = [p for p in c.keystone().get_memory_pools() if p.id == sh.pool_id][0]
    bad = bytes([data[5 << 20] ^ 0xFF])
    assert tcp.keystone() is not None
    import mmap
    fd = os.open(bytes.fromhex(pool.ucx_rkey_hex)[5:].decode(), os.O_RDWR)
    with mmap.mmap(fd, pool.size) as mm:
        off = sh.location["remote_addr"] - pool.ucx_remote_addr + (5 << 20)
        mm[off:off + 1] = bad
    os.close(fd)
    with pytest.raises(Exception):
        c.get("fast/1")


def test_cluster_token_gates_every_rpc_server(procs, tmp_path, bb):
    """With a cluster token (BB_AUTH_TOKEN / --auth-token / auth_token: in the YAMLs) bb-coord, bb-keystone and the
    worker's data server answer nothing until a connection has passed the HMAC challenge-response on it: tools with
    the token work end to end, tools without it (or with a wrong one) are refused with ACCESS_DENIED, raw frames get the
    denial marker and a close, and the token itself never crosses the wire."""
    import struct

    env_ok = dict(os.environ, BB_AUTH_TOKEN="s3cret-cluster-token")
    cport, rport, hport = free_port(), free_port(), free_port()

    def spawn(*cmd):
        p = subprocess.Popen(list(cmd), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env_ok)
        procs.procs.append(p)
        return p

    spawn(os.path.join(BIN, "bb-coord"), "--listen", f"127.0.0.1:{cport}")
    assert wait_port(cport)
    spawn(os.path.join(BIN, "bb-keystone"), os.path.join(ROOT, "configs", "keystone.yaml"), "--coord-endpoints", f"127.0.0.1:{cport}",
          "--listen-address", f"127.0.0.1:{rport}", "--http-port", str(hport), "--cluster-id", "authc")
    assert wait_port(rport)
    cfg = tmp_path / "w.yaml"
    write_worker_cfg(cfg, "wa", tmp_path / "nvme")
    spawn(os.path.join(BIN, "bb-worker"), "--config", str(cfg), "--coord-endpoints", f"127.0.0.1:{cport}", "--cluster-id", "authc")
    ks = f"127.0.0.1:{rport}"

    def cli(env, *args):
        return subprocess.run([os.path.join(BIN, "bb-cli"), *args], capture_output=True, text=True, timeout=30, env=env)

    deadline = time.time() + 10
    while time.time() < deadline:
        st = cli(env_ok, "--keystone", ks, "stats")
        if st.returncode == 0 and json.loads(st.stdout)["total_memory_pools"] == 2:
            break
        time.sleep(0.1)
    assert st.returncode == 0 and json.loads(st.stdout)["total_workers"] == 1, st.stdout + st.stderr
    r = cli(env_ok, "--keystone", ks, "smoke", "--size", "4096")
    assert r.returncode == 0 and "verify PASS" in r.stdout, r.stdout + r.stderr
    # the flag works like the environment variable
    env_none = {k: v for k, v in os.environ.items() if k != "BB_AUTH_TOKEN"}
    assert cli(env_none, "--keystone", ks, "--auth-token", "s3cret-cluster-token", "stats").returncode == 0
    for env, extra in ((env_none, []), (env_none, ["--auth-token", "wrong"]), (dict(os.environ, BB_AUTH_TOKEN="nope"), [])):
        bad = cli(env, "--keystone", ks, *extra, "stats")
        assert bad.returncode != 0, bad.stdout
    # raw frames: a put_start without the token is answered with the denial marker, then the server hangs up
    for port in (rport, cport):
        s = socket.create_connection(("127.0.0.1", port), 2.0)
        s.settimeout(2.0)
        s.sendall(struct.pack("<IIQ", 0, 8, 1))  # get_cluster_stats, empty payload
        hdr = s.recv(16)
        assert len(hdr) == 16 and struct.unpack("<IIQ", hdr)[1] == 0x7FFFFFFD
        assert s.recv(16) == b""
        s.close()
    # the handshake, played by hand: HMAC-SHA256 over fresh nonces in both directions, the token never travels
    import hashlib
    import hmac as pyhmac
    AUTH, DENIED, tok = 0x7FFFFF00, 0x7FFFFFFD, b"s3cret-cluster-token"

    def frame(method, rid, body=b""):
        return struct.pack("<IIQ", len(body), method, rid) + body

    def recv_frame(s):
        hdr = b""
        while len(hdr) < 16:
            part = s.recv(16 - len(hdr))
            if not part:
                return None
            hdr += part
        n, method, _ = struct.unpack("<IIQ", hdr)
        body = b""
        while len(body) < n:
            body += s.recv(n - len(body))
        return method, body

    def handshake(port, key, tamper=False):
        s = socket.create_connection(("127.0.0.1", port), 2.0)
        s.settimeout(2.0)
        cn = os.urandom(16)
        s.sendall(frame(AUTH, 0, b"BBA1" + cn))
        method, body = recv_frame(s)
        assert method == AUTH and len(body) == 48
        sn, srv_mac = body[:16], body[16:]
        assert srv_mac == pyhmac.new(tok, b"bb-srv" + cn + sn, hashlib.sha256).digest()  # the server proves itself first
        proof = pyhmac.new(key, b"bb-cli" + cn + sn, hashlib.sha256).digest()
        s.sendall(frame(AUTH, 1, proof[::-1] if tamper else proof))
        return s, recv_frame(s)

    for port in (rport, cport):
        s, (method, body) = handshake(port, tok)
        assert method == AUTH and body == b""
        s.sendall(frame(8, 2))  # get_cluster_stats / any method: now answered
        assert recv_frame(s)[0] != DENIED
        s.close()
        for key, tamper in ((b"wrong", False), (tok, True)):
            s, (method, _) = handshake(port, key, tamper)
            assert method == DENIED and s.recv(16) == b""
            s.close()
        s = socket.create_connection(("127.0.0.1", port), 2.0)  # presenting the token itself is not a handshake
        s.settimeout(2.0)
        s.sendall(frame(AUTH, 0, tok))
        assert recv_frame(s)[0] == DENIED
        s.close()
    # an impostor that does not hold the token: the client walks away and never sent anything derived from it
    lst = socket.socket()
    lst.bind(("127.0.0.1", 0))
    lst.listen(1)
    seen = []

    def impostor():
        c, _ = lst.accept()
        c.settimeout(2.0)
        hello = recv_frame(c)
        seen.append(hello)
        c.sendall(frame(AUTH, 0, os.urandom(16) + os.urandom(32)))
        try:
            seen.append(c.recv(64))
        except OSError:
            seen.append(b"")
        c.close()

    import threading
    th = threading.Thread(target=impostor)
    th.start()
    bad = cli(env_ok, "--keystone", f"127.0.0.1:{lst.getsockname()[1]}", "stats")
    th.join()
    lst.close()
    assert bad.returncode != 0 and "ACCESS_DENIED" in bad.stdout + bad.stderr
    assert seen[0][0] == AUTH and len(seen[0][1]) == 20 and tok not in seen[0][1] and seen[1] == b""
    # the metrics endpoint stays open (read-only)
    assert cli(env_none, "metrics", "--http", f"127.0.0.1:{hport}").returncode == 0
    # a Python client with the token in its options
    o = bb.BlackbirdClientOptions("127.0.0.1", rport, 30000, 2, "node-wa")
    o.auth_token = "s3cret-cluster-token"
    try:
        c = bb.BlackbirdClient(o)
        assert c.connect() == bb.ErrorCode.OK
        data = os.urandom(50_000)
        assert c.put("authed", data, bb.WorkerConfig(replication_factor=1, max_workers_per_copy=1)) == bb.ErrorCode.OK
        assert c.get("authed") == data
    finally:
        bb.set_cluster_token("")  # process-wide: do not leak into the other tests


def test_start_cluster_script_brings_up_a_keystone_pair(tmp_path):
    """scripts/start_cluster.sh --ha: durable coordinator, two Keystones, workers, smoke test through both endpoints;
    scripts/stop_cluster.sh stops exactly the recorded PIDs."""
    if BIN != os.path.join(ROOT, "bin"):
        pytest.skip("the scripts start the binaries under bin/")
    env = dict(os.environ, BB_COORD_PORT=str(free_port()), BB_RPC_PORT=str(free_port()), BB_HTTP_PORT=str(free_port()))
    run = tmp_path / "run"
    try:
        up = subprocess.run([os.path.join(ROOT, "scripts", "start_cluster.sh"), "-n", "2", "--ha", "-d", str(run)], env=env, capture_output=True,
                            text=True, timeout=90)
        assert up.returncode == 0, up.stdout + up.stderr
        assert "verify PASS" in up.stdout and "bb_workers 2" in up.stdout
        assert (run / "keystone2.pid").exists() and (run / "coord-data").is_dir()
        rpc2 = int(env["BB_RPC_PORT"]) + 10
        st = run_cli("--keystone", f"127.0.0.1:{env['BB_RPC_PORT']},127.0.0.1:{rpc2}", "stats")
        assert st.returncode == 0, st.stdout + st.stderr
    finally:
        down = subprocess.run([os.path.join(ROOT, "scripts", "stop_cluster.sh"), "-d", str(run)], capture_output=True, text=True, timeout=60)
    assert "stopped keystone2" in down.stdout and "stopped coord" in down.stdout


def test_start_cluster_script_secure_mode(tmp_path):
    """scripts/start_cluster.sh --secure: a generated token gates the cluster and every frame is sealed; the smoke test inside the
    script passes with it, a client without the token is refused, one with it (and encryption) works."""
    if BIN != os.path.join(ROOT, "bin"):
        pytest.skip("the scripts start the binaries under bin/")
    env = {k: v for k, v in os.environ.items() if k not in ("BB_AUTH_TOKEN", "BB_ENCRYPT_TRANSPORT")}
    env.update(BB_COORD_PORT=str(free_port()), BB_RPC_PORT=str(free_port()), BB_HTTP_PORT=str(free_port()))
    run = tmp_path / "run"
    try:
        up = subprocess.run([os.path.join(ROOT, "scripts", "start_cluster.sh"), "-n", "1", "--secure", "-d", str(run)], env=env, capture_output=True,
                            text=True, timeout=90)
        assert up.returncode == 0 and "verify PASS" in up.stdout, up.stdout + up.stderr
        token = (run / "token").read_text()
        assert len(token) >= 24 and oct((run / "token").stat().st_mode & 0o777) == "0o600"
        ks = f"127.0.0.1:{env['BB_RPC_PORT']}"
        cli = lambda e: subprocess.run([os.path.join(BIN, "bb-cli"), "--keystone", ks, "stats"], env=e, capture_output=True, text=True, timeout=30)
        assert cli(env).returncode != 0                                                        # no token
        assert cli(dict(env, BB_AUTH_TOKEN=token)).returncode != 0                             # token but plain frames
        assert cli(dict(env, BB_AUTH_TOKEN=token, BB_ENCRYPT_TRANSPORT="1")).returncode == 0
    finally:
        subprocess.run([os.path.join(ROOT, "scripts", "stop_cluster.sh"), "-d", str(run)], capture_output=True, text=True, timeout=60)
