"""Cluster bootstrap helpers.

LocalCluster   : in-process Keystone (+ optional RPC/HTTP endpoints) with N workers — the
                 plumbing configuration of BASELINE config #1, also used by tests.
GpuRankCluster : one process per GPU (torchrun).  Rank 0 hosts the Keystone; every rank runs a
                 GPU-tier worker whose HBM slab is exported to all peers over CUDA IPC, plus a
                 client with the fused-kernel device transport.
"""
from __future__ import annotations

import os
import socket
import time
from typing import Optional

from .. import _bb


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _pin_to_gpu_numa_node(device_index: int):
    """Bind this process (and every thread it creates later: RPC pool, data server, CUDA callbacks) to the
    CPUs NVML reports as local to the GPU, so pinned staging buffers and the H2D/D2H copies of the end-to-end
    path stay on the GPU's own PCIe root / NUMA node.  BB_NUMA_PIN=0 disables."""
    if os.environ.get("BB_NUMA_PIN", "1") == "0" or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        import pynvml

        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = device_index
        if visible:
            ids = [v for v in visible.split(",") if v.strip()]
            if device_index < len(ids) and ids[device_index].strip().isdigit():
                idx = int(ids[device_index])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {w * 64 + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return sorted(cpus)
    except Exception:  # no NVML, restricted cgroup, ...: run unpinned
        pass
    return None


class HostRendezvous:
    """Shared-memory rendezvous of the ranks of ONE node (torchrun workers): barrier, all-reduce of a scalar and integer
    broadcast through a small file in /dev/shm that every rank maps -- no NCCL (or any other collective library) is
    needed to bring a cluster up, and a barrier costs ~1-2 us instead of a kernel launch + stream sync.  One 64-byte slot
    per rank: [epoch u64 | value f64 | aux i64]; a rank publishes (value, epoch) and spins until every slot has reached
    the epoch.  Values are read after the barrier and before anyone can start the next epoch + 1 write (two-phase)."""

    SLOT = 64

    def __init__(self, rank: int, world: int, tag: str = ""):
        import mmap
        import struct

        self.rank, self.world, self._struct = rank, world, struct
        self.epoch = 0
        self._mm = None
        if world <= 1:
            return
        name = tag or f"{os.getppid()}-{os.environ.get('MASTER_PORT', '0')}"
        self.path = f"/dev/shm/bb-rdv-{name}"
        size = self.SLOT * (world + 1)
        if rank == 0:
            try:
                os.unlink(self.path)
            except OSError:
                pass
            fd = os.open(self.path + ".tmp", os.O_CREAT | os.O_RDWR | os.O_TRUNC, 0o600)
            os.ftruncate(fd, size)
            os.rename(self.path + ".tmp", self.path)  # appears fully sized and zeroed
        else:
            deadline = time.time() + 120
            fd = -1
            while fd < 0:
                try:
                    fd = os.open(self.path, os.O_RDWR)
                    if os.fstat(fd).st_size < size:
                        os.close(fd)
                        fd = -1
                except OSError:
                    fd = -1
                if fd < 0:
                    if time.time() > deadline:
                        raise TimeoutError(f"rendezvous file {self.path} did not appear")
                    time.sleep(0.005)
        self._mm = mmap.mmap(fd, size)
        os.close(fd)
        # 8-byte aligned numpy views: every access below is ONE load / store instruction.  (struct.pack_into with a "<"
        # format moves bytes one at a time, and a torn read of the epoch -- low byte old, high byte new -- looks like an
        # epoch from the future.)
        import numpy as np

        self._u64 = np.frombuffer(self._mm, dtype=np.uint64)
        self._f64 = np.frombuffer(self._mm, dtype=np.float64)
        self._i64 = np.frombuffer(self._mm, dtype=np.int64)
        self.barrier()

    def _put(self, value: float, aux: int):
        w = self.rank * (self.SLOT // 8)
        self._f64[w + 1] = value
        self._i64[w + 2] = aux
        self._u64[w] = self.epoch  # epoch last: it publishes the value (x86 stores are not reordered with older stores)

    def _wait(self):
        u64, step, ep = self._u64, self.SLOT // 8, self.epoch
        for r in range(self.world):
            spins = 0
            while int(u64[r * step]) < ep:
                spins += 1
                if spins > 2000:
                    time.sleep(0)  # yield: more ranks than cores (CPU tests)

    def _round(self, value: float = 0.0, aux: int = 0):
        """One publish + wait; returns every rank's (value, aux).  A second, value-less round fences the reads."""
        if self._mm is None:
            return [(value, aux)]
        self.epoch += 1
        self._put(value, aux)
        self._wait()
        step = self.SLOT // 8
        vals = [(float(self._f64[r * step + 1]), int(self._i64[r * step + 2])) for r in range(self.world)]
        self.epoch += 1
        self._u64[self.rank * step] = self.epoch
        self._wait()
        return vals

    def barrier(self):
        if self._mm is None:
            return
        self.epoch += 1
        self._u64[self.rank * (self.SLOT // 8)] = self.epoch
        self._wait()

    def allreduce(self, value: float, op: str = "max") -> float:
        vals = [v for v, _ in self._round(float(value))]
        return max(vals) if op == "max" else min(vals) if op == "min" else sum(vals)

    def broadcast_int(self, value: int, root: int = 0) -> int:
        return self._round(0.0, int(value))[root][1]

    def gather_int(self, value: int):
        return [a for _, a in self._round(0.0, int(value))]

    def close(self):
        if self._mm is not None:
            self.barrier()
            self._u64 = self._f64 = self._i64 = None  # release the buffer exports before closing the map
            self._mm.close()
            self._mm = None
            if self.rank == 0:
                try:
                    os.unlink(self.path)
                except OSError:
                    pass


class LocalCluster:
    def __init__(self, cluster_id: str = "local", n_workers: int = 2, pool_bytes: int = 64 << 20,
                 storage_class=None, serve_rpc: bool = True, coord: Optional[str] = None,
                 keystone_cfg: Optional["_bb.KeystoneConfig"] = None, lease_ttl_sec: int = 3,
                 heartbeat_interval_sec: int = 1, mount_path: str = ""):
        self.cluster_id = cluster_id
        cfg = keystone_cfg or _bb.KeystoneConfig()
        cfg.cluster_id = cluster_id
        cfg.listen_address = "127.0.0.1:0"
        cfg.http_metrics_port = "0"
        self.cfg = cfg
        self.coord_uri = coord if coord is not None else f"mem://{cluster_id}-{os.getpid()}-{id(self)}"
        self.coord = _bb.CoordService(self.coord_uri)
        assert self.coord.connect() == _bb.ErrorCode.OK
        self.keystone = _bb.KeystoneService(cfg, self.coord)
        assert self.keystone.initialize() == _bb.ErrorCode.OK
        assert self.keystone.start() == _bb.ErrorCode.OK
        self.rpc = None
        if serve_rpc:
            self.rpc = _bb.RpcService(self.keystone, cfg)
            assert self.rpc.start() == _bb.ErrorCode.OK
        self.workers = []
        sc = storage_class if storage_class is not None else _bb.StorageClass.RAM_CPU
        for i in range(n_workers):
            self.add_worker(f"worker-{i}", f"node-{i}", [(f"pool-{i}", sc, pool_bytes, mount_path)],
                            lease_ttl_sec, heartbeat_interval_sec)

    def add_worker(self, worker_id, node_id, pools, lease_ttl_sec=3, heartbeat_interval_sec=1, fabric_domain=""):
        wc = _bb.WorkerServiceConfig()
        wc.worker_id = worker_id
        wc.node_id = node_id
        wc.cluster_id = self.cluster_id
        wc.ucx_endpoint = "127.0.0.1:0"
        wc.lease_ttl_sec = lease_ttl_sec
        wc.heartbeat_interval_sec = heartbeat_interval_sec
        wc.fabric_domain = fabric_domain
        wc.storage_pools = [_bb.StoragePoolConfig(pid, sc, size, mp) for (pid, sc, size, mp) in pools]
        w = _bb.WorkerService(wc, _bb.CoordService(self.coord_uri))
        assert w.create_storage_pools_from_config() == _bb.ErrorCode.OK, "backend creation failed"
        assert w.initialize() == _bb.ErrorCode.OK
        assert w.start() == _bb.ErrorCode.OK
        self.workers.append(w)
        st = self.coord.store()
        if hasattr(st, "flush_events"):  # in-process store: registration events have been delivered when this returns
            st.flush_events()
        return w

    def client(self, node_id: str = "", io_parallelism: int = 4, in_process: bool = False, tenant: str = "", tenant_secret: str = ""):
        """A connected client.  `tenant` / `tenant_secret`: connect as that tenant (csrc/common/tenant.h) instead of as a
        member -- the identity is process-wide (like the cluster token), so it also covers this client's later data-server
        connections; pass tenant="" (the default) to go back to being a member."""
        if in_process or self.rpc is None:
            c = _bb.BlackbirdClient(_bb.LocalKeystoneApi(self.keystone), _bb.BlackbirdClientOptions(node_id=node_id, io_parallelism=io_parallelism))
        else:
            _bb.set_client_tenant(tenant, tenant_secret)
            c = _bb.BlackbirdClient(_bb.BlackbirdClientOptions("127.0.0.1", self.rpc.rpc_port, 30000, io_parallelism, node_id))
        assert c.connect() == _bb.ErrorCode.OK
        return c

    def stop(self):
        for w in self.workers:
            w.stop()
        self.workers = []
        if self.rpc is not None:
            self.rpc.stop()
        self.keystone.stop()
        if self.coord_uri.startswith("mem://"):
            _bb.drop_shared_mem_coord(self.coord_uri[len("mem://"):])

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.stop()


class GpuRankCluster:
    """One process per GPU.  Requires torch + CUDA; rendezvous through torch.distributed when world>1."""

    def __init__(self, slab_bytes: int, cluster_id: str = "gpu", keystone_port: Optional[int] = None,
                 nvls_arena_bytes: int = 0, nvls_group_size: int = 3, rpc_busy_poll_us: Optional[int] = None,
                 dram_bytes: int = 0, nvme_bytes: int = 0, nvme_path: str = "", high_watermark: float = 1.0,
                 eviction_ratio: float = 0.1, max_replicas: int = 3, use_nccl: bool = False, nvls_groups: str = "all"):
        """slab_bytes: HBM slab of this rank's GPU-tier pool.  dram_bytes / nvme_bytes add host tiers to the
        same worker (the demotion ladder GPU -> DRAM -> NVMe); high_watermark < 1 arms tier demotion."""
        import torch
        import torch.distributed as dist

        self.torch = torch
        self.nvls_groups = nvls_groups  # "all": every replica set gets a multicast group; "ring": {g, g+1, .., g+R-1} only
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.cpu_affinity = _pin_to_gpu_numa_node(self.local_rank)
        # Bring-up needs no collective library: the ranks of the node meet through a /dev/shm rendezvous (keystone port,
        # barriers, scalar reductions).  NCCL is initialised only on request (ensure_dist(): the benchmark comparators).
        self.rdv = HostRendezvous(self.rank, self.world)
        self.dist = None
        if self.world > 1 and (use_nccl or dist.is_initialized()):
            self.ensure_dist()
        _bb.install_gpu_backend_factory()
        self.node_id = f"gpu{self.rank}"
        self.keystone = None
        self.rpc = None
        port = 0
        if self.rank == 0:
            cfg = _bb.KeystoneConfig()
            cfg.cluster_id = cluster_id
            cfg.listen_address = f"127.0.0.1:{keystone_port or 0}"
            cfg.http_metrics_port = "0"
            cfg.enable_gc = False
            cfg.high_watermark = high_watermark
            cfg.eviction_ratio = eviction_ratio
            cfg.gc_interval_sec = 3600 if high_watermark >= 1.0 else 1
            cfg.rpc_threads = 4
            cfg.max_replicas = max_replicas
            if rpc_busy_poll_us is None:
                rpc_busy_poll_us = int(os.environ.get("BB_RPC_BUSY_POLL_US", "2000"))
            cfg.rpc_busy_poll_us = rpc_busy_poll_us  # clients issue RPCs between kernels a few ms apart
            self.keystone = _bb.KeystoneService(cfg, None)
            assert self.keystone.initialize() == _bb.ErrorCode.OK
            assert self.keystone.start() == _bb.ErrorCode.OK
            self.keystone.install_data_server_mover()  # demotion / repair copies run worker-to-worker (D_COPY)
            self.rpc = _bb.RpcService(self.keystone, cfg)
            assert self.rpc.start() == _bb.ErrorCode.OK
            port = self.rpc.rpc_port
        self.keystone_port = self.rdv.broadcast_int(port, 0)
        if self.rank == 0:
            self.api = _bb.LocalKeystoneApi(self.keystone)
        else:
            self.api = _bb.KeystoneRpcClient()
            assert self.api.connect("127.0.0.1", self.keystone_port, 10000) == _bb.ErrorCode.OK
        # worker: one GPU-tier pool on this rank's device, registered directly with the keystone
        wc = _bb.WorkerServiceConfig()
        wc.worker_id = f"worker-gpu{self.rank}"
        wc.node_id = self.node_id
        wc.cluster_id = cluster_id
        wc.ucx_endpoint = "127.0.0.1:0"
        wc.interconnects = ["nvlink", "tcp"]
        wc.max_bw_gbps = 7200.0
        wc.fabric_domain = "nvswitch-0"
        wc.lease_ttl_sec = 30
        wc.heartbeat_interval_sec = 5
        wc.storage_pools = [_bb.StoragePoolConfig(f"hbm{self.rank}", _bb.StorageClass.RAM_GPU, slab_bytes, "", self.local_rank)]
        if dram_bytes:
            dram = _bb.StoragePoolConfig(f"dram{self.rank}", _bb.StorageClass.RAM_CPU, dram_bytes, "")
            dram.pin_memory = True  # registered with CUDA: GPU <-> DRAM tier moves are one fused-kernel launch
            dram.shared_memory = True  # memfd-backed: GPU clients of the other ranks map it and read it over PCIe
            wc.storage_pools += [dram]
        if nvme_bytes:
            wc.storage_pools += [_bb.StoragePoolConfig(f"nvme{self.rank}", _bb.StorageClass.NVME, nvme_bytes, nvme_path or "/tmp")]
        self.worker = _bb.WorkerService(wc, None, self.api)
        assert self.worker.create_storage_pools_from_config() == _bb.ErrorCode.OK
        assert self.worker.initialize() == _bb.ErrorCode.OK
        assert self.worker.start() == _bb.ErrorCode.OK
        self.barrier()
        opts = _bb.BlackbirdClientOptions("127.0.0.1", self.keystone_port, 60000, 4, self.node_id)
        self.client_api = self.api if self.rank == 0 else self._new_api()
        self.client = _bb.BlackbirdClient(self.client_api, opts)
        assert self.client.connect() == _bb.ErrorCode.OK
        self.fabric = _bb.GpuFabric(self.local_rank, self.client_api)
        assert self.fabric.mapped_pools() >= self.world, f"mapped {self.fabric.mapped_pools()} of {self.world} slabs"
        _bb.attach_fabric(self.client, self.fabric)
        self.barrier()
        self.arena = None
        if nvls_arena_bytes:
            self._setup_nvls(cluster_id, nvls_arena_bytes, nvls_group_size)

    def _setup_nvls(self, cluster_id: str, arena_bytes: int, group_size: int):
        """Replica arenas: ring groups {g, g+1, .., g+R-1} of R GPUs, each bound to one NVSwitch multicast object."""
        R = min(group_size, self.world)
        if R < 2 or not _bb.NvlsArena.supported(self.local_rank):
            return
        import itertools
        import math

        if R == self.world:
            groups = [list(range(self.world))]
        elif self.nvls_groups == "all" and math.comb(self.world, R) <= 64:
            # one multicast group per possible replica set (C(8,3) = 56): whatever set of GPUs the placement engine
            # picks for a symmetric R-way put, there is a multicast object for exactly those GPUs -- no ring constraint
            groups = [list(c) for c in itertools.combinations(range(self.world), R)]
        else:
            groups = [[(g + k) % self.world for k in range(R)] for g in range(self.world)]
        tag = f"{cluster_id}-{self.keystone_port}"
        arena = _bb.NvlsArena(self.local_rank, self.rank, self.world, tag, groups, arena_bytes)
        for phase in (arena.phase1_create, arena.phase2_join, arena.phase3_bind, arena.phase4_map_peers):
            ec = phase()
            if self.rdv.allreduce(1.0 if ec == _bb.ErrorCode.OK else 0.0, "min") == 0.0:
                if self.rank == 0:
                    print(f"[blackbird_b200] NVLS arena unavailable ({phase.__name__}: {arena.last_error}); replicas use unicast fan-out")
                return
            self.barrier()
        self.arena = arena
        self.fabric.set_arena(arena)
        # every member registers its arena of every group it belongs to as a RAM_GPU pool of that multicast domain
        for g in range(arena.num_groups()):
            if not arena.member_of(g):
                continue
            pool = _bb.MemoryPool(_bb.NvlsArena.pool_id(g, self.rank), arena.arena_bytes, _bb.StorageClass.RAM_GPU, self.node_id,
                                  f"worker-gpu{self.rank}", self.worker.data_endpoint(), 0, "", self.local_rank, 7200.0,
                                  _bb.NvlsArena.domain(g))
            assert self.api.register_memory_pool(pool) == _bb.ErrorCode.OK
        self.barrier()

    def _new_api(self):
        api = _bb.KeystoneRpcClient()
        assert api.connect("127.0.0.1", self.keystone_port, 10000) == _bb.ErrorCode.OK
        return api

    def ensure_dist(self):
        """torch.distributed over NCCL, for callers that want collectives of their own (benchmark comparators)."""
        import torch.distributed as dist

        if self.world > 1 and not dist.is_initialized():
            dist.init_process_group("nccl", device_id=self.torch.device("cuda", self.local_rank))
        self.dist = dist if self.world > 1 else None
        return self.dist

    def barrier(self):
        """Device work of this rank done + every rank got here (host-side, shared memory)."""
        self.torch.cuda.synchronize()
        self.rdv.barrier()

    def host_barrier(self):
        """Ranks rendezvous without touching the device (calls of the synchronous client API have completed)."""
        self.rdv.barrier()

    def max_over_ranks(self, v: float) -> float:
        return self.rdv.allreduce(v, "max")

    def sum_over_ranks(self, v: float) -> float:
        return self.rdv.allreduce(v, "sum")

    def stop(self):
        self.barrier()
        self.client = None
        self.fabric = None
        self.arena = None
        self.worker.stop()
        self.barrier()
        if self.rpc is not None:
            self.rpc.stop()
        if self.keystone is not None:
            self.keystone.stop()
        self.rdv.close()


class CpuRankCluster:
    """The GpuRankCluster topology without GPUs: one process per rank (torchrun, `gloo`), rank 0 hosts the Keystone,
    every rank runs a worker with a memfd-backed DRAM pool and a client whose device batch API goes through the
    HostLoopbackTransport ("device pointers" are host buffers).  Shards placed on another rank's pool are written by
    one-sided memcpy into that rank's process (same host) or over its TCP data server.  It exists so that the multi-rank
    host-side logic -- rendezvous, registration over RPC, ring placement, cross-process pool mapping, fan-out -- runs on
    CPU-only machines (tests/test_multi_cpu.py)."""

    def __init__(self, dram_bytes: int = 64 << 20, cluster_id: str = "cpu", shared_memory: bool = True):
        import torch
        import torch.distributed as dist

        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.dist = dist if self.world > 1 else None
        self._owns_pg = False
        if self.world > 1 and not dist.is_initialized():
            dist.init_process_group("gloo")
            self._owns_pg = True
        self.node_id = f"cpu{self.rank}"
        self.keystone = self.rpc = None
        port_t = torch.zeros(1, dtype=torch.int64)
        if self.rank == 0:
            cfg = _bb.KeystoneConfig()
            cfg.cluster_id = cluster_id
            cfg.listen_address = "127.0.0.1:0"
            cfg.http_metrics_port = "0"
            cfg.enable_gc = False
            cfg.rpc_threads = 4
            self.keystone = _bb.KeystoneService(cfg, None)
            assert self.keystone.initialize() == _bb.ErrorCode.OK and self.keystone.start() == _bb.ErrorCode.OK
            self.keystone.install_data_server_mover()
            self.rpc = _bb.RpcService(self.keystone, cfg)
            assert self.rpc.start() == _bb.ErrorCode.OK
            port_t[0] = self.rpc.rpc_port
        if self.world > 1:
            dist.broadcast(port_t, 0)
        self.keystone_port = int(port_t.item())
        self.api = _bb.LocalKeystoneApi(self.keystone) if self.rank == 0 else self._new_api()
        wc = _bb.WorkerServiceConfig()
        wc.worker_id, wc.node_id, wc.cluster_id = f"worker-cpu{self.rank}", self.node_id, cluster_id
        wc.ucx_endpoint = "127.0.0.1:0"
        wc.lease_ttl_sec, wc.heartbeat_interval_sec = 30, 5
        pool = _bb.StoragePoolConfig(f"dram{self.rank}", _bb.StorageClass.RAM_CPU, dram_bytes, "")
        pool.shared_memory = shared_memory
        wc.storage_pools = [pool]
        self.worker = _bb.WorkerService(wc, None, self.api)
        assert self.worker.create_storage_pools_from_config() == _bb.ErrorCode.OK
        assert self.worker.initialize() == _bb.ErrorCode.OK and self.worker.start() == _bb.ErrorCode.OK
        self.barrier()
        opts = _bb.BlackbirdClientOptions("127.0.0.1", self.keystone_port, 60000, 4, self.node_id)
        self.client_api = self.api if self.rank == 0 else self._new_api()
        self.client = _bb.BlackbirdClient(self.client_api, opts)
        self.io_client = _bb.BlackbirdClient(self.client_api, opts)
        assert self.client.connect() == _bb.ErrorCode.OK and self.io_client.connect() == _bb.ErrorCode.OK
        _bb.attach_loopback_transport(self.client, self.io_client)
        self.barrier()

    def _new_api(self):
        api = _bb.KeystoneRpcClient()
        assert api.connect("127.0.0.1", self.keystone_port, 10000) == _bb.ErrorCode.OK
        return api

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def stop(self):
        self.barrier()
        self.client = self.io_client = None
        self.worker.stop()
        self.barrier()
        if self.rpc is not None:
            self.rpc.stop()
        if self.keystone is not None:
            self.keystone.stop()
        if self._owns_pg:
            self.dist.destroy_process_group()
