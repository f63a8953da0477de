"""HostRendezvous: the shared-memory rendezvous GpuRankCluster uses instead of NCCL for bring-up (barrier, scalar
all-reduce, integer broadcast / gather)."""
import multiprocessing as mp
import os


def _worker(rank, world, tag, q):
    from blackbird_b200.parallel.cluster import HostRendezvous

    r = HostRendezvous(rank, world, tag)
    out = []
    for it in range(200):
        out.append(r.allreduce(rank + it, "max") == world - 1 + it)
        out.append(r.allreduce(rank + 1, "sum") == world * (world + 1) / 2)
        out.append(r.broadcast_int(1000 + it if rank == 2 % world else -1, 2 % world) == 1000 + it)
        out.append(r.gather_int(rank * 7 + it) == [k * 7 + it for k in range(world)])
        r.barrier()
    r.close()
    q.put((rank, all(out)))


def test_host_rendezvous_barrier_reduce_broadcast_gather():
    world, tag = 4, f"test-{os.getpid()}"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, tag, q)) for r in range(world)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(30) for p in ps]
    assert res == [(r, True) for r in range(world)]
    assert not os.path.exists(f"/dev/shm/bb-rdv-{tag}")
