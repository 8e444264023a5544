"""world_size = 2 on CPU (gloo): the launcher-side plumbing of the N>1 path -- rendezvous on
127.0.0.1, the host-collective callbacks the engine calls, max-over-ranks timing, and the
repartition plan applied to real per-rank arrays moved with send/recv."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from gpuintegration_b200 import parallel

    r, lr, w = parallel.init_process_group(backend="gloo")
    assert (r, w) == (rank, world)
    # timing rule: max over ranks
    assert parallel.max_over_ranks(10.0 + rank) == 10.0 + world - 1

    # per-rank "parents": cols x m_local, labelled by global position so order can be checked
    counts = [1000, 37] if world == 2 else [5 * (k + 1) for k in range(world)]
    cols = 19  # 2 * 8 bounds columns + parent estimate + parent error + axis
    start = sum(counts[:rank])
    local = np.stack([np.arange(start, start + counts[rank], dtype=np.float64) + 0.001 * c for c in range(cols)])

    def exchange(sends, recv_counts):
        reqs, bufs = [], {}
        for peer, arr in sends.items():
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(arr)), peer))
        for peer, cnt in recv_counts.items():
            bufs[peer] = torch.zeros((cols, cnt), dtype=torch.float64)
            reqs.append(dist.irecv(bufs[peer], peer))
        for q in reqs:
            q.wait()
        return {p: b.numpy() for p, b in bufs.items()}

    new_local = parallel.apply_plan_numpy(local, counts, rank, exchange)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), new_local)

    # the all-reduce callback signature the engine uses (sum and max over a small double vector)
    import ctypes as C
    got = {}

    class FakeWs:
        def comm_init_host(self, nranks, rk, allreduce, alltoallv):
            got["ar"], got["a2a"], got["n"], got["r"] = allreduce, alltoallv, nranks, rk

    parallel.attach_host_collectives(FakeWs())
    assert (got["n"], got["r"]) == (world, rank)
    buf = (C.c_double * 4)(1.0 + rank, 2.0, -rank, 7.0)
    assert got["ar"](None, buf, 4, 0) == 0
    assert list(buf) == [sum(1.0 + k for k in range(world)), 2.0 * world, -sum(range(world)), 7.0 * world]
    buf = (C.c_double * 2)(float(rank), -float(rank))
    assert got["ar"](None, buf, 2, 1) == 0
    assert list(buf) == [float(world - 1), 0.0]
    # all-to-all-v callback: rank r sends (r + 1) doubles to every peer
    n_each = rank + 1
    send = (C.c_double * (n_each * world))(*[100.0 * rank + p for p in range(world) for _ in range(n_each)])
    sc = (C.c_uint64 * world)(*[0 if p == rank else n_each for p in range(world)])
    so = (C.c_uint64 * world)(*[p * n_each for p in range(world)])
    rc = (C.c_uint64 * world)(*[0 if p == rank else p + 1 for p in range(world)])
    ro_list, acc = [], 0
    for p in range(world):
        ro_list.append(acc)
        acc += 0 if p == rank else p + 1
    ro = (C.c_uint64 * world)(*ro_list)
    recv = (C.c_double * max(acc, 1))()
    assert got["a2a"](None, send, sc, so, recv, rc, ro) == 0
    for p in range(world):
        if p != rank:
            assert list(recv[ro_list[p]:ro_list[p] + p + 1]) == [100.0 * p + rank] * (p + 1)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_two_rank_partition_plumbing(tmp_path, world):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"rank{r}.npy") for r in range(world)]
    sizes = [p.shape[1] for p in parts]
    assert max(sizes) - min(sizes) <= 1
    whole = np.concatenate(parts, axis=1)
    total = whole.shape[1]
    for c in range(whole.shape[0]):
        assert np.array_equal(whole[c], np.arange(total, dtype=np.float64) + 0.001 * c)
