"""Multi-GPU plumbing: one process per GPU, launched by torchrun; torch.distributed is only the
rendezvous and (for the host back-end used in tests) the transport.  The data path of the
partitioner itself lives in the C++ engine (csrc/engine.cu, csrc/comm.cc):

  * every rank holds a contiguous slice of the region list and runs evaluate / classify / compact /
    split on it with no data-path collective;
  * per iteration ONE all-reduce of a 32-double vector (sums, counts, one-hot per-rank active counts)
    enqueued on the compute stream right behind the classify kernel (ncclAllReduce over NVLink);
  * when max/mean of the per-rank active counts exceeds the threshold, surviving parents are
    repartitioned order-preservingly with grouped ncclSend/ncclRecv (plan: pgn_plan_rebalance).

The reference has no multi-GPU path at all; this is new.
"""
import ctypes as C
import os

import numpy as np

from . import pagani


def env_rank():
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) when absent."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init_process_group(backend=None):
    """Join the torchrun rendezvous (127.0.0.1 by default, as the container hostname may not resolve)."""
    import torch
    import torch.distributed as dist

    rank, local_rank, world = env_rank()
    if world == 1:
        return rank, local_rank, world
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local_rank)
    if not dist.is_initialized():
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def attach_nccl(ws, group=None):
    """Give the engine its own NCCL communicator (same ranks as the torch process group)."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if rank == 0:
        uid = torch.tensor(list(pagani.nccl_unique_id()), dtype=torch.uint8)
    else:
        uid = torch.zeros(pagani.capi.PGN_NCCL_ID_BYTES, dtype=torch.uint8)
    if dist.get_backend(group) == "nccl":
        uid = uid.cuda()
    dist.broadcast(uid, src=0, group=group)
    ws.comm_init_nccl(bytes(uid.cpu().tolist()), world, rank)


def attach_host_collectives(ws, group=None):
    """Route the engine's collectives through torch.distributed on host buffers (gloo).

    Lets several ranks share one GPU (tests) and exercises the same partition / rebalance logic as
    the NCCL back-end.
    """
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)

    def allreduce(_user, buf, n, op):
        a = np.ctypeslib.as_array(buf, shape=(n,))
        t = torch.from_numpy(a)
        dist.all_reduce(t, op=dist.ReduceOp.SUM if op == 0 else dist.ReduceOp.MAX, group=group)
        return 0

    def alltoallv(_user, send, scount, soff, recv, rcount, roff):
        sc = [scount[r] for r in range(world)]
        so = [soff[r] for r in range(world)]
        rc = [rcount[r] for r in range(world)]
        ro = [roff[r] for r in range(world)]
        stot, rtot = sum(sc), sum(rc)
        # a rank that sends (receives) nothing gets a NULL pointer from the engine
        empty = torch.empty(0, dtype=torch.float64)
        sbuf = torch.from_numpy(np.ctypeslib.as_array(send, shape=(stot,))) if stot else empty
        rbuf = torch.from_numpy(np.ctypeslib.as_array(recv, shape=(rtot,))) if rtot else empty
        ops = []
        for peer in range(world):
            if peer == rank:
                continue
            if sc[peer]:
                ops.append(dist.P2POp(dist.isend, sbuf[so[peer]:so[peer] + sc[peer]], peer, group))
            if rc[peer]:
                ops.append(dist.P2POp(dist.irecv, rbuf[ro[peer]:ro[peer] + rc[peer]], peer, group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return 0

    ws.comm_init_host(world, rank, allreduce, alltoallv)


def max_over_ranks(value, group=None):
    """Max of a host scalar over all ranks (the timing rule: report the slowest rank)."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def apply_plan_numpy(local, counts, rank, exchange):
    """Reference implementation of the repartition on host arrays (used by the gloo CPU test).

    local: [cols][m_local] array of this rank; exchange(sends) -> recvs moves {peer: array}.
    Returns the rank's new [cols][m_new] array.
    """
    new_counts, transfers = pagani.plan_rebalance(counts)
    cols = local.shape[0]
    sends = {}
    for src, dst, so, _do, cnt in transfers:
        if src == rank and dst != rank:
            sends.setdefault(dst, []).append(local[:, so:so + cnt])
    sends = {p: np.concatenate(v, axis=1) for p, v in sends.items()}
    recvs = exchange(sends, {src: cnt for src, dst, _so, _do, cnt in transfers if dst == rank and src != rank})
    out = np.zeros((cols, new_counts[rank]), dtype=local.dtype)
    taken = {p: 0 for p in recvs}
    for src, dst, so, do, cnt in transfers:
        if dst != rank:
            continue
        if src == rank:
            out[:, do:do + cnt] = local[:, so:so + cnt]
        else:
            out[:, do:do + cnt] = recvs[src][:, taken[src]:taken[src] + cnt]
            taken[src] += cnt
    return out
