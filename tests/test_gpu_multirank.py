"""The partitioned path on real kernels: several ranks (processes) drive the engine on this box's
GPU(s); collectives go through the host back-end (gloo) so that a single-GPU box can run it, and
through NCCL when the box has at least two GPUs.  Checks: the partitioned run follows the
single-rank trajectory (region counts per iteration, final result), every rank returns the same
global result, and the repartition actually moves regions."""
import json
import os
import socket

import pytest
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir, backend, family, ndim, epsrel, rebalance):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch

    import gpuintegration_b200 as pg
    from gpuintegration_b200 import parallel

    ngpu = torch.cuda.device_count()
    device = rank % ngpu
    torch.cuda.set_device(device)
    parallel.init_process_group(backend="gloo" if backend == "host" else "nccl")
    ws = pg.Workspace(ndim, device=device, trace=True)
    if backend == "host":
        parallel.attach_host_collectives(ws)
    else:
        parallel.attach_nccl(ws)
    ws.set_rebalance(*rebalance)
    r = ws.integrate(pg.Genz(family, ndim), epsrel, 1e-40)
    tr = ws.trace()
    with open(os.path.join(out_dir, f"rank{rank}.json"), "w") as f:
        json.dump({"estimate": r.estimate, "errorest": r.errorest, "nregions": r.nregions, "status": r.status,
                   "iters": r.iters, "neval": r.neval,
                   "n_global": [t["nregions_global"] for t in tr], "n_local": [t["nregions"] for t in tr],
                   "nactive": [t["nactive"] for t in tr], "rebalanced": [t["rebalanced"] for t in tr],
                   "heuristic_ran": [t["heuristic_ran"] for t in tr], "threshold": [t["threshold"] for t in tr],
                   "heuristic_probes": [t["heuristic_probes"] for t in tr]}, f)
    ws.close()
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()


def _run(tmp_path, world, backend, family, ndim, epsrel, rebalance=(1.10, 4096)):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path), backend, family, ndim, epsrel, rebalance),
             nprocs=world, join=True)
    return [json.load(open(tmp_path / f"rank{r}.json")) for r in range(world)]


def _single(pg, family, ndim, epsrel):
    with pg.Workspace(ndim, trace=True) as ws:
        r = ws.integrate(pg.Genz(family, ndim), epsrel, 1e-40)
        return r, ws.trace()


@pytest.mark.parametrize("world,family,ndim,epsrel", [(2, 5, 8, 1e-3), (3, 6, 6, 1e-3), (2, 1, 8, 1e-3)])
def test_partitioned_run_follows_single_rank_trajectory(pg, tmp_path, world, family, ndim, epsrel):
    ranks = _run(tmp_path, world, "host", family, ndim, epsrel)
    r1, t1 = _single(pg, family, ndim, epsrel)
    for rk in ranks:
        assert rk["n_global"] == [t["nregions_global"] for t in t1]
        assert rk["nactive"] == [t["nactive"] for t in t1]
        assert (rk["status"], rk["nregions"], rk["iters"], rk["neval"]) == (r1.status, r1.nregions, r1.iters, r1.neval)
        assert abs(rk["estimate"] - r1.estimate) <= 1e-12 * max(abs(r1.estimate), r1.errorest)
        assert rk["estimate"] == ranks[0]["estimate"] and rk["errorest"] == ranks[0]["errorest"]
    # the local lists partition the global one at every iteration
    for i in range(len(t1)):
        assert sum(rk["n_local"][i] for rk in ranks) == t1[i]["nregions_global"]


def test_rebalance_moves_regions_and_keeps_the_trajectory(pg, tmp_path):
    # the discontinuous integrand finishes most of the cube early: one rank's slice dries up
    ranks = _run(tmp_path, 2, "host", 6, 6, 1e-3, rebalance=(1.0, 2))
    r1, t1 = _single(pg, 6, 6, 1e-3)
    assert any(any(rk["rebalanced"]) for rk in ranks)
    for rk in ranks:
        assert rk["n_global"] == [t["nregions_global"] for t in t1]
        assert (rk["status"], rk["nregions"]) == (r1.status, r1.nregions)
    # after a rebalance the ranks' local counts differ by at most 2 (pairs of children)
    for i, flag in enumerate(ranks[0]["rebalanced"][:-1]):
        if flag:
            a, b = ranks[0]["n_local"][i + 1], ranks[1]["n_local"][i + 1]
            assert abs(a - b) <= 2


def test_nccl_backend_when_two_gpus(pg, tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (NCCL refuses two ranks on one device)")
    ranks = _run(tmp_path, 2, "nccl", 5, 8, 1e-3, rebalance=(1.0, 2))
    r1, t1 = _single(pg, 5, 8, 1e-3)
    for rk in ranks:
        assert rk["n_global"] == [t["nregions_global"] for t in t1]
        assert (rk["status"], rk["nregions"]) == (r1.status, r1.nregions)
        assert abs(rk["estimate"] - r1.estimate) <= 1e-12 * max(abs(r1.estimate), r1.errorest)


def test_nccl_in_kernel_search_follows_single_rank(pg, tmp_path):
    """Heuristic threshold search with its per-probe all-reduce done INSIDE the kernel over NVLink peer
    mailboxes (csrc/threshold_search.cuh, k_search<true>): the two ranks must walk through the same probes
    and accept the same thresholds as the single-rank device search."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ranks = _run(tmp_path, 2, "nccl", 1, 8, 1e-4)
    r1, t1 = _single(pg, 1, 8, 1e-4)
    assert any(t["heuristic_ran"] for t in t1)
    for rk in ranks:
        assert rk["n_global"] == [t["nregions_global"] for t in t1]
        assert rk["nactive"] == [t["nactive"] for t in t1]
        assert rk["heuristic_ran"] == [t["heuristic_ran"] for t in t1]
        assert rk["heuristic_probes"] == [t["heuristic_probes"] for t in t1]
        assert rk["threshold"] == pytest.approx([t["threshold"] for t in t1], rel=1e-12)
        assert (rk["status"], rk["nregions"], rk["iters"]) == (r1.status, r1.nregions, r1.iters)
        assert abs(rk["estimate"] - r1.estimate) <= 1e-12 * max(abs(r1.estimate), r1.errorest)
        assert rk["estimate"] == ranks[0]["estimate"] and rk["errorest"] == ranks[0]["errorest"]


def _golden(family, ndim, eps):
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "trajectories", f"f{family}_{ndim}_{eps:g}.json")
    return json.load(open(path))


# PGN_TEST_WORLDS=4,8 restricts the rank counts (a multi-GPU box is charged per GPU-minute)
@pytest.mark.parametrize("world", [int(w) for w in os.environ.get("PGN_TEST_WORLDS", "2,4,8").split(",")])
@pytest.mark.parametrize("family,ndim,epsrel", [(1, 8, 1e-4), (4, 8, 1e-9), (5, 10, 1e-3)])
def test_nccl_ranks_follow_the_reference_golden(tmp_path, world, family, ndim, epsrel):
    """One process per GPU over NCCL + the in-kernel NVLink exchanges (fused statistics, threshold search): every rank
    must walk through the trajectory the UNMODIFIED single-GPU reference recorded (tests/golden/trajectories) --
    global region counts, survivors and heuristic-filter decisions per iteration, status / nregions / iters, estimate
    to 1e-12.  (4, 8, 1e-9) is BASELINE config 5: F_4_8D, epsrel 1e-9, partitioned region list."""
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    doc = _golden(family, ndim, epsrel)
    its, s = doc["iterations"], doc["summary"]
    ranks = _run(tmp_path, world, "nccl", family, ndim, epsrel)
    for rk in ranks:
        assert rk["n_global"] == [r["n"] for r in its]
        assert rk["heuristic_ran"] == [r["heuristic_ran"] for r in its]
        for i, r in enumerate(its):
            if r["outcome"] == "split":
                assert rk["nactive"][i] == r["nactive"], i
        assert (rk["status"], rk["nregions"], rk["iters"]) == (s["status"], s["nregions"], s["iters"])
        assert abs(rk["estimate"] - s["estimate"]) <= 1e-12 * max(abs(s["estimate"]), s["errorest"])
        assert rk["estimate"] == ranks[0]["estimate"] and rk["errorest"] == ranks[0]["errorest"]
    for i in range(len(its)):
        assert sum(rk["n_local"][i] for rk in ranks) == its[i]["n"]
