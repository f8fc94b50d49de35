"""N > 1 host logic on CPU (gloo, world_size 2): column sharding of the Monte-Carlo samples, the all-reduce of the
per-evaluation sufficient sums and the parameter broadcast reproduce the single-rank result.  The per-rank sums come
from the CPU oracle here; on the GPU box the same algebra runs in gmb_model_loglik / gmb_model_mcnr over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def shard(m_total, rank, world):
    """Consecutive column ranges, the layout gmb_model_set_u expects (trailing rank drops the niter remainder)."""
    per = (m_total + world - 1) // world
    lo = min(rank * per, m_total)
    return lo, min(lo + per, m_total)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from glmmrmcml_b200 import synth
    cfg = synth.config2(m=37)
    fl = oracle.flink(cfg["family"], cfg["link"])
    m = cfg["U"].shape[1]
    lo, hi = shard(m, rank, world)
    Uloc = cfg["U"][:, lo:hi]
    # rank 0 owns the parameters and broadcasts them (beta, theta, sigma), SURVEY §8e
    par = torch.tensor(np.concatenate([cfg["beta"], cfg["theta"], [1.0]])) if rank == 0 else torch.zeros(cfg["P"] + 3, dtype=torch.float64)
    dist.broadcast(par, src=0)
    beta = par[: cfg["P"]].numpy()
    # E-step objective: local sum over the rank's columns, all-reduced, divided by m_total
    zd = oracle.gemm(cfg["Z"], Uloc)
    xb = cfg["X"] @ beta
    _, per_sample = oracle.loglik_zd(zd, xb, cfg["y"], 1.0, fl, per_sample=True)
    s = torch.tensor([per_sample.sum(), float(hi - lo)], dtype=torch.float64)
    dist.all_reduce(s)
    ll = s[0].item() / s[1].item()
    # MCNR: P^2 + P + 1 sufficient sums
    loc = oracle.mcnr(cfg["X"], cfg["Z"], Uloc, cfg["y"], beta, 1.0, fl)
    nl = hi - lo
    pack = torch.tensor(np.concatenate([loc["xtwx"].ravel(order="F") * nl, loc["score"] * nl, [loc["sigma"] * nl]]))
    dist.all_reduce(pack)
    P = cfg["P"]
    xtwx = pack[: P * P].numpy().reshape(P, P, order="F") / m
    score = pack[P * P: P * P + P].numpy() / m
    sigma = pack[-1].item() / m
    # mvn_ll
    dsum = torch.tensor([oracle.mvn_loglik(cfg["cov"], cfg["data"], cfg["eff_range"], cfg["theta"], Uloc) * nl], dtype=torch.float64)
    dist.all_reduce(dsum)
    if rank == 0:
        full = oracle.mcnr(cfg["X"], cfg["Z"], cfg["U"], cfg["y"], cfg["beta"], 1.0, fl)
        q.put(dict(ll=ll, ll_full=oracle.loglik_faithful(cfg["X"], cfg["Z"], cfg["U"], cfg["y"], cfg["beta"], 1.0, fl),
                   xtwx_err=float(np.max(np.abs(xtwx - full["xtwx"]))), score_err=float(np.max(np.abs(score - full["score"]))),
                   sigma_err=abs(sigma - full["sigma"]), d=dsum.item() / m,
                   d_full=oracle.mvn_loglik(cfg["cov"], cfg["data"], cfg["eff_range"], cfg["theta"], cfg["U"])))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_estep_equals_single_rank():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert abs(res["ll"] - res["ll_full"]) <= 1e-12 * abs(res["ll_full"])
    assert res["xtwx_err"] < 1e-12 and res["score_err"] < 1e-12 and res["sigma_err"] < 1e-13
    assert abs(res["d"] - res["d_full"]) <= 1e-12 * abs(res["d_full"])


def _chain_worker(rank, world, port, q):
    """One MCML iteration of src/mcml_full.cpp:83-126 with the CHAINS split over the ranks the way gmb_mcml_full does it
    (csrc/api.cpp: C_local = ceil(C / world), rank r runs the global chains [r C_local, (r+1) C_local) on the Philox streams of their
    global index, every rank all-reduces its MCNR sums): same samples and same Newton step as all chains on one rank."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from glmmrmcml_b200 import synth
    cfg = synth.config2(m=8)
    fl = oracle.flink(cfg["family"], cfg["link"])
    X, Z, y, L, beta = cfg["X"], cfg["Z"], cfg["y"], cfg["L"], cfg["beta"]
    ZL = oracle.gemm(Z, L); xb = X @ beta
    C_total, per, seed = 6, 4, 20221208
    C_local = (C_total + world - 1) // world

    def draw(chains):
        cols = [oracle.hmc_chain(ZL, L, xb, y, 1.0, fl, 12, per - 1, 0.3, 20, 0.9, seed, chain=c)["u"] for c in chains]
        return np.asfortranarray(np.concatenate(cols, axis=1))          # chain-major, per columns each (column 0 = state after warm-up)

    Uloc = draw(range(rank * C_local, (rank + 1) * C_local))
    nl, tot = Uloc.shape[1], C_total * per
    loc = oracle.mcnr(X, Z, Uloc, y, beta, 1.0, fl)
    P = cfg["P"]
    pack = torch.tensor(np.concatenate([loc["xtwx"].ravel(order="F") * nl, loc["score"] * nl]))
    dist.all_reduce(pack)
    xtwx = pack[: P * P].numpy().reshape(P, P, order="F") / tot
    step = np.linalg.solve(xtwx, pack[P * P:].numpy() / tot)
    gathered = [torch.zeros(Uloc.shape, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.tensor(np.ascontiguousarray(Uloc)))
    if rank == 0:
        Uall = draw(range(C_total))
        full = oracle.mcnr(X, Z, Uall, y, beta, 1.0, fl)
        q.put(dict(u_equal=bool(np.array_equal(np.concatenate([g.numpy() for g in gathered], axis=1), Uall)),
                   step_err=float(np.max(np.abs(step - full["beta_incr"]))), cols=int(Uall.shape[1]), tot=tot))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_chain_sharding_reproduces_the_single_rank_iteration():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_chain_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res["cols"] == res["tot"] and res["u_equal"]          # same chains, same streams, chain-major order
    assert res["step_err"] < 1e-11


def test_shard_covers_all_columns():
    for m in (1, 7, 10_000, 100_003):
        for w in (1, 2, 4, 8):
            cols = [shard(m, r, w) for r in range(w)]
            assert cols[0][0] == 0 and cols[-1][1] == m
            assert all(cols[i][1] == cols[i + 1][0] for i in range(w - 1))
