"""Worker of tests/test_gpu_multirank.py (launched by torch.distributed.run, one rank per GPU): the NCCL-sharded E-step, covariance
objective and MCML loop against the SAME quantities computed by one rank alone.  Rank 0 prints one JSON object."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def main():
    import torch
    import torch.distributed as dist
    import glmmrmcml_b200 as g
    from glmmrmcml_b200 import synth
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    ctx = g.Context(lr)
    ids = [g.Context.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.comm_init(ids[0], rank, world)
    solo = g.Context(lr)                       # no communicator: the single-rank reference on the same GPU
    res = {"world": world}
    cases = {"C2": synth.config2(m=2001), "C4": synth.config4(ncl=40, nt=10, k=1, m=1500), "C3": synth.config3(nloc=200, m=130),
             "C1": synth.config1(m=251)}
    rng = np.random.default_rng(3)
    for name, cfg in cases.items():
        m = cfg["m"]; P = cfg["P"]; sig = 1.3 if cfg["family"] == "gaussian" else 1.0
        lo, hi = rank * m // world, (rank + 1) * m // world
        B = np.asfortranarray(cfg["beta"][:, None] + 1e-3 * rng.standard_normal((P, 24)))
        T = np.asfortranarray(cfg["theta"][:, None] * (1 + 1e-3 * rng.standard_normal((cfg["theta"].size, 7))))
        vals = []
        for c, cols, mt in ((ctx, slice(lo, hi), m), (solo, slice(0, m), None)):
            mdl = g.Model(c, cfg["X"], cfg["Z"], cfg["y"], cfg["family"], cfg["link"])
            cv = g.Covariance(c, cfg["cov"], cfg["data"], cfg["eff_range"])
            mdl.set_u(cfg["U"][:, cols], m_total=mt)
            nr = mdl.mcnr(cfg["beta"], sig)
            vals.append(dict(ll=mdl.log_likelihood(cfg["beta"], sig), llb=mdl.log_likelihood_batch(B, np.full(24, sig)), xtwx=nr["xtwx"], score=nr["score"],
                             sigma=nr["sigma"], incr=nr["beta_incr"], dl=cv.loglik_model(cfg["theta"], mdl), dlb=cv.loglik_model_batch(T, mdl),
                             dlu=cv.loglik(cfg["theta"], cfg["U"][:, cols], m_total=mt)))
            mdl.close(); cv.close()
        a, b = vals
        res[name] = {k: rel(a[k], b[k]) for k in a}
        res[name]["score"] = float(np.max(np.abs(a["score"] - b["score"])) / np.max(np.abs(b["xtwx"])))
    # mcml_full: 3 MCML iterations with chains split over the ranks against all chains on one rank (same Philox streams by chain index)
    cfg = synth.config2(m=8, ncl=8, nt=4, nind=6)
    start = np.concatenate([cfg["beta"] * 0.5, [0.4, 0.5], [1.0]])
    kw = dict(mcnr=True, m=1599, maxiter=3, warmup=60, tol=1e-9, verbose=False, lam=1.0, maxsteps=20, target_accept=0.9, n_chains=64, seed=5)
    ctx.make_default()
    fm = g.mcml_full(cfg["cov"], cfg["data"], cfg["eff_range"], cfg["Z"], cfg["X"], cfg["y"], cfg["family"], cfg["link"], start, **kw)
    dist.barrier()
    solo.make_default()
    fs = g.mcml_full(cfg["cov"], cfg["data"], cfg["eff_range"], cfg["Z"], cfg["X"], cfg["y"], cfg["family"], cfg["link"], start, **kw)
    per = -(-1600 // 64)                       # columns per chain
    mloc = (64 // world) * per
    have = min(1600 - rank * mloc, mloc)
    du = float(np.max(np.abs(fm["u"][:, :have] - fs["u"][:, rank * mloc: rank * mloc + have]))) if rank * mloc < 1600 and have > 0 else 0.0
    t = torch.tensor([du], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res["mcml_full"] = {"beta": float(np.max(np.abs(fm["beta"] - fs["beta"]))), "theta": float(np.max(np.abs(fm["theta"] - fs["theta"]))),
                        "iter": [fm["iter"], fs["iter"]], "u_max_abs": float(t.item())}
    worst = max(max(v for k, v in res[nm].items() if k != "incr") for nm in cases)           # the all-reduced sums themselves
    worst_incr = max(res[nm]["incr"] for nm in cases)                                          # Newton increment: the sums through a P x P solve
    res["worst_sum_rel_err"] = worst; res["worst_incr_rel_err"] = worst_incr
    res["ok"] = bool(worst <= 1e-12 and worst_incr <= 1e-9 and res["mcml_full"]["beta"] <= 1e-9 and res["mcml_full"]["theta"] <= 1e-7   # the theta step is an optimiser with xtol = 1e-8
                     and res["mcml_full"]["u_max_abs"] <= 1e-6 and fm["iter"] == fs["iter"])
    dist.barrier()
    if rank == 0:
        print("NCCL_PARITY " + json.dumps(res), flush=True)
        try:                                                      # keep the record next to the other GPU outputs
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", "nccl_parity.json"), "w") as f:
                json.dump(res, f, indent=1)
        except OSError:
            pass
    dist.destroy_process_group()
    sys.exit(0 if res["ok"] else 1)


if __name__ == "__main__":
    main()
