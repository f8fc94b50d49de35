import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import numpy as np
    import glmmrmcml_b200 as g
    from glmmrmcml_b200 import synth
    ctx = g.Context(0); cfg = synth.config2(m=64)
    mdl = g.Model(ctx, cfg["X"], cfg["Z"], cfg["y"], "binomial", "logit")
    for rep in range(2):
        out = mdl.hmc_sample(cfg["L"], cfg["beta"], 1.0, warmup=25, nsamp_per_chain=5, lam=5.0, max_steps=100, target_accept=0.95,
                             n_chains=1184, seed=3, keep_on_device=True, want_u=False)
    st = out["stats"]
    print(f"tune {os.environ.get('GMB_FUSED_TUNE')} kernel_ms {st['kernel_ms']:.2f}  us/leapfrog-step {st['kernel_ms']*1e3/(30*st['steps_mean']):.2f}")
else:
    for t in (0, 1, 2, 3):
        env = dict(os.environ, GMB_FUSED_TUNE=str(t))
        subprocess.run([sys.executable, __file__, "x"], env=env)
