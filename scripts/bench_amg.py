"""Time-to-solve of Jacobi-PCG vs AMG-FCG on Cubic N^3 (one GPU)."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
import openpnm_b200 as b2
import torch

N = int(sys.argv[1]) if len(sys.argv) > 1 else 200
which = sys.argv[2].split(",") if len(sys.argv) > 2 else ["amg", "jacobi"]
pn = b2.Cubic([N, N, N], coords=False)
g = 10.0 ** np.random.default_rng(0).uniform(-15, -12, pn.Nt)
bcv = np.full(pn.Np, np.nan)
bcv[pn.pores("left")] = 2e5
bcv[pn.pores("right")] = 1e5
for precond in which:
    ctx = b2.B200PCG(precond=precond).ctx
    conns = ctx.to_device(pn.conns, torch.int64)
    gd = ctx.to_device(g, torch.float64)
    ctx.build_laplacian(conns, gd, pn.Np)
    ctx.apply_bcs(bcv, None)
    for rep in range(2):
        ctx.apply_bcs(bcv, None)           # new epoch: the hierarchy is rebuilt, as in a fresh run
        ctx.synchronize()
        t0 = time.perf_counter()
        x, it, rel, info = ctx.pcg(rtol=1e-10)
        ctx.synchronize()
        wall = time.perf_counter() - t0
        st = ctx.stats()
        print(json.dumps(dict(N=N, precond=precond, rep=rep, iters=it, relres=rel, info=info,
                              wall_s=round(wall, 4), t_setup_ms=round(st["t_setup_ms"], 2),
                              t_solve_ms=round(st["t_solve_ms"], 2))), flush=True)
    if precond == "amg":
        print("levels", ctx.amg_setup(), flush=True)
    del ctx, x
    torch.cuda.empty_cache()
