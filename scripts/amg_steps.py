"""Per-step breakdown of the multigrid step at Cubic n^3 on one GPU: assembly, hierarchy
setup, solve (CUDA events inside the library) and the wall time of the whole step."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import openpnm_b200 as b2
from openpnm_b200 import dist as bdist

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
Np, Nt = n ** 3, 3 * n * n * (n - 1)
pn = b2.Cubic([n, n, n], coords=False)
g = 10.0 ** np.random.default_rng(0).uniform(-15, -12, Nt)
bcv = np.full(Np, np.nan)
bcv[:n * n] = 2e5
bcv[Np - n * n:] = 1e5
ss = bdist.SlabSolver(Np, 0, Np, device=0)
ctx = ss.ctx
cd, gd, bd = ctx.to_device(pn.conns), ctx.to_device(g), ctx.to_device(bcv)
ctx.set_precond("amg")
for rep in range(reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ss.assemble(cd, gd, bd, None)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    x, it, rel, info = ss.solve(rtol=1e-10)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    st = ctx.stats()
    print(f"rep {rep}: assemble {1e3*(t1-t0):7.1f} ms | solve call {1e3*(t2-t1):7.1f} ms "
          f"(setup {st['t_setup_ms']:.1f} + solve {st['t_solve_ms']:.1f}, {it} it) | "
          f"launches {st['kernel_launches']}", flush=True)
