"""Host-side stage times of StokesFlow.run()+rate() at Cubic n^3 (B200_PROFILE=1 stage
timer of openpnm_b200.algorithms), on 1 GPU or under torchrun.
    B200_PROFILE=1 python scripts/e2e_stages.py 400 [amg|jacobi]"""
import os
import sys
import time

os.environ["B200_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import openpnm_b200 as b2
from openpnm_b200 import dist as bdist
from openpnm_b200.device import pinned_copy

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
precond = sys.argv[2] if len(sys.argv) > 2 else "amg"
rank, world, local = bdist.init_process_group()
Np, Nt = n ** 3, 3 * n * n * (n - 1)
pn = b2.Cubic([n, n, n], coords=False, lazy=world > 1)
g = pinned_copy(10.0 ** np.random.default_rng(0).uniform(-15, -12, Nt))
if world == 1:
    dict.__setitem__(pn, "throat.conns", pinned_copy(pn.conns))
bcv = np.full(Np, np.nan)
bcv[:n * n] = 2e5
bcv[Np - n * n:] = 1e5
ph = b2.Phase(pn)
dict.__setitem__(ph, "throat.hydraulic_conductance", g)
sf = b2.StokesFlow(network=pn, phase=ph)
dict.__setitem__(sf, "pore.bc.value", pinned_copy(bcv))
dict.__setitem__(sf, "pore.bc.rate", pinned_copy(np.full(Np, np.nan)))
sf.settings["cache"] = False
solver = b2.B200PCG(tol=1e-10, precond=precond, device=local)
left = np.arange(n * n)
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sf.run(solver=solver)
    t1 = time.perf_counter()
    q = sf.rate(pores=left)[0]
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    if rank == 0:
        st = {k: round(v, 1) for k, v in sf.stats.get("host_stage_ms", {}).items()}
        print(f"rep {rep} world {world} {precond}: run {1e3*(t1-t0):.1f} ms, rate {1e3*(t2-t1):.1f} ms, "
              f"device setup {sf.stats['t_setup_ms']:.1f} + solve {sf.stats['t_solve_ms']:.1f}; "
              f"stages {st}", flush=True)
