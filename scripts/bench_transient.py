"""Throughput of the device transient path: Cubic n^3 TransientFickianDiffusion."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import openpnm_b200 as b2

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
pn = b2.Cubic([n, n, n], spacing=1e-4, coords=False)
rng = np.random.default_rng(0)
pn["pore.volume"] = rng.uniform(0.5, 1.5, pn.Np) * 1e-12
ph = b2.Phase(pn)
ph["throat.diffusive_conductance"] = 10.0 ** rng.uniform(-15, -12, pn.Nt)
alg = b2.TransientFickianDiffusion(network=pn, phase=ph)
alg.set_value_BC(pores=pn.pores("left"), values=1.0)
alg.set_value_BC(pores=pn.pores("right"), values=0.0)
alg.settings["validate_topology"] = False
integ = b2.B200RK45(atol=1e-6, rtol=1e-6)
for tend in (200.0, 1500.0):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    alg.run(x0=0.0, tspan=(0, tend), saveat=np.array([tend]), integrator=integ)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ns, nf = alg.soln.nsteps, alg.soln.nfev
    print(f"n={n} tend={tend}: {ns} steps, {nf} rhs evals, {dt:.3f} s wall, "
          f"{1e3*dt/max(nf,1):.3f} ms/rhs, {pn.Np*nf/dt:.3e} DOF*rhs/s, "
          f"{112*pn.Np*nf/dt/1e9:.0f} GB/s (112 B/row/rhs model), mean c {alg.x.mean():.6f}", flush=True)
