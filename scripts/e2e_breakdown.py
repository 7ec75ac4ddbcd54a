"""Where the host-side time of StokesFlow.run() goes at Cubic n^3 (pinned host inputs)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import openpnm_b200 as b2
from openpnm_b200.device import pinned_copy

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
pn = b2.Cubic([n, n, n], coords=False)
g = pinned_copy(10.0 ** np.random.default_rng(0).uniform(-15, -12, pn.Nt))
conns = pinned_copy(pn.conns)
bcv = np.full(pn.Np, np.nan)
bcv[pn.pores("left")] = 2e5
bcv[pn.pores("right")] = 1e5
bcv = pinned_copy(bcv)
bcr = pinned_copy(np.full(pn.Np, np.nan))
s = b2.B200PCG(tol=1e-10)
ctx = s.ctx


def T(label, fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    print(f"{label:28s} {1e3*(time.perf_counter()-t0):9.1f} ms", flush=True)
    return out


t = torch.from_numpy(conns)
print("from_numpy(pinned).is_pinned():", t.is_pinned())
for rep in range(2):
    print("--- rep", rep)
    cd = T("H2D conns (3.06 GB)", lambda: ctx.to_device(conns))
    gd = T("H2D g", lambda: ctx.to_device(g))
    bd = T("H2D bc_value", lambda: ctx.to_device(bcv))
    rd = T("H2D bc_rate", lambda: ctx.to_device(bcr))
    T("build_laplacian", lambda: ctx.build_laplacian(cd, gd, pn.Np))
    T("apply_bcs", lambda: ctx.apply_bcs(bd, rd))
    x, ni, conv, cgi, info = T("picard (PCG)", lambda: ctx.picard(cg_rtol=1e-10))
    stage = torch.empty(pn.Np, dtype=torch.float64, pin_memory=True) if rep == 0 else stage
    T("D2H x -> pinned", lambda: stage.copy_(x))
    xh = T("host copy of x", lambda: stage.numpy().copy())
    T("rate_pores(left)", lambda: ctx.rate_pores(x, pn.pores("left")).sum().item())
ph = b2.Phase(pn)
dict.__setitem__(ph, "throat.hydraulic_conductance", g)
dict.__setitem__(pn, "throat.conns", conns)
sf = b2.StokesFlow(network=pn, phase=ph)
dict.__setitem__(sf, "pore.bc.value", bcv)
dict.__setitem__(sf, "pore.bc.rate", bcr)
sf.settings["validate_topology"] = False
sf.settings["cache"] = False
for rep in range(2):
    T("StokesFlow.run total", lambda: sf.run(solver=s))
    print("   cg_iters", sf.soln.cg_iters, "solve ms", sf.stats["t_solve_ms"])
