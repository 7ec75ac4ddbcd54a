"""Assembly-only run (K1 + K2 + operator setup) for a per-kernel ncu launch list."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import openpnm_b200 as b2

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
pn = b2.Cubic([n, n, n], coords=False)
g = 10.0 ** np.random.default_rng(0).uniform(-15, -12, pn.Nt)
bcv = np.full(pn.Np, np.nan)
bcv[pn.pores("left")] = 2e5
bcv[pn.pores("right")] = 1e5
ctx = b2.B200PCG().ctx
conns_d, g_d, bcv_d = ctx.to_device(pn.conns), ctx.to_device(g), ctx.to_device(bcv)
for rep in range(2):
    ctx.build_laplacian(conns_d, g_d, pn.Np)
    ctx.apply_bcs(bcv_d, None)
    ctx.pcg_prepare(0)
nc = ctx.check_topology(conns_d, pn.Np, bcv_d, None)
print("done", nc)
