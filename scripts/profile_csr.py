"""Short CSR-operator run for ncu: Cubic n^3 forced onto the CSR-stream kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import openpnm_b200 as b2

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pn = b2.Cubic([n, n, n], coords=False)
g = 10.0 ** np.random.default_rng(0).uniform(-15, -12, pn.Nt)
bcv = np.full(pn.Np, np.nan)
bcv[pn.pores("left")] = 1.0
bcv[pn.pores("right")] = 0.0
ctx = b2.B200PCG(fmt="csr").ctx
ctx.set_format(1)
ctx.build_laplacian(pn.conns, g, pn.Np)
ctx.apply_bcs(bcv, None)
x, it, rel, info = ctx.pcg(rtol=1e-10, maxiter=24)
print("iters", it, "info", info)
