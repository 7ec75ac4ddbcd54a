"""K_A tuning sweep: time the SpMV+dots kernel on a resident Cubic n^3 operator."""
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
for chunk in [int(c) for c in (sys.argv[2:] or ["1", "2", "4", "8", "16"])]:
    os.environ["B200_KA_CHUNK"] = str(chunk)
    s = b2.B200PCG(tol=1e-10)
    ctx = s.ctx
    ctx.build_laplacian(pn.conns, g, pn.Np)
    ctx.apply_bcs(bcv, None)
    x, it, rel, info = ctx.pcg(rtol=1e-10, maxiter=40)
    t = [ctx.time_kernel(0, 30) for _ in range(3)]
    tb = ctx.time_kernel(1, 30)
    print(f"n={n} chunk={chunk}: K_A {min(t)*1e3:.1f} us ({48*pn.Np/min(t)/1e6:.0f} GB/s format, "
          f"{(12*ctx.csr_shape(1)[1]+20*pn.Np)/min(t)/1e6:.0f} GB/s model)  K_BC {tb*1e3:.1f} us",
          flush=True)
    ctx.close()
    del s, ctx
