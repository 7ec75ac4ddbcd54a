"""CSR-stream K_A timing: cubic lattice forced to CSR, and a Delaunay network."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import openpnm_b200 as b2
from openpnm_b200 import dist as bdist
from _gen import delaunay_conns


def delaunay(npts):
    pts = np.random.default_rng(0).random((npts, 3))
    c = delaunay_conns(pts)
    L = np.linalg.norm(pts[c[:, 0]] - pts[c[:, 1]], axis=1)
    c = c[L < 4.0 * npts ** (-1 / 3)]
    perm, inv = bdist.coordinate_order(pts, axis=0)
    return pts[perm], np.sort(inv[c], axis=1)


def run(name, conns, Np, bcv, block):
    os.environ["B200_CSR_BLOCK"] = "1" if block else "0"
    g = 10.0 ** np.random.default_rng(0).uniform(-15, -12, conns.shape[0])
    s = b2.B200PCG(tol=1e-10, fmt="csr")
    ctx = s.ctx
    ctx.set_format(1)
    ctx.build_laplacian(conns, g, Np)
    ctx.apply_bcs(bcv, None)
    x, it, rel, info = ctx.pcg(rtol=1e-10)
    st = ctx.stats()
    nnz = ctx.csr_shape(1)[1]
    t = min(ctx.time_kernel(0, 30) for _ in range(3))
    fmt_bytes = 12 * (nnz - Np) + 4 * Np + 24 * Np
    model = 12 * nnz + 4 * Np + 16 * Np
    print(f"{name} block={block}: Np={Np} nnz={nnz} iters={it} info={info} ms/iter="
          f"{st['t_solve_ms']/max(it,1):.4f} K_A {t*1e3:.1f} us format {fmt_bytes/t/1e6:.0f} GB/s "
          f"model {model/t/1e6:.0f} GB/s", flush=True)
    ctx.close()


if __name__ == "__main__":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    pn = b2.Cubic([n, n, n], coords=False)
    bcv = np.full(pn.Np, np.nan)
    bcv[pn.pores("left")] = 1.0
    bcv[pn.pores("right")] = 0.0
    run(f"cubic{n}", pn.conns, pn.Np, bcv, True)
    t0 = time.time()
    pts, conns = delaunay(int(sys.argv[1]) if len(sys.argv) > 1 else 400000)
    print(f"delaunay built in {time.time()-t0:.1f}s, band {bdist.bandwidth(conns)}", flush=True)
    bcv = np.full(len(pts), np.nan)
    bcv[pts[:, 0] < 0.01] = 1.0
    bcv[pts[:, 0] > 0.99] = 0.0
    run("delaunay", conns, len(pts), bcv, True)
