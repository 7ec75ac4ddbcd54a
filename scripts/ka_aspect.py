"""K_A (DIA SpMV + dots) and K_BC at 64 M pores for lattices of different aspect ratios: the
+-k_y / +-k_x neighbour reads of K_A are served by L1 / L2, not by shared-memory staging
(DESIGN section 4), so their reuse distance -- 2 k 48 B -- matters once it exceeds the caches.
    python scripts/ka_aspect.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from openpnm_b200 import dist as bdist

PEAK = 6535.4
shapes = [(400, 400, 400), (1600, 200, 200), (100, 800, 800), (25, 1600, 1600), (64000, 100, 10),
          (10, 100, 64000)]
for shape in shapes:
    nx, ny, nz = shape
    Np = nx * ny * nz
    ids, conns = bdist.cubic_slab_throats(shape, 0, nx)
    g = 10.0 ** np.random.default_rng(0).uniform(-15, -12, ids.size)
    bcv = np.full(Np, np.nan)
    bcv[:ny * nz] = 2e5
    bcv[Np - ny * nz:] = 1e5
    ss = bdist.SlabSolver(Np, 0, Np, device=0)
    ss.assemble(conns, g, bcv, None)
    del conns, g, ids
    ss.ctx.pcg(rtol=1e-1, maxiter=8)
    fmt, nd = ss.ctx.pcg_format()
    ta, tb = ss.ctx.time_kernel(0, 30), ss.ctx.time_kernel(1, 30)
    ob, _ = ss.ctx.operator_bytes()
    print(json.dumps({"shape": shape, "fmt": fmt, "nd": nd, "k_y": nz, "k_x": ny * nz,
                      "reuse_distance_MB_x": round(2 * ny * nz * 48 / 1e6, 1),
                      "K_A_ms": round(ta, 4), "K_A_frac_format_bytes": round(ob / ta / 1e6 / PEAK, 3),
                      "K_BC_ms": round(tb, 4),
                      "K_BC_frac": round(56 * Np / tb / 1e6 / PEAK, 3)}), flush=True)
    ss.ctx.close()
    torch.cuda.empty_cache()
