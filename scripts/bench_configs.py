"""Timings of the BASELINE.json configurations other than the headline one
(recorded under profiles/; bench.py stays on Cubic 400^3 StokesFlow)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch

import openpnm_b200 as b2
from openpnm_b200 import dist as bdist
from _gen import delaunay_conns


def hetero(nt, seed=0):
    return 10.0 ** np.random.default_rng(seed).uniform(-15, -12, nt)


def timed(label, fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, time.perf_counter() - t0


PRECOND = sys.argv[1] if len(sys.argv) > 1 else "jacobi"
print(f"== preconditioner: {PRECOND}", flush=True)
solver = b2.B200PCG(tol=1e-10, precond=PRECOND)

# config 1: Cubic 50^3 StokesFlow
pn = b2.Cubic([50, 50, 50], spacing=1e-4, coords=False)
ph = b2.Phase(pn)
ph["throat.hydraulic_conductance"] = hetero(pn.Nt)
sf = b2.StokesFlow(network=pn, phase=ph)
sf.set_value_BC(pores=pn.pores("left"), values=2e5)
sf.set_value_BC(pores=pn.pores("right"), values=1e5)
sf.run(solver=solver)
_, dt = timed("c1", lambda: sf.run(solver=solver))
print(f"config1 Cubic 50^3 StokesFlow run(): {dt*1e3:.1f} ms end to end ({sf.soln.cg_iters} PCG its, "
      f"solve {sf.stats['t_solve_ms']:.1f} ms); reference ScipySpsolve 163 s, Jacobi-CG ~1 s", flush=True)

# config 2: Cubic 200^3 FickianDiffusion
pn = b2.Cubic([200, 200, 200], spacing=1e-4, coords=False)
ph = b2.Phase(pn)
ph["throat.diffusive_conductance"] = hetero(pn.Nt)
fd = b2.FickianDiffusion(network=pn, phase=ph)
fd.set_value_BC(pores=pn.pores("left"), values=1.0)
fd.set_value_BC(pores=pn.pores("right"), values=0.0)
fd.run(solver=solver)
_, dt = timed("c2", lambda: fd.run(solver=solver))
print(f"config2 Cubic 200^3 FickianDiffusion run(): {dt*1e3:.1f} ms end to end incl. topology check "
      f"({fd.soln.cg_iters} its, solve {fd.stats['t_solve_ms']:.1f} ms, "
      f"{pn.Np*fd.soln.cg_iters/(fd.stats['t_solve_ms']*1e-3):.3e} DOF*iter/s)", flush=True)

# config 5: Cubic 200^3 ReactiveTransport, power-law sink on every free pore
ph["pore.A1"], ph["pore.A2"], ph["pore.A3"] = -1e-13, 2.0, 0.0
ph.add_model(propname="pore.rxn", model="power_law", X="pore.concentration", A1="pore.A1",
             A2="pore.A2", A3="pore.A3")
rt = b2.ReactiveTransport(network=pn, phase=ph)
rt.settings._update({"quantity": "pore.concentration",
                     "conductance": "throat.diffusive_conductance"})
left = pn.pores("left")
rt.set_value_BC(pores=left, values=1.0)
rt.set_source(pores=np.setdiff1d(pn.Ps, left), propname="pore.rxn")
_, dt = timed("c5", lambda: rt.run(solver=solver))
print(f"config5 Cubic 200^3 ReactiveTransport (power law) run(): {dt:.2f} s, {rt.soln.num_iter} "
      f"outer iterations, {rt.soln.cg_iters} PCG its in total, converged={rt.soln.is_converged}",
      flush=True)

# config 4: Delaunay, 1M points, trimmed + sorted by x (CSR operator)
t0 = time.perf_counter()
npts = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
pts = np.random.default_rng(0).random((npts, 3))
c = delaunay_conns(pts)
L = np.linalg.norm(pts[c[:, 0]] - pts[c[:, 1]], axis=1)
c = c[L < 4.0 * npts ** (-1 / 3)]
perm, inv = bdist.coordinate_order(pts, axis=0)
pts, conns = pts[perm], np.sort(inv[c], axis=1)
t_gen = time.perf_counter() - t0
pn = b2.Network(coords=pts, conns=conns)
ph = b2.Phase(pn)
ph["throat.diffusive_conductance"] = hetero(pn.Nt, 2)
fd = b2.FickianDiffusion(network=pn, phase=ph)
fd.set_value_BC(pores=np.flatnonzero(pts[:, 0] < 0.01), values=1.0)
fd.set_value_BC(pores=np.flatnonzero(pts[:, 0] > 0.99), values=0.0)
fd.run(solver=solver)
_, dt = timed("c4", lambda: fd.run(solver=solver))
nnz = fd.stats["nnz_system"]
print(f"config4 Delaunay {pn.Np} pores / {pn.Nt} throats (host triangulation {t_gen:.0f} s): run() "
      f"{dt*1e3:.1f} ms, {fd.soln.cg_iters} its, solve {fd.stats['t_solve_ms']:.1f} ms = "
      f"{fd.stats['t_solve_ms']/fd.soln.cg_iters*1e3:.1f} us/iter, "
      f"{(12*nnz+16*pn.Np+88*pn.Np)*fd.soln.cg_iters/(fd.stats['t_solve_ms']*1e-3)/1e9:.0f} GB/s model, "
      f"band {bdist.bandwidth(conns)}", flush=True)
