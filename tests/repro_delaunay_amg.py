"""TEST INFRASTRUCTURE (lives under tests/ because it uses the oracle's Delaunay helper).
Minimal repro: slab multigrid on the small x-sorted Delaunay case of tests/dist_worker.py,
in a fresh process with the library's debug trace on.
    torchrun --nproc-per-node 4 tests/repro_delaunay_amg.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def main():
    import torch
    import torch.distributed as dist
    import openpnm_b200 as b2
    from openpnm_b200 import dist as bdist, topotools
    from oracle import oracle
    rank, world, local = bdist.init_process_group()
    pts = np.random.default_rng(7).random((12000, 3))
    pts = pts[bdist.coordinate_order(pts)[0]]
    conns = oracle.delaunay_network(pts)
    L = np.linalg.norm(pts[conns[:, 0]] - pts[conns[:, 1]], axis=1)
    pn3 = b2.Network(coords=pts, conns=conns[L < 0.06])
    hl = topotools.check_network_health(pn3, ctx=b2.B200PCG().ctx)
    topotools.trim(pn3, pores=hl["disconnected_pores"], ctx=b2.B200PCG().ctx)
    g3 = 10.0 ** np.random.default_rng(8).uniform(-15, -12, pn3.Nt)
    ph3 = b2.Phase(pn3)
    ph3["throat.diffusive_conductance"] = g3
    bcv = np.full(pn3.Np, np.nan)
    bcv[pn3["pore.coords"][:, 0] < 0.05] = 1.0
    bcv[pn3["pore.coords"][:, 0] > 0.95] = 0.0
    xs = {}
    for precond in ("jacobi", "amg"):
        if precond == "amg":
            os.environ["B200_DEBUG"] = "1"
        fd = b2.FickianDiffusion(network=pn3, phase=ph3)
        fd["pore.bc.value"] = bcv
        fd.settings["validate_topology"] = False
        fd.run(solver=b2.B200PCG(tol=1e-13, precond=precond, maxiter=400))
        xs[precond] = np.array(fd.x)
        if rank == 0:
            print(precond, "iters", fd.soln.cg_iters, "nan", int(np.isnan(fd.x).sum()),
                  "first nan row", int(np.argmax(np.isnan(fd.x))) if np.isnan(fd.x).any() else -1,
                  flush=True)
    if rank == 0:
        print("max diff", float(np.nanmax(np.abs(xs["amg"] - xs["jacobi"]))), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
