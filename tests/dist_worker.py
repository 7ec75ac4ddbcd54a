"""Worker for the multi-GPU slab tests (launched by torchrun, one rank per GPU).

Solves a Cubic network split into x-slabs and checks, on every rank, the
gathered solution against the CPU oracle: x within 1e-8, rate conservation,
and the local CSR rows bit-exact against the oracle's global matrix."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from openpnm_b200 import dist as bdist
    from oracle import oracle

    rank, world, local = bdist.init_process_group()
    # (8,5,6) on 8 ranks and the 'z' cases: one-plane slabs, some of which have x-couplings
    # only downwards (no mirrored upper entry among the owned rows)
    cases = [((8, 5, 6), "x"), ((13, 4, 7), "x"), ((24, 24, 24), "x"), ((2, 5, 6), "z"),
             ((8, 3, 5), "z")]
    for shape, bcdir in cases:
        if shape[0] < world:
            continue
        for fmt in ("auto", "csr"):
            nx, ny, nz = shape
            net = oracle.cubic_network(shape)
            Np, Nt = net["Np"], net["Nt"]
            g_all = 10.0 ** np.random.default_rng(3).uniform(-15, -12, Nt)
            bcv = np.full(Np, np.nan)
            lo_lbl, hi_lbl, side = ("left", "right", "top") if bcdir == "x" else \
                                   ("top", "bottom", "front")
            bcv[net["labels"][lo_lbl]] = 2.0
            bcv[net["labels"][hi_lbl]] = 1.0
            bcr = np.full(Np, np.nan)
            bcr[net["labels"][side][5:9]] = 1e-13         # a few rate BCs too
            bcr[np.isfinite(bcv)] = np.nan
            rb, re = bdist.cubic_slab_ranges(shape, world)[rank]
            ids, conns = bdist.cubic_slab_throats(shape, rb // (ny * nz), re // (ny * nz))
            # the analytic slab throat list must equal the mask over the full list
            mask = bdist.slab_throat_mask(net["conns"], rb, re)
            assert np.array_equal(np.sort(ids), np.flatnonzero(mask))
            assert np.array_equal(conns, net["conns"][ids])
            ss = bdist.SlabSolver(Np, rb, re, device=local, fmt=fmt)
            f, nnz0, nnz = ss.assemble(conns, g_all[ids], bcv, bcr)
            A, b, f_ref = oracle.apply_bcs(oracle.build_A(net["conns"], g_all, Np), Np, bcv, bcr)
            M = oracle.to_csr(A, Np)
            assert abs(f - f_ref) <= 1e-12 * abs(f_ref), (f, f_ref)
            L = ss.ctx.export_csr_scipy(1, ncols=Np)
            Mloc = M[rb:re]
            assert np.array_equal(L.indptr, Mloc.indptr), "slab indptr differs"
            assert np.array_equal(L.indices, Mloc.indices), "slab indices differ"
            np.testing.assert_allclose(L.data, Mloc.data, rtol=1e-12, atol=0)
            xr, _ = oracle.spsolve(M, b)
            x_loc, it, rel, info = ss.solve(rtol=1e-13)
            assert info == 0, (info, it, rel)
            x = ss.gather(x_loc).cpu().numpy()
            err = np.max(np.abs(x - xr)) / np.max(np.abs(xr))
            assert err < 1e-8, (shape, fmt, err)
            # same solve with the collectives fused into the kernels over peer memory
            p2p = ss.enable_p2p() if world > 1 else False
            if p2p:
                x2_loc, it2, rel2, info2 = ss.solve(rtol=1e-13)
                assert info2 == 0, (info2, it2, rel2)
                x2 = ss.gather(x2_loc).cpu().numpy()
                err2 = np.max(np.abs(x2 - xr)) / np.max(np.abs(xr))
                assert err2 < 1e-8, (shape, fmt, "p2p", err2)
                x3_loc, it3, _, _ = ss.solve(rtol=1e-13)          # repeatable bit for bit
                assert it3 == it2 and bool((x3_loc == x2_loc).all())
                it = (it, it2)
            xg = torch.from_numpy(x).to(ss.ctx.tdev)
            q = torch.tensor([ss.rate(xg, net["labels"][lo_lbl]),
                              ss.rate(xg, net["labels"][hi_lbl])], dtype=torch.float64,
                             device=ss.ctx.tdev)
            if world > 1:
                dist.all_reduce(q)
            qref = oracle.rate(net["conns"], g_all, xr, pores=net["labels"][lo_lbl])[0]
            assert abs(float(q[0]) - qref) <= 1e-8 * abs(qref)
            if rank == 0:
                print(f"dist ok shape={shape} fmt={fmt} world={world} iters={it} err={err:.2e} "
                      f"p2p={p2p} fmtcode={ss.ctx.pcg_format()}", flush=True)
            ss.ctx.close()
    # ---- the algorithm API under torchrun: the same script, slab-parallel run()
    import openpnm_b200 as b2
    shape = (16, 9, 10)
    pn = b2.Cubic(shape, spacing=1e-4)
    ph = b2.Phase(pn)
    g = 10.0 ** np.random.default_rng(5).uniform(-15, -12, pn.Nt)
    ph["throat.hydraulic_conductance"] = g
    sf = b2.StokesFlow(network=pn, phase=ph)
    sf.set_value_BC(pores=pn.pores("left"), values=2e5)
    sf.set_value_BC(pores=pn.pores("right"), values=1e5)
    sf.run(solver=b2.B200PCG(tol=1e-13))
    ref = oracle.transport_run(pn.conns, g, pn.Np, bc_value=sf["pore.bc.value"])
    err = np.max(np.abs(sf.x - ref["x"]) / np.abs(ref["x"]))
    assert sf.soln.is_converged and err < 1e-8, err
    q = sf.rate(pores=pn.pores("left"))[0]
    qref = oracle.rate(pn.conns, g, ref["x"], pores=pn.pores("left"))[0]
    assert abs(q - qref) <= 1e-8 * abs(qref)
    qs = sf.rate(pores=pn.pores("right"), mode="single")
    assert qs.size == pn.pores("right").size and abs(qs.sum() + q) <= 1e-8 * abs(q)
    qt = sf.rate(throats=[0, 5, 17], mode="single")
    np.testing.assert_allclose(qt, oracle.rate(pn.conns, g, ref["x"], throats=[0, 5, 17],
                                               mode="single"), rtol=1e-8)
    if rank == 0:
        print(f"dist ok algorithm-API world={world} err={err:.2e} cg_iters={sf.soln.cg_iters} "
              f"p2p={getattr(sf, '_slab', None) is not None}", flush=True)
    # nonlinear source, Picard loop slab-parallel
    ph["throat.diffusive_conductance"] = g
    ph["pore.A1"], ph["pore.A2"], ph["pore.A3"] = -1e-13, 2.0, 0.0
    ph.add_model(propname="pore.rxn", model="power_law", X="pore.concentration",
                 A1="pore.A1", A2="pore.A2", A3="pore.A3")
    rt = b2.ReactiveTransport(network=pn, phase=ph)
    rt.settings._update({"quantity": "pore.concentration",
                         "conductance": "throat.diffusive_conductance"})
    left = pn.pores("left")
    rt.set_value_BC(pores=left, values=1.0)
    src = np.setdiff1d(pn.Ps, left)
    rt.set_source(pores=src, propname="pore.rxn")
    rt.run(solver=b2.B200PCG(tol=1e-12))
    smask = np.zeros(pn.Np, dtype=bool)
    smask[src] = True
    refp = oracle.picard_run(pn.conns, g, pn.Np, bc_value=rt["pore.bc.value"],
                             sources=[(smask, "power_law", {"A1": -1e-13, "A2": 2.0, "A3": 0.0})])
    assert rt.soln.is_converged and rt.soln.num_iter == refp["num_iter"], \
        (rt.soln.num_iter, refp["num_iter"])
    errp = np.max(np.abs(rt.x - refp["x"])) / np.max(np.abs(refp["x"]))
    assert errp < 1e-8, errp
    if rank == 0:
        print(f"dist ok reactive world={world} num_iter={rt.soln.num_iter} err={errp:.2e}",
              flush=True)
    # ---- multigrid through the algorithm API (one hierarchy across the slabs)
    shape = (32, 12, 12)
    pn = b2.Cubic(shape, spacing=1e-4)
    ph = b2.Phase(pn)
    g = 10.0 ** np.random.default_rng(6).uniform(-15, -12, pn.Nt)
    ph["throat.hydraulic_conductance"] = g
    sf = b2.StokesFlow(network=pn, phase=ph)
    sf.set_value_BC(pores=pn.pores("left"), values=2e5)
    sf.set_value_BC(pores=pn.pores("right"), values=1e5)
    sf.run(solver=b2.B200PCG(tol=1e-13))
    it_jacobi = sf.soln.cg_iters
    sf.run(solver=b2.B200PCG(tol=1e-13, precond="amg"))
    ref = oracle.transport_run(pn.conns, g, pn.Np, bc_value=sf["pore.bc.value"])
    err = np.max(np.abs(sf.x - ref["x"]) / np.abs(ref["x"]))
    assert sf.soln.is_converged and err < 1e-8, err
    assert sf.soln.cg_iters < it_jacobi / 3, (sf.soln.cg_iters, it_jacobi)
    if rank == 0:
        print(f"dist ok algorithm-API multigrid world={world} err={err:.2e} "
              f"iters={sf.soln.cg_iters} (jacobi {it_jacobi})", flush=True)
    # ---- topology check of a slab run: a block cut off from every BC pore is refused
    cut = ~((pn.conns[:, 0] // 144 == 19) & (pn.conns[:, 1] // 144 == 20))
    pn2 = b2.Network(conns=pn.conns[cut], Np=pn.Np)
    ph2 = b2.Phase(pn2)
    ph2["throat.hydraulic_conductance"] = g[cut]
    sf2 = b2.StokesFlow(network=pn2, phase=ph2)
    sf2["pore.bc.value"] = np.where(np.arange(pn.Np) < 144, 2e5, np.nan)
    try:
        sf2.run(solver=b2.B200PCG(tol=1e-10))
        raise AssertionError("a clustered network was accepted")
    except Exception as e:
        assert "clustered" in str(e), e
    sf2["pore.bc.value"][pn.Np - 1] = 1e5          # one BC pore in the far block: healthy
    sf2.run(solver=b2.B200PCG(tol=1e-10))
    assert sf2.soln.is_converged
    if rank == 0:
        print(f"dist ok slab topology check world={world}", flush=True)
    # ---- irregular network: Delaunay, pores ordered by x, over-long throats cut, isolated
    # pores and minor clusters trimmed on the device, then the slab-parallel solve
    from openpnm_b200 import topotools
    pts = np.random.default_rng(7).random((12000, 3))
    pts = pts[bdist.coordinate_order(pts)[0]]
    conns = oracle.delaunay_network(pts)
    L = np.linalg.norm(pts[conns[:, 0]] - pts[conns[:, 1]], axis=1)
    pn3 = b2.Network(coords=pts, conns=conns[L < 0.06])
    hl = topotools.check_network_health(pn3, ctx=b2.B200PCG().ctx)
    assert len(hl["disconnected_pores"]) > 0
    topotools.trim(pn3, pores=hl["disconnected_pores"], ctx=b2.B200PCG().ctx)
    g3 = 10.0 ** np.random.default_rng(8).uniform(-15, -12, pn3.Nt)
    ph3 = b2.Phase(pn3)
    ph3["throat.diffusive_conductance"] = g3
    bcv = np.full(pn3.Np, np.nan)
    bcv[pn3["pore.coords"][:, 0] < 0.05] = 1.0
    bcv[pn3["pore.coords"][:, 0] > 0.95] = 0.0
    ref = oracle.transport_run(pn3.conns, g3, pn3.Np, bc_value=bcv)
    its = {}
    for precond in ("jacobi", "amg"):
        fd = b2.FickianDiffusion(network=pn3, phase=ph3)
        fd["pore.bc.value"] = bcv
        fd.run(solver=b2.B200PCG(tol=1e-13, precond=precond))
        err = np.max(np.abs(fd.x - ref["x"]))
        assert fd.soln.is_converged and err < 1e-8, (precond, err)
        qi = fd.rate(pores=np.flatnonzero(bcv == 1.0))[0]
        qo = fd.rate(pores=np.flatnonzero(bcv == 0.0))[0]
        assert abs(qi + qo) <= 1e-8 * abs(qi)
        its[precond] = fd.soln.cg_iters
    assert its["amg"] < its["jacobi"] / 2, its
    if rank == 0:
        print(f"dist ok delaunay Np={pn3.Np} Nt={pn3.Nt} band={bdist.bandwidth(pn3.conns)} "
              f"world={world} iters={its} fmt={fd._ctx.pcg_format()}", flush=True)
    # ---- AdvectionDiffusion, slab-parallel: (Nt,2) upwind conductance from a pressure field,
    # outflow BC, non-symmetric rows per slab, BiCGStab with halo exchange + all-reduced dots
    pna = b2.Cubic((14, 9, 8), spacing=1e-4, coords=False)
    rng = np.random.default_rng(0)
    gh = 10.0 ** rng.uniform(-15, -14, pna.Nt)
    gd = 10.0 ** rng.uniform(-15.5, -14.5, pna.Nt)
    pha = b2.Phase(pna)
    pha["throat.hydraulic_conductance"] = gh
    sfa = b2.StokesFlow(network=pna, phase=pha)
    sfa.set_value_BC(pores=pna.pores("right"), values=1.0)
    sfa.set_value_BC(pores=pna.pores("left"), values=0.0)
    sfa.run(solver=b2.B200PCG(tol=1e-13))
    P = sfa.x
    pha["pore.pressure"] = P
    ca = pna.conns
    Q = -gh * (P[ca[:, 1]] - P[ca[:, 0]])
    g2 = np.column_stack((np.maximum(-Q, 0) + gd, np.maximum(Q, 0) + gd))
    dict.__setitem__(pha, "throat.ad_dif_conductance", g2)
    ad = b2.AdvectionDiffusion(network=pna, phase=pha)
    ad.set_value_BC(pores=pna.pores("right"), values=2.0)
    ad.set_outflow_BC(pores=pna.pores("left"))
    taken = np.union1d(pna.pores("right"), pna.pores("left"))
    ad.set_value_BC(pores=np.setdiff1d(pna.pores("top"), taken)[::3], values=0.3)
    ad.run(solver=b2.B200PCG(tol=1e-13))
    refa = oracle.picard_run(ca, g2, pna.Np, bc_value=ad["pore.bc.value"],
                             bc_outflow=ad["pore.bc.outflow"])
    erra = np.max(np.abs(ad.x - refa["x"])) / np.max(np.abs(refa["x"]))
    assert ad.soln.is_converged and erra <= 1e-8, erra
    Qt = g2[:, 0] * refa["x"][ca[:, 1]] - g2[:, 1] * refa["x"][ca[:, 0]]
    np.testing.assert_allclose(ad.rate(throats=[0, 7, 50], mode="single"),
                               np.abs(Qt[[0, 7, 50]]), rtol=1e-6, atol=1e-30)
    if rank == 0:
        print(f"dist ok advection-diffusion world={world} err={erra:.2e} "
              f"iters={ad.soln.cg_iters} fmt={ad._ctx.pcg_format()}", flush=True)
    # ---- transient run, slab-parallel (RK45 on the device: halo exchange per right-hand
    # side, all-reduced error norms): same accepted steps and states as scipy's RK45
    shape = (16, 10, 8)
    pnt = b2.Cubic(shape, spacing=1e-4, coords=False)
    rng = np.random.default_rng(3)
    gt = 10.0 ** rng.uniform(-15, -14, pnt.Nt)
    V = rng.uniform(0.5, 1.5, pnt.Np) * 1e-12
    pnt["pore.volume"] = V
    pht = b2.Phase(pnt)
    pht["throat.diffusive_conductance"] = gt
    pht["pore.A1"], pht["pore.A2"], pht["pore.A3"] = -2e-15, 1.5, 0.0
    pht.add_model(propname="pore.rxn", model="power_law", X="pore.concentration",
                  A1="pore.A1", A2="pore.A2", A3="pore.A3")
    for with_source in (False, True):
        alg = b2.TransientFickianDiffusion(network=pnt, phase=pht)
        alg.set_value_BC(pores=pnt.pores("left"), values=1.0)
        alg.set_rate_BC(pores=pnt.pores("right")[:20], rates=-1e-16)
        srcs = ()
        if with_source:
            sp = np.setdiff1d(pnt.Ps, np.union1d(pnt.pores("left"), pnt.pores("right")))
            alg.set_source(pores=sp, propname="pore.rxn")
            sm = np.zeros(pnt.Np, dtype=bool)
            sm[sp] = True
            srcs = [(sm, "power_law", {"A1": -2e-15, "A2": 1.5, "A3": 0.0})]
        alg.run(x0=0.1, tspan=(0, 300.0), integrator=b2.B200RK45(atol=1e-9, rtol=1e-7))
        sol = alg.soln["pore.concentration"]
        t_o, Y_o = oracle.transient_run(pnt.conns, gt, pnt.Np, V, 0.1, (0, 300.0),
                                        bc_value=alg["pore.bc.value"], bc_rate=alg["pore.bc.rate"],
                                        sources=srcs, rtol=1e-7, atol=1e-9)
        assert sol.t.size == t_o.size, (sol.t.size, t_o.size)      # same accepted steps
        np.testing.assert_allclose(sol.t, t_o, rtol=1e-9)
        assert np.max(np.abs(np.asarray(sol) - Y_o)) <= 1e-9
        assert np.array_equal(alg.x, np.asarray(sol)[:, -1])
    if rank == 0:
        print(f"dist ok transient world={world} steps={alg.soln.nsteps} nfev={alg.soln.nfev}",
              flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except BaseException:
        import traceback
        msg = traceback.format_exc().strip().splitlines()
        print(f"WORKER-ERR rank={os.environ.get('RANK')}: " + " | ".join(msg[-6:]), flush=True)
        os._exit(1)
