"""GPU: the BASELINE.json configurations at (or near) their stated sizes, checked
through size-independent properties because no CPU oracle finishes there:
independent residual, conservation, maximum principle, agreement between the
two operator formats, Picard contraction, and repeatability."""
import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu


def hetero(nt, seed=0):
    return 10.0 ** np.random.default_rng(seed).uniform(-15, -12, nt)


def test_config2_fickian_200_cubed(b2):
    """Cubic 200^3 FickianDiffusion (8M pores), Jacobi-PCG on one B200."""
    n = 200
    pn = b2.Cubic([n, n, n], spacing=1e-4, coords=False)
    ph = b2.Phase(pn)
    ph["throat.diffusive_conductance"] = hetero(pn.Nt)
    fd = b2.FickianDiffusion(network=pn, phase=ph)
    fd.set_value_BC(pores=pn.pores("left"), values=1.0)
    fd.set_value_BC(pores=pn.pores("right"), values=0.0)
    fd.settings["validate_topology"] = False
    s = b2.B200PCG(tol=1e-12)
    fd.run(solver=s)
    assert fd.soln.is_converged and fd.soln.num_iter == 1
    assert fd.soln.relres <= 1e-12 * (1 + 1e-6)
    assert s.ctx.pcg_format() == (2, 3)                       # 3 symmetric diagonals
    ctx = s.ctx
    x = fd._x_dev
    b = ctx.get_b()
    res = ctx.spmv(1, x) - b                                   # canonical CSR, independent kernel
    assert float(res.norm() / b.norm()) <= 2e-12
    rin, rout = fd.rate(pores=pn.pores("left"))[0], fd.rate(pores=pn.pores("right"))[0]
    assert rin > 0 and abs(rin + rout) <= 1e-8 * abs(rin)
    c = fd["pore.concentration"]
    assert c.min() >= -1e-10 and c.max() <= 1 + 1e-10
    # net rate of every interior pore is zero (steady state) to the solver tolerance
    interior = np.arange(n * n, 2 * n * n)
    q = fd.rate(pores=interior, mode="single")
    assert np.max(np.abs(q)) <= 1e-9 * abs(rin)
    # the generic CSR operator gives the same field
    fd2 = b2.FickianDiffusion(network=pn, phase=ph)
    fd2["pore.bc.value"] = fd["pore.bc.value"]
    fd2.settings["validate_topology"] = False
    fd2.run(solver=b2.B200PCG(tol=1e-12, fmt="csr"))
    assert np.max(np.abs(fd2.x - c)) <= 1e-8
    assert abs(fd2.soln.cg_iters - fd.soln.cg_iters) <= 0.05 * fd.soln.cg_iters + 10


def test_config5_reactive_power_law_on_device(b2):
    """Cubic ReactiveTransport with a nonlinear power-law sink on every free pore;
    same outer-iteration count as the oracle on 30^3, contraction at 120^3."""
    def build(n):
        pn = b2.Cubic([n, n, n], coords=False)
        ph = b2.Phase(pn)
        ph["throat.diffusive_conductance"] = hetero(pn.Nt)
        ph["pore.A1"], ph["pore.A2"], ph["pore.A3"] = -1e-13, 2.0, 0.0
        ph.add_model(propname="pore.rxn", model="power_law", X="pore.concentration",
                     A1="pore.A1", A2="pore.A2", A3="pore.A3")
        rt = b2.ReactiveTransport(network=pn, phase=ph)
        rt.settings._update({"quantity": "pore.concentration",
                             "conductance": "throat.diffusive_conductance",
                             "validate_topology": False})
        left = pn.pores("left")
        rt.set_value_BC(pores=left, values=1.0)
        rt.set_source(pores=np.setdiff1d(pn.Ps, left), propname="pore.rxn")
        return pn, ph, rt, left

    pn, ph, rt, left = build(30)
    rt.run(solver=b2.B200PCG(tol=1e-12))
    mask = np.ones(pn.Np, dtype=bool)
    mask[left] = False
    bcv = np.full(pn.Np, np.nan)
    bcv[left] = 1.0
    ref = oracle.picard_run(pn.conns, ph["throat.diffusive_conductance"], pn.Np, bc_value=bcv,
                            sources=[(mask, "power_law", {"A1": -1e-13, "A2": 2.0, "A3": 0.0})])
    assert rt.soln.is_converged and ref["is_converged"]
    assert rt.soln.num_iter == ref["num_iter"]
    assert np.max(np.abs(rt.x - ref["x"])) <= 1e-8 * np.max(np.abs(ref["x"]))
    # mass balance: what enters through the inlet is consumed by the reaction
    consumed = ph["pore.rxn.rate"][mask].sum()
    q_in = rt.rate(pores=left)[0]
    assert abs(q_in + consumed) <= 1e-6 * abs(q_in)

    pn, ph, rt, left = build(120)
    rt.run(solver=b2.B200PCG(tol=1e-10))
    assert rt.soln.is_converged and 2 <= rt.soln.num_iter <= 40
    c = rt.x
    assert c.min() >= -1e-9 and c.max() <= 1 + 1e-9
    S1, S2, r = (ph[f"pore.rxn.{k}"] for k in ("S1", "S2", "rate"))
    np.testing.assert_allclose(r, S1 * c + S2, rtol=1e-9, atol=1e-30)   # SourceTermTest.py:41-89
    consumed = r[np.setdiff1d(pn.Ps, left)].sum()
    q_in = rt.rate(pores=left)[0]
    assert abs(q_in + consumed) <= 1e-5 * abs(q_in)


def test_config5_at_the_stated_size(b2):
    """BASELINE config 5 at its stated size: Cubic 200^3 (8 M pores) ReactiveTransport with the
    power-law sink, Picard loop on the device, with the multigrid and with Jacobi inside the
    loop: same outer iterations, same x; size-independent properties (mass balance,
    r = S1 c + S2, maximum principle)."""
    n = 200
    pn = b2.Cubic([n, n, n], coords=False)
    g = hetero(pn.Nt)
    left = pn.pores("left")
    free = np.setdiff1d(pn.Ps, left)
    out = {}
    for precond in ("amg", "jacobi"):
        ph = b2.Phase(pn)
        dict.__setitem__(ph, "throat.diffusive_conductance", g)
        ph["pore.A1"], ph["pore.A2"], ph["pore.A3"] = -1e-13, 2.0, 0.0
        ph.add_model(propname="pore.rxn", model="power_law", X="pore.concentration",
                     A1="pore.A1", A2="pore.A2", A3="pore.A3")
        rt = b2.ReactiveTransport(network=pn, phase=ph)
        rt.settings._update({"quantity": "pore.concentration",
                             "conductance": "throat.diffusive_conductance"})
        rt.set_value_BC(pores=left, values=1.0)
        rt.set_source(pores=free, propname="pore.rxn")
        rt.run(solver=b2.B200PCG(tol=1e-10, precond=precond))
        assert rt.soln.is_converged
        c = rt.x
        assert c.min() >= -1e-9 and c.max() <= 1 + 1e-9
        S1, S2, r = (ph[f"pore.rxn.{k}"] for k in ("S1", "S2", "rate"))
        np.testing.assert_allclose(r, S1 * c + S2, rtol=1e-9, atol=1e-30)
        q_in = rt.rate(pores=left)[0]
        assert abs(q_in + r[free].sum()) <= 1e-5 * abs(q_in)
        out[precond] = (np.array(c), rt.soln.num_iter, rt.soln.cg_iters)
    assert out["amg"][1] == out["jacobi"][1]
    assert np.max(np.abs(out["amg"][0] - out["jacobi"][0])) <= 1e-6
    assert out["amg"][2] < out["jacobi"][2] / 4


def test_config4_irregular_delaunay_sorted(b2):
    """Delaunay network (irregular coordination), isolated pores trimmed, pores
    renumbered by x: exercises the CSR-stream operator at 150k pores and is checked
    against SuperLU on a 20k-pore instance."""
    from openpnm_b200 import dist as bdist

    def make(npts, seed):
        pts = np.random.default_rng(seed).random((npts, 3))
        c = oracle.delaunay_network(pts)
        L = np.linalg.norm(pts[c[:, 0]] - pts[c[:, 1]], axis=1)
        c = c[L < 4.0 * npts ** (-1 / 3)]                 # trim the long hull edges
        used = np.zeros(npts, dtype=bool)
        used[c.ravel()] = True                            # drop isolated pores (trim semantics)
        remap = np.cumsum(used) - 1
        pts, c = pts[used], remap[c]
        perm, inv = bdist.coordinate_order(pts, axis=0)
        return pts[perm], np.sort(inv[c], axis=1)

    for npts, check_direct in ((20000, True), (150000, False)):
        pts, conns = make(npts, 0)
        Np = pts.shape[0]
        g = hetero(conns.shape[0], 2)
        pn = b2.Network(coords=pts, conns=conns)
        ph = b2.Phase(pn)
        ph["throat.diffusive_conductance"] = g
        fd = b2.FickianDiffusion(network=pn, phase=ph)
        inlet, outlet = np.flatnonzero(pts[:, 0] < 0.02), np.flatnonzero(pts[:, 0] > 0.98)
        fd.set_value_BC(pores=inlet, values=1.0)
        fd.set_value_BC(pores=outlet, values=0.0)
        s = b2.B200PCG(tol=1e-12)
        fd.run(solver=s)                                   # includes the topology validation
        assert fd.soln.is_converged
        assert s.ctx.pcg_format()[0] == 1                  # irregular -> CSR operator
        rin, rout = fd.rate(pores=inlet)[0], fd.rate(pores=outlet)[0]
        assert rin > 0 and abs(rin + rout) <= 1e-8 * abs(rin)
        assert fd.x.min() >= -1e-9 and fd.x.max() <= 1 + 1e-9
        if check_direct:
            bcv = np.full(Np, np.nan)
            bcv[inlet], bcv[outlet] = 1.0, 0.0
            ref = oracle.transport_run(conns, g, Np, bc_value=bcv)
            assert np.max(np.abs(fd.x - ref["x"])) <= 1e-8
            M = oracle.to_csr(ref["A"], Np)
            A = fd.A
            assert np.array_equal(A.indptr, M.indptr) and np.array_equal(A.indices, M.indices)
            rref = oracle.rate(conns, g, ref["x"], pores=inlet)[0]
            assert abs(rin - rref) <= 1e-8 * abs(rref)


def test_golden_cubic50_direct_solve(b2, golden, golden_meta):
    """Config 1 scale: Cubic 50^3 against the reference's own ScipySpsolve run
    (163 s of SuperLU, tests/golden/make_golden.py --big)."""
    g, m = golden("g8_cubic50_hetero"), golden_meta["g8_cubic50_hetero"]
    pn = b2.Cubic([50, 50, 50], spacing=1e-4, coords=False)
    ph = b2.Phase(pn)
    ph["throat.hydraulic_conductance"] = hetero(pn.Nt)
    sf = b2.StokesFlow(network=pn, phase=ph)
    sf.set_value_BC(pores=pn.pores("left"), values=2e5)
    sf.set_value_BC(pores=pn.pores("right"), values=1e5)
    sf.run(solver=b2.B200PCG(tol=1e-13))
    assert sf.A.nnz == m["nnz"]
    assert np.max(np.abs(sf.x - g["x"]) / np.abs(g["x"])) < 1e-8
    for lbl in ("left", "right"):
        r = sf.rate(pores=g["pores_" + lbl])[0]
        assert abs(r - m["rates"][lbl]) <= 1e-8 * abs(m["rates"][lbl])


def test_cg_vs_cg_cubic_100(b2):
    """above 50^3 there is no direct-solve oracle, so the GPU
    PCG is checked against the oracle's numpy Jacobi-PCG at rtol 1e-13 (1 M pores)."""
    n = 100
    pn = b2.Cubic([n, n, n], spacing=1e-4, coords=False)
    g = hetero(pn.Nt)
    bcv = np.full(pn.Np, np.nan)
    bcv[pn.pores("left")] = 2e5
    bcv[pn.pores("right")] = 1e5
    A, b, f = oracle.apply_bcs(oracle.build_A(pn.conns, g, pn.Np), pn.Np, bcv, None)
    M = oracle.to_csr(A, pn.Np)
    xr, info, its_ref = oracle.jacobi_pcg(M, b, rtol=1e-13)
    assert info == 0
    ph = b2.Phase(pn)
    ph["throat.hydraulic_conductance"] = g
    sf = b2.StokesFlow(network=pn, phase=ph)
    sf["pore.bc.value"] = bcv
    sf.run(solver=b2.B200PCG(tol=1e-13))
    assert sf.soln.is_converged
    assert np.max(np.abs(sf.x - xr) / np.abs(xr)) < 1e-8
    assert abs(sf.soln.cg_iters - its_ref) <= 0.1 * its_ref + 70
    A_gpu = sf.A
    assert np.array_equal(A_gpu.indptr, M.indptr) and np.array_equal(A_gpu.indices, M.indices)
    rl = sf.rate(pores=pn.pores("left"))[0]
    rref = oracle.rate(pn.conns, g, xr, pores=pn.pores("left"))[0]
    assert abs(rl - rref) <= 1e-8 * abs(rref)
