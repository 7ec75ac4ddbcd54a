"""GPU: AdvectionDiffusion: non-symmetric assembly from an
(Nt,2) conductance, outflow BC, BiCGStab; against the reference's golden vectors."""
import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu

AD_CASES = ["g14_ad_4x3x1_powerlaw", "g14_ad_4x3x1_upwind", "g14_ad_9x7x5_outflow_powerlaw",
            "g14_ad_9x7x5_outflow_hybrid"]


def build(b2, g):
    Np = g["x"].size
    pn = b2.Network(conns=g["conns"], Np=Np)
    ph = b2.Phase(pn)
    dict.__setitem__(ph, "throat.ad_dif_conductance", g["g2"])
    ad = b2.AdvectionDiffusion(network=pn, phase=ph)
    ad["pore.bc.value"] = g["bc_value"]
    if "left" in g:
        ph["throat.hydraulic_conductance"] = g["gh"]
        ph["pore.pressure"] = g["pressure"]
        ad.set_outflow_BC(pores=g["left"])
        m = np.isfinite(g["bc_outflow"])
        np.testing.assert_allclose(ad["pore.bc.outflow"][m], g["bc_outflow"][m], rtol=1e-12)
    return pn, ph, ad


@pytest.mark.parametrize("name", AD_CASES)
def test_advection_diffusion_golden(name, b2, golden):
    g = golden(name)
    pn, ph, ad = build(b2, g)
    s = b2.B200PCG(tol=1e-13)
    ad.run(solver=s)
    assert ad.soln.is_converged and ad.soln.num_iter == 1
    A = ad.A
    assert np.array_equal(A.indptr, g["indptr"]) and np.array_equal(A.indices, g["indices"])
    np.testing.assert_allclose(A.data, g["data"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(ad.b, g["b"], rtol=1e-12, atol=1e-30)
    scale = np.max(np.abs(g["x"]))
    assert np.max(np.abs(ad.x - g["x"])) <= 1e-8 * scale
    assert s.ctx.pcg_format()[0] == 1                       # CSR operator, BiCGStab
    # rates against the reference formula on the golden x
    c, g2, x = g["conns"], g["g2"], g["x"]
    Qt = g2[:, 0] * x[c[:, 1]] - g2[:, 1] * x[c[:, 0]]
    Qp = np.zeros(x.size)
    np.add.at(Qp, c[:, 0], -Qt)
    np.add.at(Qp, c[:, 1], Qt)
    inlet = np.flatnonzero(g["bc_value"] == 2)
    r = ad.rate(pores=inlet)[0]
    assert abs(r - Qp[inlet].sum()) <= 1e-7 * abs(Qp[inlet].sum())
    rt = ad.rate(throats=[0, 3, 5], mode="single")
    np.testing.assert_allclose(rt, np.abs(Qt[[0, 3, 5]]), rtol=1e-7, atol=1e-30)


def test_advection_diffusion_medium_vs_direct(b2):
    """30^3 with a pressure-driven flow field (upwind conductance restated here)
    against SuperLU on the oracle's assembly."""
    n = 30
    pn = b2.Cubic([n, n, n], spacing=1e-4, coords=False)
    rng = np.random.default_rng(0)
    gh = 10.0 ** rng.uniform(-15, -14, pn.Nt)
    gd = 10.0 ** rng.uniform(-15.5, -14.5, pn.Nt)
    ph = b2.Phase(pn)
    ph["throat.hydraulic_conductance"] = gh
    sf = b2.StokesFlow(network=pn, phase=ph)
    sf.set_value_BC(pores=pn.pores("right"), values=1.0)
    sf.set_value_BC(pores=pn.pores("left"), values=0.0)
    sf.run(solver=b2.B200PCG(tol=1e-13))
    P = sf.x
    ph["pore.pressure"] = P
    c = pn.conns
    Q = -gh * (P[c[:, 1]] - P[c[:, 0]])                 # flow from c0 to c1
    # upwind scheme (models/physics/ad_dif_conductance/_funcs.py): w = max(Q, 0) + gd etc.
    g2 = np.column_stack((np.maximum(-Q, 0) + gd, np.maximum(Q, 0) + gd))
    dict.__setitem__(ph, "throat.ad_dif_conductance", g2)
    ad = b2.AdvectionDiffusion(network=pn, phase=ph)
    ad.set_value_BC(pores=pn.pores("right"), values=2.0)
    ad.set_outflow_BC(pores=pn.pores("left"))
    taken = np.union1d(pn.pores("right"), pn.pores("left"))
    ad.set_value_BC(pores=np.setdiff1d(pn.pores("top"), taken)[::3], values=0.3)
    ad.run(solver=b2.B200PCG(tol=1e-13))
    assert ad.soln.is_converged and ad.x.std() > 0.1
    ref = oracle.picard_run(c, g2, pn.Np, bc_value=ad["pore.bc.value"],
                            bc_outflow=ad["pore.bc.outflow"])
    assert np.max(np.abs(ad.x - ref["x"])) <= 1e-8 * np.max(np.abs(ref["x"]))
    M = oracle.to_csr(ref["A"], pn.Np)
    A = ad.A
    assert np.array_equal(A.indptr, M.indptr) and np.array_equal(A.indices, M.indices)


def test_b200pcg_solve_on_nonsymmetric_coo(b2, golden):
    """The drop-in solver detects a non-symmetric A (the reference's AdvectionDiffusion
    hands one to solver.solve) and switches from CG to BiCGStab."""
    import scipy.sparse as sprs
    g = golden("g14_ad_9x7x5_outflow_powerlaw")
    Np = g["x"].size
    A = sprs.csr_matrix((g["data"], g["indices"], g["indptr"]), shape=(Np, Np))
    s = b2.B200PCG(tol=1e-13)
    x, code = s.solve(A=A.tocoo(), b=g["b"], x0=np.zeros(Np))
    assert code == 0
    assert np.max(np.abs(x - g["x"])) <= 1e-8 * np.max(np.abs(g["x"]))
    assert s.info["fmt"] == "csr"


def test_transient_advection_diffusion_vs_oracle(b2):
    """TransientAdvectionDiffusion = the RK45 driver on the non-symmetric assembly."""
    shape = (12, 6, 5)
    pn = b2.Cubic(shape, spacing=1e-4)
    rng = np.random.default_rng(4)
    gd = 10.0 ** rng.uniform(-15.5, -14.5, pn.Nt)
    Q = rng.normal(0, 3e-15, pn.Nt)
    g2 = np.column_stack((np.maximum(-Q, 0) + gd, np.maximum(Q, 0) + gd))
    V = rng.uniform(0.5, 1.5, pn.Np) * 1e-13
    pn["pore.volume"] = V
    ph = b2.Phase(pn)
    dict.__setitem__(ph, "throat.ad_dif_conductance", g2)
    alg = b2.algorithms.TransientAdvectionDiffusion(network=pn, phase=ph)
    alg.set_value_BC(pores=pn.pores("right"), values=2.0)
    alg.set_value_BC(pores=pn.pores("left"), values=0.0)
    saveat = np.array([0.0, 1.0, 4.0, 10.0])
    alg.run(x0=0.5, tspan=(0, 10.0), saveat=saveat, integrator=b2.B200RK45(atol=1e-9, rtol=1e-8))
    sol = alg.soln["pore.concentration"]
    t_o, Y_o = oracle.transient_run(pn.conns, g2, pn.Np, V, 0.5, (0, 10.0), saveat=saveat,
                                    bc_value=alg["pore.bc.value"],
                                    bc_outflow=alg["pore.bc.outflow"], rtol=1e-8, atol=1e-9)
    assert np.array_equal(sol.t, t_o)
    assert np.max(np.abs(np.asarray(sol) - Y_o)) <= 1e-8 * np.max(np.abs(Y_o))
