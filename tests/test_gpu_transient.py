"""GPU: explicit transient path against the reference's
golden vectors and the CPU oracle (scipy.solve_ivp RK45).

Tolerance: the device integrator restates SciPy's RK45 controller, so the
step sequence is the same and results agree far below the integrator's own
rtol; the bar is 1e-9 relative to the field scale (1e-7 when the integration
rtol is the default 1e-6, where one accept/reject flip would matter)."""
import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu


def test_transient_fickian_reference_test(b2, golden):
    # tests/unit/algorithms/TransientFickianDiffusionTest.py:8-36
    g = golden("g13_transient_4x3x1")
    pn = b2.Cubic([4, 3, 1], spacing=1.0)
    pn["pore.volume"] = 1e-14
    ph = b2.Phase(pn)
    ph["throat.diffusive_conductance"] = 1e-15
    alg = b2.TransientFickianDiffusion(network=pn, phase=ph)
    alg.set_value_BC(pores=pn.pores("right"), values=1)
    alg.set_value_BC(pores=pn.pores("left"), values=0)
    alg.run(x0=0, tspan=(0, 10))
    np.testing.assert_allclose(alg.x.mean(), 0.40803, rtol=1e-5)
    np.testing.assert_allclose(alg.x, g["x10"], rtol=0, atol=1e-7)
    sol = alg.soln["pore.concentration"]
    assert sol.shape[0] == pn.Np and sol.t[0] == 0 and sol.t[-1] == 10
    mid = sol(5.0)                                   # interpolation like TransientSolution
    assert mid.shape == (pn.Np,) and 0 <= mid.mean() <= alg.x.mean()
    alg.run(x0=0, tspan=(0, 200))
    np.testing.assert_allclose(alg.x.mean(), 0.5, rtol=1e-5)
    # long after steady state the controller hovers at the stability limit and the
    # state only carries the integrator's own tolerance (rtol = atol = 1e-6)
    np.testing.assert_allclose(alg.x, g["x200"], rtol=0, atol=1e-5)


@pytest.mark.parametrize("fmt", ["auto", "csr"])
def test_transient_power_law_golden(fmt, b2, golden, golden_meta):
    g = golden("g13_transient_8x7x6_power_law")
    m = golden_meta["g13_transient_8x7x6_power_law"]
    Np = g["volume"].size
    pn = b2.Network(conns=g["conns"], Np=Np)
    pn["pore.volume"] = g["volume"]
    ph = b2.Phase(pn)
    ph["throat.diffusive_conductance"] = g["g"]
    for k, v in m["source"]["params"].items():
        ph["pore." + k] = v
    ph.add_model(propname="pore.rxn", model=m["source"]["kind"], X="pore.concentration",
                 A1="pore.A1", A2="pore.A2", A3="pore.A3")
    alg = b2.TransientFickianDiffusion(network=pn, phase=ph)
    bc = np.flatnonzero(np.isfinite(g["bc_value"]))
    alg.set_value_BC(pores=bc, values=g["bc_value"][bc])
    alg.set_source(pores=g["src_pores"], propname="pore.rxn")
    alg.run(x0=m["x0"], tspan=tuple(m["tspan"]), saveat=g["t"],
            integrator=b2.B200RK45(atol=m["atol"], rtol=m["rtol"], fmt=fmt))
    sol = alg.soln["pore.concentration"]
    assert np.array_equal(sol.t, g["t"])
    assert np.max(np.abs(np.asarray(sol) - g["Y"])) <= 1e-9 * np.max(np.abs(g["Y"]))
    np.testing.assert_allclose(alg.x.mean(), m["y_end_mean"], rtol=1e-8)


def test_transient_vs_oracle_medium(b2):
    """20^3 heterogeneous diffusion, every accepted step stored (t_eval=None):
    same number of steps as scipy and the same states."""
    n = 20
    pn = b2.Cubic([n, n, n], spacing=1e-4, coords=False)
    rng = np.random.default_rng(3)
    g = 10.0 ** rng.uniform(-15, -14, pn.Nt)
    V = rng.uniform(0.5, 1.5, pn.Np) * 1e-12
    pn["pore.volume"] = V
    ph = b2.Phase(pn)
    ph["throat.diffusive_conductance"] = g
    alg = b2.TransientFickianDiffusion(network=pn, phase=ph)
    alg.set_value_BC(pores=pn.pores("left"), values=1.0)
    alg.set_rate_BC(pores=pn.pores("right")[:50], rates=-1e-16)
    alg.settings["validate_topology"] = False
    alg.run(x0=0.1, tspan=(0, 400.0), integrator=b2.B200RK45(atol=1e-9, rtol=1e-7))
    sol = alg.soln["pore.concentration"]
    t_o, Y_o = oracle.transient_run(pn.conns, g, pn.Np, V, 0.1, (0, 400.0),
                                    bc_value=alg["pore.bc.value"], bc_rate=alg["pore.bc.rate"],
                                    rtol=1e-7, atol=1e-9)
    assert sol.t.size == t_o.size                       # same accepted steps
    np.testing.assert_allclose(sol.t, t_o, rtol=1e-9)
    assert np.max(np.abs(np.asarray(sol) - Y_o)) <= 1e-9
    assert alg.soln.nsteps == t_o.size - 1
