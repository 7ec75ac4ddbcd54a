"""tests/unit/algorithms/AdvectionDiffusionTest.py of the reference, restated with
openpnm_b200's StokesFlow / AdvectionDiffusion running on REAL OpenPNM network and
phase objects (the `ad_dif` conductance model stays the reference's own code:
openpnm/models/physics/ad_dif_conductance/_funcs.py).  Known answers are the
reference's.  Skipped when the reference is not importable."""
import os
import sys

import numpy as np
import pytest
from numpy.testing import assert_allclose

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_shim  # noqa: E402

pytestmark = pytest.mark.gpu

SCHEMES = ["upwind", "hybrid", "powerlaw", "exponential"]


class Case:
    pass


@pytest.fixture(scope="module")
def case(b2):
    """AdvectionDiffusionTest.setup_class (:10-34)"""
    if ref_shim.find_reference() is None:
        pytest.skip("reference OpenPNM not available")
    op = ref_shim.import_reference()
    c = Case()
    c.op = op
    np.random.seed(0)
    c.net = op.network.Cubic(shape=[4, 3, 1], spacing=1.)
    c.phase = op.phase.Phase(network=c.net)
    c.phase["throat.diffusive_conductance"] = 1e-15
    c.phase["throat.hydraulic_conductance"] = 1e-15
    c.sf = b2.StokesFlow(network=c.net, phase=c.phase)
    c.sf.set_value_BC(pores=c.net.pores("right"), values=1)
    c.sf.set_value_BC(pores=c.net.pores("left"), values=0)
    c.sf.run(solver=b2.B200PCG(tol=1e-13))
    c.phase.update(c.sf.soln)
    mod = op.models.physics.ad_dif_conductance.ad_dif
    c.phase.add_model(propname="throat.ad_dif_conductance", model=mod, s_scheme="powerlaw")
    for s in SCHEMES:
        c.phase.add_model(propname=f"throat.ad_dif_conductance_{s}", model=mod, s_scheme=s)
    c.phase.regenerate_models()
    c.ad = b2.AdvectionDiffusion(network=c.net, phase=c.phase)
    c.ad.set_value_BC(pores=c.net.pores("right"), values=2)
    c.ad.set_value_BC(pores=c.net.pores("left"), values=0)
    c.ad.run(solver=b2.B200PCG(tol=1e-13))
    yield c
    op.Workspace().clear()


def test_settings(case):
    """:36-55"""
    ad = case.ad
    keep = {k: ad.settings[k] for k in ("quantity", "conductance", "diffusive_conductance",
                                        "hydraulic_conductance", "pressure")}
    ad.settings._update({"quantity": "pore.blah", "conductance": "throat.foo",
                         "diffusive_conductance": "throat.foo2",
                         "hydraulic_conductance": "throat.foo3", "pressure": "pore.bar"})
    assert ad.settings["quantity"] == "pore.blah"
    assert ad.settings["conductance"] == "throat.foo"
    assert ad.settings["diffusive_conductance"] == "throat.foo2"
    assert ad.settings["hydraulic_conductance"] == "throat.foo3"
    assert ad.settings["pressure"] == "pore.bar"
    ad.settings._update(keep)
    assert keep["quantity"] == "pore.concentration"
    assert keep["conductance"] == "throat.ad_dif_conductance"


@pytest.mark.parametrize("scheme,inner", [
    ("powerlaw", (0.89653, 1.53924)),          # :80-95
    ("upwind", (0.86486, 1.51351)),            # :97-112
    ("hybrid", (0.89908, 1.54128)),            # :114-130
    ("exponential", (0.89688173, 1.53952557)),  # :132-148
])
def test_scheme_known_answers(case, b2, scheme, inner):
    ad = case.ad
    ad.settings["conductance"] = f"throat.ad_dif_conductance_{scheme}"
    ad.run(solver=b2.B200PCG(tol=1e-13))
    x = [0., 0., 0.] + [inner[0]] * 3 + [inner[1]] * 3 + [2., 2., 2.]
    assert_allclose(actual=ad["pore.concentration"], desired=x, rtol=1e-5, atol=1e-12)
    ad.settings["conductance"] = "throat.ad_dif_conductance"


def test_conductance_follows_the_pressure_field(case, b2):
    """:57-78: a new StokesFlow solution changes `ad_dif`; the next AD run uses it."""
    g_old = np.copy(case.phase["throat.ad_dif_conductance"])
    x_old = np.copy(case.ad["pore.concentration"])
    case.sf.set_value_BC(pores=case.net.pores("right"), values=1.5, mode="overwrite")
    case.sf.run(solver=b2.B200PCG(tol=1e-13))
    case.phase.update(case.sf.soln)
    case.phase.regenerate_models()
    g_new = case.phase["throat.ad_dif_conductance"]
    assert g_old.mean() != g_new.mean()
    assert_allclose(g_new.mean(), 1.01258990e-15)
    case.ad.run(solver=b2.B200PCG(tol=1e-13))
    assert np.max(np.abs(case.ad["pore.concentration"] - x_old)) > 1e-3
    # back to the original pressure field for the other tests
    case.sf.set_value_BC(pores=case.net.pores("right"), values=1, mode="overwrite")
    case.sf.run(solver=b2.B200PCG(tol=1e-13))
    case.phase.update(case.sf.soln)
    case.phase.regenerate_models()


def test_outflow_BC(case, b2):
    """:150-161"""
    ad = b2.AdvectionDiffusion(network=case.net, phase=case.phase)
    ad.settings["cache"] = False
    ad.set_value_BC(pores=case.net.pores("right"), values=2)
    ad.set_outflow_BC(pores=case.net.pores("left"))
    for s in SCHEMES:
        ad.settings._update({"quantity": "pore.concentration",
                             "conductance": f"throat.ad_dif_conductance_{s}"})
        ad.run(solver=b2.B200PCG(tol=1e-13))
        assert_allclose(actual=ad["pore.concentration"].mean(), desired=2, rtol=1e-5)


def test_outflow_bc_modes(case, b2):
    """:163-172"""
    ad = b2.AdvectionDiffusion(network=case.net, phase=case.phase)
    ad.set_value_BC(pores=[2, 3], values=1)
    with pytest.raises(Exception):
        ad.set_rate_BC(pores=[2, 3], rates=2)
    ad.set_rate_BC(pores=[2, 3], rates=2, mode="overwrite")
    assert (~np.isnan(ad["pore.bc.rate"])).sum() == 2
    ad.set_value_BC(pores=[2, 3], mode="remove")
    ad.set_rate_BC(pores=[2, 3], rates=2, mode="overwrite")
    assert (~np.isnan(ad["pore.bc.rate"])).sum() == 2


def test_value_BC_does_not_overwrite_outflow(case, b2):
    """:174-179"""
    ad = b2.AdvectionDiffusion(network=case.net, phase=case.phase)
    ad.set_outflow_BC(pores=[0, 1])
    with pytest.raises(Exception):
        ad.set_value_BC(pores=[0, 1], values=1)
    assert np.isfinite(ad["pore.bc.value"]).sum() == 0


def test_add_rate_BC_fails_when_outflow_BC_present(case, b2):
    """:181-188"""
    ad = b2.AdvectionDiffusion(network=case.net, phase=case.phase)
    ad.set_outflow_BC(pores=[0, 1])
    with pytest.raises(Exception):
        ad.set_rate_BC(pores=[0, 1], rates=0.5)
    assert np.isfinite(ad["pore.bc.rate"]).sum() == 0
    ad.set_rate_BC(pores=[2, 3], rates=0.5)
    assert np.all(ad["pore.bc.rate"][[2, 3]] == 0.5)


def test_outflow_BC_rigorous(case, b2):
    """:190-217: with a sink in the interior the gradient at the outflow face is ~0."""
    op = case.op
    ad = b2.AdvectionDiffusion(network=case.net, phase=case.phase)
    ad.settings["cache"] = False
    case.phase["pore.A1"] = -5e-16
    case.phase.add_model(propname="pore.rxn", model=op.models.physics.source_terms.linear,
                         A1="pore.A1", X="pore.concentration", regen_mode="deferred")
    internal = case.net.pores(["left", "right"], mode="not")
    ad.set_source(pores=internal, propname="pore.rxn")
    ad.set_value_BC(pores=case.net.pores("right"), values=2)
    ad.set_outflow_BC(pores=case.net.pores("left"))
    for s in SCHEMES:
        ad.settings._update({"quantity": "pore.concentration",
                             "conductance": f"throat.ad_dif_conductance_{s}"})
        ad.run(solver=b2.B200PCG(tol=1e-13))
        y = ad["pore.concentration"].reshape(4, 3).mean(axis=1)
        dydx_left = (y[1] - y[0]) / (1 - 0)
        dydx_total = (y[1] - y[-1]) / (3 - 0)
        assert_allclose(actual=dydx_total + dydx_left, desired=dydx_total)
    case.phase.models.pop("pore.rxn@all", None)


def test_rate(case, b2):
    """:219-241"""
    ad = b2.AdvectionDiffusion(network=case.net, phase=case.phase)
    ad.settings["cache"] = False
    ad.set_value_BC(pores=case.net.pores("right"), values=2)
    ad.set_value_BC(pores=case.net.pores("left"), values=0)
    rng = np.random.default_rng(0)
    for s in SCHEMES:
        ad.settings._update({"quantity": "pore.concentration",
                             "conductance": f"throat.ad_dif_conductance_{s}"})
        ad.run(solver=b2.B200PCG(tol=1e-13))
        mdot_inlet = ad.rate(pores=case.net.pores("right"))[0]
        mdot_outlet = ad.rate(pores=case.net.pores("left"))[0]
        temp = rng.choice(case.net.pores(["right", "left"], mode="not"), size=3, replace=False)
        mdot_internal = ad.rate(pores=temp)[0]
        assert_allclose(mdot_inlet - mdot_internal, mdot_inlet)      # nothing generated inside
        assert_allclose(mdot_inlet, -mdot_outlet)
