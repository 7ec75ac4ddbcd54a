"""The reference's transient unit tests restated against openpnm_b200 (device RK45):

  tests/unit/algorithms/TransientReactiveTransportTest.py
  tests/unit/algorithms/TransientAdvectionDiffusionTest.py   (real OpenPNM objects:
      the `ad_dif` conductance model is the reference's own code)

Known answers (1.13133, 0.55642) are the reference's."""
import os
import sys

import numpy as np
import numpy.testing as nt
import pytest
from scipy.interpolate import interp1d

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_shim  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def alg(b2):
    """TransientReactiveTransportTest.setup_class (:9-33)"""
    net = b2.Cubic(shape=[3, 3, 1], spacing=1e-6)
    net["pore.volume"] = 1e-12
    phase = b2.Phase(net)
    phase["pore.A"] = -1e-13
    phase["pore.k"] = 2
    phase["throat.diffusive_conductance"] = 1e-12
    phase.add_model(propname="pore.reaction", model="standard_kinetics", prefactor="pore.A",
                    exponent="pore.k", X="pore.concentration", regen_mode="deferred")
    a = b2.TransientReactiveTransport(network=net, phase=phase)
    a.settings._update({"quantity": "pore.concentration",
                        "conductance": "throat.diffusive_conductance"})
    a.set_value_BC(pores=net.pores("front"), values=2)
    a.set_source(propname="pore.reaction", pores=net.pores("back"))
    a.net = net
    return a


def test_transient_reactive_transport(alg):
    """:35-39"""
    alg.run(x0=0, tspan=(0, 1))
    nt.assert_allclose(alg.x.mean(), 1.13133, rtol=1e-5)


def test_transient_solution(alg, b2):
    """:41-58"""
    alg.run(x0=0, tspan=(0, 1), saveat=0.1)
    q = alg.settings["quantity"]
    assert isinstance(alg.soln[q], b2.TransientSolution)
    nt.assert_array_equal(alg.soln[q].shape, (alg.Np, 11))
    nt.assert_array_equal(alg.soln[q].t, np.arange(0, 1.1, 0.1))
    assert hasattr(alg.soln[q], "__call__")
    f = interp1d(alg.soln[q].t, alg.soln[q])
    nt.assert_allclose(f(0.05), alg.soln[q](0.05))
    with nt.assert_raises(Exception):
        alg.soln[q](1.01)                       # no extrapolation
    with nt.assert_raises(Exception):
        alg.soln(1.01)


def test_consecutive_runs_preserves_solution(alg):
    """:60-70"""
    alg.run(x0=0, tspan=(0, 0.1))
    alg.run(x0=alg.x, tspan=(0.1, 0.3))
    out1 = alg.soln
    alg.run(x0=0, tspan=(0, 0.3))
    out2 = alg.soln
    q = alg.settings["quantity"]
    nt.assert_allclose(out1[q](0.3), out2[q](0.3), rtol=1e-5)


def test_adding_bc_over_sources(alg):
    """:72-74"""
    with nt.assert_raises(Exception):
        alg.set_value_BC(pores=alg.net.pores("back"), values=0.3)


def test_adding_sources_over_bc(alg):
    """:76-79"""
    with nt.assert_raises(Exception):
        alg.set_source(propname="pore.reaction", pores=alg.net.pores("front"))


def test_transient_advection_diffusion(b2):
    """TransientAdvectionDiffusionTest.py:8-44"""
    if ref_shim.find_reference() is None:
        pytest.skip("reference OpenPNM not available")
    op = ref_shim.import_reference()
    try:
        net = op.network.Cubic(shape=[4, 3, 1], spacing=1.0)
        phase = op.phase.Phase(network=net)
        phase["throat.diffusive_conductance"] = 1e-15
        phase["throat.hydraulic_conductance"] = 1e-15
        net["pore.volume"] = 1e-14
        sf = b2.StokesFlow(network=net, phase=phase)
        sf.settings._update({"quantity": "pore.pressure",
                             "conductance": "throat.hydraulic_conductance"})
        sf.set_value_BC(pores=net.pores("right"), values=1)
        sf.set_value_BC(pores=net.pores("left"), values=0)
        sf.run(solver=b2.B200PCG(tol=1e-13))
        phase[sf.settings["quantity"]] = sf.x
        phase.add_model(propname="throat.ad_dif_conductance",
                        model=op.models.physics.ad_dif_conductance.ad_dif, s_scheme="powerlaw")
        phase.regenerate_models()
        ad = b2.TransientAdvectionDiffusion(network=net, phase=phase)
        ad.settings._update({"quantity": "pore.concentration",
                             "conductance": "throat.ad_dif_conductance",
                             "diffusive_conductance": "throat.diffusive_conductance",
                             "hydraulic_conductance": "throat.hydraulic_conductance",
                             "pressure": "pore.pressure"})
        ad.set_value_BC(pores=net.pores("right"), values=2)
        ad.set_value_BC(pores=net.pores("left"), values=0)
        ad.run(x0=0, tspan=(0, 1))
        nt.assert_allclose(ad.x.mean(), 0.55642, rtol=1e-5)
    finally:
        op.Workspace().clear()
