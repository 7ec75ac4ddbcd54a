"""The reference's own algorithm unit tests for the transport path, restated
against openpnm_b200 with the same shapes, settings and known answers:

  tests/unit/algorithms/GenericTransportTest.py    (Transport)
  tests/unit/algorithms/ReactiveTransportTest.py   (ReactiveTransport)

Each test cites the reference test it follows.  The known-answer constants
(0.717129, 0.666667, the linear profiles) are the reference's, not ours.
"""
import numpy as np
import numpy.testing as nt
import pytest

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------ Transport
@pytest.fixture(scope="module")
def tnet(b2):
    """GenericTransportTest.setup_class (:9-19): 9^3 cubic, unit conductance."""
    net = b2.Cubic(shape=[9, 9, 9])
    phase = b2.Phase(net)
    phase["pore.mole_fraction"] = 0
    phase["throat.diffusive_conductance"] = 1.0
    return net, phase


def _transport(b2, net, phase, cls=None):
    alg = (cls or b2.Transport)(network=net, phase=phase)
    alg.settings["conductance"] = "throat.diffusive_conductance"
    alg.settings["quantity"] = "pore.mole_fraction"
    return alg


def test_remove_boundary_conditions(b2, tnet):
    """GenericTransportTest.py:29-37"""
    net, phase = tnet
    alg = b2.Transport(network=net, phase=phase)
    alg.set_value_BC(pores=net.pores("top"), values=1)
    alg.set_value_BC(pores=net.pores("bottom"), values=0)
    assert np.sum(np.isfinite(alg["pore.bc.value"])) > 0
    alg.set_value_BC(pores=net.pores("top"), mode="remove")
    assert np.sum(np.isfinite(alg["pore.bc.value"])) > 0
    alg.set_value_BC(pores=net.pores("bottom"), mode="remove")
    assert np.sum(np.isfinite(alg["pore.bc.value"])) == 0


def test_generic_transport_attaches_solution(b2, tnet):
    """GenericTransportTest.py:39-49"""
    net, phase = tnet
    alg = _transport(b2, net, phase)
    alg.set_value_BC(pores=net.pores("top"), values=1)
    alg.set_value_BC(pores=net.pores("bottom"), values=0)
    alg.run()
    assert isinstance(alg.soln[alg.settings["quantity"]], b2.SteadyStateSolution)


def test_ill_defined_topology(b2):
    """GenericTransportTest.py:51-62: 5x1x1 with the middle pore trimmed leaves
    a cluster without a boundary condition."""
    net = b2.Network(conns=[[0, 1], [2, 3]], Np=4)
    phase = b2.Phase(net)
    phase["throat.diffusive_conductance"] = 1.0
    alg = _transport(b2, net, phase)
    alg.set_value_BC(pores=[0], values=1)
    with pytest.raises(Exception):
        alg.run()


def test_linear_system_with_nans_or_infs(b2):
    """GenericTransportTest.py:64-91"""
    net = b2.Cubic(shape=[5, 1, 1])
    phase = b2.Phase(net)
    phase["throat.diffusive_conductance"] = 1.0
    alg = _transport(b2, net, phase)
    alg.set_value_BC(pores=net.pores("left"), values=1)
    for bad in (np.inf, np.nan, -np.inf):
        phase["throat.diffusive_conductance"][1] = bad
        with pytest.raises(Exception):
            alg.run()
        phase["throat.diffusive_conductance"] = 1.0
    alg.run()
    nt.assert_allclose(alg["pore.mole_fraction"], 1.0, rtol=1e-8)


def test_ill_defined_settings(b2):
    """GenericTransportTest.py:93-113"""
    net = b2.Cubic(shape=[5, 1, 1])
    phase = b2.Phase(net)
    phase["throat.diffusive_conductance"] = 1.0
    alg = _transport(b2, net, phase)
    alg.set_value_BC(pores=net.pores("left"), values=1)
    for key in ("conductance", "quantity", "phase"):
        keep = alg.settings[key]
        alg.settings[key] = None
        with pytest.raises(Exception):
            alg.run()
        alg.settings[key] = keep
    alg.run()


def test_two_value_conditions(b2, tnet):
    """GenericTransportTest.py:115-125"""
    net, phase = tnet
    alg = _transport(b2, net, phase)
    alg.set_value_BC(pores=net.pores("top"), values=1)
    alg.set_value_BC(pores=net.pores("bottom"), values=0)
    alg.run()
    x = [0.0, 0.125, 0.25, 0.375, 0.5, 0.625, 0.75, 0.875, 1.0]
    y = np.unique(np.around(alg["pore.mole_fraction"], decimals=3))
    assert np.all(x == y)


def test_one_value_one_rate(b2, tnet):
    """GenericTransportTest.py:127-135"""
    net, phase = tnet
    alg = _transport(b2, net, phase)
    alg.set_rate_BC(pores=net.pores("bottom"), rates=1)
    alg.set_value_BC(pores=net.pores("top"), values=0)
    alg.run()
    x = [0., 1., 2., 3., 4., 5., 6., 7., 8.]
    y = np.unique(np.around(alg["pore.mole_fraction"], decimals=3))
    assert np.all(x == y)


def test_set_value_bc_where_rate_is_already_set(b2, tnet):
    """GenericTransportTest.py:136-148"""
    net, phase = tnet
    alg = _transport(b2, net, phase)
    alg.set_rate_BC(pores=[0, 1], rates=1, mode="add")
    with pytest.raises(Exception):
        alg.set_value_BC(pores=[1, 2], values=0, mode="add")
    assert np.isfinite(alg["pore.bc.rate"]).sum() == 2
    assert np.isfinite(alg["pore.bc.value"]).sum() == 0
    alg.set_rate_BC(pores=[0, 1], rates=1, mode="overwrite")
    alg.set_value_BC(pores=[1, 2], values=0, mode="overwrite")
    assert np.isfinite(alg["pore.bc.rate"]).sum() == 1
    assert np.isfinite(alg["pore.bc.value"]).sum() == 2


def test_cache(b2):
    """GenericTransportTest.py:150-170: with settings.cache the pure Laplacian is
    kept across runs even if the conductance array was edited in place."""
    net = b2.Cubic(shape=[9, 9, 9])
    phase = b2.Phase(net)
    phase["throat.diffusive_conductance"] = 1.0
    alg = _transport(b2, net, phase)
    alg.set_rate_BC(pores=net.pores("bottom"), rates=1)
    alg.set_value_BC(pores=net.pores("top"), values=0)
    alg.settings["cache"] = True
    alg.run()
    x_before = alg["pore.mole_fraction"].mean()
    phase["throat.diffusive_conductance"][1] = 50.0
    alg.run()
    assert x_before == alg["pore.mole_fraction"].mean()
    alg.settings["cache"] = False
    alg.run()
    assert x_before != alg["pore.mole_fraction"].mean()


def test_rate_single_pore(b2, tnet):
    """GenericTransportTest.py:172-185"""
    net, phase = tnet
    alg = _transport(b2, net, phase, b2.ReactiveTransport)
    pores = net.pores("left")
    alg.set_rate_BC(pores=pores, rates=1.235 * np.ones(pores.size))
    alg.set_value_BC(pores=net.pores("right"), values=0.0)
    alg.run()
    rate = alg.rate(pores=net.pores("right"))[0]
    assert np.isclose(rate, -1.235 * net.pores("right").size)
    assert np.isclose(alg.rate(pores=net.Ps), 0.0)


def test_rate_multiple_pores(b2, tnet):
    """GenericTransportTest.py:187-201"""
    net, phase = tnet
    alg = _transport(b2, net, phase)
    alg.set_rate_BC(pores=[0, 1, 2, 3], rates=1.235)
    alg.set_rate_BC(pores=[5, 6, 19, 35, 0], rates=3.455, mode="overwrite")
    alg.set_value_BC(pores=[50, 51, 52, 53], values=0.0)
    alg.run()
    rate = alg.rate(pores=[50, 51, 52, 53])[0]
    assert np.isclose(rate, -(1.235 * 3 + 3.455 * 5))
    assert np.isclose(alg.rate(pores=net.Ps), 0.0)


def test_rate_multiple_values(b2, tnet):
    """GenericTransportTest.py:203-214"""
    net, phase = tnet
    alg = _transport(b2, net, phase)
    alg.set_rate_BC(pores=[0, 1, 2, 3], rates=[0, 3.5, 0.4, -12])
    alg.set_value_BC(pores=[50, 51, 52, 53], values=0.0)
    alg.run()
    rate_individual = alg.rate(pores=[0, 1, 2, 3], mode="single")
    rate_net = alg.rate(pores=[0, 1, 2, 3], mode="group")[0]
    nt.assert_allclose(rate_individual, [0, 3.5, 0.4, -12], atol=1e-10)
    nt.assert_allclose(rate_net, sum([0, 3.5, 0.4, -12]))


# ---------------------------------------------------------- ReactiveTransport
@pytest.fixture()
def rnet(b2):
    """ReactiveTransportTest.setup_class (:11-21)"""
    net = b2.Cubic(shape=[4, 4, 4])
    phase = b2.Phase(net)
    phase["throat.diffusive_conductance"] = 1e-15
    phase["pore.A"] = -1e-15
    phase["pore.k"] = 2
    phase.add_model(propname="pore.reaction", model="standard_kinetics",
                    prefactor="pore.A", exponent="pore.k",
                    X="pore.concentration", regen_mode="deferred")
    alg = b2.ReactiveTransport(network=net, phase=phase)
    alg.settings._update({"conductance": "throat.diffusive_conductance",
                          "quantity": "pore.concentration"})
    return net, phase, alg


def test_reactive_settings(rnet):
    """ReactiveTransportTest.py:23-33"""
    _, _, alg = rnet
    alg.settings._update({"conductance": "throat.cond", "quantity": "pore.test",
                          "newton_maxiter": 123, "relaxation_quantity": 3.21})
    assert alg.settings["conductance"] == "throat.cond"
    assert alg.settings["quantity"] == "pore.test"
    assert alg.settings["newton_maxiter"] == 123
    assert alg.settings["relaxation_quantity"] == 3.21


def test_multiple_set_source_with_same_name_keeps_one(rnet):
    """ReactiveTransportTest.py:57-67"""
    net, _, alg = rnet
    for _ in range(3):
        alg.set_source(pores=net.pores("bottom"), propname="pore.reaction")
    alg.set_value_BC(pores=net.pores("top"), values=1.0)
    alg.run()
    nt.assert_allclose(alg["pore.concentration"].mean(), 0.717129, rtol=1e-5)


def test_one_value_one_source(rnet):
    """ReactiveTransportTest.py:69-77"""
    net, _, alg = rnet
    alg.set_source(pores=net.pores("bottom"), propname="pore.reaction")
    alg.set_value_BC(pores=net.pores("top"), values=1.0)
    alg.run()
    nt.assert_allclose(alg["pore.concentration"].mean(), 0.717129, rtol=1e-6)


def test_source_over_BCs(rnet):
    """ReactiveTransportTest.py:79-87"""
    net, _, alg = rnet
    alg.set_value_BC(pores=net.pores("left"), values=1.0)
    alg.set_value_BC(pores=net.pores("right"), values=0.5)
    with pytest.raises(Exception):
        alg.set_source(pores=net.pores(["left", "right"]), propname="pore.reaction")


def test_multiple_source_terms_same_location(rnet):
    """ReactiveTransportTest.py:97-112"""
    net, phase, alg = rnet
    phase.add_model(propname="pore.another_reaction", model="standard_kinetics",
                    prefactor="pore.A", exponent="pore.k", X="pore.concentration",
                    regen_mode="deferred")
    alg.set_source(pores=net.pores("left"), propname="pore.reaction")
    alg.set_source(pores=net.pores("left"), propname="pore.another_reaction")
    alg.set_value_BC(pores=net.pores("right"), values=1.0)
    alg.run()
    nt.assert_allclose(alg["pore.concentration"].mean(), 0.666667, rtol=1e-5)


def test_source_term_is_set_as_iterative_prop(rnet):
    """ReactiveTransportTest.py:114-119"""
    net, _, alg = rnet
    alg.set_source(pores=net.pores("left"), propname="pore.reaction")
    assert "pore.reaction" in alg.iterative_props


def test_reset_by_assignment_and_pop(rnet):
    """The reset idiom every ReactiveTransportTest case starts with (:70-72,
    :80-82): nan the BC arrays and pop the nested 'pore.source' prefix."""
    net, _, alg = rnet
    alg.set_source(pores=net.pores("left"), propname="pore.reaction")
    alg.set_value_BC(pores=net.pores("right"), values=1.0)
    alg["pore.bc.rate"] = np.nan
    alg["pore.bc.value"] = np.nan
    alg.pop("pore.source", None)
    assert alg["pore.source"] == {}
    assert alg.pop("pore.source", None) is None
    assert np.isfinite(alg["pore.bc.value"]).sum() == 0
    # and the object is usable again
    alg.set_source(pores=net.pores("bottom"), propname="pore.reaction")
    alg.set_value_BC(pores=net.pores("top"), values=1.0)
    alg.run()
    nt.assert_allclose(alg["pore.concentration"].mean(), 0.717129, rtol=1e-6)


def test_quantity_relaxation_consistency_w_base_solution(rnet):
    """ReactiveTransportTest.py:121-132"""
    net, _, alg = rnet
    alg.set_source(pores=net.pores("bottom"), propname="pore.reaction")
    alg.set_value_BC(pores=net.pores("top"), values=1.0)
    alg.run()
    c_mean_base = alg["pore.concentration"].mean()
    alg.settings["relaxation_factor"] = 0.5
    alg.run()
    nt.assert_allclose(c_mean_base, alg["pore.concentration"].mean(), rtol=1e-6)


def test_set_source_with_modes(rnet):
    """ReactiveTransportTest.py:134-147"""
    net, _, alg = rnet
    alg.set_source(pores=net.pores("left"), propname="pore.reaction", mode="add")
    assert alg["pore.source.reaction"].sum() == net.num_pores("left")
    alg.set_source(pores=[0, 1, 2], propname="pore.reaction", mode=["clear", "add"])
    assert alg["pore.source.reaction"].sum() == 3
    alg.set_source(pores=[2, 3, 4], propname="pore.reaction", mode="add")
    assert alg["pore.source.reaction"].sum() == 5


def test_solution_should_diverge_w_large_relaxation(rnet):
    """ReactiveTransportTest.py:162-174"""
    net, _, alg = rnet
    alg.set_source(pores=net.pores("bottom"), propname="pore.reaction")
    alg.set_value_BC(pores=net.pores("top"), values=1.0)
    alg.settings._update({"relaxation_factor": 20.0, "newton_maxiter": 25})
    alg.run()
    assert not alg.soln.is_converged


# ------------------------------------------- models evaluated in Python (hook)
def _py_standard_kinetics(phase, X, prefactor, exponent):
    """A user-written source model (same algebra as source_terms/_funcs.py:21-69);
    being a Python callable it has no device kernel and goes through the
    b200_picard_cb update hook."""
    x = phase[X]
    A, b = phase[prefactor], phase[exponent]
    return {"S1": A * b * x ** (b - 1), "S2": A * (1 - b) * x ** b, "rate": A * x ** b}


def test_python_source_model_matches_device_kernel(b2, rnet):
    """The host-evaluated path (regenerate on the phase, S1/S2 handed to the device
    every outer iteration) walks the same iterates as the device kernel."""
    net, phase, alg = rnet
    alg.set_source(pores=net.pores("bottom"), propname="pore.reaction")
    alg.set_value_BC(pores=net.pores("top"), values=1.0)
    alg.run(solver=b2.B200PCG(tol=1e-13))
    x_dev, it_dev = np.array(alg.x), alg.soln.num_iter

    phase2 = b2.Phase(net)
    phase2["throat.diffusive_conductance"] = 1e-15
    phase2["pore.A"] = -1e-15
    phase2["pore.k"] = 2
    phase2.add_model(propname="pore.reaction", model=_py_standard_kinetics,
                     prefactor="pore.A", exponent="pore.k", X="pore.concentration")
    alg2 = b2.ReactiveTransport(network=net, phase=phase2)
    alg2.settings._update({"conductance": "throat.diffusive_conductance",
                           "quantity": "pore.concentration"})
    alg2.set_source(pores=net.pores("bottom"), propname="pore.reaction")
    alg2.set_value_BC(pores=net.pores("top"), values=1.0)
    assert alg2.iterative_props == ["pore.reaction"]
    assert alg2._update_hook(None) is not None and alg._update_hook(None) is None
    alg2.run(solver=b2.B200PCG(tol=1e-13))
    assert alg2.soln.is_converged and alg2.soln.num_iter == it_dev
    nt.assert_allclose(alg2.x, x_dev, rtol=1e-10)
    nt.assert_allclose(alg2.x.mean(), 0.717129, rtol=1e-6)
    # the phase is left regenerated at the converged x (_algorithm.py:89-91)
    nt.assert_allclose(phase2["pore.reaction.rate"], -1e-15 * alg2.x ** 2, rtol=1e-12)
    nt.assert_allclose(alg2.rate(pores=net.pores("top"))[0],
                       -phase2["pore.reaction.rate"][net.pores("bottom")].sum(), rtol=1e-6)


def test_python_model_error_propagates(b2, rnet):
    """An exception inside a host model aborts the device loop and surfaces as is."""
    net, _, _ = rnet

    def broken(phase, X):
        raise ZeroDivisionError("model blew up")
    phase = b2.Phase(net)
    phase["throat.diffusive_conductance"] = 1.0
    phase.add_model(propname="pore.reaction", model=broken, X="pore.concentration")
    alg = b2.ReactiveTransport(network=net, phase=phase)
    alg.settings._update({"conductance": "throat.diffusive_conductance",
                          "quantity": "pore.concentration"})
    alg.set_source(pores=net.pores("bottom"), propname="pore.reaction")
    alg.set_value_BC(pores=net.pores("top"), values=1.0)
    with pytest.raises(ZeroDivisionError):
        alg.run()


def test_variable_conductance(b2):
    """ReactiveTransportTest.py:192-214 (kept there as a commented-out case; the
    known answer is the reference's): diffusivity depends on the concentration,
    the conductance on the diffusivity, so A is rebuilt every outer iteration
    (_transport.py:105-107 turns the cache off)."""
    net = b2.Cubic(shape=[4, 4, 4])
    phase = b2.Phase(net)

    def D_var(target, pore_concentration="pore.concentration"):
        X = target[pore_concentration]
        return 1e-9 * (1 + 5 * (X / 10) ** 2)

    def g_var(target, pore_diffusivity="pore.diffusivity"):
        D = target[pore_diffusivity]
        Dt = D[target.network.conns].mean(axis=1)     # interpolate to throats
        return 1e-6 * Dt

    phase["pore.concentration"] = 0.0
    phase.add_model(propname="pore.diffusivity", model=D_var,
                    pore_concentration="pore.concentration")
    phase.add_model(propname="throat.diffusive_conductance", model=g_var,
                    pore_diffusivity="pore.diffusivity")
    phase.regenerate_models()
    alg = b2.ReactiveTransport(network=net, phase=phase)
    alg.settings._update({"conductance": "throat.diffusive_conductance",
                          "quantity": "pore.concentration"})
    assert alg.iterative_props == ["pore.diffusivity", "throat.diffusive_conductance"]
    alg.set_value_BC(pores=net.pores("front"), values=10.0)
    alg.set_value_BC(pores=net.pores("back"), values=0.0)
    alg.run(solver=b2.B200PCG(tol=1e-12))
    assert alg.soln.is_converged and alg.soln.num_iter > 2
    c_avg = alg["pore.concentration"].reshape(4, 4, 4).mean(axis=(0, 2))
    nt.assert_allclose(c_avg, [10.0, 8.18175755, 5.42194391, 0.0], rtol=1e-5)
    # net flow is conserved with the final conductance
    nt.assert_allclose(alg.rate(pores=net.pores("front"))[0],
                       -alg.rate(pores=net.pores("back"))[0], rtol=1e-5)


def _variable_g_case(b2, device):
    """diffusivity = polynomial of the concentration, throat value = mean of the pores,
    conductance = generic_diffusive with (Nt,3) size factors: the chain that has device
    kernels (b200_set_conductance_model); device=False forces the host hook instead."""
    rng = np.random.default_rng(4)
    net = b2.Cubic(shape=[6, 5, 4])
    net["throat.diffusive_size_factors"] = rng.uniform(0.5, 2.0, (net.Nt, 3)) * 1e-3
    phase = b2.Phase(net)
    phase["pore.concentration"] = 0.0

    def polynomial(target, a, prop):          # openpnm/models/misc/_simple_equations.py:88-117
        x = target[prop].astype(float)
        return sum(a[i] * x ** i for i in range(len(a)))

    def from_neighbor_pores(target, prop, mode="min"):
        v = target[prop][target.network.conns]
        return {"min": v.min(axis=1), "max": v.max(axis=1), "mean": v.mean(axis=1)}[mode]

    def generic_diffusive(target, pore_diffusivity="pore.diffusivity",
                          throat_diffusivity="throat.diffusivity",
                          size_factors="throat.diffusive_size_factors"):
        cn = target.network.conns       # models/physics/_utils.py:46-61
        Dt = target[throat_diffusivity]
        D1, D2 = target[pore_diffusivity][cn].T
        F1, Ft, F2 = target.network[size_factors].T
        return 1 / (1 / (D1 * F1) + 1 / (Dt * Ft) + 1 / (D2 * F2))

    phase.add_model(propname="pore.diffusivity", model=polynomial,
                    a=[1e-9, 0.0, 5e-11], prop="pore.concentration")
    phase.add_model(propname="throat.diffusivity", model=from_neighbor_pores,
                    prop="pore.diffusivity", mode="mean")
    phase.add_model(propname="throat.diffusive_conductance", model=generic_diffusive,
                    pore_diffusivity="pore.diffusivity", throat_diffusivity="throat.diffusivity",
                    size_factors="throat.diffusive_size_factors")
    phase.regenerate_models()
    alg = b2.ReactiveTransport(network=net, phase=phase)
    alg.settings._update({"conductance": "throat.diffusive_conductance",
                          "quantity": "pore.concentration", "device_conductance": device})
    alg.set_value_BC(pores=net.pores("front"), values=10.0)
    alg.set_value_BC(pores=net.pores("back"), values=0.0)
    return net, phase, alg


def test_variable_conductance_on_the_device(b2):
    """g(x) through polynomial -> from_neighbor_pores -> generic_diffusive is
    evaluated, re-assembled and re-linearised on the device; same iterates as the host
    hook (the reference's `_update_iterative_props` path), phase left as regenerate_models
    leaves it."""
    out = {}
    for device in (False, True):
        net, phase, alg = _variable_g_case(b2, device)
        assert alg.iterative_props == ["pore.diffusivity", "throat.diffusivity",
                                       "throat.diffusive_conductance"]
        plan = alg._device_conductance_plan(alg.iterative_props)
        assert plan is not None and plan["kind"] == "poisson" and plan["interp"] == "mean"
        alg.run(solver=b2.B200PCG(tol=1e-13))
        assert alg.soln.is_converged and alg.soln.num_iter > 2
        out[device] = (np.array(alg.x), alg.soln.num_iter,
                       np.array(phase["throat.diffusive_conductance"]),
                       alg.rate(pores=net.pores("front"))[0])
        launches = alg.stats["kernel_launches"]
    assert out[True][1] == out[False][1]                      # same outer iterations
    nt.assert_allclose(out[True][0], out[False][0], rtol=1e-9, atol=1e-12)
    nt.assert_allclose(out[True][2], out[False][2], rtol=1e-9)   # final g on the phase
    nt.assert_allclose(out[True][3], out[False][3], rtol=1e-8)
    # the nonlinearity bites: the profile is not the linear one
    c = out[True][0].reshape(6, 5, 4).mean(axis=(0, 2))
    nt.assert_allclose([c[0], c[-1]], [10.0, 0.0], atol=1e-9)
    assert abs(c[2] - 5.0) > 0.3, c
