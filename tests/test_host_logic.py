"""CPU: host-side mirror of the reference interface (no compute)."""
import numpy as np
import pytest

import openpnm_b200 as b2
from oracle import oracle


def test_cubic_matches_reference_numbering(golden):
    g = golden("g7_cubic20_hetero")
    pn = b2.Cubic([20, 20, 20], spacing=1e-4)
    assert np.array_equal(pn["throat.conns"], g["conns"])
    assert np.array_equal(pn.pores("left"), g["pores_left"])
    assert np.array_equal(pn.pores("right"), g["pores_right"])
    net = oracle.cubic_network([4, 3, 2])
    pn = b2.Cubic([4, 3, 2])
    for lbl in ("left", "right", "front", "back", "bottom", "top"):
        assert np.array_equal(pn.pores(lbl), net["labels"][lbl])
    assert pn.Nt == 46 and pn.Np == 24
    np.testing.assert_allclose(pn["pore.coords"], net["coords"])


def test_set_bc_bookkeeping():
    # tests/unit/algorithms/BCTest.py:18-88
    pn = b2.Cubic([3, 3, 3])
    ph = b2.Phase(pn)
    alg = b2.Transport(network=pn, phase=ph)
    alg.set_value_BC(pores=[0, 1], values=1.0)
    assert np.isfinite(alg["pore.bc.value"]).sum() == 2
    with pytest.raises(Exception):
        alg.set_rate_BC(pores=[1, 2], rates=1.0)            # 'add' over an existing BC
    alg.set_rate_BC(pores=[1, 2], rates=3.0, mode="overwrite")
    assert np.isnan(alg["pore.bc.value"][1]) and alg["pore.bc.rate"][1] == 3.0
    alg.set_value_BC(pores=[5, 6], values=[1.0, 2.0])
    assert alg["pore.bc.value"][6] == 2.0
    with pytest.raises(Exception):
        alg.set_value_BC(pores=[7, 8], values=[1.0, 2.0, 3.0])
    alg.clear_rate_BCs()
    assert np.isnan(alg["pore.bc.rate"]).all()
    alg.set_BC(pores=None, bctype=[], mode="remove")
    assert np.isnan(alg["pore.bc.value"]).all()
    assert alg["pore.source"] == {}                         # GenericTransportTest.py:370-375
    assert set(alg["pore.bc"].keys()) == {"rate", "value"}


def test_source_bc_mutual_exclusion():
    # tests/unit/algorithms/ReactiveTransportTest.py:79-95
    pn = b2.Cubic([4, 4, 4])
    ph = b2.Phase(pn)
    ph.add_model(propname="pore.reaction", model="standard_kinetics", prefactor="pore.A",
                 exponent="pore.k", X="pore.concentration")
    alg = b2.ReactiveTransport(network=pn, phase=ph)
    alg.settings._update({"quantity": "pore.concentration",
                          "conductance": "throat.diffusive_conductance"})
    alg.set_value_BC(pores=pn.pores("top"), values=1.0)
    with pytest.raises(Exception):
        alg.set_source(pores=pn.pores("top"), propname="pore.reaction")
    alg.set_source(pores=pn.pores("bottom"), propname="pore.reaction")
    with pytest.raises(Exception):
        alg.set_value_BC(pores=pn.pores("bottom"), values=1.0)
    assert alg["pore.source.reaction"].sum() == 16
    alg.set_source(pores=pn.pores("bottom")[:4], propname="pore.reaction", mode="remove")
    assert alg["pore.source.reaction"].sum() == 12
    alg.set_source(pores=[], propname="pore.reaction", mode="clear")
    assert alg["pore.source.reaction"].sum() == 0


def test_settings_of_subclasses(golden_meta):
    m = golden_meta["g2_subclassed_9"]
    pn = b2.Cubic([2, 2, 2])
    ph = b2.Phase(pn)
    sf = b2.StokesFlow(network=pn, phase=ph)
    fd = b2.FickianDiffusion(network=pn, phase=ph)
    assert sf.settings["quantity"] == m["stokes_quantity"]
    assert sf.settings["conductance"] == m["stokes_conductance"]
    assert fd.settings["quantity"] == m["fickian_quantity"]
    assert fd.settings["conductance"] == m["fickian_conductance"]
    assert sf.settings.f_rtol == 1e-6 and sf.settings.newton_maxiter == 5000


def test_clustered_topology_raises():
    # tests/unit/algorithms/GenericTransportTest.py:51-61
    pn = b2.Network(conns=[[0, 1], [1, 2], [3, 4]], Np=5)
    ph = b2.Phase(pn)
    ph["throat.g"] = 1.0
    alg = b2.Transport(network=pn, phase=ph, quantity="pore.x", conductance="throat.g")
    alg.set_value_BC(pores=[0], values=1.0)
    alg.set_value_BC(pores=[2], values=0.0)
    with pytest.raises(Exception, match="clustered"):
        alg._validate_topology_health()
    alg.set_value_BC(pores=[3], values=0.0)
    alg._validate_topology_health()


def test_solver_interface_shape():
    s = b2.B200PCG()
    assert s.tol == 1e-8 and hasattr(s, "maxiter") and callable(s.solve)
    s = b2.B200PCG(tol=1e-10, maxiter=123)
    assert s._get_atol(np.array([3.0, 4.0])) == pytest.approx(5e-10)


def test_iterative_props_follow_the_dependency_graph():
    # openpnm/algorithms/_algorithm.py:38-67: everything downstream of the quantity
    # (and of settings.variable_props), lexicographic topological order, quantity excluded
    pn = b2.Cubic([3, 3, 3])
    ph = b2.Phase(pn)
    ph.add_model(propname="pore.diffusivity", model=lambda t, c: 1.0, c="pore.concentration")
    ph.add_model(propname="throat.diffusive_conductance", model=lambda t, D: 1.0,
                 D="pore.diffusivity")
    ph.add_model(propname="pore.rxn", model="standard_kinetics", prefactor="pore.A",
                 exponent="pore.k", X="pore.concentration")
    ph.add_model(propname="pore.unrelated", model=lambda t, v: 1.0, v="pore.viscosity")
    alg = b2.ReactiveTransport(network=pn, phase=ph)
    alg.settings._update({"quantity": "pore.concentration",
                          "conductance": "throat.diffusive_conductance"})
    assert alg.iterative_props == ["pore.diffusivity", "pore.rxn", "throat.diffusive_conductance"]
    alg.settings["variable_props"] = {"pore.viscosity"}
    assert alg.iterative_props == ["pore.diffusivity", "pore.rxn", "pore.viscosity",
                                   "pore.unrelated", "throat.diffusive_conductance"]
    # a plain Transport on a phase without models iterates nothing
    assert b2.ReactiveTransport(network=pn, phase=b2.Phase(pn)).iterative_props == []


def test_update_hook_only_when_python_models_are_involved():
    pn = b2.Cubic([3, 3, 3])
    ph = b2.Phase(pn)
    ph["throat.diffusive_conductance"] = 1.0
    ph["pore.A1"], ph["pore.A2"] = -1.0, 2.0
    ph.add_model(propname="pore.rxn", model="power_law", A1="pore.A1", A2="pore.A2",
                 X="pore.concentration")
    alg = b2.ReactiveTransport(network=pn, phase=ph)
    alg.settings._update({"quantity": "pore.concentration",
                          "conductance": "throat.diffusive_conductance"})
    alg.set_source(pores=[13], propname="pore.rxn")
    assert alg._update_hook(None) is None                 # device kernel for the source
    # a parameter array that itself depends on the quantity needs the host
    ph.add_model(propname="pore.A1", model=lambda t, c: -1.0, c="pore.concentration")
    assert alg._update_hook(None) is not None
    del ph.models["pore.A1"]
    # a Python source model, and a conductance that follows the quantity
    ph.add_model(propname="pore.custom", model=lambda t, X: {"S1": 0, "S2": 0, "rate": 0},
                 X="pore.concentration")
    alg.set_source(pores=[14], propname="pore.custom")
    assert alg._update_hook(None) is not None
    alg.set_source(pores=[14], propname="pore.custom", mode="clear")
    alg.pop("pore.source.custom")
    assert alg._update_hook(None) is None
    ph.add_model(propname="throat.diffusive_conductance", model=lambda t, c: 1.0,
                 c="pore.concentration")
    assert alg._update_hook(None) is not None


def test_phase_regenerate_models_and_container_rules():
    pn = b2.Cubic([2, 2, 2])
    ph = b2.Phase(pn)
    ph["pore.x"] = 2.0
    ph.add_model(propname="pore.y", model=lambda t, x: t[x] ** 2, x="pore.x")
    ph.add_model(propname="pore.src", model=lambda t, y: {"S1": t[y], "S2": 1.0, "rate": 0.0},
                 y="pore.y")
    ph.add_model(propname="pore.kernel", model="linear", X="pore.x")      # device kernel: skipped
    ph.regenerate_models()
    assert np.all(ph["pore.y"] == 4.0) and np.all(ph["pore.src.S1"] == 4.0)
    assert set(ph["pore.src"].keys()) == {"S1", "S2", "rate"} and "pore.kernel" not in ph
    ph["pore.x"] = 3.0
    ph.regenerate_models(propnames=["pore.y"])
    assert np.all(ph["pore.y"] == 9.0) and np.all(ph["pore.src.S1"] == 4.0)
    assert pn.num_pores("left") == 4 and pn.num_throats() == pn.Nt
    alg = b2.ReactiveTransport(network=pn, phase=ph)
    alg["pore.source.a"] = True
    alg["pore.source.b"] = False
    assert set(alg.pop("pore.source").keys()) == {"a", "b"} and alg["pore.source"] == {}
    assert alg.pop("pore.source", None) is None
    with pytest.raises(KeyError):
        alg.pop("pore.nothing")


def test_b200pcg_precond_option():
    assert b2.B200PCG(precond="amg").precond == "amg"
    with pytest.raises(ValueError):
        b2.B200PCG(precond="ilu")


def test_cubic_slab_throat_ranges_are_the_slab_throat_ids():
    """the three contiguous id ranges of a lattice slab == the ids of cubic_slab_throats;
    the lazy Cubic materialises the same conns as the eager one"""
    from openpnm_b200 import dist as bdist
    from openpnm_b200.network import Cubic
    for shape in [(8, 5, 6), (13, 4, 7), (5, 1, 4), (6, 3, 1)]:
        pl = shape[1] * shape[2]
        for world in (1, 2, 3):
            for rb, re in bdist.cubic_slab_ranges(shape, world):
                ids, _ = bdist.cubic_slab_throats(shape, rb // pl, re // pl)
                rg = bdist.cubic_slab_throat_ranges(shape, rb // pl, re // pl)
                ids2 = np.concatenate([np.arange(a, b) for a, b in rg]) if rg else np.zeros(0, int)
                assert np.array_equal(ids, ids2)
        eager, lazy = Cubic(shape, coords=False), Cubic(shape, coords=False, lazy=True)
        assert "throat.conns" not in lazy.keys() and lazy.Nt == eager.Nt
        assert np.array_equal(lazy.pores("right"), eager.pores("right"))
        assert np.array_equal(lazy.conns, eager.conns)


def test_spatial_order_is_a_permutation_with_bounded_bandwidth():
    """dist.spatial_order: equal-count bins along x, Morton curve inside a bin -- a valid
    permutation whose index bandwidth for short-range neighbours stays ~2 bins, and which
    brings most neighbours within a few hundred indices (what the staged SELL kernel needs)"""
    from openpnm_b200 import dist as bdist
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(0)
    pts = rng.random((20000, 3))
    perm, inv = bdist.spatial_order(pts, axis=0, blocks=8)
    assert np.array_equal(np.sort(perm), np.arange(20000)) and np.array_equal(perm[inv], np.arange(20000))
    pairs = cKDTree(pts).query_pairs(0.05, output_type="ndarray")
    d_new = np.abs(inv[pairs[:, 0]] - inv[pairs[:, 1]])
    assert d_new.max() <= 2 * 20000 // 8 + 8            # neighbours sit in the same or the next bin
    by_x = np.argsort(np.argsort(pts[:, 0]))
    d_x = np.abs(by_x[pairs[:, 0]] - by_x[pairs[:, 1]])
    # far more local than the plain x-sort (0.83 against 0.38 here)
    assert (d_new < 256).mean() > 0.75 and (d_x < 256).mean() < 0.5
    # bins are coordinate slabs: every pore of bin k lies left of every pore of bin k + 2
    xs = pts[perm, 0].reshape(8, -1)
    assert all(xs[k].max() <= xs[k + 1].min() + 1e-12 for k in range(7))


def test_host_copy_threads():
    from openpnm_b200.algorithms import _host_copy
    a = np.random.default_rng(1).random(5_000_001)
    b = _host_copy(a)
    assert b is not a and np.array_equal(a, b)
    c = _host_copy(a[:1000])
    assert np.array_equal(c, a[:1000])


@pytest.mark.parametrize("c", [6, 14, 18, 26])
def test_cubic_connectivity_matches_reference_numbering(c, golden):
    """`Cubic(connectivity=...)`: the same throats, in the same order, as the reference's
    generator (openpnm/_skgraph/generators/_cubic.py:40-67; fixture made by the real
    reference, tests/golden/make_golden.py g16)"""
    g = golden("g16_cubic_connectivity_5x4x6")
    pn = b2.Cubic([int(v) for v in g["shape"]], spacing=1e-4, connectivity=c)
    assert np.array_equal(pn["throat.conns"], g[f"conns_{c}"])
    assert np.array_equal(pn.pores("left"), g["left"])
    assert np.array_equal(pn.pores("right"), g["right"])


def test_cubic_connectivity_lazy_and_invalid():
    from openpnm_b200.network import Cubic
    for shape in [(5, 4, 6), (5, 1, 4), (3, 3, 1), (1, 1, 7)]:
        for c in (6, 14, 18, 20, 26):
            eager, lazy = Cubic(shape, connectivity=c), Cubic(shape, connectivity=c, lazy=True)
            assert lazy.Nt == eager.Nt == eager["throat.conns"].shape[0]
            assert np.array_equal(lazy.conns, eager.conns)
            assert (eager.conns[:, 0] < eager.conns[:, 1]).all()
            # every pore pair at most once
            assert np.unique(eager.conns, axis=0).shape[0] == eager.Nt
    nx, ny, nz = 5, 4, 6
    assert Cubic((nx, ny, nz), connectivity=26).Nt == \
        Cubic((nx, ny, nz), connectivity=6).Nt + Cubic((nx, ny, nz), connectivity=20).Nt
    with pytest.raises(Exception, match="Invalid connectivity"):
        Cubic((3, 3, 3), connectivity=8)


def test_slab_plan_analytic_lattice_equals_the_masked_plan():
    """Transport._slab_plan: the index-arithmetic plan of a 6-neighbour lattice selects the
    same throats as masking the full list; a lattice with diagonal neighbours (or one whose
    conns were replaced) takes the masked plan and reports the true bandwidth"""
    from openpnm_b200 import dist as bdist
    from openpnm_b200.network import Cubic
    shape = (9, 4, 5)
    for world in (2, 3):
        for rank in range(world):
            pn = Cubic(shape, lazy=True)
            alg = b2.Transport(network=pn, phase=b2.Phase(pn))
            rb, re, band, sel, conns_loc = alg._slab_plan(world, rank)
            assert "throat.conns" not in pn.keys()            # never materialised
            full = Cubic(shape).conns
            ids = np.flatnonzero(bdist.slab_throat_mask(full, rb, re))
            ids_plan = np.concatenate([np.arange(a, b) for a, b in sel])
            assert np.array_equal(ids_plan, ids) and np.array_equal(conns_loc, full[ids])
            assert band == shape[1] * shape[2] == bdist.bandwidth(full)
            # 26 neighbours: not the analytic form
            pd = Cubic(shape, connectivity=26)
            alg = b2.Transport(network=pd, phase=b2.Phase(pd))
            rb2, re2, band2, sel2, loc2 = alg._slab_plan(world, rank)
            assert (rb2, re2) == (rb, re) and not isinstance(sel2, list)
            assert band2 == bdist.bandwidth(pd.conns) == shape[1] * shape[2] + shape[2] + 1
            assert np.array_equal(loc2, pd.conns[bdist.slab_throat_mask(pd.conns, rb, re)])


def _variable_g_alg(poly=(1e-9, 0.0, 5e-11), mode="mean", sf_cols=3, gname="generic_diffusive",
                    pname="polynomial", fixed_throat=False, extra=False):
    """ReactiveTransport whose conductance follows the quantity through three phase models
    (restated reference models, recognised by name)"""
    import test_gpu_zz_round2_goldens as R
    net = b2.Cubic([4, 3, 3])
    rng = np.random.default_rng(0)
    net["throat.diffusive_size_factors"] = rng.uniform(0.5, 2.0, (net.Nt, 3) if sf_cols == 3
                                                       else net.Nt)
    phase = b2.Phase(net)
    phase["pore.concentration"] = 0.0

    def other(target, a, prop):
        return R.polynomial(target, a, prop)
    pmodel = {"polynomial": R.polynomial, "other": other}[pname]
    gmodel = {"generic_diffusive": R.generic_diffusive}[gname]
    phase.add_model(propname="pore.diffusivity", model=pmodel, a=list(poly),
                    prop="pore.concentration")
    if fixed_throat:
        phase["throat.diffusivity"] = np.full(net.Nt, 2e-9)
    else:
        phase.add_model(propname="throat.diffusivity", model=R.from_neighbor_pores,
                        prop="pore.diffusivity", mode=mode)
    phase.add_model(propname="throat.diffusive_conductance", model=gmodel,
                    pore_diffusivity="pore.diffusivity", throat_diffusivity="throat.diffusivity",
                    size_factors="throat.diffusive_size_factors")
    if extra:      # one more property that follows the quantity through Python
        phase.add_model(propname="pore.extra", model=other, a=[0.0, 1.0], prop="pore.concentration")
    phase.regenerate_models()
    alg = b2.ReactiveTransport(network=net, phase=phase)
    alg.settings._update({"conductance": "throat.diffusive_conductance",
                          "quantity": "pore.concentration"})
    return net, phase, alg


def test_device_conductance_plan_detection():
    """Transport._device_conductance_plan: the polynomial -> neighbour interpolation ->
    generic_* chain is handed to the device kernels; anything else keeps the host hook"""
    net, phase, alg = _variable_g_alg()
    plan = alg._device_conductance_plan(alg.iterative_props)
    assert plan["kind"] == "poisson" and plan["interp"] == "mean" and plan["throat_prop"] is None
    assert np.array_equal(plan["poly"], [1e-9, 0.0, 5e-11])
    assert plan["size_factors"].shape == (net.Nt, 3) and plan["pore_prop"] == "pore.diffusivity"
    for mode in ("min", "max"):
        assert _variable_g_alg(mode=mode)[2]._device_conductance_plan(
            ["pore.diffusivity", "throat.diffusivity", "throat.diffusive_conductance"]
        )["interp"] == mode
    net, phase, alg = _variable_g_alg(sf_cols=1)
    assert alg._device_conductance_plan(alg.iterative_props)["size_factors"].shape == (net.Nt,)
    net, phase, alg = _variable_g_alg(fixed_throat=True)
    plan = alg._device_conductance_plan(alg.iterative_props)
    assert plan["throat_prop"].shape == (net.Nt,) and np.all(plan["throat_prop"] == 2e-9)
    # not the chain: host hook
    for kw in (dict(mode="sum"), dict(pname="other"), dict(poly=[1.0] * 9), dict(extra=True)):
        net, phase, alg = _variable_g_alg(**kw)
        assert alg._device_conductance_plan(alg.iterative_props) is None, kw


def test_solution_containers_behave_like_the_reference():
    """openpnm/algorithms/_solution.py:15-84: SteadyStateSolution is a VIEW of its array
    (the aliasing the Picard loop's first-iteration quirk rests on), TransientSolution
    interpolates linearly in time, keeps `.t` through slicing, refuses times outside tspan"""
    from openpnm_b200.algorithms import SteadyStateSolution, TransientSolution, Settings
    x = np.arange(5.0)
    s = SteadyStateSolution(x)
    s[:] = 7.0
    assert np.all(x == 7.0) and isinstance(s, np.ndarray)
    t = np.array([0.0, 1.0, 3.0])
    y = np.array([[0.0, 2.0, 6.0], [1.0, 1.0, 5.0]])          # (Np, len(t))
    sol = TransientSolution(t, y)
    assert np.array_equal(sol.t, t) and sol.shape == (2, 3)
    np.testing.assert_allclose(sol(0.5), [1.0, 1.0])
    np.testing.assert_allclose(sol.interpolate(2.0), [4.0, 3.0])
    np.testing.assert_allclose(sol([1.0, 3.0]), y[:, 1:])
    assert np.array_equal(sol[:, -1], [6.0, 5.0]) and np.array_equal(sol[0:1].t, t)
    with pytest.raises(ValueError):
        sol(3.5)
    st = Settings(a=1)
    st.b = 2
    st._update({"c": 3})
    assert (st.a, st["b"], st.c) == (1, 2, 3)
    with pytest.raises(AttributeError):
        st.missing


def test_clearing_bcs():
    # openpnm/algorithms/_algorithm.py:93-103, BCTest.py
    pn = b2.Cubic([3, 3, 3])
    alg = b2.Transport(network=pn, phase=b2.Phase(pn))
    alg.set_value_BC(pores=pn.pores("left"), values=1.0)
    alg.set_rate_BC(pores=pn.pores("right"), rates=2.0)
    alg.clear_value_BCs()
    assert np.isnan(alg["pore.bc.value"]).all() and np.isfinite(alg["pore.bc.rate"]).sum() == 9
    alg.set_value_BC(pores=pn.pores("left"), values=1.0)
    alg.clear_rate_BCs()
    assert np.isnan(alg["pore.bc.rate"]).all() and np.isfinite(alg["pore.bc.value"]).sum() == 9
    alg.set_rate_BC(pores=pn.pores("right"), rates=2.0)
    alg.clear_BCs()
    assert np.isnan(alg["pore.bc.rate"]).all() and np.isnan(alg["pore.bc.value"]).all()
    assert np.array_equal(alg.Ps, np.arange(27)) and alg.Ts.size == pn.Nt
