"""Device path against the round-2 golden vectors of the REAL reference (g15 network
health + trim, g16 Cubic connectivity 6/14/18/26, g17 conductance that follows the
quantity), made by tests/golden/make_golden.py.

These fixtures were generated -- and the oracle pinned on them, tests/test_oracle_golden.py
-- after the round's GPU minutes were spent, so this file has not run on a B200 yet.  The
same device functions are covered by files that did run green (tests/test_gpu_topology.py,
test_gpu_dropin_openpnm.py, test_gpu_reference_suite.py: device == oracle / live reference /
host hook).  Until a first hardware run the tests are non-strict xfail: a pass is reported
as XPASS, a failure cannot turn the verified suite red, and the file is collected last."""
import numpy as np
import numpy.testing as nt
import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False,
                                reason="written after the round's GPU budget was spent; "
                                       "not yet run on hardware")]


def test_device_health_and_trim_against_the_reference_fixture(b2, golden, golden_meta):
    """openpnm/utils/_health.py:40-98 and openpnm/topotools/_topotools.py:157-229"""
    from openpnm_b200 import topotools
    g = golden("g15_health_trim_1500")
    m = golden_meta["g15_health_trim_1500"]
    pn = b2.Network(coords=g["coords"], conns=g["conns"])
    h = topotools.check_network_health(pn)
    assert h["isolated_pores"] == g["isolated"].tolist()
    assert h["disconnected_pores"] == g["disconnected"].tolist()
    assert h["looped_throats"] == g["looped"].tolist()
    assert h["n_clusters"] == m["n_clusters"]
    ctx = b2.B200PCG().ctx
    size = ctx.network_health(g["conns"], pn.Np)["cluster_size"].cpu().numpy()
    assert np.array_equal(size, g["cluster_size"])
    topotools.trim(pn, pores=h["disconnected_pores"])
    assert np.array_equal(pn.conns, g["trim1_conns"])
    nt.assert_array_equal(pn["pore.coords"], g["trim1_coords"])
    assert topotools.check_network_health(pn)["disconnected_pores"] == []
    topotools.trim(pn, throats=g["trim2_throats"])
    assert np.array_equal(pn.conns, g["trim2_conns"]) and pn.Np == m["Np_after"][1]
    topotools.trim(pn, pores=g["trim3_pores"], throats=g["trim3_throats"])
    assert np.array_equal(pn.conns, g["trim3_conns"])
    nt.assert_array_equal(pn["pore.coords"], g["trim3_coords"])
    assert (pn.Np, pn.Nt) == (m["Np_after"][2], m["Nt_after"][2])


@pytest.mark.parametrize("c,nd", [(6, 3), (14, 7), (18, 9), (26, 13)])
def test_cubic_connectivity_against_the_reference_fixture(b2, golden, c, nd):
    """openpnm/_skgraph/generators/_cubic.py:40-67 + StokesFlow with SuperLU: this
    package's lattice generator, device assembly, DIA-sym operator with nd super-diagonals
    (and the sliced-ELL one), x and rates"""
    g = golden("g16_cubic_connectivity_5x4x6")
    pn = b2.Cubic([int(v) for v in g["shape"]], spacing=1e-4, connectivity=c)
    assert np.array_equal(pn["throat.conns"], g[f"conns_{c}"])
    ph = b2.Phase(pn)
    ph["throat.hydraulic_conductance"] = g[f"g_{c}"]
    for fmt in ("auto", "csr"):
        for precond in ("jacobi", "amg"):
            sf = b2.StokesFlow(network=pn, phase=ph)
            sf.set_value_BC(pores=pn.pores("left"), values=2e5)
            sf.set_value_BC(pores=pn.pores("right"), values=1e5)
            sf.run(solver=b2.B200PCG(tol=1e-13, fmt=fmt, precond=precond))
            assert sf.soln.is_converged
            nt.assert_allclose(sf.x, g[f"x_{c}"], rtol=1e-8, atol=0)
            if fmt == "auto":
                assert sf.stats["ndiags"] == nd
            for side in ("left", "right"):
                nt.assert_allclose(sf.rate(pores=pn.pores(side)), g[f"rate_{side}_{c}"],
                                   rtol=1e-7)


# the three reference models, restated for this package's Phase stand-in; the device plan
# recognises them by name (algorithms._COND_KINDS), as it does the reference's own
def polynomial(target, a, prop):                  # models/misc/_simple_equations.py:113-117
    x = target[prop].astype(float)
    return sum(a[i] * x ** i for i in range(len(a)))


def from_neighbor_pores(target, prop, mode="min"):     # models/misc/_neighbor_lookups.py:121-138
    v = target[prop][target.network.conns]
    return {"min": np.amin, "max": np.amax, "mean": np.mean, "sum": np.sum}[mode](v, axis=1)


def generic_diffusive(target, pore_diffusivity="pore.diffusivity",
                      throat_diffusivity="throat.diffusivity",
                      size_factors="throat.diffusive_size_factors"):
    cn = target.network.conns                          # models/physics/_utils.py:46-61
    Dt = target[throat_diffusivity]
    D1, D2 = target[pore_diffusivity][cn].T
    SF = target.network[size_factors]
    if SF.ndim != 2:
        return Dt * SF
    F1, Ft, F2 = SF.T
    return 1 / (1 / (D1 * F1) + 1 / (Dt * Ft) + 1 / (D2 * F2))


def generic_hydraulic(target, pore_viscosity="pore.viscosity",
                      throat_viscosity="throat.viscosity",
                      size_factors="throat.hydraulic_size_factors"):
    cn = target.network.conns          # models/physics/hydraulic_conductance/_funcs.py:37-53
    mu1, mu2 = target[pore_viscosity][cn].T
    mut = target[throat_viscosity]
    SF = target.network[size_factors]
    F1, Ft, F2 = SF.T if SF.ndim > 1 else (np.inf, SF, np.inf)
    return 1 / (1 / (F1 / mu1) + 1 / (Ft / mut) + 1 / (F2 / mu2))


VARIABLE_G = ["diffusive_mean_sf3", "diffusive_min_sf1", "hydraulic_max_sf3",
              "hydraulic_fixed_throat_sf3", "hydraulic_mean_sf1"]


@pytest.mark.parametrize("device", [True, False])
@pytest.mark.parametrize("name", VARIABLE_G)
def test_variable_conductance_against_the_reference_fixture(b2, golden, golden_meta, name,
                                                            device):
    """the reference's ReactiveTransport with misc.polynomial ->
    misc.from_neighbor_pores -> generic_diffusive / generic_hydraulic: same Picard count,
    x, final conductance on the phase and rates -- through the device chain
    (b200_set_conductance_model) and through the host hook."""
    g = golden("g17_variable_g_" + name)
    meta = golden_meta["g17_variable_g"][name]
    net = b2.Cubic(meta["shape"])
    assert np.array_equal(net["throat.conns"], g["conns"])
    phase = b2.Phase(net)
    q = meta["quantity"]
    phase[q] = 0.0
    if meta["kind"] == "poisson":
        pkey, tkey, gkey, skey = ("pore.diffusivity", "throat.diffusivity",
                                  "throat.diffusive_conductance", "throat.diffusive_size_factors")
        gmodel = generic_diffusive
        gargs = dict(pore_diffusivity=pkey, throat_diffusivity=tkey, size_factors=skey)
    else:
        pkey, tkey, gkey, skey = ("pore.viscosity", "throat.viscosity",
                                  "throat.hydraulic_conductance", "throat.hydraulic_size_factors")
        gmodel = generic_hydraulic
        gargs = dict(pore_viscosity=pkey, throat_viscosity=tkey, size_factors=skey)
    net[skey] = g["size_factors"]
    phase.add_model(propname=pkey, model=polynomial, a=[float(v) for v in g["poly"]], prop=q)
    if g["throat_fixed"].size:
        phase[tkey] = g["throat_fixed"]
    else:
        phase.add_model(propname=tkey, model=from_neighbor_pores, prop=pkey, mode=meta["mode"])
    phase.add_model(propname=gkey, model=gmodel, **gargs)
    phase.regenerate_models()
    alg = b2.ReactiveTransport(network=net, phase=phase)
    alg.settings._update({"conductance": gkey, "quantity": q, "device_conductance": device})
    alg.set_value_BC(pores=net.pores("front"), values=10.0)
    alg.set_value_BC(pores=net.pores("back"), values=0.0)
    nt.assert_array_equal(np.nan_to_num(alg["pore.bc.value"], nan=-1.0),
                          np.nan_to_num(g["bc_value"], nan=-1.0))
    if device:
        assert alg._device_conductance_plan(alg.iterative_props) is not None
    alg.run(solver=b2.B200PCG(tol=1e-13))
    assert alg.soln.is_converged and alg.soln.num_iter == meta["num_iter"]
    nt.assert_allclose(alg.x, g["x"], rtol=1e-8, atol=1e-11)
    nt.assert_allclose(phase[gkey], g["g_final"], rtol=1e-8)
    nt.assert_allclose(phase[pkey], g["pore_prop_final"], rtol=1e-8)
    nt.assert_allclose(alg.rate(pores=net.pores("front")), g["rate_front"], rtol=1e-7)
    nt.assert_allclose(alg.rate(pores=net.pores("back")), g["rate_back"], rtol=1e-7)


@pytest.mark.parametrize("precond", ["jacobi", "amg"])
def test_config1_as_stated_against_the_reference_fixture(b2, golden, golden_meta, precond):
    """BASELINE config 1, literally (spheres_and_cylinders geometry, physics.basic, SuperLU
    by the reference): x within 1e-8 relative, both rates, the system's nnz."""
    name = "g18_config1_cubic50_spheres_and_cylinders"
    g, m = golden(name), golden_meta[name]
    pn = b2.Cubic([50, 50, 50], spacing=1e-4, coords=False)
    ph = b2.Phase(pn)
    ph["throat.hydraulic_conductance"] = g["g"]
    sf = b2.StokesFlow(network=pn, phase=ph)
    sf.set_value_BC(pores=pn.pores("left"), values=2e5)
    sf.set_value_BC(pores=pn.pores("right"), values=1e5)
    sf.run(solver=b2.B200PCG(tol=1e-13, precond=precond))
    assert sf.soln.is_converged and sf.A.nnz == m["nnz"]
    assert np.max(np.abs(sf.x - g["x"]) / np.abs(g["x"])) < 1e-8
    for side in ("left", "right"):
        r = sf.rate(pores=pn.pores(side))[0]
        assert abs(r - m["rate_" + side]) <= 1e-8 * abs(m["rate_" + side])
