"""CPU: the oracle restatement against the golden vectors generated from the
real reference (tests/golden/make_golden.py) and against the known-answer
constants of the reference's own tests."""
import numpy as np
import pytest

from oracle import oracle

LINEAR = ["g1_solvers_3x4x5", "g4_csr_4x3x2", "g6_two_values_9", "g6_value_rate_9",
          "g7_cubic20_hetero", "g7_cubic_7x5x11_zero_g", "g11_multigraph_40",
          "g12_delaunay_600"]


@pytest.mark.parametrize("name", LINEAR)
def test_linear_assembly_and_solution(name, golden, golden_meta):
    g = golden(name)
    Np = int(g["Np"])
    pure = oracle.build_A(g["conns"], g["g"], Np)
    A, b, f = oracle.apply_bcs(pure, Np, g["bc_value"], g["bc_rate"])
    M = oracle.to_csr(A, Np)
    assert np.array_equal(M.indptr, g["indptr"])          # bit-exact structure
    assert np.array_equal(M.indices, g["indices"])
    np.testing.assert_allclose(M.data, g["data"], rtol=1e-13, atol=0)
    np.testing.assert_allclose(b, g["b"], rtol=1e-12, atol=0)
    assert np.isclose(f, float(g["f"]), rtol=1e-14)
    assert M.nnz == golden_meta[name]["nnz"]
    x, code = oracle.spsolve(M, b)
    assert code == 0
    np.testing.assert_allclose(x, g["x"], rtol=1e-9, atol=0)
    for lbl, val in golden_meta[name]["rates"].items():
        r = oracle.rate(g["conns"], g["g"], x, pores=g["pores_" + lbl])[0]
        assert np.isclose(r, val, rtol=1e-8)


def test_known_answers_of_reference_tests(golden, golden_meta):
    # tests/unit/algorithms/SolversTest.py:26-49
    assert np.isclose(golden_meta["g1_solvers_3x4x5"]["x_mean"], 0.624134, rtol=1e-5)
    assert np.isclose(golden_meta["g1_solvers_3x4x5"]["x_mean_scipycg"], 0.624134, rtol=1e-5)
    # tests/unit/algorithms/SubclassedTransportTest.py:18-82
    m = golden_meta["g2_subclassed_9"]
    assert np.isclose(m["rate_in_stokes"], 1025915.625, rtol=1e-10)
    assert np.isclose(m["rate_in_fickian"], 10.125, rtol=1e-10)
    assert np.isclose(m["rate_in_stokes"], -m["rate_out_stokes"], rtol=1e-10)
    # tests/unit/algorithms/ReactiveTransportTest.py:69-77, 97-112
    assert np.isclose(golden_meta["g3_reactive_4_standard_kinetics"]["c_mean"], 0.717129, rtol=1e-6)
    assert np.isclose(golden_meta["g3_reactive_4_two_sources"]["c_mean"], 0.666667, rtol=1e-6)
    # tests/unit/algorithms/GenericTransportTest.py:114-135
    y = np.unique(np.around(golden("g6_two_values_9")["x"], decimals=3))
    assert np.allclose(y, np.linspace(0, 1, 9))
    y = np.unique(np.around(golden("g6_value_rate_9")["x"], decimals=3))
    assert np.allclose(y, np.arange(9.0))


def test_stokes_9_oracle(golden, golden_meta):
    net = oracle.cubic_network([9, 9, 9])
    bcv = np.full(net["Np"], np.nan)
    bcv[net["labels"]["top"]] = 101325
    bcv[net["labels"]["bottom"]] = 0
    out = oracle.transport_run(net["conns"], np.ones(net["Nt"]), net["Np"], bc_value=bcv)
    np.testing.assert_allclose(out["x"], golden("g2_subclassed_9")["x_stokes"], rtol=1e-9, atol=1e-6)
    r = oracle.rate(net["conns"], np.ones(net["Nt"]), out["x"], pores=net["labels"]["top"])[0]
    assert np.isclose(r, 1025915.625, rtol=1e-9)
    assert out["num_iter"] == golden_meta["g2_subclassed_9"]["num_iter"] == 1


REACTIVE = ["g3_reactive_4_standard_kinetics", "g3_reactive_4_power_law", "g3_reactive_4_linear",
            "g3_reactive_4_two_sources", "g9_reactive_10_power_law", "g3_reactive_4_diverge"]


def reactive_sources(g, meta):
    srcs = []
    Np = g["x"].size
    for n, s in enumerate(meta["sources"]):
        mask = np.zeros(Np, dtype=bool)
        mask[g[f"src{n}_pores"]] = True
        srcs.append((mask, s["kind"], s["params"]))
    return srcs


@pytest.mark.parametrize("name", REACTIVE)
def test_reactive_outer_loop(name, golden, golden_meta):
    g, m = golden(name), golden_meta[name]
    st = m["settings"]
    out = oracle.picard_run(g["conns"], g["g"], g["x"].size, g["bc_value"], g["bc_rate"],
                            reactive_sources(g, m), relaxation_factor=st["relaxation_factor"],
                            newton_maxiter=st["newton_maxiter"], f_rtol=st["f_rtol"],
                            x_rtol=st["x_rtol"])
    assert out["num_iter"] == m["num_iter"]
    assert out["is_converged"] == m["is_converged"]
    if m["is_converged"]:
        np.testing.assert_allclose(out["x"], g["x"], rtol=1e-9, atol=0)
        assert np.isclose(out["x"].mean(), m["c_mean"], rtol=1e-10)
        M = oracle.to_csr(out["A"], g["x"].size)
        assert np.array_equal(M.indptr, g["indptr"]) and np.array_equal(M.indices, g["indices"])


def test_cubic_generator_matches_reference_conns(golden):
    g = golden("g7_cubic20_hetero")
    net = oracle.cubic_network([20, 20, 20], spacing=1e-4)
    assert np.array_equal(net["conns"], g["conns"])
    assert np.array_equal(net["labels"]["left"], g["pores_left"])
    assert np.array_equal(net["labels"]["right"], g["pores_right"])


def test_conduit_analytic(golden_meta):
    # tests/integration/Flow_in_straight_conduit.py:8-55
    for row in golden_meta["g5_conduit"]:
        g = np.array(row["g"])
        conns = np.array([[0, 1], [1, 2], [2, 3], [3, 4]])
        bcv = np.full(5, np.nan)
        bcv[0], bcv[4] = 1.0, 0.0
        out = oracle.transport_run(conns, g, 5, bc_value=bcv)
        r = oracle.rate(conns, g, out["x"], pores=[0])[0]
        assert np.isclose(r, row["Q"], rtol=1e-10)


def test_jacobi_pcg_oracle_matches_direct(golden):
    g = golden("g7_cubic20_hetero")
    Np = int(g["Np"])
    import scipy.sparse as sprs
    M = sprs.csr_matrix((g["data"], g["indices"], g["indptr"]), shape=(Np, Np))
    x, info, its = oracle.jacobi_pcg(M, g["b"], rtol=1e-13)
    assert info == 0 and its > 10
    np.testing.assert_allclose(x, g["x"], rtol=1e-8, atol=0)


def test_transient_oracle_against_reference(golden, golden_meta):
    """TransientFickianDiffusionTest.py:8-36 and a heterogeneous power-law case."""
    g = golden("g13_transient_4x3x1")
    net = oracle.cubic_network([4, 3, 1])
    assert np.array_equal(net["conns"], g["conns"])
    t, Y = oracle.transient_run(net["conns"], 1e-15 * np.ones(net["Nt"]), net["Np"], 1e-14, 0.0,
                                (0, 10), bc_value=g["bc_value"])
    np.testing.assert_allclose(Y[:, -1], g["x10"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(Y[:, -1].mean(), 0.40803, rtol=1e-5)
    g = golden("g13_transient_8x7x6_power_law")
    m = golden_meta["g13_transient_8x7x6_power_law"]
    Np = g["volume"].size
    mask = np.zeros(Np, dtype=bool)
    mask[g["src_pores"]] = True
    t, Y = oracle.transient_run(g["conns"], g["g"], Np, g["volume"], m["x0"], tuple(m["tspan"]),
                                saveat=g["t"], bc_value=g["bc_value"],
                                sources=[(mask, m["source"]["kind"], m["source"]["params"])],
                                rtol=m["rtol"], atol=m["atol"])
    assert np.array_equal(t, g["t"])
    np.testing.assert_allclose(Y, g["Y"], rtol=1e-10, atol=0)


AD_CASES = ["g14_ad_4x3x1_powerlaw", "g14_ad_4x3x1_upwind", "g14_ad_9x7x5_outflow_powerlaw",
            "g14_ad_9x7x5_outflow_hybrid"]


@pytest.mark.parametrize("name", AD_CASES)
def test_advection_diffusion_oracle(name, golden):
    """AdvectionDiffusionTest.py:10-115 and heterogeneous cases with an outflow BC."""
    g = golden(name)
    Np = g["x"].size
    res = oracle.picard_run(g["conns"], g["g2"], Np, bc_value=g["bc_value"], bc_outflow=g["bc_outflow"])
    M = oracle.to_csr(res["A"], Np)
    assert np.array_equal(M.indptr, g["indptr"]) and np.array_equal(M.indices, g["indices"])
    np.testing.assert_allclose(M.data, g["data"], rtol=1e-13, atol=0)
    np.testing.assert_allclose(res["x"], g["x"], rtol=1e-9, atol=1e-12)
    assert (abs(M - M.T) > 0).nnz > 0                      # genuinely non-symmetric
    if "left" in g:
        ov = oracle.outflow_values(g["conns"], g["gh"], g["pressure"], g["left"], Np)
        m = np.isfinite(g["bc_outflow"])
        np.testing.assert_allclose(ov[m], g["bc_outflow"][m], rtol=1e-12)
    if name == "g14_ad_4x3x1_powerlaw":
        np.testing.assert_allclose(g["x"][3:9], [0.89653] * 3 + [1.53924] * 3, rtol=1e-5)


# ----------------------------------------------------------------- round-2 rows
def test_network_health_and_trim(golden, golden_meta):
    """oracle.network_health / oracle.trim against the reference's
    check_network_health, cluster_size model and three successive topotools.trim calls."""
    g = golden("g15_health_trim_1500")
    m = golden_meta["g15_health_trim_1500"]
    conns, Np = g["conns"], g["coords"].shape[0]
    h = oracle.network_health(conns, Np)
    assert h["n_clusters"] == m["n_clusters"]
    assert np.array_equal(np.flatnonzero(h["isolated"]), g["isolated"])
    assert np.array_equal(h["disconnected"], g["disconnected"])
    assert np.array_equal(h["looped"], g["looped"])
    assert np.array_equal(h["cluster_size"], g["cluster_size"])
    assert g["isolated"].size == m["n_isolated"] > 0
    c1, Pk, Tk = oracle.trim(conns, Np, pores=g["disconnected"])
    assert np.array_equal(c1, g["trim1_conns"])
    assert np.array_equal(g["coords"][Pk], g["trim1_coords"])
    assert (Pk.size, Tk.size) == (m["Np_after"][0], m["Nt_after"][0])
    assert oracle.network_health(c1, Pk.size)["disconnected"].size == 0
    c2, Pk2, Tk2 = oracle.trim(c1, Pk.size, throats=g["trim2_throats"])
    assert np.array_equal(c2, g["trim2_conns"]) and Pk2.size == m["Np_after"][1]
    c3, Pk3, Tk3 = oracle.trim(c2, Pk2.size, pores=g["trim3_pores"], throats=g["trim3_throats"])
    assert np.array_equal(c3, g["trim3_conns"])
    assert np.array_equal(g["trim1_coords"][Pk3], g["trim3_coords"])
    assert (Pk3.size, Tk3.size) == (m["Np_after"][2], m["Nt_after"][2])


@pytest.mark.parametrize("c", [6, 14, 18, 26])
def test_cubic_connectivity_solve(c, golden, golden_meta):
    """Lattices with 14 / 18 / 26 neighbours (openpnm/_skgraph/generators/_cubic.py:40-67):
    the reference's throat list, its SuperLU solution and rates; the operator has
    7 / 9 / 13 super-diagonals, the widest cases of the DIA-sym format."""
    g = golden("g16_cubic_connectivity_5x4x6")
    conns, gv, bcv = g[f"conns_{c}"], g[f"g_{c}"], g[f"bc_value_{c}"]
    Np = int(np.prod(g["shape"]))
    out = oracle.transport_run(conns, gv, Np, bc_value=bcv)
    np.testing.assert_allclose(out["x"], g[f"x_{c}"], rtol=1e-9, atol=0)
    for side in ("left", "right"):
        r = oracle.rate(conns, gv, out["x"], pores=g[side])
        np.testing.assert_allclose(r, g[f"rate_{side}_{c}"], rtol=1e-8)
    M = oracle.to_csr(oracle.apply_bcs(oracle.build_A(conns, gv, Np), Np, bcv, None)[0], Np)
    offs = np.unique(M.indices - np.repeat(np.arange(Np), np.diff(M.indptr)))
    assert int((offs > 0).sum()) == golden_meta["g16_cubic_connectivity_5x4x6"][str(c)][
        "super_diagonals"] == {6: 3, 14: 7, 18: 9, 26: 13}[c]


VARIABLE_G = ["diffusive_mean_sf3", "diffusive_min_sf1", "hydraulic_max_sf3",
              "hydraulic_fixed_throat_sf3", "hydraulic_mean_sf1"]


def variable_g_chain(g, meta):
    """g(x) of a g17 fixture: polynomial -> neighbour interpolation (or a fixed throat
    property) -> conduit conductance, from the oracle's restatements."""
    conns, poly, sf = g["conns"], g["poly"], g["size_factors"]
    fixed = g["throat_fixed"] if g["throat_fixed"].size else None

    def g_of_x(x):
        pv = oracle.polynomial(poly, x)
        tv = fixed if fixed is not None else oracle.from_neighbor_pores(conns, pv, meta["mode"])
        return oracle.conduit_conductance(meta["kind"], conns, pv, tv, sf)
    return g_of_x


@pytest.mark.parametrize("name", VARIABLE_G)
def test_variable_conductance_picard(name, golden, golden_meta):
    """a conductance that follows the quantity (the reference's
    misc.polynomial -> misc.from_neighbor_pores -> generic_diffusive / generic_hydraulic,
    run by its ReactiveTransport with SuperLU): same number of Picard iterations, same
    x, same conductance left on the phase, same rates."""
    g = golden("g17_variable_g_" + name)
    meta = golden_meta["g17_variable_g"][name]
    Np = int(np.prod(meta["shape"]))
    chain = variable_g_chain(g, meta)
    out = oracle.picard_run(g["conns"], chain, Np, bc_value=g["bc_value"])
    assert out["is_converged"] and out["num_iter"] == meta["num_iter"] > 2
    np.testing.assert_allclose(out["x"], g["x"], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(oracle.polynomial(g["poly"], g["x"]), g["pore_prop_final"],
                               rtol=1e-14)
    gfin = chain(g["x"])
    np.testing.assert_allclose(gfin, g["g_final"], rtol=1e-13, atol=0)
    front = np.flatnonzero(g["bc_value"] == 10.0)
    back = np.flatnonzero(g["bc_value"] == 0.0)
    np.testing.assert_allclose(oracle.rate(g["conns"], gfin, g["x"], pores=front),
                               g["rate_front"], rtol=1e-9)
    np.testing.assert_allclose(oracle.rate(g["conns"], gfin, g["x"], pores=back),
                               g["rate_back"], rtol=1e-9)
    # the nonlinearity bites: the linear problem with g(0) has a different solution
    lin = oracle.transport_run(g["conns"], chain(np.zeros(Np)), Np, bc_value=g["bc_value"])
    assert np.max(np.abs(lin["x"] - g["x"])) > 1e-2


def test_config1_as_stated(golden, golden_meta):
    """BASELINE config 1 / SURVEY 8d, literally: Cubic 50^3 with the reference's
    spheres_and_cylinders geometry (np.random.seed(0)) and physics.basic conductances,
    left 2e5 / right 1e5, solved by the reference with SuperLU (245 s).  SURVEY's probe
    values, the oracle's assembly, and the reference's x as the solution of the oracle's
    system (a 125 000-unknown SuperLU solve is too slow for the CPU suite; the residual
    and the rates are not)."""
    name = "g18_config1_cubic50_spheres_and_cylinders"
    g, m = golden(name), golden_meta[name]
    assert (m["Np"], m["Nt"], m["nnz"]) == (125000, 367500, 830400)
    assert np.isclose(m["rate_left"], 3.06977613e-07, rtol=1e-8)
    assert np.isclose(g["g"].min(), 2.9e-15, rtol=0.02) and np.isclose(g["g"].max(), 7.3e-13, rtol=0.01)
    net = oracle.cubic_network([50, 50, 50], spacing=1e-4)
    Np = net["Np"]
    bcv = np.full(Np, np.nan)
    bcv[net["labels"]["left"]] = 2e5
    bcv[net["labels"]["right"]] = 1e5
    A, b, f = oracle.apply_bcs(oracle.build_A(net["conns"], g["g"], Np), Np, bcv, None)
    M = oracle.to_csr(A, Np)
    assert M.nnz == m["nnz"] and np.isclose(f, m["f"], rtol=1e-13)
    assert np.linalg.norm(M @ g["x"] - b) <= 1e-13 * np.linalg.norm(b)
    for side in ("left", "right"):
        r = oracle.rate(net["conns"], g["g"], g["x"], pores=net["labels"][side])[0]
        assert np.isclose(r, m["rate_" + side], rtol=1e-10)
    assert np.isclose(m["rate_left"], -m["rate_right"], rtol=1e-10)


def test_known_answers_of_the_reference_model_tests():
    """Constants asserted by the reference's own model tests, on the oracle's restatements.
    tests/unit/models/physics/HydraulicConductanceTest.py:8-37,
    DiffusiveConductanceTest.py:8-36 (Cubic 5^3, size factors (0.123, 0.981, 0.551) or one
    factor 0.896; viscosity 1e-5; diffusivity 1.3), SourceTermTest.py:41-89 (rate = S1 X + S2
    at the linearisation point; the closed forms)."""
    net = oracle.cubic_network([5, 5, 5])
    conns, Np, Nt = net["conns"], net["Np"], net["Nt"]
    sf3 = np.ones((Nt, 3)) * (0.123, 0.981, 0.551)
    mu_p, mu_t = np.full(Np, 1e-5), np.full(Nt, 1e-5)
    g = oracle.conduit_conductance("hydraulic", conns, mu_p, mu_t, sf3)
    np.testing.assert_allclose(g.mean(), 9120.483231751232, rtol=1e-7)
    g = oracle.conduit_conductance("hydraulic", conns, mu_p, mu_t, np.full(Nt, 0.896))
    np.testing.assert_allclose(g.mean(), 89600.0, rtol=1e-7)
    D_p, D_t = np.full(Np, 1.3), np.full(Nt, 1.3)
    g = oracle.conduit_conductance("poisson", conns, D_p, D_t, sf3)
    np.testing.assert_allclose(g.mean(), 0.091204832 * 1.3, rtol=1e-7)
    g = oracle.conduit_conductance("poisson", conns, D_p, D_t, np.full(Nt, 0.896))
    np.testing.assert_allclose(g.mean(), 0.896 * 1.3, rtol=1e-7)
    # misc.from_neighbor_pores / polynomial on something one can do by hand
    v = np.arange(Np, dtype=float)
    c = conns[:5]
    assert np.array_equal(oracle.from_neighbor_pores(c, v, "min"), c.min(axis=1))
    assert np.array_equal(oracle.from_neighbor_pores(c, v, "max"), c.max(axis=1))
    assert np.array_equal(oracle.from_neighbor_pores(c, v, "mean"), c.mean(axis=1))
    np.testing.assert_allclose(oracle.polynomial([1.0, -2.0, 0.5], np.array([0.0, 2.0, 4.0])),
                               [1.0, -1.0, 1.0])
    # source terms
    X = np.random.default_rng(0).uniform(0.1, 0.9, 50)
    S1, S2, r = oracle.source_term("linear", X, dict(A1=0.5e-11, A2=1.5e-12))
    np.testing.assert_allclose(r, 0.5e-11 * X + 1.5e-12, rtol=1e-15)
    np.testing.assert_allclose(S1 * X + S2, r, rtol=1e-13)
    S1, S2, r = oracle.source_term("power_law", X, dict(A1=0.5e-12, A2=2.5, A3=-1.4e-11))
    np.testing.assert_allclose(r, 0.5e-12 * X ** 2.5 - 1.4e-11, rtol=1e-15)
    np.testing.assert_allclose(S1 * X + S2, r, rtol=1e-12)
    np.testing.assert_allclose(S1, 0.5e-12 * 2.5 * X ** 1.5, rtol=1e-14)
