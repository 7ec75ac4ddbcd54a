"""Generate the golden vectors under tests/golden/ from the REAL reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py [--big]

Every case runs the unmodified reference (imported through tests/ref_shim.py),
stores inputs + outputs in `tests/golden/<case>.npz`, and at the same time
checks the oracle restatement (oracle/oracle.py) against the reference's
output -- this is the step that *pins* the oracle.  `--big` adds the Cubic 50^3
SuperLU case (~3 minutes of CPU).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

import ref_shim  # noqa: E402
from oracle import oracle  # noqa: E402

op = ref_shim.import_reference()
assert op is not None, "reference not found"
META = {}


def save(name, **arrays):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)


def csr_of(alg):
    A = alg.A.tocsr()
    A.sort_indices()
    return A


def check_oracle_linear(name, pn, g, bcv, bcr, alg, rtol=1e-9):
    """Oracle vs reference on the assembled system and the solution."""
    Np = pn.Np
    conns = pn["throat.conns"]
    pure = oracle.build_A(conns, g, Np)
    A, b, f = oracle.apply_bcs(pure, Np, bcv, bcr)
    Ao = oracle.to_csr(A, Np)
    Ar = csr_of(alg)
    assert np.array_equal(Ao.indptr, Ar.indptr), name
    assert np.array_equal(Ao.indices, Ar.indices), name
    assert np.allclose(Ao.data, Ar.data, rtol=1e-13, atol=0), name
    assert np.allclose(b, alg.b, rtol=1e-12, atol=0), name
    x, _ = oracle.spsolve(Ao, b)
    assert np.allclose(x, alg.x, rtol=rtol, atol=0), name
    return Ao, b, f


def linear_case(name, pn, g, value_bcs=(), rate_bcs=(), rate_pores=None,
                store_full=True, solver=None):
    ph = op.phase.Phase(network=pn)
    ph["throat.conductance"] = g
    alg = op.algorithms.Transport(network=pn, phase=ph)
    alg.settings._update({"quantity": "pore.x", "conductance": "throat.conductance"})
    for pores, vals, mode in value_bcs:
        alg.set_value_BC(pores=pores, values=vals, mode=mode)
    for pores, vals, mode in rate_bcs:
        alg.set_rate_BC(pores=pores, rates=vals, mode=mode)
    t0 = time.time()
    alg.run(solver=solver or op.solvers.ScipySpsolve())
    dt = time.time() - t0
    bcv = np.array(alg["pore.bc.value"], dtype=float)
    bcr = np.array(alg["pore.bc.rate"], dtype=float)
    g = np.array(ph["throat.conductance"], dtype=float)
    Ao, b, f = check_oracle_linear(name, pn, g, bcv, bcr, alg)
    A = csr_of(alg)
    out = dict(conns=pn["throat.conns"].astype(np.int64), g=g, bc_value=bcv,
               bc_rate=bcr, Np=np.int64(pn.Np), f=np.float64(f),
               b=np.array(alg.b), x=np.array(alg.x),
               indptr=A.indptr.astype(np.int32), indices=A.indices.astype(np.int32),
               data=A.data)
    rates = {}
    if rate_pores is not None:
        for lbl, pores in rate_pores.items():
            r = alg.rate(pores=pores)[0]
            ro = oracle.rate(pn["throat.conns"], g, np.array(alg.x), pores=pores)[0]
            assert np.isclose(r, ro, rtol=1e-10, atol=1e-300), (name, lbl, r, ro)
            rates[lbl] = float(r)
            out["pores_" + lbl] = np.asarray(pores, dtype=np.int64)
    if not store_full:
        for k in ("conns", "g", "indptr", "indices", "data", "b"):
            out.pop(k)
    save(name, **out)
    META[name] = dict(Np=int(pn.Np), Nt=int(pn.Nt), nnz=int(A.nnz), f=float(f),
                      x_mean=float(np.mean(alg.x)), rates=rates,
                      num_iter=int(getattr(alg.soln, "num_iter", 1)),
                      is_converged=bool(alg.soln.is_converged),
                      ref_seconds=round(dt, 3))
    print(name, META[name])
    return alg


def g1_solvers():
    # tests/unit/algorithms/SolversTest.py:12-24
    pn = op.network.Cubic(shape=[3, 4, 5])
    g = np.linspace(1, 5, pn.Nt)
    alg = linear_case("g1_solvers_3x4x5", pn, g,
                      value_bcs=[(pn.pores("front"), 1.0, "add"),
                                 (pn.pores("bottom"), 0.0, "overwrite")],
                      rate_pores={"front": pn.pores("front"), "bottom": pn.pores("bottom")})
    assert np.isclose(alg.x.mean(), 0.6241341116805926, rtol=1e-12)
    alg.run(solver=op.solvers.ScipyCG())
    META["g1_solvers_3x4x5"]["x_mean_scipycg"] = float(alg.x.mean())


def g2_stokes():
    # tests/unit/algorithms/SubclassedTransportTest.py:18-82 (Stokes leg)
    pn = op.network.Cubic(shape=[9, 9, 9])
    ph = op.phase.Phase(network=pn)
    ph["throat.hydraulic_conductance"] = 1.0
    ph["throat.diffusive_conductance"] = 1.0
    sf = op.algorithms.StokesFlow(network=pn, phase=ph)
    sf.set_value_BC(pores=pn.pores("top"), values=101325)
    sf.set_value_BC(pores=pn.pores("bottom"), values=0)
    sf.run(solver=op.solvers.ScipySpsolve())
    r_in = sf.rate(pores=pn.pores("top"))[0]
    r_out = sf.rate(pores=pn.pores("bottom"))[0]
    surf = sf.rate(pores=pn.pores("surface"))[0] if "pore.surface" in pn.keys() else np.nan
    fd = op.algorithms.FickianDiffusion(network=pn, phase=ph)
    fd.set_value_BC(pores=pn.pores("top"), values=1)
    fd.set_value_BC(pores=pn.pores("bottom"), values=0)
    fd.run(solver=op.solvers.ScipySpsolve())
    r_fd = fd.rate(pores=pn.pores("top"))[0]
    save("g2_subclassed_9", x_stokes=np.array(sf.x), x_fickian=np.array(fd.x),
         top=pn.pores("top").astype(np.int64), bottom=pn.pores("bottom").astype(np.int64))
    META["g2_subclassed_9"] = dict(rate_in_stokes=float(r_in), rate_out_stokes=float(r_out),
                                   rate_surface=float(surf), rate_in_fickian=float(r_fd),
                                   num_iter=int(sf.soln.num_iter),
                                   stokes_quantity=sf.settings["quantity"],
                                   stokes_conductance=sf.settings["conductance"],
                                   fickian_quantity=fd.settings["quantity"],
                                   fickian_conductance=fd.settings["conductance"])
    print("g2", META["g2_subclassed_9"])


def reactive_case(name, shape, g, kind, params, src_label, bc_label, bc_val,
                  extra_sources=(), x0=None, settings=None):
    from openpnm.models.physics import source_terms
    pn = op.network.Cubic(shape=shape)
    ph = op.phase.Phase(network=pn)
    if np.isscalar(g):
        ph["throat.diffusive_conductance"] = g
    else:
        ph["throat.diffusive_conductance"] = g(pn) if callable(g) else g
    srcs = [(src_label, kind, params)] + list(extra_sources)
    alg = op.algorithms.ReactiveTransport(network=pn, phase=ph)
    alg.settings._update({"conductance": "throat.diffusive_conductance",
                          "quantity": "pore.concentration"})
    if settings:
        alg.settings._update(settings)
    out_src = {}
    for n, (lbl, knd, prm) in enumerate(srcs):
        kw = {}
        for k, v in prm.items():
            key = f"pore.src{n}_{k}"
            ph[key] = v
            kw[k] = key
        ph.add_model(propname=f"pore.reaction{n}", model=getattr(source_terms, knd),
                     X="pore.concentration", regen_mode="deferred", **kw)
        pores = pn.pores(lbl) if isinstance(lbl, str) else lbl
        alg.set_source(pores=pores, propname=f"pore.reaction{n}")
        out_src[f"src{n}_pores"] = np.asarray(pores, dtype=np.int64)
    bc_pores = pn.pores(bc_label)
    alg.set_value_BC(pores=bc_pores, values=bc_val)
    alg.run(solver=op.solvers.ScipySpsolve(), x0=x0)
    c = np.array(alg["pore.concentration"])
    gv = np.array(ph["throat.diffusive_conductance"], dtype=float)
    # oracle cross-check
    osrc = []
    for n, (lbl, knd, prm) in enumerate(srcs):
        mask = np.zeros(pn.Np, dtype=bool)
        mask[out_src[f"src{n}_pores"]] = True
        osrc.append((mask, knd, prm))
    res = oracle.picard_run(pn["throat.conns"], gv, pn.Np,
                            bc_value=np.array(alg["pore.bc.value"], dtype=float),
                            bc_rate=np.array(alg["pore.bc.rate"], dtype=float),
                            sources=osrc, x0=x0,
                            relaxation_factor=alg.settings["relaxation_factor"],
                            newton_maxiter=alg.settings["newton_maxiter"],
                            f_rtol=alg.settings["f_rtol"], x_rtol=alg.settings["x_rtol"])
    assert res["num_iter"] == alg.soln.num_iter, (name, res["num_iter"], alg.soln.num_iter)
    assert res["is_converged"] == alg.soln.is_converged, name
    if alg.soln.is_converged:
        assert np.allclose(res["x"], c, rtol=1e-9, atol=0), name
    Ao = oracle.to_csr(res["A"], pn.Np)
    Ar = csr_of(alg)
    assert np.array_equal(Ao.indptr, Ar.indptr) and np.array_equal(Ao.indices, Ar.indices), name
    r_src = alg.rate(pores=out_src["src0_pores"])[0]
    save(name, conns=pn["throat.conns"].astype(np.int64), g=gv,
         bc_value=np.array(alg["pore.bc.value"], dtype=float),
         bc_rate=np.array(alg["pore.bc.rate"], dtype=float), x=c,
         x0=np.zeros(pn.Np) if x0 is None else x0,
         indptr=Ar.indptr.astype(np.int32), indices=Ar.indices.astype(np.int32),
         data=Ar.data, b=np.array(alg.b), **out_src)
    META[name] = dict(Np=int(pn.Np), c_mean=float(c.mean()), num_iter=int(alg.soln.num_iter),
                      is_converged=bool(alg.soln.is_converged), rate_src0=float(r_src),
                      sources=[dict(kind=k, params={a: float(b) for a, b in p.items()})
                               for _, k, p in srcs],
                      settings=dict(relaxation_factor=float(alg.settings["relaxation_factor"]),
                                    newton_maxiter=int(alg.settings["newton_maxiter"]),
                                    f_rtol=float(alg.settings["f_rtol"]),
                                    x_rtol=float(alg.settings["x_rtol"])))
    print(name, META[name])


def g3_reactive():
    # tests/unit/algorithms/ReactiveTransportTest.py:11-21, 69-77
    reactive_case("g3_reactive_4_standard_kinetics", [4, 4, 4], 1e-15,
                  "standard_kinetics", {"prefactor": -1e-15, "exponent": 2},
                  "bottom", "top", 1.0)
    assert np.isclose(META["g3_reactive_4_standard_kinetics"]["c_mean"],
                      0.7171292729553325, rtol=1e-10)
    reactive_case("g3_reactive_4_power_law", [4, 4, 4], 1e-15,
                  "power_law", {"A1": -1e-15, "A2": 2, "A3": 0.0},
                  "bottom", "top", 1.0)
    reactive_case("g3_reactive_4_linear", [4, 4, 4], 1e-15,
                  "linear", {"A1": -2e-15, "A2": 1e-16},
                  "bottom", "top", 1.0)
    # ReactiveTransportTest.py:97-112: two sources on 'left', value on 'right'
    reactive_case("g3_reactive_4_two_sources", [4, 4, 4], 1e-15,
                  "standard_kinetics", {"prefactor": -1e-15, "exponent": 2},
                  "left", "right", 1.0,
                  extra_sources=[("left", "standard_kinetics",
                                  {"prefactor": -1e-15, "exponent": 2})])
    # ReactiveTransportTest.py:162-174: divergence with relaxation_factor=20
    reactive_case("g3_reactive_4_diverge", [4, 4, 4], 1e-15,
                  "standard_kinetics", {"prefactor": -1e-15, "exponent": 2},
                  "bottom", "top", 1.0,
                  settings={"relaxation_factor": 20.0, "newton_maxiter": 25})
    # heterogeneous 10^3 power-law (config 5 in miniature)

    def ghet(pn):
        return 10.0 ** np.random.default_rng(0).uniform(-15, -12, pn.Nt)

    pn_tmp = op.network.Cubic(shape=[10, 10, 10])
    interior = np.setdiff1d(pn_tmp.Ps, pn_tmp.pores("left"))
    reactive_case("g9_reactive_10_power_law", [10, 10, 10], ghet,
                  "power_law", {"A1": -1e-13, "A2": 2, "A3": 0.0},
                  interior, "left", 1.0)


def g4_csr():
    # Cubic 4x3x2, g = 1..46, left=2.0 / right=1.0
    pn = op.network.Cubic(shape=[4, 3, 2])
    g = np.arange(1, pn.Nt + 1, dtype=float)
    alg = linear_case("g4_csr_4x3x2", pn, g,
                      value_bcs=[(pn.pores("left"), 2.0, "add"),
                                 (pn.pores("right"), 1.0, "add")],
                      rate_pores={"left": pn.pores("left"), "right": pn.pores("right")})
    A = csr_of(alg)
    assert A.nnz == 64
    assert np.isclose(alg.rate(pores=pn.pores("left"))[0], 73.70713723, rtol=1e-8)


def g6_profiles():
    # tests/unit/algorithms/GenericTransportTest.py:114-135
    pn = op.network.Cubic(shape=[9, 9, 9])
    g = np.ones(pn.Nt)
    a = linear_case("g6_two_values_9", pn, g,
                    value_bcs=[(pn.pores("top"), 1.0, "add"), (pn.pores("bottom"), 0.0, "add")],
                    rate_pores={"top": pn.pores("top"), "bottom": pn.pores("bottom")})
    y = np.unique(np.around(a.x, decimals=3))
    assert np.allclose(y, np.linspace(0, 1, 9))
    a = linear_case("g6_value_rate_9", pn, g,
                    value_bcs=[(pn.pores("top"), 0.0, "add")],
                    rate_bcs=[(pn.pores("bottom"), 1.0, "add")],
                    rate_pores={"top": pn.pores("top"), "bottom": pn.pores("bottom")})
    y = np.unique(np.around(a.x, decimals=3))
    assert np.allclose(y, np.arange(9.0))


def g5_conduit():
    # tests/integration/Flow_in_straight_conduit.py:8-55 (analytic Hagen-Poiseuille)
    mu, Pin, Pout = 1.0, 1.0, 0.0
    rows = []
    for D in [0.9, 1.0, 1.1, 1.5]:
        Q = np.pi * (D / 2) ** 4 / (8 * mu * 4) * (Pin - Pout)
        pn = op.network.Cubic(shape=[5, 1, 1], spacing=1)
        pn["pore.diameter"] = D
        pn["throat.diameter"] = D
        pn.regenerate_models()
        pn.add_model(propname="throat.hydraulic_size_factors",
                     model=op.models.geometry.hydraulic_size_factors.spheres_and_cylinders)
        ph = op.phase.Phase(network=pn)
        ph["pore.viscosity"] = mu
        ph.add_model(propname="throat.hydraulic_conductance",
                     model=op.models.physics.hydraulic_conductance.generic_hydraulic)
        sf = op.algorithms.StokesFlow(network=pn, phase=ph)
        sf.set_value_BC(pores=pn.pores("xmin"), values=Pin)
        sf.set_value_BC(pores=pn.pores("xmax"), values=Pout)
        sf.run(solver=op.solvers.ScipySpsolve())
        r = sf.rate(pores=pn.pores("xmin"))[0]
        assert np.isclose(r, Q, rtol=1e-10)
        rows.append(dict(D=D, Q=float(Q), rate=float(r),
                         g=[float(v) for v in ph["throat.hydraulic_conductance"]]))
    META["g5_conduit"] = rows
    print("g5", rows)


def g7_hetero():
    for n in (20,):
        pn = op.network.Cubic(shape=[n, n, n], spacing=1e-4)
        g = 10.0 ** np.random.default_rng(0).uniform(-15, -12, pn.Nt)
        linear_case(f"g7_cubic{n}_hetero", pn, g,
                    value_bcs=[(pn.pores("left"), 2e5, "add"), (pn.pores("right"), 1e5, "add")],
                    rate_pores={"left": pn.pores("left"), "right": pn.pores("right")})
    # non-cubic aspect, rate + value BCs, some zero conductances
    pn = op.network.Cubic(shape=[7, 5, 11])
    rng = np.random.default_rng(1)
    g = rng.uniform(0.5, 2.0, pn.Nt)
    g[rng.choice(pn.Nt, 40, replace=False)] = 0.0
    linear_case("g7_cubic_7x5x11_zero_g", pn, g,
                value_bcs=[(pn.pores("back"), 3.0, "add")],
                rate_bcs=[(pn.pores("front"), -0.25, "add")],
                rate_pores={"front": pn.pores("front"), "back": pn.pores("back")})


def g11_irregular():
    # duplicate throats, a self-loop-free multigraph and zero conductances on a
    # hand-made network (semantics of tocsr / eliminate_zeros / laplacian)
    rng = np.random.default_rng(5)
    Np = 40
    coords = rng.random((Np, 3))
    c = rng.integers(0, Np, size=(150, 2))
    c = c[c[:, 0] != c[:, 1]]
    c = np.sort(c, axis=1)
    chain = np.vstack((np.arange(Np - 1), np.arange(1, Np))).T   # keep it connected
    conns = np.vstack((c, chain, c[:10]))                        # 10 duplicated throats
    pn = op.network.Network(coords=coords, conns=conns)
    g = rng.uniform(1, 2, pn.Nt)
    g[::17] = 0.0
    linear_case("g11_multigraph_40", pn, g,
                value_bcs=[(np.arange(0, 4), 1.5, "add"), (np.arange(36, 40), -0.5, "add")],
                rate_pores={"in": np.arange(0, 4), "out": np.arange(36, 40)})
    # Delaunay network, BC by coordinate (config 4 in miniature)
    pts = np.random.default_rng(0).random((600, 3))
    dn = op.network.Delaunay(points=pts, shape=[1, 1, 1], reflect=False)
    co = oracle.delaunay_network(np.array(dn["pore.coords"]))
    cr = np.array(dn["throat.conns"])
    cr = cr[np.lexsort((cr[:, 1], cr[:, 0]))]
    assert np.array_equal(co, cr), "oracle delaunay edges differ from reference"
    x = dn["pore.coords"][:, 0]
    g = 10.0 ** np.random.default_rng(2).uniform(-15, -12, dn.Nt)
    linear_case("g12_delaunay_600", dn, g,
                value_bcs=[(np.flatnonzero(x < 0.1), 1.0, "add"),
                           (np.flatnonzero(x > 0.9), 0.0, "add")],
                rate_pores={"in": np.flatnonzero(x < 0.1), "out": np.flatnonzero(x > 0.9)})
    save("g12_delaunay_600_coords", coords=np.array(dn["pore.coords"]))


def g13_transient():
    # tests/unit/algorithms/TransientFickianDiffusionTest.py:8-36 (volume, g, BCs as there)
    from openpnm.models.physics import source_terms
    pn = op.network.Cubic(shape=[4, 3, 1], spacing=1.0)
    pn["pore.volume"] = 1e-14
    ph = op.phase.Phase(network=pn)
    ph["throat.diffusive_conductance"] = 1e-15
    alg = op.algorithms.TransientFickianDiffusion(network=pn, phase=ph)
    alg.settings._update({"quantity": "pore.concentration",
                          "conductance": "throat.diffusive_conductance"})
    alg.set_value_BC(pores=pn.pores("right"), values=1)
    alg.set_value_BC(pores=pn.pores("left"), values=0)
    alg.run(x0=0, tspan=(0, 10))
    assert np.isclose(alg.x.mean(), 0.40803, rtol=1e-5)
    x10 = np.array(alg.x)
    alg.run(x0=0, tspan=(0, 200))
    assert np.isclose(alg.x.mean(), 0.5, rtol=1e-5)
    x200 = np.array(alg.x)
    save("g13_transient_4x3x1", conns=pn["throat.conns"].astype(np.int64),
         bc_value=np.array(alg["pore.bc.value"], dtype=float), x10=x10, x200=x200)
    META["g13_transient_4x3x1"] = dict(x10_mean=float(x10.mean()), x200_mean=float(x200.mean()))
    # heterogeneous 8x7x6 with a power-law sink, saved at given times
    pn = op.network.Cubic(shape=[8, 7, 6], spacing=1e-4)
    rng = np.random.default_rng(7)
    pn["pore.volume"] = rng.uniform(0.5, 2.0, pn.Np) * 1e-13
    ph = op.phase.Phase(network=pn)
    ph["throat.diffusive_conductance"] = 10.0 ** rng.uniform(-15, -13, pn.Nt)
    ph["pore.A1"], ph["pore.A2"], ph["pore.A3"] = -2e-14, 1.5, 0.0
    ph.add_model(propname="pore.rxn", model=source_terms.power_law, A1="pore.A1", A2="pore.A2",
                 A3="pore.A3", X="pore.concentration", regen_mode="deferred")
    alg = op.algorithms.TransientFickianDiffusion(network=pn, phase=ph)
    alg.settings._update({"quantity": "pore.concentration",
                          "conductance": "throat.diffusive_conductance"})
    alg.set_value_BC(pores=pn.pores("left"), values=1.0)
    src = np.setdiff1d(pn.Ps, pn.pores("left"))
    alg.set_source(pores=src, propname="pore.rxn")
    saveat = np.array([0.0, 0.5, 1.0, 2.5, 5.0])
    alg.run(x0=0.2, tspan=(0, 5.0), saveat=saveat,
            integrator=op.integrators.ScipyRK45(atol=1e-8, rtol=1e-8))
    sol = alg.soln["pore.concentration"]
    t_ref, Y_ref = np.array(sol.t), np.array(sol)
    mask = np.zeros(pn.Np, dtype=bool)
    mask[src] = True
    gv = np.array(ph["throat.diffusive_conductance"], dtype=float)
    t_o, Y_o = oracle.transient_run(pn["throat.conns"], gv, pn.Np, np.array(pn["pore.volume"]),
                                    0.2, (0, 5.0), saveat=saveat,
                                    bc_value=np.array(alg["pore.bc.value"], dtype=float),
                                    sources=[(mask, "power_law",
                                              {"A1": -2e-14, "A2": 1.5, "A3": 0.0})],
                                    rtol=1e-8, atol=1e-8)
    assert np.array_equal(t_o, t_ref)
    assert np.allclose(Y_o, Y_ref, rtol=1e-10, atol=0), np.max(np.abs(Y_o - Y_ref))
    save("g13_transient_8x7x6_power_law", conns=pn["throat.conns"].astype(np.int64), g=gv,
         volume=np.array(pn["pore.volume"]), bc_value=np.array(alg["pore.bc.value"], dtype=float),
         src_pores=src.astype(np.int64), t=t_ref, Y=Y_ref)
    META["g13_transient_8x7x6_power_law"] = dict(
        x0=0.2, tspan=[0.0, 5.0], rtol=1e-8, atol=1e-8,
        source=dict(kind="power_law", params={"A1": -2e-14, "A2": 1.5, "A3": 0.0}),
        y_end_mean=float(Y_ref[:, -1].mean()))
    print("g13", META["g13_transient_4x3x1"], META["g13_transient_8x7x6_power_law"])


def g14_advection_diffusion():
    # tests/unit/algorithms/AdvectionDiffusionTest.py:10-115 (4x3x1, powerlaw / upwind)
    mod = op.models.physics.ad_dif_conductance.ad_dif

    def make(shape, gh, gd, schemes, outflow_label=None, seed=None):
        pn = op.network.Cubic(shape=shape, spacing=1.0)
        pn["throat.length"] = 1.0 if seed is None else np.random.default_rng(seed).uniform(
            0.5, 1.5, pn.Nt)
        ph = op.phase.Phase(network=pn)
        ph["throat.diffusive_conductance"] = gd(pn) if callable(gd) else gd
        ph["throat.hydraulic_conductance"] = gh(pn) if callable(gh) else gh
        sf = op.algorithms.StokesFlow(network=pn, phase=ph)
        sf.set_value_BC(pores=pn.pores("right"), values=1)
        sf.set_value_BC(pores=pn.pores("left"), values=0)
        sf.run(solver=op.solvers.ScipySpsolve())
        ph.update(sf.soln)
        out = {}
        for sch in schemes:
            ph.add_model(propname=f"throat.ad_{sch}", model=mod, s_scheme=sch)
            ph.regenerate_models()
            ad = op.algorithms.AdvectionDiffusion(network=pn, phase=ph)
            ad.settings["conductance"] = f"throat.ad_{sch}"
            ad.set_value_BC(pores=pn.pores("right"), values=2)
            if outflow_label is None:
                ad.set_value_BC(pores=pn.pores("left"), values=0)
            else:
                ad.set_outflow_BC(pores=pn.pores(outflow_label))
                # a second Dirichlet patch so the field is not trivially uniform
                taken = np.union1d(pn.pores("right"), pn.pores(outflow_label))
                ad.set_value_BC(pores=np.setdiff1d(pn.pores("top"), taken)[::2], values=0.3)
            ad.run(solver=op.solvers.ScipySpsolve())
            g2 = np.array(ph[f"throat.ad_{sch}"], dtype=float)
            bcv = np.array(ad["pore.bc.value"], dtype=float)
            bco = np.array(ad["pore.bc.outflow"], dtype=float)
            # oracle cross-check: assembly bit-exact, outflow values, solution
            if outflow_label is not None:
                ov = oracle.outflow_values(pn["throat.conns"], ph["throat.hydraulic_conductance"],
                                           ph["pore.pressure"], pn.pores(outflow_label), pn.Np)
                assert np.allclose(ov[np.isfinite(ov)], bco[np.isfinite(bco)], rtol=1e-12)
            res = oracle.picard_run(pn["throat.conns"], g2, pn.Np, bc_value=bcv,
                                    bc_rate=np.array(ad["pore.bc.rate"], dtype=float),
                                    bc_outflow=bco)
            Ao, Ar = oracle.to_csr(res["A"], pn.Np), csr_of(ad)
            assert np.array_equal(Ao.indptr, Ar.indptr) and np.array_equal(Ao.indices, Ar.indices)
            assert np.allclose(Ao.data, Ar.data, rtol=1e-13, atol=0)
            assert np.allclose(res["x"], ad.x, rtol=1e-9, atol=1e-12)
            assert res["num_iter"] == ad.soln.num_iter
            out[sch] = dict(g2=g2, bc_value=bcv, bc_outflow=bco, x=np.array(ad.x),
                            indptr=Ar.indptr.astype(np.int32), indices=Ar.indices.astype(np.int32),
                            data=Ar.data, b=np.array(ad.b))
        return pn, ph, out

    pn, ph, out = make([4, 3, 1], 1e-15, 1e-15, ["powerlaw", "upwind"])
    assert np.allclose(out["powerlaw"]["x"][3:9],
                       [0.89653] * 3 + [1.53924] * 3, rtol=1e-5)
    assert np.allclose(out["upwind"]["x"][3:9], [0.86486] * 3 + [1.51351] * 3, rtol=1e-5)
    for sch, d in out.items():
        save(f"g14_ad_4x3x1_{sch}", conns=pn["throat.conns"].astype(np.int64), **d)
    # heterogeneous 3-D case with an outflow boundary
    pn, ph, out = make([9, 7, 5],
                       lambda pn: 10.0 ** np.random.default_rng(1).uniform(-15, -14, pn.Nt),
                       lambda pn: 10.0 ** np.random.default_rng(2).uniform(-15.5, -14.5, pn.Nt),
                       ["powerlaw", "hybrid"], outflow_label="left", seed=3)
    for sch, d in out.items():
        save(f"g14_ad_9x7x5_outflow_{sch}", conns=pn["throat.conns"].astype(np.int64),
             gh=np.array(ph["throat.hydraulic_conductance"]),
             pressure=np.array(ph["pore.pressure"]), left=pn.pores("left").astype(np.int64), **d)
    META["g14_ad"] = dict(cases=["g14_ad_4x3x1_powerlaw", "g14_ad_4x3x1_upwind",
                                 "g14_ad_9x7x5_outflow_powerlaw", "g14_ad_9x7x5_outflow_hybrid"])
    print("g14 ok", {k: float(v["x"].mean()) for k, v in out.items()})


def g15_health_trim():
    # check_network_health (openpnm/utils/_health.py:40-98) and topotools.trim
    # (openpnm/topotools/_topotools.py:157-229) on a Delaunay network whose long throats
    # were cut (isolated pores + small clusters), with two looped throats added
    pts = np.random.default_rng(8).random((1500, 3))
    pts = pts[np.argsort(pts[:, 0], kind="stable")]
    conns = oracle.delaunay_network(pts)
    L = np.linalg.norm(pts[conns[:, 0]] - pts[conns[:, 1]], axis=1)
    conns = np.vstack((conns[L < 0.11], [[7, 7], [900, 900]]))
    pn = op.network.Network(coords=pts, conns=conns)
    h = op.utils.check_network_health(pn)
    size = np.array(op.models.network.cluster_size(pn))
    ref = oracle.network_health(conns, pts.shape[0])
    assert np.flatnonzero(ref["isolated"]).tolist() == list(h["isolated_pores"])
    assert ref["disconnected"].tolist() == list(h["disconnected_pores"])
    assert ref["looped"].tolist() == list(h["looped_throats"])
    assert np.array_equal(ref["cluster_size"], size)
    assert len(h["isolated_pores"]) > 0 and ref["n_clusters"] > 3, "case does not bite"
    out = dict(coords=pts, conns=conns, cluster_size=size,
               isolated=np.array(h["isolated_pores"], dtype=np.int64),
               disconnected=np.array(h["disconnected_pores"], dtype=np.int64),
               looped=np.array(h["looped_throats"], dtype=np.int64))
    # trim 1: the suggested fix (everything outside the largest cluster); trim 2: some
    # throats of what is left; trim 3: pores and throats together
    op.topotools.trim(network=pn, pores=h["disconnected_pores"])
    new, Pk, Tk = oracle.trim(conns, pts.shape[0], ref["disconnected"])
    assert np.array_equal(new, pn["throat.conns"]) and np.allclose(pts[Pk], pn["pore.coords"])
    out.update(trim1_conns=np.array(pn["throat.conns"]), trim1_coords=np.array(pn["pore.coords"]))
    Ts = np.arange(3, pn.Nt, 11)
    c1, n1 = np.array(pn["throat.conns"]), pn.Np
    op.topotools.trim(network=pn, throats=Ts)
    new, Pk, Tk = oracle.trim(c1, n1, throats=Ts)
    assert np.array_equal(new, pn["throat.conns"]) and Pk.size == pn.Np
    out.update(trim2_throats=Ts, trim2_conns=np.array(pn["throat.conns"]))
    Ps, Ts3 = np.arange(5, pn.Np, 13), np.arange(0, pn.Nt, 29)
    c2, n2, xyz2 = np.array(pn["throat.conns"]), pn.Np, np.array(pn["pore.coords"])
    op.topotools.trim(network=pn, pores=Ps, throats=Ts3)
    new, Pk, Tk = oracle.trim(c2, n2, pores=Ps, throats=Ts3)
    assert np.array_equal(new, pn["throat.conns"]) and np.allclose(xyz2[Pk], pn["pore.coords"])
    out.update(trim3_pores=Ps, trim3_throats=Ts3, trim3_conns=np.array(pn["throat.conns"]),
               trim3_coords=np.array(pn["pore.coords"]))
    save("g15_health_trim_1500", **out)
    META["g15_health_trim_1500"] = dict(
        n_clusters=int(ref["n_clusters"]), n_isolated=len(h["isolated_pores"]),
        n_disconnected=len(h["disconnected_pores"]), Np_after=[int(n1), int(n2), int(pn.Np)],
        Nt_after=[int(c1.shape[0]), int(c2.shape[0]), int(pn.Nt)])
    print("g15", META["g15_health_trim_1500"])


def g16_cubic_connectivity():
    # openpnm/_skgraph/generators/_cubic.py:40-67: throat numbering of the 6 / 14 / 18 / 26
    # neighbour lattices, and a StokesFlow solve on each (rows have 7 / 9 / 13 distinct
    # super-diagonals: the DIA-sym operator's widest cases)
    shape = [5, 4, 6]
    out, meta = {}, {}
    for c in (6, 14, 18, 26):
        pn = op.network.Cubic(shape=shape, spacing=1e-4, connectivity=c)
        conns = np.array(pn["throat.conns"])
        g = 10.0 ** np.random.default_rng(c).uniform(-15, -12, pn.Nt)
        ph = op.phase.Phase(network=pn)
        ph["throat.hydraulic_conductance"] = g
        sf = op.algorithms.StokesFlow(network=pn, phase=ph)
        sf.set_value_BC(pores=pn.pores("left"), values=2e5)
        sf.set_value_BC(pores=pn.pores("right"), values=1e5)
        sf.run(solver=op.solvers.ScipySpsolve())
        bcv = np.array(sf["pore.bc.value"], dtype=float)
        check_oracle_linear(f"g16_c{c}", pn, g, bcv, None, sf)
        A = csr_of(sf)
        offs = np.unique(A.indices - np.repeat(np.arange(pn.Np), np.diff(A.indptr)))
        out.update({f"conns_{c}": conns, f"g_{c}": g, f"x_{c}": np.array(sf.x),
                    f"bc_value_{c}": bcv,
                    f"rate_left_{c}": sf.rate(pores=pn.pores("left")),
                    f"rate_right_{c}": sf.rate(pores=pn.pores("right"))})
        meta[str(c)] = dict(Nt=int(pn.Nt), super_diagonals=int((offs > 0).sum()))
        op.Workspace().clear()
    out["shape"] = np.array(shape)
    out["left"] = np.array(op.network.Cubic(shape=shape).pores("left"))
    out["right"] = np.array(op.network.Cubic(shape=shape).pores("right"))
    save("g16_cubic_connectivity_5x4x6", **out)
    META["g16_cubic_connectivity_5x4x6"] = meta
    print("g16", meta)


def _oracle_chain(kind, conns, poly, mode, sf, throat_fixed=None):
    def g_of_x(x):
        pv = oracle.polynomial(poly, x)
        tv = throat_fixed if throat_fixed is not None else \
            oracle.from_neighbor_pores(conns, pv, mode)
        return oracle.conduit_conductance(kind, conns, pv, tv, sf)
    return g_of_x


def g17_variable_conductance():
    # the conductance follows the quantity through
    # misc.polynomial (openpnm/models/misc/_simple_equations.py:88-117) ->
    # misc.from_neighbor_pores (_neighbor_lookups.py:82-138) -> generic_diffusive /
    # generic_hydraulic (physics/_utils.py:46-61, hydraulic_conductance/_funcs.py:37-53);
    # ReactiveTransport re-assembles every Picard iteration (_transport.py:107-108)
    mods = op.models
    cases = [
        ("diffusive_mean_sf3", "poisson", [1e-9, 0.0, 5e-11], "mean", 3, False),
        ("diffusive_min_sf1", "poisson", [2e-9, 1e-10], "min", 1, False),
        ("hydraulic_max_sf3", "hydraulic", [1e-3, 2e-5, 1e-6], "max", 3, False),
        ("hydraulic_fixed_throat_sf3", "hydraulic", [1e-3, 0.0, 3e-6], None, 3, True),
        ("hydraulic_mean_sf1", "hydraulic", [1e-3, 4e-5], "mean", 1, False),
    ]
    meta = {}
    for name, kind, poly, mode, sfc, fixed in cases:
        rng = np.random.default_rng(4)
        pn = op.network.Cubic(shape=[6, 5, 4])
        conns = np.array(pn["throat.conns"])
        sf = rng.uniform(0.5, 2.0, (pn.Nt, 3) if sfc == 3 else pn.Nt) * 1e-3
        ph = op.phase.Phase(network=pn)
        q = "pore.concentration" if kind == "poisson" else "pore.pressure"
        ph[q] = 0.0
        if kind == "poisson":
            pkey, tkey, gkey, skey = ("pore.diffusivity", "throat.diffusivity",
                                      "throat.diffusive_conductance",
                                      "throat.diffusive_size_factors")
            gmodel = mods.physics.diffusive_conductance.generic_diffusive
            gargs = dict(pore_diffusivity=pkey, throat_diffusivity=tkey, size_factors=skey)
        else:
            pkey, tkey, gkey, skey = ("pore.viscosity", "throat.viscosity",
                                      "throat.hydraulic_conductance",
                                      "throat.hydraulic_size_factors")
            gmodel = mods.physics.hydraulic_conductance.generic_hydraulic
            gargs = dict(pore_viscosity=pkey, throat_viscosity=tkey, size_factors=skey)
        pn[skey] = sf
        ph.add_model(propname=pkey, model=mods.misc.polynomial, a=poly, prop=q)
        tfix = None
        if fixed:
            tfix = rng.uniform(1.0, 2.0, pn.Nt) * poly[0]
            ph[tkey] = tfix
        else:
            ph.add_model(propname=tkey, model=mods.misc.from_neighbor_pores, prop=pkey,
                         mode=mode)
        ph.add_model(propname=gkey, model=gmodel, **gargs)
        ph.regenerate_models()
        alg = op.algorithms.ReactiveTransport(network=pn, phase=ph)
        alg.settings._update({"conductance": gkey, "quantity": q})
        alg.set_value_BC(pores=pn.pores("front"), values=10.0)
        alg.set_value_BC(pores=pn.pores("back"), values=0.0)
        alg.run(solver=op.solvers.ScipySpsolve())
        assert alg.soln.is_converged and alg.soln.num_iter > 2, name
        bcv = np.array(alg["pore.bc.value"], dtype=float)
        o = oracle.picard_run(conns, _oracle_chain(kind, conns, poly, mode, sf, tfix),
                              pn.Np, bc_value=bcv)
        assert o["num_iter"] == alg.soln.num_iter, (name, o["num_iter"], alg.soln.num_iter)
        assert np.allclose(o["x"], alg.x, rtol=1e-10, atol=1e-13), name
        gfin = np.array(ph[gkey], dtype=float)
        assert np.allclose(_oracle_chain(kind, conns, poly, mode, sf, tfix)(alg.x), gfin,
                           rtol=1e-13, atol=0), name
        save("g17_variable_g_" + name, conns=conns, size_factors=sf, poly=np.array(poly),
             throat_fixed=np.zeros(0) if tfix is None else tfix, bc_value=bcv,
             x=np.array(alg.x), g_final=gfin, pore_prop_final=np.array(ph[pkey]),
             rate_front=alg.rate(pores=pn.pores("front")),
             rate_back=alg.rate(pores=pn.pores("back")))
        meta[name] = dict(kind=kind, mode=mode, num_iter=int(alg.soln.num_iter),
                          quantity=q, shape=[6, 5, 4])
        op.Workspace().clear()
    META["g17_variable_g"] = meta
    print("g17", meta)


def g8_big():
    # config 1 scale: Cubic 50^3, SuperLU (~3 min)
    n = 50
    pn = op.network.Cubic(shape=[n, n, n], spacing=1e-4)
    g = 10.0 ** np.random.default_rng(0).uniform(-15, -12, pn.Nt)
    alg = linear_case("g8_cubic50_hetero", pn, g,
                      value_bcs=[(pn.pores("left"), 2e5, "add"), (pn.pores("right"), 1e5, "add")],
                      rate_pores={"left": pn.pores("left"), "right": pn.pores("right")},
                      store_full=False)
    del alg


def g18_config1():
    # BASELINE config 1 / SURVEY 8d, literally: Cubic 50^3, spheres_and_cylinders geometry with
    # np.random.seed(0), viscosity 1e-3, physics.basic -> generic_hydraulic, left 2e5 / right 1e5,
    # ScipySpsolve (SuperLU, ~3 min).  The conductances come out of the reference's geometry
    # models, so they are stored (float64) with the solution.
    n = 50
    np.random.seed(0)
    pn = op.network.Cubic(shape=[n, n, n], spacing=1e-4)
    pn.add_model_collection(op.models.collections.geometry.spheres_and_cylinders)
    pn.regenerate_models()
    ph = op.phase.Phase(network=pn)
    ph["pore.viscosity"] = 1e-3
    ph.add_model_collection(op.models.collections.physics.basic)
    ph.regenerate_models()
    g = np.array(ph["throat.hydraulic_conductance"], dtype=float)
    sf = op.algorithms.StokesFlow(network=pn, phase=ph)
    sf.set_value_BC(pores=pn.pores("left"), values=2e5)
    sf.set_value_BC(pores=pn.pores("right"), values=1e5)
    t0 = time.time()
    sf.run(solver=op.solvers.ScipySpsolve())
    dt = time.time() - t0
    x = np.array(sf.x)
    bcv = np.array(sf["pore.bc.value"], dtype=float)
    A = csr_of(sf)
    # oracle: same system bit for bit in structure, same b; the reference's x solves it
    pure = oracle.build_A(pn["throat.conns"], g, pn.Np)
    Ao, b, f = oracle.apply_bcs(pure, pn.Np, bcv, None)
    Ao = oracle.to_csr(Ao, pn.Np)
    assert np.array_equal(Ao.indptr, A.indptr) and np.array_equal(Ao.indices, A.indices)
    assert np.allclose(Ao.data, A.data, rtol=1e-13, atol=0) and np.allclose(b, sf.b, rtol=1e-12)
    relres = np.linalg.norm(Ao @ x - b) / np.linalg.norm(b)
    rl, rr = sf.rate(pores=pn.pores("left"))[0], sf.rate(pores=pn.pores("right"))[0]
    assert np.isclose(oracle.rate(pn["throat.conns"], g, x, pores=pn.pores("left"))[0], rl,
                      rtol=1e-10)
    save("g18_config1_cubic50_spheres_and_cylinders", g=g, x=x)
    META["g18_config1_cubic50_spheres_and_cylinders"] = dict(
        Np=int(pn.Np), Nt=int(pn.Nt), nnz=int(A.nnz), f=float(f), g_min=float(g.min()),
        g_max=float(g.max()), g_mean=float(g.mean()), rate_left=float(rl), rate_right=float(rr),
        relres_of_reference_x=float(relres), ref_seconds=round(dt, 1))
    print("g18", META["g18_config1_cubic50_spheres_and_cylinders"])


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    meta_path = os.path.join(HERE, "golden.json")
    if os.path.exists(meta_path):
        META.update(json.load(open(meta_path)))
    cases = [g1_solvers, g2_stokes, g3_reactive, g4_csr, g5_conduit, g6_profiles,
             g7_hetero, g11_irregular, g13_transient, g14_advection_diffusion,
             g15_health_trim, g16_cubic_connectivity, g17_variable_conductance]
    if args.big:
        cases = [g8_big, g18_config1]
    for fn in cases:
        if args.only and args.only not in fn.__name__:
            continue
        fn()
        op.Workspace().clear()
    import scipy
    META["_versions"] = dict(numpy=np.__version__, scipy=scipy.__version__,
                             openpnm=str(op.__version__))
    json.dump(META, open(meta_path, "w"), indent=1, sort_keys=True)
    print("wrote", meta_path)
