"""GPU + the REAL reference: `openpnm` (unmodified, from /root/reference or the
offline install under baseline/_ref) driven with the B200 solver and with the
B200 algorithms.  Skipped when the reference is not importable."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_shim  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def op():
    if ref_shim.find_reference() is None:
        pytest.skip("reference OpenPNM not available")
    mod = ref_shim.import_reference()
    yield mod
    mod.Workspace().clear()


def hetero(nt, seed=0):
    return 10.0 ** np.random.default_rng(seed).uniform(-15, -12, nt)


def test_reference_stokesflow_with_b200pcg(op, b2):
    """alg.run(solver=B200PCG()) on the reference's own StokesFlow, against the
    reference's own ScipySpsolve answer (north star: x and rate within 1e-8)."""
    pn = op.network.Cubic(shape=[15, 15, 15], spacing=1e-4)
    ph = op.phase.Phase(network=pn)
    ph["throat.hydraulic_conductance"] = hetero(pn.Nt)
    sf = op.algorithms.StokesFlow(network=pn, phase=ph)
    sf.set_value_BC(pores=pn.pores("left"), values=2e5)
    sf.set_value_BC(pores=pn.pores("right"), values=1e5)
    sf.run(solver=op.solvers.ScipySpsolve())
    x_ref = np.array(sf.x)
    q_ref = sf.rate(pores=pn.pores("left"))[0]
    solver = b2.B200PCG(tol=1e-13)
    sf.run(solver=solver)                       # _transport.py:213 calls solver.solve(A=, b=, x0=)
    assert sf.soln.is_converged and sf.soln.num_iter == 1
    assert np.max(np.abs(sf.x - x_ref) / np.abs(x_ref)) < 1e-8
    assert abs(sf.rate(pores=pn.pores("left"))[0] - q_ref) <= 1e-8 * abs(q_ref)
    assert solver.info["fmt"] == "dia" and solver.info["exit_code"] == 0


def test_reference_stokesflow_with_b200_multigrid(op, b2):
    """`B200PCG(precond='amg')` in the seat of the reference's PyamgRugeStubenSolver
    (solvers/_pyamg.py:10-17): the reference assembles, the device builds the
    hierarchy from the COO it is handed and solves."""
    pn = op.network.Cubic(shape=[15, 15, 15], spacing=1e-4)
    ph = op.phase.Phase(network=pn)
    ph["throat.hydraulic_conductance"] = hetero(pn.Nt, 4)
    sf = op.algorithms.StokesFlow(network=pn, phase=ph)
    sf.set_value_BC(pores=pn.pores("left"), values=2e5)
    sf.set_rate_BC(pores=pn.pores("right"), rates=-1e-9)
    sf.run(solver=op.solvers.ScipySpsolve())
    x_ref = np.array(sf.x)
    solver = b2.B200PCG(tol=1e-13, precond="amg")
    sf.run(solver=solver)
    assert sf.soln.is_converged
    assert np.max(np.abs(sf.x - x_ref)) <= 1e-8 * np.max(np.abs(x_ref))
    assert solver.info["exit_code"] == 0
    # iteration *counts* are performance, not parity: they are bounded relative to
    # Jacobi-PCG in tests/test_gpu_zz_perf_bounds.py so that they can never stop -x
    # before the parity files have run


@pytest.mark.parametrize("conn,nd", [(14, 7), (18, 9), (26, 13)])
def test_reference_cubic_with_diagonal_neighbours(op, b2, conn, nd):
    """`Cubic(connectivity=14/18/26)` (openpnm/_skgraph/generators/_cubic.py:40-67): 7 / 9 / 13
    super-diagonals, all on the implicit-index DIA fast path (B200_MAX_DIAGS = 13)."""
    pn = op.network.Cubic(shape=[12, 11, 10], spacing=1e-4, connectivity=conn)
    ph = op.phase.Phase(network=pn)
    ph["throat.hydraulic_conductance"] = hetero(pn.Nt, conn)
    sf = op.algorithms.StokesFlow(network=pn, phase=ph)
    sf.set_value_BC(pores=pn.pores("left"), values=2e5)
    sf.set_value_BC(pores=pn.pores("right"), values=1e5)
    sf.run(solver=op.solvers.ScipySpsolve())
    x_ref = np.array(sf.x)
    for precond in ("jacobi", "amg"):
        solver = b2.B200PCG(tol=1e-13, precond=precond)
        sf.run(solver=solver)
        assert sf.soln.is_converged
        assert np.max(np.abs(sf.x - x_ref) / np.abs(x_ref)) < 1e-8
        assert solver.info["fmt"] == "dia" and solver.info["ndiags"] == nd
    # the same network through the B200 algorithm (device assembly), forced to the
    # general sliced-ELL operator: same answer
    alg = b2.StokesFlow(network=pn, phase=ph)
    alg.set_value_BC(pores=pn.pores("left"), values=2e5)
    alg.set_value_BC(pores=pn.pores("right"), values=1e5)
    alg.run(solver=b2.B200PCG(tol=1e-13, fmt="csr"))
    assert np.max(np.abs(alg.x - x_ref) / np.abs(x_ref)) < 1e-8


def test_default_solver_registration(op, b2):
    """op.Workspace().settings.default_solver = 'B200PCG' (utils/_workspace.py:44-56)."""
    b2.register_with_openpnm(op)
    ws = op.Workspace()
    old = ws.settings.default_solver
    try:
        ws.settings.default_solver = "B200PCG"
        pn = op.network.Cubic(shape=[9, 9, 9])
        ph = op.phase.Phase(network=pn)
        ph["throat.diffusive_conductance"] = 1.0
        fd = op.algorithms.FickianDiffusion(network=pn, phase=ph)
        fd.set_value_BC(pores=pn.pores("top"), values=1)
        fd.set_value_BC(pores=pn.pores("bottom"), values=0)
        fd.run()                                  # no solver argument: getattr(solvers, name)()
        np.testing.assert_allclose(fd.rate(pores=pn.pores("top"))[0], 10.125, rtol=1e-6)
    finally:
        ws.settings.default_solver = old


def test_reference_reactive_loop_with_b200pcg(op, b2):
    """The reference's own Picard loop (_reactive_transport.py:150-213) calling
    the B200 solver once per outer iteration: ReactiveTransportTest.py:69-77."""
    from openpnm.models.physics import source_terms
    pn = op.network.Cubic(shape=[4, 4, 4])
    ph = op.phase.Phase(network=pn)
    ph["throat.diffusive_conductance"] = 1e-15
    ph["pore.A"] = -1e-15
    ph["pore.k"] = 2
    ph.add_model(propname="pore.reaction", model=source_terms.standard_kinetics,
                 prefactor="pore.A", exponent="pore.k", X="pore.concentration",
                 regen_mode="deferred")
    alg = op.algorithms.ReactiveTransport(network=pn, phase=ph)
    alg.settings._update({"conductance": "throat.diffusive_conductance",
                          "quantity": "pore.concentration"})
    alg.set_source(pores=pn.pores("bottom"), propname="pore.reaction")
    alg.set_value_BC(pores=pn.pores("top"), values=1.0)
    alg.run(solver=b2.B200PCG(tol=1e-13))
    np.testing.assert_allclose(alg["pore.concentration"].mean(), 0.7171292729553325, rtol=1e-8)
    assert alg.soln.num_iter == 6 and alg.soln.is_converged


def test_b200_algorithms_accept_reference_objects(op, b2):
    """openpnm_b200.ReactiveTransport on real OpenPNM network/phase objects (device
    assembly + device Picard loop) against the reference's own run."""
    from openpnm.models.physics import source_terms
    pn = op.network.Cubic(shape=[12, 12, 12], spacing=1e-4)
    ph = op.phase.Phase(network=pn)
    ph["throat.diffusive_conductance"] = hetero(pn.Nt, 3)
    ph["pore.A1"] = -1e-13
    ph["pore.A2"] = 2.0
    ph["pore.A3"] = 0.0
    ph.add_model(propname="pore.rxn", model=source_terms.power_law, A1="pore.A1", A2="pore.A2",
                 A3="pore.A3", X="pore.concentration", regen_mode="deferred")
    src = np.setdiff1d(pn.Ps, pn.pores("left"))
    ref = op.algorithms.ReactiveTransport(network=pn, phase=ph)
    ref.settings._update({"conductance": "throat.diffusive_conductance",
                          "quantity": "pore.concentration"})
    ref.set_source(pores=src, propname="pore.rxn")
    ref.set_value_BC(pores=pn.pores("left"), values=1.0)
    ref.run(solver=op.solvers.ScipySpsolve())
    c_ref = np.array(ref["pore.concentration"])
    q_ref = ref.rate(pores=pn.pores("left"))[0]
    A_ref = ref.A.tocsr()
    A_ref.sort_indices()

    alg = b2.ReactiveTransport(network=pn, phase=ph)
    alg.settings._update({"conductance": "throat.diffusive_conductance",
                          "quantity": "pore.concentration"})
    alg.set_source(pores=src, propname="pore.rxn")
    alg.set_value_BC(pores=pn.pores("left"), values=1.0)
    alg.run(solver=b2.B200PCG(tol=1e-13))
    assert alg.soln.num_iter == ref.soln.num_iter and alg.soln.is_converged
    assert np.max(np.abs(alg.x - c_ref)) <= 1e-8 * np.max(np.abs(c_ref))
    assert abs(alg.rate(pores=pn.pores("left"))[0] - q_ref) <= 1e-8 * abs(q_ref)
    A = alg.A
    assert np.array_equal(A.indptr, A_ref.indptr) and np.array_equal(A.indices, A_ref.indices)
    np.testing.assert_allclose(A.data, A_ref.data, rtol=1e-7, atol=0)


def test_host_evaluated_models_on_reference_objects(op, b2):
    """A source model without a device kernel (exponential,
    source_terms/_funcs.py:172-223) and a conductance that depends on the
    concentration, both defined on a REAL OpenPNM phase: openpnm_b200 regenerates
    them through the b200_picard_cb hook exactly where the reference calls
    _update_iterative_props, so the two outer loops walk the same iterates."""
    from openpnm.models.physics import source_terms
    pn = op.network.Cubic(shape=[8, 8, 8], spacing=1e-4)
    ph = op.phase.Phase(network=pn)

    def D_var(phase, c="pore.concentration"):
        return 1e-9 * (1 + 0.5 * phase[c] ** 2)

    def g_var(phase, D="pore.diffusivity"):
        return 1e-6 * phase[D][phase.network.conns].mean(axis=1)

    ph["pore.concentration"] = 0.0
    ph.add_model(propname="pore.diffusivity", model=D_var)
    ph.add_model(propname="throat.diffusive_conductance", model=g_var)
    for k, v in dict(A1=-2e-16, A2=np.e, A3=1.5, A4=1.0, A5=0.0, A6=0.0).items():
        ph["pore." + k] = v            # r = A1 * A2**(A3 * x**A4 + A5) + A6
    ph.add_model(propname="pore.rxn", model=source_terms.exponential, X="pore.concentration",
                 regen_mode="deferred", **{k: "pore." + k for k in
                                           ("A1", "A2", "A3", "A4", "A5", "A6")})
    src = pn.pores("right")

    def setup(cls):
        ph["pore.concentration"] = 0.0
        ph.regenerate_models()
        alg = cls(network=pn, phase=ph)
        alg.settings._update({"conductance": "throat.diffusive_conductance",
                              "quantity": "pore.concentration"})
        alg.set_source(pores=src, propname="pore.rxn")
        alg.set_value_BC(pores=pn.pores("left"), values=1.0)
        return alg

    ref = setup(op.algorithms.ReactiveTransport)
    ref.run(solver=op.solvers.ScipySpsolve())
    c_ref, it_ref = np.array(ref["pore.concentration"]), ref.soln.num_iter
    q_ref = ref.rate(pores=pn.pores("left"))[0]
    assert ref.soln.is_converged and it_ref > 3

    alg = setup(b2.ReactiveTransport)
    assert set(alg.iterative_props) == set(ref.iterative_props)
    alg.run(solver=b2.B200PCG(tol=1e-13))
    assert alg.soln.is_converged and alg.soln.num_iter == it_ref
    assert np.max(np.abs(alg.x - c_ref)) <= 1e-8 * np.max(np.abs(c_ref))
    assert abs(alg.rate(pores=pn.pores("left"))[0] - q_ref) <= 1e-8 * abs(q_ref)


def test_reference_advection_diffusion_with_b200(op, b2):
    """The reference's own AdvectionDiffusion (AdvectionDiffusionTest.py:10-93) solved
    with the B200 solver object: non-symmetric A -> BiCGStab behind the same API."""
    pn = op.network.Cubic(shape=[4, 3, 1], spacing=1.0)
    pn["throat.length"] = 1.0
    ph = op.phase.Phase(network=pn)
    ph["throat.diffusive_conductance"] = 1e-15
    ph["throat.hydraulic_conductance"] = 1e-15
    sf = op.algorithms.StokesFlow(network=pn, phase=ph)
    sf.set_value_BC(pores=pn.pores("right"), values=1)
    sf.set_value_BC(pores=pn.pores("left"), values=0)
    sf.run(solver=b2.B200PCG(tol=1e-13))
    ph.update(sf.soln)
    ad = op.algorithms.AdvectionDiffusion(network=pn, phase=ph)
    ad.set_value_BC(pores=pn.pores("right"), values=2)
    ad.set_value_BC(pores=pn.pores("left"), values=0)
    ph.add_model(propname="throat.ad_dif_conductance",
                 model=op.models.physics.ad_dif_conductance.ad_dif, s_scheme="powerlaw")
    ad.run(solver=b2.B200PCG(tol=1e-13))
    x = [0., 0., 0., 0.89653, 0.89653, 0.89653, 1.53924, 1.53924, 1.53924, 2., 2., 2.]
    np.testing.assert_allclose(ad["pore.concentration"], x, rtol=1e-5, atol=1e-12)
    assert ad.soln.is_converged


def test_integration_md_stub_runs_verbatim(op, b2):
    """The ctypes binding printed in INTEGRATION.md section 1, executed as written
    (only the library path is made absolute) inside the real OpenPNM."""
    import re
    import types
    from openpnm_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    md = open(os.path.join(root, "INTEGRATION.md")).read()
    code = re.search(r"```python\n(# openpnm/solvers/_b200\.py\n.*?)```", md, re.S).group(1)
    code = code.replace("'libb200pnm.so'", repr(_lib.LIB_PATH))
    mod = types.ModuleType("openpnm.solvers._b200")
    exec(compile(code, "INTEGRATION.md", "exec"), mod.__dict__)
    op.solvers.B200PCG_stub = mod.B200PCG
    pn = op.network.Cubic(shape=[9, 9, 9])
    ph = op.phase.Phase(network=pn)
    ph["throat.hydraulic_conductance"] = 1.0
    sf = op.algorithms.StokesFlow(network=pn, phase=ph)
    sf.set_value_BC(pores=pn.pores("top"), values=101325)
    sf.set_value_BC(pores=pn.pores("bottom"), values=0)
    sf.run(solver=mod.B200PCG(tol=1e-12))
    np.testing.assert_allclose(sf.rate(pores=pn.pores("top"))[0], 1025915.625, rtol=1e-9)
    assert sf.soln.is_converged
