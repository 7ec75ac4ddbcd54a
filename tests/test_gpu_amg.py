"""GPU: aggregation-multigrid preconditioner (csrc/amg.cu).

What is checked: the hierarchy is a true Galerkin hierarchy (A_{l+1} = P^T A_l P
for the exported piecewise-constant P, to rounding), boundary rows stay off the
coarse grids, the AMG-preconditioned solve returns the same x as Jacobi-PCG and
as the reference's direct solve (north-star tolerance 1e-8 at CG rtol 1e-12), the
outer Picard loop walks the same iterates, the iteration count does not grow
with the lattice size, and repeated runs are bit-identical."""
import numpy as np
import pytest
import scipy.sparse as sprs

from oracle import oracle

pytestmark = pytest.mark.gpu


def lattice(b2, shape, seed=0):
    pn = b2.Cubic(list(shape), coords=False)
    g = 10.0 ** np.random.default_rng(seed).uniform(-15, -12, pn.Nt)
    bcv = np.full(pn.Np, np.nan)
    bcv[pn.pores("left")] = 2e5
    bcv[pn.pores("right")] = 1e5
    return pn, g, bcv


def level_csr(ctx, l, with_agg):
    ip, ix, dt, ag = ctx.amg_export_level(l, with_agg=with_agg)
    n = ip.numel() - 1
    return (ip.cpu().numpy(), ix.cpu().numpy(), dt.cpu().numpy(),
            ag.cpu().numpy() if ag is not None else None, n)


def test_hierarchy_is_galerkin(b2):
    pn, g, bcv = lattice(b2, (24, 20, 18))
    ctx = b2.B200PCG(precond="amg").ctx
    ctx.build_laplacian(pn.conns, g, pn.Np)
    ctx.apply_bcs(bcv, None)
    levels = ctx.amg_setup()
    assert len(levels) >= 2 and levels[0] == (pn.Np, ctx.csr_shape(1)[1])
    assert levels[-1][0] <= 2048
    nbc = int(np.isfinite(bcv).sum())
    for l in range(len(levels) - 1):
        ip, ix, dt, ag, n = level_csr(ctx, l, True)
        A = sprs.csr_matrix((dt, ix, ip), shape=(n, n))
        ipc, ixc, dtc, _, nc = level_csr(ctx, l + 1, False)
        Ac = sprs.csr_matrix((dtc, ixc, ipc), shape=(nc, nc))
        assert levels[l + 1] == (nc, Ac.nnz)
        # every coarse column appears once per row (stored in first-arrival order)
        rows = np.repeat(np.arange(nc), np.diff(ipc))
        assert np.unique(rows.astype(np.int64) * nc + ixc).size == ixc.size
        Ac.sort_indices()
        keep = ag >= 0
        assert ag.max() == nc - 1 and set(np.unique(ag[keep])) == set(range(nc))
        if l == 0:
            # rows decoupled by _apply_BCs are the only ones left off the coarse grid
            assert (~keep).sum() == nbc and np.all(~keep[np.isfinite(bcv)])
        sizes = np.bincount(ag[keep])
        assert 2.0 <= sizes.mean() <= 40.0 and n / nc >= 1.5
        P = sprs.csr_matrix((np.ones(keep.sum()), (np.flatnonzero(keep), ag[keep])), shape=(n, nc))
        G = (P.T @ A @ P).tocsr()
        G.sort_indices()
        assert np.array_equal(G.indptr, Ac.indptr) and np.array_equal(G.indices, Ac.indices)
        np.testing.assert_allclose(Ac.data, G.data, rtol=1e-12, atol=0)
        # coarse operators stay symmetric M-matrices with zero-or-positive row sums
        assert abs(Ac - Ac.T).max() <= 1e-12 * abs(Ac).max()
        assert Ac.diagonal().min() > 0


@pytest.mark.parametrize("fmt", ["auto", "csr"])
def test_amg_solution_matches_jacobi_and_direct(fmt, b2):
    pn, g, bcv = lattice(b2, (20, 18, 16), seed=3)
    xs = {}
    for precond in ("jacobi", "amg"):
        sf = b2.StokesFlow(network=pn, phase=_phase(b2, pn, g))
        sf["pore.bc.value"] = bcv
        s = b2.B200PCG(tol=1e-12, precond=precond, fmt=fmt)
        sf.run(solver=s)
        assert sf.soln.is_converged
        xs[precond] = (np.array(sf.x), sf.stats["cg_iters"], sf.rate(pores=pn.pores("left"))[0])
    A = oracle.build_A(pn.conns, g, pn.Np)
    A, b, _ = oracle.apply_bcs(A, pn.Np, bcv, None)
    x_ref, _ = oracle.spsolve(oracle.to_csr(A, pn.Np), b)
    for k in xs:
        assert np.max(np.abs(xs[k][0] - x_ref) / np.abs(x_ref)) < 1e-8
    assert abs(xs["amg"][2] - xs["jacobi"][2]) <= 1e-8 * abs(xs["jacobi"][2])
    assert xs["amg"][1] < xs["jacobi"][1]


def _phase(b2, pn, g, key="throat.hydraulic_conductance"):
    ph = b2.Phase(pn)
    ph[key] = g
    return ph


def test_amg_on_golden_and_small_fallback(b2, golden, golden_meta):
    """cubic20 (8000 rows) goes through the hierarchy; the 60-pore SolversTest case is
    below the threshold and silently stays with Jacobi."""
    for name in ("g7_cubic20_hetero", "g1_solvers_3x4x5"):
        g = golden(name)
        Np = g["x"].size
        pn = b2.Network(conns=g["conns"], Np=Np)
        ph = b2.Phase(pn)
        ph["throat.conductance"] = g["g"]
        alg = b2.Transport(network=pn, phase=ph)
        alg.settings._update({"quantity": "pore.x", "conductance": "throat.conductance"})
        alg["pore.bc.value"] = g["bc_value"]
        alg["pore.bc.rate"] = g["bc_rate"]
        alg.run(solver=b2.B200PCG(tol=1e-12, precond="amg"))
        scale = np.max(np.abs(g["x"]))
        assert np.max(np.abs(alg.x - g["x"])) <= 1e-8 * scale


def test_amg_irregular_network(b2):
    pts = np.random.default_rng(7).random((6000, 3))
    order = np.argsort(pts[:, 0], kind="stable")
    pts = pts[order]
    conns = oracle.delaunay_network(pts)
    conns = conns["conns"] if isinstance(conns, dict) else conns
    g = 10.0 ** np.random.default_rng(8).uniform(-15, -12, conns.shape[0])
    bcv = np.full(6000, np.nan)
    bcv[pts[:, 0] < 0.05] = 1.0
    bcv[pts[:, 0] > 0.95] = 0.0
    pn = b2.Network(conns=conns, Np=6000)
    fd = b2.FickianDiffusion(network=pn, phase=_phase(b2, pn, g, "throat.diffusive_conductance"))
    fd["pore.bc.value"] = bcv
    fd.run(solver=b2.B200PCG(tol=1e-12, precond="amg"))
    A = oracle.build_A(conns, g, 6000)
    A, b, _ = oracle.apply_bcs(A, 6000, bcv, None)
    x_ref, _ = oracle.spsolve(oracle.to_csr(A, 6000), b)
    assert np.max(np.abs(fd.x - x_ref)) <= 1e-8


def test_amg_inside_the_picard_loop(b2):
    """Reactive transport: the hierarchy follows every re-linearisation of the
    source term (it moves the diagonal); the outer loop is the same."""
    pn = b2.Cubic([16, 16, 16], coords=False)
    g = 10.0 ** np.random.default_rng(2).uniform(-15, -12, pn.Nt)
    out = {}
    for precond in ("jacobi", "amg"):
        ph = _phase(b2, pn, g, "throat.diffusive_conductance")
        ph["pore.A1"], ph["pore.A2"], ph["pore.A3"] = -1e-13, 2.0, 0.0
        ph.add_model(propname="pore.rxn", model="power_law", A1="pore.A1", A2="pore.A2",
                     A3="pore.A3", X="pore.concentration")
        rt = b2.ReactiveTransport(network=pn, phase=ph)
        rt.settings._update({"quantity": "pore.concentration",
                             "conductance": "throat.diffusive_conductance"})
        rt.set_value_BC(pores=pn.pores("left"), values=1.0)
        rt.set_source(pores=np.setdiff1d(pn.Ps, pn.pores("left")), propname="pore.rxn")
        rt.run(solver=b2.B200PCG(tol=1e-13, precond=precond))
        assert rt.soln.is_converged
        out[precond] = (np.array(rt.x), rt.soln.num_iter, rt.soln.cg_iters)
    assert out["amg"][1] == out["jacobi"][1]
    assert np.max(np.abs(out["amg"][0] - out["jacobi"][0])) <= 1e-9 * np.max(np.abs(out["jacobi"][0]))
    assert out["amg"][2] < out["jacobi"][2] / 2


def test_amg_iterations_do_not_grow_and_runs_repeat(b2):
    its = {}
    for n in (32, 64):
        pn, g, bcv = lattice(b2, (n, n, n), seed=1)
        res = []
        for rep in range(2):
            ctx = b2.B200PCG(precond="amg").ctx
            ctx.build_laplacian(pn.conns, g, pn.Np)
            ctx.apply_bcs(bcv, None)
            x, it, rel, info = ctx.pcg(rtol=1e-10)
            assert info == 0 and rel <= 1e-10 * (1 + 1e-6)
            b = ctx.get_b()
            assert float((ctx.spmv(1, x) - b).norm() / b.norm()) <= 1.01e-10
            res.append((x.cpu().numpy().view(np.uint64), it))
            if n == 64:
                break
        its[n] = res[0][1]
        if len(res) == 2:
            assert res[0][1] == res[1][1] and np.array_equal(res[0][0], res[1][0])
    assert its[64] <= 2 * its[32]      # absolute counts: tests/test_gpu_zz_perf_bounds.py


def test_amg_with_dead_throats_and_production_source(b2):
    """Zero conductances (rows that lose neighbours, some pores isolated from the
    hierarchy) and a source with S1 > 0 (production: the diagonal shrinks)."""
    pn = b2.Cubic([18, 17, 16], coords=False)
    rng = np.random.default_rng(11)
    g = 10.0 ** rng.uniform(-15, -12, pn.Nt)
    g[rng.random(pn.Nt) < 0.15] = 0.0
    d = np.zeros(pn.Np)
    np.add.at(d, pn.conns[:, 0], g)
    np.add.at(d, pn.conns[:, 1], g)
    assert (d == 0).sum() >= 1                    # at least one pore lost all its throats
    A1 = np.where(d > 4e-16, 2e-16, -2e-16)       # production wherever the diagonal stays positive
    out = {}
    for precond in ("jacobi", "amg"):
        ph = _phase(b2, pn, g, "throat.diffusive_conductance")
        ph["pore.A1"], ph["pore.A2"] = A1, 1e-15
        ph.add_model(propname="pore.rxn", model="linear", A1="pore.A1", A2="pore.A2",
                     X="pore.concentration")
        rt = b2.ReactiveTransport(network=pn, phase=ph)
        rt.settings._update({"quantity": "pore.concentration",
                             "conductance": "throat.diffusive_conductance",
                             "validate_topology": False})
        rt.set_value_BC(pores=pn.pores("left"), values=1.0)
        rt.set_value_BC(pores=pn.pores("right"), values=0.0)
        interior = np.setdiff1d(pn.Ps, np.r_[pn.pores("left"), pn.pores("right")])
        rt.set_source(pores=interior, propname="pore.rxn")
        rt.run(solver=b2.B200PCG(tol=1e-12, precond=precond))
        assert rt.soln.is_converged
        out[precond] = np.array(rt.x)
    assert np.max(np.abs(out["amg"] - out["jacobi"])) <= 1e-8 * np.max(np.abs(out["jacobi"]))


def test_amg_copes_with_a_non_m_matrix(b2):
    """Symmetric positive definite, but half of the couplings are positive: the
    aggregation only sees the negative ones.  The call must still return the solution
    (by multigrid if it converges, by the Jacobi hand-over otherwise)."""
    n = 3000
    rng = np.random.default_rng(5)
    off = rng.uniform(0.5, 1.0, n - 1) * np.where(rng.random(n - 1) < 0.5, 1.0, -1.0)
    far = rng.uniform(0.5, 1.0, n - 37) * np.where(rng.random(n - 37) < 0.5, 1.0, -1.0)
    A = sprs.diags([off, off, far, far], [1, -1, 37, -37], shape=(n, n)).tocsr()
    A = A + sprs.diags(np.asarray(abs(A).sum(axis=1)).ravel() + 0.01)
    b = rng.standard_normal(n)
    x, info = b2.B200PCG(tol=1e-12, precond="amg").solve(A.tocoo(), b)
    assert info == 0
    assert np.linalg.norm(A @ x - b) <= 1e-11 * np.linalg.norm(b)


def test_amg_hands_over_to_jacobi(b2, monkeypatch):
    """With the multigrid iteration capped below what it needs (B200_AMG_MAXITER,
    a test knob read at every call) the same call finishes with Jacobi-PCG from the
    last iterate: converged, same answer, fewer Jacobi iterations than from zero."""
    pn, g, bcv = lattice(b2, (20, 18, 16), seed=4)
    ref = b2.B200PCG(precond="jacobi").ctx
    ref.build_laplacian(pn.conns, g, pn.Np)
    ref.apply_bcs(bcv, None)
    x_ref, it_ref, _, info = ref.pcg(rtol=1e-12)
    assert info == 0
    monkeypatch.setenv("B200_AMG_MAXITER", "6")
    ctx = b2.B200PCG(precond="amg").ctx
    ctx.build_laplacian(pn.conns, g, pn.Np)
    ctx.apply_bcs(bcv, None)
    x, it, rel, info = ctx.pcg(rtol=1e-12)
    assert info == 0 and rel <= 1e-12 * (1 + 1e-6)
    assert 40 < it < it_ref                         # Jacobi iterations after the hand-over
    assert float((x - x_ref).abs().max() / x_ref.abs().max()) < 1e-8
