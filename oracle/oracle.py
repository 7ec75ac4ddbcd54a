"""CPU oracle for the OpenPNM transport-solve hot path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy/scipy *restatement* of the reference algorithm.  It is the
checker for the CUDA path, never the thing shipped or measured as the product:
only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import it.  `openpnm_b200/` must never import it.

Parity status: PINNED.  `tests/golden/make_golden.py` runs the real reference
(/root/reference, through `tests/ref_shim.py`) in the build container, checks
every function below against it, and commits the resulting vectors under
`tests/golden/`; `tests/test_oracle_golden.py` re-checks the oracle against
those vectors and against the known-answer constants of the reference's own
tests (SolversTest, SubclassedTransportTest, GenericTransportTest,
ReactiveTransportTest, Flow_in_straight_conduit).  Round 2 added the reference's
check_network_health / cluster_size / topotools.trim (g15), Cubic connectivity
6/14/18/26 with its SuperLU solves (g16) and ReactiveTransport with a conductance
that follows the quantity through misc.polynomial -> misc.from_neighbor_pores ->
generic_diffusive / generic_hydraulic (g17: same Picard count, x to 1e-10).

Third-party arithmetic the reference delegates to (not vendored, not pinned by
the reference: setup.py:46-64): scipy (this image: 1.18.1) --
`scipy.sparse.csgraph.laplacian`, `coo_matrix.tocsr`, `linalg.spsolve`
(SuperLU), `linalg.cg`, `optimize._nonlin.TerminationCondition`.  scipy is
present on both the build container and the GPU box, so the oracle calls
`spsolve`/`tocsr` directly (exactly what openpnm/solvers/_scipy.py:13-15 does) and
restates the rest in numpy so that each step can be compared separately.

All citations are file:line under /root/reference/.
"""
import numpy as np
import scipy.sparse as sprs
import scipy.sparse.linalg as spla

__all__ = [
    "cubic_network", "adjacency_coo", "laplacian_coo", "build_A", "apply_bcs",
    "apply_sources", "to_csr", "source_term", "picard_run", "rate",
    "jacobi_pcg", "spsolve", "transport_run", "delaunay_network", "transient_run",
    "apply_outflow", "outflow_values", "polynomial", "from_neighbor_pores",
    "conduit_conductance", "network_health", "trim",
]


# --------------------------------------------------------------------------
# network generation (input side; openpnm/_skgraph/generators/_cubic.py:25-82)
# --------------------------------------------------------------------------
def cubic_network(shape, spacing=1.0):
    """Cubic lattice exactly as `op.network.Cubic(shape, spacing)` numbers it.

    Pore index = x*Ny*Nz + y*Nz + z (z fastest; _cubic.py:33-36); throats are
    the z-neighbour pairs, then y-, then x- (_cubic.py:40-42, 69-73); conns are
    int64, [low, high].  Face labels follow
    _skgraph/generators/tools/_funcs.py:216-250: left/right = x min/max,
    front/back = y min/max, bottom/top = z min/max.
    """
    shape = np.array(shape, ndmin=1, dtype=int)
    if shape.size < 3:
        shape = np.concatenate((shape, np.ones(3 - shape.size, dtype=int)))
    nx, ny, nz = (int(s) for s in shape)
    idx = np.arange(nx * ny * nz, dtype=np.int64).reshape(nx, ny, nz)
    tails = [idx[:, :, :-1].ravel(), idx[:, :-1, :].ravel(), idx[:-1, :, :].ravel()]
    heads = [idx[:, :, 1:].ravel(), idx[:, 1:, :].ravel(), idx[1:, :, :].ravel()]
    conns = np.vstack((np.concatenate(tails), np.concatenate(heads))).T
    conns = np.ascontiguousarray(conns, dtype=np.int64)
    sp = np.ones(3) * np.asarray(spacing, dtype=float)
    ii, jj, kk = np.unravel_index(np.arange(nx * ny * nz), (nx, ny, nz))
    coords = (np.vstack((ii, jj, kk)).T + 0.5) * sp
    labels = {
        "left": np.flatnonzero(ii == 0), "right": np.flatnonzero(ii == nx - 1),
        "front": np.flatnonzero(jj == 0), "back": np.flatnonzero(jj == ny - 1),
        "bottom": np.flatnonzero(kk == 0), "top": np.flatnonzero(kk == nz - 1),
    }
    return {"shape": (nx, ny, nz), "Np": nx * ny * nz, "Nt": conns.shape[0],
            "conns": conns, "coords": coords, "labels": labels}


def delaunay_network(points):
    """Delaunay edges of `points` as upper-triangular, de-duplicated conns.

    Same edge set as `tri_to_am` (_skgraph/tools/_funcs.py:629-669) followed by
    the triu + coo conversion in `generators.delaunay` (_delaunay.py:8-59), but
    built from `vertex_neighbor_vertices` without the per-point Python loop.
    """
    import scipy.spatial as sptl
    tri = sptl.Delaunay(points)
    indptr, indices = tri.vertex_neighbor_vertices
    rows = np.repeat(np.arange(points.shape[0]), np.diff(indptr))
    keep = rows < indices
    conns = np.vstack((rows[keep], indices[keep])).T.astype(np.int64)
    order = np.lexsort((conns[:, 1], conns[:, 0]))
    return np.ascontiguousarray(conns[order])


# --------------------------------------------------------------------------
# assembly
# --------------------------------------------------------------------------
def adjacency_coo(conns, g, Np):
    """`Network.create_adjacency_matrix(weights=g, fmt='coo')`.

    openpnm/network/_network.py:274-293: row=[c0;c1], col=[c1;c0]; an (Nt,)
    weight gives data=[g;g] (symmetric), an (Nt,2) weight gives
    data=weights.flatten(order='F') = [g[:,0]; g[:,1]] (:281-285), i.e. entry
    (c0,c1) carries g[:,0] and entry (c1,c0) carries g[:,1] (AdvectionDiffusion).
    """
    conns = np.asarray(conns)
    g = np.asarray(g, dtype=float)
    row = np.append(conns[:, 0], conns[:, 1])
    col = np.append(conns[:, 1], conns[:, 0])
    data = g.flatten(order="F") if g.ndim == 2 else np.append(g, g)
    return row, col, data


def laplacian_coo(row, col, data, Np):
    """`scipy.sparse.csgraph.laplacian(am)` on a COO adjacency, in numpy.

    scipy/sparse/csgraph/_laplacian.py:449-484 (`_laplacian_sparse`):
    w = colsum - diagonal; data *= -1; COO `setdiag(w)` drops every existing
    diagonal entry and appends (i, i, w_i) for i = 0..Np-1 (_coo.py:511-547).
    """
    colsum = np.zeros(Np)
    np.add.at(colsum, col, data)
    dmask = row == col
    diag = np.zeros(Np)
    np.add.at(diag, row[dmask], data[dmask])
    w = colsum - diag
    keep = ~dmask
    ar = np.arange(Np, dtype=row.dtype)
    return (np.concatenate((row[keep], ar)),
            np.concatenate((col[keep], ar)),
            np.concatenate((-data[keep], w)))


def build_A(conns, g, Np):
    """`Transport._build_A` (openpnm/algorithms/_transport.py:95-115) -> COO triple."""
    row, col, data = adjacency_coo(conns, g, Np)
    return laplacian_coo(row, col, data, Np)


def _coo_diagonal(row, col, data, Np):
    d = np.zeros(Np)
    m = row == col
    np.add.at(d, row[m], data[m])
    return d


def _coo_setdiag(row, col, data, values):
    keep = row != col
    ar = np.arange(values.size, dtype=row.dtype)
    return (np.concatenate((row[keep], ar)), np.concatenate((col[keep], ar)),
            np.concatenate((data[keep], values)))


def apply_bcs(A, Np, bc_value=None, bc_rate=None):
    """`Transport._build_b` + `_apply_BCs` (_transport.py:117-121, 145-169).

    A is the COO triple of the pure Laplacian.  NaN in bc_value / bc_rate means
    "no BC here" (Transport.__init__, _transport.py:68-69).  Returns
    (A_coo_after, b, f).
    """
    row, col, data = (np.array(a) for a in A)
    b = np.zeros(Np)
    f = np.nan
    if bc_rate is not None:
        ind = np.isfinite(bc_rate)
        b[ind] = bc_rate[ind]
    if bc_value is not None:
        f = _coo_diagonal(row, col, data, Np).mean()
        ind = np.isfinite(bc_value)
        b[ind] = bc_value[ind] * f
        x_bc = np.zeros(Np)
        x_bc[ind] = bc_value[ind]
        Ax = np.zeros(Np)
        np.add.at(Ax, row, data * x_bc[col])
        b[~ind] -= Ax[~ind]
        mask = ind[row] | ind[col]
        data[mask] = 0
        datadiag = _coo_diagonal(row, col, data, Np)
        datadiag[ind] = f
        row, col, data = _coo_setdiag(row, col, data, datadiag)
        nz = data != 0
        row, col, data = row[nz], col[nz], data[nz]
    return (row, col, data), b, f


def apply_outflow(A, Np, bc_outflow):
    """`AdvectionDiffusion._apply_BCs`, the part after the parent's
    (openpnm/algorithms/_advection_diffusion.py:86-105): diag[ind] += outflow[ind]
    through COO `setdiag` (all Np diagonal entries become explicit)."""
    row, col, data = A
    diag = _coo_diagonal(row, col, data, Np)
    ind = np.isfinite(bc_outflow)
    diag[ind] += bc_outflow[ind]
    return _coo_setdiag(row, col, data, diag)


def outflow_values(conns, gh, pressure, pores, Np):
    """`AdvectionDiffusion.set_outflow_BC` (_advection_diffusion.py:56-84): the net
    hydraulic flow rate into each listed pore, stored as its 'outflow' BC value."""
    conns = np.asarray(conns)
    sel = np.isin(conns[:, 0], pores) | np.isin(conns[:, 1], pores)
    C12 = conns[sel]
    P12 = np.asarray(pressure)[C12]
    Q12 = -np.asarray(gh)[sel] * np.diff(P12, axis=1).squeeze(axis=1)
    Qp = np.zeros(Np)
    np.add.at(Qp, C12[:, 0], -Q12)
    np.add.at(Qp, C12[:, 1], Q12)
    out = np.full(Np, np.nan)
    out[pores] = Qp[pores]
    return out


def apply_sources(A, b, Np, sources):
    """`ReactiveTransport._apply_sources` (_reactive_transport.py:125-148).

    sources: iterable of (mask_bool[Np], S1[Np], S2[Np]).
    """
    row, col, data = A
    b = b.copy()
    for mask, S1, S2 in sources:
        diag = _coo_diagonal(row, col, data, Np)
        diag[mask] += -S1[mask]
        row, col, data = _coo_setdiag(row, col, data, diag)
        b[mask] += S2[mask]
    return (row, col, data), b


def to_csr(A, Np):
    """`A.tocsr()` as the solvers do (openpnm/solvers/_scipy.py:13-14)."""
    row, col, data = A
    return sprs.coo_matrix((data, (row, col)), shape=(Np, Np)).tocsr()


# --------------------------------------------------------------------------
# source-term models (openpnm/models/physics/source_terms/_funcs.py)
# --------------------------------------------------------------------------
def source_term(kind, X, params):
    """Return (S1, S2, rate) for the three models on the hot path.

    standard_kinetics: _funcs.py:21-69   r = A X^b
    linear:            _funcs.py:73-116  r = A1 X + A2
    power_law:         _funcs.py:120-168 r = A1 X^A2 + A3
    """
    X = np.asarray(X, dtype=float)
    with np.errstate(all="ignore"):
        if kind == "standard_kinetics":
            A, b = params["prefactor"], params["exponent"]
            r = A * (X ** b)
            S1 = A * b * (X ** (b - 1))
            S2 = A * (1 - b) * (X ** b)
        elif kind == "linear":
            A1, A2 = params.get("A1", 0.0), params.get("A2", 0.0)
            r = A1 * X + A2
            S1 = A1 * np.ones_like(X)
            S2 = A2 * np.ones_like(X)
        elif kind == "power_law":
            A, B, C = params.get("A1", 0.0), params.get("A2", 0.0), params.get("A3", 0.0)
            r = A * X ** B + C
            S1 = A * B * X ** (B - 1)
            S2 = A * X ** B * (1 - B) + C
        else:
            raise ValueError(kind)
    return S1, S2, r


# --------------------------------------------------------------------------
# solvers
# --------------------------------------------------------------------------
def spsolve(A_csr, b):
    """`ScipySpsolve.solve` (openpnm/solvers/_scipy.py:8-15): SuperLU, exit code 0."""
    return spla.spsolve(A_csr, b), 0


def jacobi_pcg(A_csr, b, x0=None, rtol=1e-8, maxiter=None, precond=True):
    """Textbook (Jacobi-)preconditioned CG, stop when ||b - A x||_2 <= rtol*||b||_2.

    Same stopping rule as `ScipyCG.solve` (_scipy.py:21-26: tol=self.tol,
    atol=norm(b)*tol).  Restated in numpy so the iteration count is observable.
    Returns (x, info, iters) with scipy's info convention (0 ok, >0 maxiter).
    """
    n = b.size
    x = np.zeros(n) if x0 is None else np.array(x0, dtype=float)
    d = A_csr.diagonal()
    Minv = 1.0 / d if precond else np.ones(n)
    if maxiter is None:
        maxiter = 10 * n
    atol = rtol * np.linalg.norm(b)
    r = b - A_csr @ x
    if np.linalg.norm(r) <= atol:
        return x, 0, 0
    z = Minv * r
    p = z.copy()
    rz = r @ z
    for k in range(1, maxiter + 1):
        Ap = A_csr @ p
        alpha = rz / (p @ Ap)
        x += alpha * p
        r -= alpha * Ap
        if np.linalg.norm(r) <= atol:
            return x, 0, k
        z = Minv * r
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, maxiter, maxiter


# --------------------------------------------------------------------------
# outer loops
# --------------------------------------------------------------------------
def _termination_check(state, f, x, dx, f_rtol, x_rtol):
    """scipy.optimize._nonlin.TerminationCondition.check with f_tol=x_tol=inf,
    norm=numpy.linalg.norm, as constructed at _reactive_transport.py:183-185."""
    f_norm, x_norm, dx_norm = (np.linalg.norm(v) for v in (f, x, dx))
    if state.get("f0") is None:
        state["f0"] = f_norm
    if f_norm == 0:
        return True
    return bool(f_norm / f_rtol <= state["f0"] and dx_norm / x_rtol <= x_norm)


def picard_run(conns, g, Np, bc_value=None, bc_rate=None, sources=(), x0=None,
               solver=None, relaxation_factor=1.0, newton_maxiter=5000,
               f_rtol=1e-6, x_rtol=1e-6, bc_outflow=None):
    """`Transport.run` + `ReactiveTransport._run_special` for one algorithm.

    _transport.py:171-221 and _reactive_transport.py:150-213, including the
    first-iteration aliasing quirk: `SteadyStateSolution(x0)` is a *view* of x0
    (_solution.py:17-20) and `xold = self.x` is that buffer, so after the first
    solve `self.soln[q][:] = self.x` (_transport.py:220) overwrites xold in
    place and the first `dx = self.x - xold` is identically zero.

    sources: iterable of (mask_bool[Np], kind, params) evaluated at the
    current x each outer iteration (`_update_iterative_props`,
    _algorithm.py:69-91).  solver(A_csr, b, x0) -> (x, exit_code).
    Returns dict(x, num_iter, is_converged, A (COO), b).
    """
    if solver is None:
        solver = lambda A, b, x0: spsolve(A, b)  # noqa: E731
    pure = None if callable(g) else build_A(conns, g, Np)

    def update(x):
        # a conductance that is itself an iterative property (g callable: g(x)) switches
        # the cache off and the Laplacian is rebuilt from the regenerated values at every
        # update (_transport.py:106-114 after _update_iterative_props, _algorithm.py:69-91)
        A0 = build_A(conns, g(x), Np) if callable(g) else pure
        A, b, _ = apply_bcs(A0, Np, bc_value, bc_rate)
        if bc_outflow is not None:
            A = apply_outflow(A, Np, bc_outflow)
        lin = []
        for mask, kind, params in sources:
            S1, S2, _ = source_term(kind, x, params)
            lin.append((mask, S1, S2))
        if lin:
            A, b = apply_sources(A, b, Np, lin)
        return A, b

    x = np.zeros(Np) if x0 is None else np.array(x0, dtype=float)
    if not np.isfinite(x).all():
        raise Exception("x0 contains inf/nan values")
    w = relaxation_factor
    A, b = update(x)
    xold = x
    dx = np.zeros(Np)
    state = {}
    num_iter = 0
    is_converged = False
    first = True
    for i in range(newton_maxiter):
        res = to_csr(A, Np) @ x - b
        if _termination_check(state, res, xold, dx, f_rtol, x_rtol):
            is_converged = True
            break
        if not (np.isfinite(A[2]).all() and np.isfinite(b).all()):
            raise Exception("A or b contains inf/nan values")
        x_new, exit_code = solver(to_csr(A, Np), b, xold)
        x = w * x_new + (1 - w) * x
        A, b = update(x)
        if first:
            dx = np.zeros(Np)      # aliasing quirk, see docstring
            first = False
        else:
            dx = x - xold
        xold = x
        num_iter = i + 1
    return {"x": x, "num_iter": num_iter, "is_converged": is_converged,
            "A": A, "b": b}


def transport_run(conns, g, Np, bc_value=None, bc_rate=None, solver=None, x0=None):
    """Linear problem: one assembly + one solve (what `run` amounts to when
    there are no sources; num_iter == 1 for a converged direct solve)."""
    out = picard_run(conns, g, Np, bc_value, bc_rate, (), x0, solver)
    return out


def transient_run(conns, g, Np, volume, x0, tspan, saveat=None, bc_value=None, bc_rate=None,
                  sources=(), rtol=1e-6, atol=1e-6, bc_outflow=None):
    """`TransientReactiveTransport.run` with `ScipyRK45`
    (openpnm/algorithms/_transient_reactive_transport.py:51-123,
    openpnm/integrators/_scipy.py:18-56): dy/dt = (-A y + b) / V with A, b
    re-assembled (BCs + sources linearised at y) at every right-hand-side call,
    integrated by scipy.integrate.solve_ivp(method="RK45", t_eval=saveat).
    Returns (t, Y) with Y of shape (Np, len(t)) as SciPy does."""
    from scipy.integrate import solve_ivp
    if np.isscalar(saveat):
        saveat = np.arange(*tspan, saveat)
    if (saveat is not None) and (tspan[1] not in saveat):
        saveat = np.hstack((saveat, [tspan[1]]))
    y0 = np.ones(Np, dtype=float) * x0
    if bc_value is not None:                      # _merge_inital_and_boundary_values
        m = ~np.isnan(bc_value)
        y0[m] = bc_value[m]
    pure = build_A(conns, g, Np)
    V = np.asarray(volume, dtype=float) * np.ones(Np)

    def ode_func(t, y):
        A, b, _ = apply_bcs(pure, Np, bc_value, bc_rate)
        if bc_outflow is not None:
            A = apply_outflow(A, Np, bc_outflow)
        lin = [(mask, *source_term(kind, y, params)[:2]) for mask, kind, params in sources]
        if lin:
            A, b = apply_sources(A, b, Np, lin)
        return (-(to_csr(A, Np) @ y) + b) / V

    sol = solve_ivp(ode_func, tspan, y0, method="RK45", atol=atol, rtol=rtol, t_eval=saveat)
    if not sol.success:
        raise Exception(sol.message)
    return sol.t, sol.y


def rate(conns, g, x, pores=None, throats=None, mode="group"):
    """`Transport.rate` (_transport.py:259-330), always with the pure conductances."""
    conns = np.asarray(conns)
    g = np.asarray(g, dtype=float)
    Qt = g * (x[conns[:, 1]] - x[conns[:, 0]])
    if throats is not None and np.size(throats):
        R = np.abs(Qt[np.asarray(throats)])
    else:
        Qp = np.zeros(x.size)
        np.add.at(Qp, conns[:, 0], -Qt)
        np.add.at(Qp, conns[:, 1], Qt)
        R = Qp[np.asarray(pores)]
    if mode == "group":
        R = np.sum(R)
    return np.array(R, ndmin=1)


# ------------------------------------------------- conductance models
def polynomial(a, x):
    """models.misc.polynomial (openpnm/models/misc/_simple_equations.py:113-117):
    sum of a[i] * x**i, accumulated in that order."""
    x = np.asarray(x).astype(float)
    value = 0.0
    for i in range(len(a)):
        value = value + a[i] * (x ** i)
    return value


def from_neighbor_pores(conns, pore_values, mode="min"):
    """models.misc.from_neighbor_pores on all throats
    (openpnm/models/misc/_neighbor_lookups.py:121-138), nan-free input."""
    v = np.asarray(pore_values)[np.asarray(conns)]
    return {"min": np.amin, "max": np.amax, "mean": np.mean, "sum": np.sum}[mode](v, axis=1)


def conduit_conductance(kind, conns, pore_prop, throat_prop, size_factors):
    """Series conductance of pore 1 | throat | pore 2.
    kind 'hydraulic': generic_hydraulic (models/physics/hydraulic_conductance/_funcs.py:37-53),
    g_i = F_i / mu_i, a 1-D size factor array is the throat's with F1 = F2 = inf;
    kind 'poisson': _poisson_conductance (models/physics/_utils.py:46-61) behind
    generic_diffusive / _thermal / _electrical, g_i = D_i * F_i, a 1-D size factor array is
    the whole conduit's: g = D_throat * F."""
    conns = np.asarray(conns)
    p1, p2 = np.asarray(pore_prop)[conns].T
    pt = np.asarray(throat_prop)
    SF = np.asarray(size_factors)
    if kind == "hydraulic":
        F1, Ft, F2 = SF.T if SF.ndim > 1 else (np.inf, SF, np.inf)
        g1, gt, g2 = F1 / p1, Ft / pt, F2 / p2
    elif kind == "poisson":
        if SF.ndim != 2:
            return pt * SF
        F1, Ft, F2 = SF.T
        g1, gt, g2 = p1 * F1, pt * Ft, p2 * F2
    else:
        raise ValueError(kind)
    return 1 / (1 / g1 + 1 / gt + 1 / g2)


# ---------------------------------------------------------------- health + trim
def network_health(conns, Np):
    """check_network_health restricted to what the path needs
    (openpnm/utils/_health.py:40-98; models/network/_health.py:44-76): per-pore cluster
    size (scipy connected_components on the triu adjacency), isolated pores
    (np.unique(conns) misses them), pores outside the largest cluster, looped throats."""
    from scipy.sparse import csgraph
    conns = np.asarray(conns, dtype=np.int64).reshape(-1, 2)
    am = sprs.coo_matrix((np.ones(conns.shape[0], dtype=np.int8), (conns[:, 0], conns[:, 1])),
                         shape=(Np, Np))
    ncomp, cluster_num = csgraph.connected_components(am, directed=False)
    _, ind, N = np.unique(cluster_num, return_inverse=True, return_counts=True)
    size = N[ind]
    isolated = np.ones(Np, dtype=bool)
    isolated[np.unique(conns)] = False
    return {"cluster_size": size, "isolated": isolated, "n_clusters": int(ncomp),
            "disconnected": np.flatnonzero(size < size.max()),
            "looped": np.flatnonzero(conns[:, 0] == conns[:, 1]),
            "cluster_num": cluster_num}


def trim(conns, Np, pores=(), throats=()):
    """topotools.trim (openpnm/topotools/_topotools.py:157-229): returns the remapped
    conns of the kept throats, the kept pore indices and the kept throat indices."""
    conns = np.asarray(conns, dtype=np.int64).reshape(-1, 2)
    Pkeep = np.ones(Np, dtype=bool)
    Tkeep = np.ones(conns.shape[0], dtype=bool)
    if np.size(pores):
        Pkeep[np.asarray(pores)] = False
        Tkeep[~(Pkeep[conns[:, 0]] & Pkeep[conns[:, 1]])] = False   # find_neighbor_throats(union)
    if np.size(throats):
        Tkeep[np.asarray(throats)] = False
    Pmap = -np.ones(Np, dtype=np.int64)
    Pmap[Pkeep] = np.arange(Pkeep.sum())
    new = np.stack([Pmap[conns[Tkeep, 0]], Pmap[conns[Tkeep, 1]]], axis=1)
    return new, np.flatnonzero(Pkeep), np.flatnonzero(Tkeep)
