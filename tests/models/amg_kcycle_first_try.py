"""TEST INFRASTRUCTURE (uses oracle/), kept for the record: the first CPU prototype (numpy/scipy) of an aggregation AMG with K-cycle for the
Jacobi-scaled graph Laplacian systems of the transport path: measures outer
iteration counts before any CUDA is written.  Not product code."""
import sys, time
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle


def handshake_pairs(A, rounds=6):
    """Parallel heavy-edge matching: every unmatched node points at its strongest
    unmatched neighbour (ties -> smaller index); mutual pointers match."""
    n = A.shape[0]
    A = A.tocsr()
    indptr, indices = A.indptr, A.indices
    w = -A.data.copy()                       # off-diagonals are negative
    rows = np.repeat(np.arange(n), np.diff(indptr))
    w[indices == rows] = -np.inf
    w[w <= 0] = -np.inf
    agg = -np.ones(n, dtype=np.int64)
    for _ in range(rounds):
        un = agg < 0
        ww = np.where(un[indices] & un[rows], w, -np.inf)
        # argmax per row
        best = np.full(n, -np.inf)
        np.maximum.at(best, rows, ww)
        cand = (ww == best[rows]) & np.isfinite(ww)
        ptr = np.full(n, n, dtype=np.int64)
        np.minimum.at(ptr, rows[cand], indices[cand])
        has = ptr < n
        idx = np.flatnonzero(has)
        mutual = idx[ptr[ptr[idx]] == idx]
        lead = mutual[mutual < ptr[mutual]]
        agg[lead] = lead
        agg[ptr[lead]] = lead
        if not len(lead):
            break
    # leftovers: join strongest neighbour's aggregate if any, else singleton
    un = np.flatnonzero(agg < 0)
    if len(un):
        ww = np.where((agg[indices] >= 0) & (agg[rows] < 0), w, -np.inf)
        best = np.full(n, -np.inf)
        np.maximum.at(best, rows, ww)
        cand = (ww == best[rows]) & np.isfinite(ww)
        ptr = np.full(n, n, dtype=np.int64)
        np.minimum.at(ptr, rows[cand], indices[cand])
        j = un[ptr[un] < n]
        agg[j] = agg[ptr[j]]
        rest = np.flatnonzero(agg < 0)
        deg = np.zeros(n); np.add.at(deg, rows, np.isfinite(w))
        iso = rest[deg[rest] == 0]
        conn = rest[deg[rest] > 0]
        agg[conn] = conn
        agg[iso] = -1                      # decoupled rows (BC pores): not on the coarse grid
    keep = agg >= 0
    inv = -np.ones(n, dtype=np.int64)
    _, inv_k = np.unique(agg[keep], return_inverse=True)
    inv[keep] = inv_k
    return inv


def coarsen(A, passes):
    n = A.shape[0]
    agg = np.arange(n)
    Ac = A
    for _ in range(passes):
        a = handshake_pairs(Ac)
        nc = a.max() + 1
        k = a >= 0
        P = sp.csr_matrix((np.ones(k.sum()), (np.flatnonzero(k), a[k])), shape=(Ac.shape[0], nc))
        Ac = (P.T @ Ac @ P).tocsr()
        agg = np.where(agg >= 0, a[np.maximum(agg, 0)], -1)
    k = agg >= 0
    P = sp.csr_matrix((np.ones(k.sum()), (np.flatnonzero(k), agg[k])), shape=(n, agg.max() + 1))
    return P, Ac


class Level:
    pass


def setup(A, passes=2, min_size=400, max_levels=12):
    levels = []
    while True:
        L = Level()
        L.A = A.tocsr()
        L.Dinv = 1.0 / A.diagonal()
        levels.append(L)
        if A.shape[0] <= min_size or len(levels) >= max_levels:
            break
        P, Ac = coarsen(A, passes)
        if Ac.shape[0] >= 0.8 * A.shape[0]:
            break
        L.P = P
        A = Ac
    levels[-1].lu = spla.splu(levels[-1].A.tocsc())
    return levels


def smooth(L, x, b, kind, nsweeps=1):
    if kind == 'jacobi':
        for _ in range(nsweeps):
            x = x + 0.67 * L.Dinv * (b - L.A @ x)
        return x
    if kind == 'l1':
        if not hasattr(L, 'l1inv'):
            L.l1inv = 1.0 / np.asarray(abs(L.A).sum(axis=1)).ravel()
        for _ in range(nsweeps):
            x = x + L.l1inv * (b - L.A @ x)
        return x
    if kind == 'cheb':
        # Chebyshev on D^-1 A for [lmax/a, lmax], lmax<=2 bound
        lmax, lmin = 2.0, 2.0 / 6.0
        d, c = (lmax + lmin) / 2, (lmax - lmin) / 2
        r = L.Dinv * (b - L.A @ x)
        p = r / d
        x = x + p
        alpha = 1.0 / d
        for k in range(1, nsweeps):
            r = L.Dinv * (b - L.A @ x)
            beta = (c * alpha / 2) ** 2 * (2 if k == 1 else 1)
            alpha = 1.0 / (d - beta / alpha) if k > 1 else 1.0 / (d - c * c / (2 * d))
            beta = alpha * d - 1.0
            p = alpha * r + beta * p
            x = x + p
        return x
    raise ValueError(kind)


COUNT = {}


def cycle(levels, l, b, kind, kc=2, sweeps=1):
    L = levels[l]
    COUNT[l] = COUNT.get(l, 0) + 1
    if l == len(levels) - 1:
        return L.lu.solve(b)
    x = smooth(L, np.zeros_like(b), b, kind, sweeps)
    r = b - L.A @ x
    rc = L.P.T @ r
    if l + 1 == len(levels) - 1 or (kc >= 0 and l >= kc):
        ec = cycle(levels, l + 1, rc, kind, kc, sweeps)
    else:
        # K-cycle: up to 2 FCG iterations on the coarse problem (Notay)
        Ac = levels[l + 1].A
        c1 = cycle(levels, l + 1, rc, kind, kc, sweeps)
        v1 = Ac @ c1
        rho1 = c1 @ v1
        a1 = c1 @ rc
        r1 = rc - (a1 / rho1) * v1
        if kc < 0 and np.linalg.norm(r1) <= 0.25 * np.linalg.norm(rc):
            ec = (a1 / rho1) * c1
        else:
            c2 = cycle(levels, l + 1, r1, kind, kc, sweeps)
            v2 = Ac @ c2
            g = c2 @ v1
            bta = c2 @ v2
            a2 = c2 @ r1
            rho2 = bta - g * g / rho1
            ec = (a1 / rho1 - g * a2 / (rho1 * rho2)) * c1 + (a2 / rho2) * c2
    x = x + L.P @ ec
    x = x + (smooth(L, np.zeros_like(b), b - L.A @ x, kind, sweeps))
    return x


def fcg(A, b, M, rtol=1e-10, maxiter=500):
    """flexible CG (Notay): one direction of memory"""
    x = np.zeros_like(b)
    r = b.copy()
    bn = np.linalg.norm(b)
    d = None
    for it in range(1, maxiter + 1):
        z = M(r)
        if d is None:
            d = z
        else:
            d = z - ((z @ q) / dq) * d
        q = A @ d
        dq = d @ q
        al = (d @ r) / dq
        x += al * d
        r -= al * q
        if np.linalg.norm(r) <= rtol * bn:
            return x, it
    return x, maxiter


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    kind = sys.argv[2] if len(sys.argv) > 2 else 'jacobi'
    passes = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    sweeps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    kc = int(sys.argv[5]) if len(sys.argv) > 5 else 2
    net = oracle.cubic_network([N, N, N])
    conns = net['conns']
    Nt = conns.shape[0]
    g = 10.0 ** np.random.default_rng(0).uniform(-15, -12, Nt)
    Np = N ** 3
    bcv = np.full(Np, np.nan)
    idx = np.arange(Np).reshape(N, N, N)
    bcv[idx[0].ravel()] = 2e5
    bcv[idx[-1].ravel()] = 1e5
    A = oracle.build_A(conns, g, Np)
    A, b = oracle.apply_bcs(A, Np, bcv, None)[:2]
    A = oracle.to_csr(A, Np)
    d = A.diagonal()
    s = 1 / np.sqrt(d)
    At = sp.diags(s) @ A @ sp.diags(s)
    bt = s * b
    t0 = time.time()
    levels = setup(A.tocsr(), passes=passes)
    print('levels', [L.A.shape[0] for L in levels], 'nnz', [L.A.nnz for L in levels], 'setup %.1fs' % (time.time() - t0))
    opc = sum(L.A.nnz for L in levels) / levels[0].A.nnz
    print('operator complexity %.2f' % opc)
    t0 = time.time()
    x, it = fcg(A, b, lambda r: cycle(levels, 0, r, kind, kc, sweeps))
    print('AMG-FCG iterations', it, 'time %.1fs' % (time.time() - t0), 'visits', dict(sorted(COUNT.items())))
    res = np.linalg.norm(b - A @ x) / np.linalg.norm(b)
    print('relres', res)
    # plain Jacobi-CG for comparison
    xs, itj = oracle.jacobi_pcg(A, b, rtol=1e-10, maxiter=20000)[:2] if False else (None, None)


main()
