"""TEST INFRASTRUCTURE (uses oracle/): CPU model (numpy/scipy) of the aggregation-AMG K-cycle that csrc/amg.cu
implements -- same matching rule, same Chebyshev recurrences, same truncated
K-cycle, CG on the coarsest level.  Used to choose parameters; not product code."""
import sys, time
import numpy as np, scipy.sparse as sp
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle

ROUNDS = 5


def pair_pass(A):
    """One pairwise pass.  Returns agg (n,) coarse ids, -1 for decoupled rows."""
    n = A.shape[0]
    A = A.tocsr(); A.sort_indices()
    indptr, indices = A.indptr, A.indices
    rows = np.repeat(np.arange(n), np.diff(indptr))
    w = -A.data.copy()
    w[(indices == rows) | ~(w > 0)] = -np.inf
    match = -np.ones(n, dtype=np.int64)

    def strongest(ok_edge):
        ww = np.where(ok_edge, w, -np.inf)
        best = np.full(n, -np.inf)
        np.maximum.at(best, rows, ww)
        cand = (ww == best[rows]) & np.isfinite(ww)
        ptr = np.full(n, n, dtype=np.int64)
        np.minimum.at(ptr, rows[cand], indices[cand])
        ptr[ptr == n] = -1
        return ptr

    for _ in range(ROUNDS):
        un = match < 0
        prop = strongest(un[rows] & un[indices])
        i = np.flatnonzero(prop >= 0)
        ok = i[prop[prop[i]] == i]
        match[ok] = prop[ok]
    un = match < 0
    deg = np.zeros(n); np.add.at(deg, rows, np.isfinite(w))
    join = strongest(un[rows] & ~un[indices])          # strongest matched neighbour
    leader = np.where(match >= 0, np.minimum(np.arange(n), match), -1)
    j = np.flatnonzero(un & (join >= 0))
    leader[j] = leader[join[j]]
    s = np.flatnonzero(un & (join < 0) & (deg > 0))    # coupled only to unmatched nodes: singleton
    leader[s] = s
    flag = leader == np.arange(n)
    ids = np.cumsum(flag) - 1
    agg = np.where(leader >= 0, ids[np.maximum(leader, 0)], -1)
    return agg


def galerkin(A, agg):
    k = agg >= 0
    nc = agg.max() + 1
    P = sp.csr_matrix((np.ones(k.sum()), (np.flatnonzero(k), agg[k])), shape=(A.shape[0], nc))
    return P, (P.T @ A @ P).tocsr()


class Level: pass


def setup(A, passes=3, coarsest=2000, max_levels=10, passes0=None):
    levels = []
    while True:
        L = Level(); L.A = A.tocsr(); L.n = A.shape[0]
        d = A.diagonal(); L.dinv = 1.0 / d
        L.lmax = float(np.max(np.asarray(abs(A).sum(axis=1)).ravel() / d))   # Gershgorin bound of D^-1 A
        levels.append(L)
        if L.n <= coarsest or len(levels) >= max_levels:
            break
        agg = np.arange(L.n); Ac = L.A
        for _ in range(passes0 if (passes0 and len(levels) == 1) else passes):
            a = pair_pass(Ac)
            _, Ac = galerkin(Ac, a)
            agg = np.where(agg >= 0, a[np.maximum(agg, 0)], -1)
        if Ac.shape[0] > 0.7 * L.n:
            break
        k = agg >= 0
        L.P = sp.csr_matrix((np.ones(k.sum()), (np.flatnonzero(k), agg[k])), shape=(L.n, agg.max() + 1))
        A = Ac
    return levels


def cheb(L, b, x, from_zero, deg, ratio):
    """Saad Alg. 12.1 with M = D: eigenvalues of D^-1 A in [lmax/ratio, lmax]."""
    lmax = 1.05 * L.lmax; lmin = lmax / ratio
    theta, delta = (lmax + lmin) / 2, (lmax - lmin) / 2
    sigma = theta / delta
    rho = 1.0 / sigma
    if from_zero:
        d = L.dinv * b / theta
        x = d.copy()
    else:
        d = L.dinv * (b - L.A @ x) / theta
        x = x + d
    for _ in range(deg - 1):
        rho_new = 1.0 / (2 * sigma - rho)
        d = rho_new * rho * d + (2 * rho_new / delta) * L.dinv * (b - L.A @ x)
        x = x + d
        rho = rho_new
    return x


def coarse_cg(L, b, rtol, maxit):
    x = np.zeros_like(b); r = b.copy(); z = L.dinv * r; p = z.copy(); rz = r @ z
    b2 = b @ b
    if b2 == 0: return x
    for it in range(maxit):
        q = L.A @ p
        al = rz / (p @ q)
        x += al * p; r -= al * q
        if r @ r <= rtol * rtol * b2: break
        z = L.dinv * r; rzn = r @ z; p = z + (rzn / rz) * p; rz = rzn
    return x


VIS = {}


def cycle(levels, l, b, prm):
    L = levels[l]
    VIS[l] = VIS.get(l, 0) + 1
    if l == len(levels) - 1:
        return coarse_cg(L, b, prm['crtol'], prm['cmax'])
    dg = prm['deg'] if l == 0 else prm.get('cdeg', prm['deg'])
    x = cheb(L, b, None, True, prm.get('pre', dg), prm['ratio'])
    rc = L.P.T @ (b - L.A @ x)
    Lc = levels[l + 1]
    if l < prm['klev'] and l + 1 < len(levels) - 1:
        c1 = cycle(levels, l + 1, rc, prm)
        v1 = Lc.A @ c1; rho1 = c1 @ v1; a1 = c1 @ rc
        r1 = rc - (a1 / rho1) * v1
        c2 = cycle(levels, l + 1, r1, prm)
        v2 = Lc.A @ c2; gam = c2 @ v1; bet = c2 @ v2; a2 = c2 @ r1
        rho2 = bet - gam * gam / rho1
        ec = (a1 / rho1 - gam * a2 / (rho1 * rho2)) * c1 + (a2 / rho2) * c2
    else:
        ec = cycle(levels, l + 1, rc, prm)
    x = x + L.P @ ec
    return cheb(L, b, x, False, prm.get('post', dg), prm['ratio'])


def fcg(A, b, M, rtol=1e-10, maxiter=300):
    x = np.zeros_like(b); r = b.copy(); bn = np.linalg.norm(b); d = None
    for it in range(1, maxiter + 1):
        z = M(r)
        d = z if d is None else z - ((z @ q) / dq) * d
        q = A @ d; dq = d @ q; al = (d @ r) / dq
        x += al * d; r -= al * q
        if np.linalg.norm(r) <= rtol * bn: return x, it
    return x, maxiter


def problem(kind, N):
    rng = np.random.default_rng(0)
    if kind == 'cubic':
        net = oracle.cubic_network([N, N, N]); conns = net['conns']; Np = N ** 3
        idx = np.arange(Np).reshape(N, N, N)
        left, right = idx[0].ravel(), idx[-1].ravel()
    else:
        pts = rng.random((N, 3)); conns = oracle.delaunay_network(pts); conns = conns['conns'] if isinstance(conns, dict) else conns; Np = N
        xs = pts[:, 0]
        left, right = np.flatnonzero(xs < 0.05), np.flatnonzero(xs > 0.95)
    g = 10.0 ** rng.uniform(-15, -12, conns.shape[0])
    bcv = np.full(Np, np.nan); bcv[left] = 2e5; bcv[right] = 1e5
    A = oracle.build_A(conns, g, Np)
    A, b = oracle.apply_bcs(A, Np, bcv, None)[:2]
    return oracle.to_csr(A, Np), b


def main():
    kind = sys.argv[1]; N = int(sys.argv[2])
    prm = dict(deg=int(sys.argv[3]), ratio=float(sys.argv[4]), klev=int(sys.argv[5]),
               crtol=float(sys.argv[6]), cmax=int(sys.argv[7]))
    passes = int(sys.argv[8]) if len(sys.argv) > 8 else 3
    if len(sys.argv) > 10: prm['pre'] = int(sys.argv[9]); prm['post'] = int(sys.argv[10])
    A, b = problem(kind, N)
    import os
    p0 = int(os.environ.get('P0', '0')) or None
    if os.environ.get('CDEG'): prm['cdeg'] = int(os.environ['CDEG'])
    t0 = time.time(); levels = setup(A, passes=passes, passes0=p0)
    print('levels', [L.n for L in levels], 'nnz', [L.A.nnz for L in levels], 'lmax', ['%.2f' % L.lmax for L in levels], 'setup %.1fs' % (time.time() - t0))
    t0 = time.time()
    x, it = fcg(A, b, lambda r: cycle(levels, 0, r, prm))
    print(prm, 'iterations', it, 'time %.1fs' % (time.time() - t0), 'visits/it', {k: round(v / it, 1) for k, v in sorted(VIS.items())},
          'relres %.2e' % (np.linalg.norm(b - A @ x) / np.linalg.norm(b)))


if __name__ == '__main__':
    main()
