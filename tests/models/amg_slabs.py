"""numpy model for the round-2 distributed multigrid: aggregates never cross a slab
boundary (matching ignores cross-slab edges), the Galerkin operators stay GLOBAL
(cross-slab couplings between coarse nodes are kept), so the hierarchy is a true
multigrid for the whole problem and every level can be row-partitioned by slab.
Question answered here: what does forbidding cross-slab aggregates cost in
iterations?  Not product code."""
import os
import sys
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import amg_v2
ns = vars(amg_v2)
from oracle import oracle


def setup_slabs(A, owner, passes=3, passes0=4, coarsest=2000):
    """like amg_v2.setup, but pair_pass only sees edges inside a slab"""
    levels = []
    while True:
        L = ns['Level'](); L.A = A.tocsr(); L.n = A.shape[0]
        d = A.diagonal(); L.dinv = 1.0 / d
        L.lmax = float(np.max(np.asarray(abs(A).sum(axis=1)).ravel() / d))
        levels.append(L)
        if L.n <= coarsest or len(levels) >= 10:
            break
        agg = np.arange(L.n); Ac = L.A; own = owner
        for _ in range(passes0 if len(levels) == 1 else passes):
            C = Ac.tocoo()
            keep = own[C.row] == own[C.col]
            Alocal = sp.coo_matrix((C.data[keep], (C.row[keep], C.col[keep])), shape=Ac.shape).tocsr()
            a = ns['pair_pass'](Alocal)
            _, Ac = ns['galerkin'](Ac, a)                  # global Galerkin product
            k = a >= 0
            own_c = np.zeros(a.max() + 1, dtype=own.dtype); own_c[a[k]] = own[k]
            own = own_c
            agg = np.where(agg >= 0, a[np.maximum(agg, 0)], -1)
        if Ac.shape[0] > 0.7 * L.n:
            break
        k = agg >= 0
        L.P = sp.csr_matrix((np.ones(k.sum()), (np.flatnonzero(k), agg[k])), shape=(L.n, agg.max() + 1))
        A = Ac; owner = own
    return levels


def main():
    shape = tuple(int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (48, 24, 24)
    net = oracle.cubic_network(shape); Np = net['Np']
    g = 10.0 ** np.random.default_rng(3).uniform(-15, -12, net['Nt'])
    bcv = np.full(Np, np.nan); bcv[net['labels']['left']] = 2.0; bcv[net['labels']['right']] = 1.0
    A, b, _ = oracle.apply_bcs(oracle.build_A(net['conns'], g, Np), Np, bcv, None)
    A = oracle.to_csr(A, Np)
    prm = dict(deg=2, ratio=8.0, klev=2, crtol=0.1, cmax=30)
    plane = shape[1] * shape[2]
    for world in (1, 2, 4, 8):
        owner = (np.arange(Np) // plane) * world // shape[0]
        levels = setup_slabs(A, owner)
        x, it = ns['fcg'](A, b, lambda r: ns['cycle'](levels, 0, r, prm), rtol=1e-12)
        print('slabs', world, 'iterations', it, 'levels', [L.n for L in levels],
              'relres %.1e' % (np.linalg.norm(b - A @ x) / np.linalg.norm(b)), flush=True)


if __name__ == '__main__':
    main()
