"""TEST INFRASTRUCTURE (uses oracle/ through amg_v2): CPU model of a *box* aggregation for
lattice operators -- the next step DESIGN.md section 10 names for the single-GPU multigrid.

On a 6-neighbour lattice the aggregates can be the 2x2x2 boxes of the index lattice
instead of the result of pairwise matching.  With a piecewise-constant prolongation only
face-adjacent boxes couple, so every Galerkin operator is again a 7-point lattice operator:
three super-diagonals at offsets (1, cz, cy*cz) -- the DIA-sym format and the K_A /
k_dia_smooth kernels of the fine level apply to every level, no column indices anywhere,
and the hierarchy setup is a structured reduction (sum of the conductances crossing a box
face) with no matching rounds, no sort and no gather.

Measured with this model (same cycle as csrc/amg.cu: Chebyshev(2) on [lmax/8, lmax], K-cycle on
the first levels, CG on the coarsest), g = 10^U(-15,-12), outer flexible CG to 1e-10:

    lattice     boxes 2x2x2   pairwise matching (4 + 3 passes, what csrc/amg.cu does)
    50^3            30              33
    64^3            27              30
    96^3            28              30
    128^3           30              34
    45x37x61        28              31

K-cycle on level 1 only gives 29 at 96^3.  The box form is NOT better on strongly
anisotropic conductances (z throats x1000: 148 against 93) and equal on six decades of
i.i.d. scatter (183 / 184, both poor): it would be the lattice fast path, chosen like the
DIA operator itself, with the matching hierarchy kept for everything else.  Not product code.
"""
import numpy as np
import scipy.sparse as sp

import amg_v2 as M


def box_aggregates(A, shape, box=(2, 2, 2)):
    """coarse id of every row of a lattice operator (index = (x*ny + y)*nz + z), -1 for rows
    without off-diagonal entries (eliminated Dirichlet rows), and the coarse lattice shape"""
    cs = tuple(-(-s // b) for s, b in zip(shape, box))
    i, j, k = np.unravel_index(np.arange(A.shape[0]), shape)
    agg = ((i // box[0]) * cs[1] + (j // box[1])) * cs[2] + (k // box[2])
    offd = np.asarray(abs(A).sum(axis=1)).ravel() - abs(A.diagonal())
    return np.where(offd > 0, agg, -1), cs


def setup(A, shape, box=(2, 2, 2), coarsest=2000):
    levels = []
    while True:
        L = M.Level()
        L.A, L.n, L.shape = A.tocsr(), A.shape[0], tuple(shape)
        d = A.diagonal()
        L.dinv = 1.0 / d
        L.lmax = float(np.max(np.asarray(abs(A).sum(axis=1)).ravel() / d))
        levels.append(L)
        if L.n <= coarsest:
            return levels
        agg, cs = box_aggregates(L.A, shape, box)
        keep = agg >= 0
        L.P = sp.csr_matrix((np.ones(keep.sum()), (np.flatnonzero(keep), agg[keep])),
                            shape=(L.n, int(np.prod(cs))))
        Ac = (L.P.T @ L.A @ L.P).tocsr()
        empty = Ac.diagonal() == 0           # a box of Dirichlet rows only: decoupled unknown
        if empty.any():
            Ac = (Ac + sp.diags(empty.astype(float))).tocsr()
        A, shape = Ac, cs


def super_diagonals(A):
    A = A.tocoo()
    k = A.col - A.row
    return np.unique(k[(k > 0) & (A.data != 0)])
