"""TEST INFRASTRUCTURE: executable specification of the distributed multigrid setup
planned for round 2 ("Multigrid on several GPUs").

Every rank owns a contiguous block of rows on every level.  What a rank holds and
what it has to receive is made explicit:

  level data of rank r   rows [off[r], off[r+1]) of A_l as CSR with GLOBAL column ids
  matching               sees owned columns only (aggregates never cross ranks)
  numbering              local aggregate ids + exclusive scan of the counts over the
                         ranks -> global coarse ids (one all-gather of R integers)
  halo of the map        for every non-owned column j that appears in its rows, the
                         rank needs agg_global[j] from the owner of j: an index-list
                         exchange with the neighbouring ranks (`needs` below)
  Galerkin rows          then purely local: coarse row I = sum of the member rows,
                         columns mapped through (own map | received halo map)
  cycle products         y = A_l x needs x at the same non-owned columns: the same
                         index lists, reused by every product on that level

`build` runs that procedure rank by rank (a loop stands in for the ranks) and records
the traffic; tests/test_amg_model_cpu.py checks that the result IS the hierarchy of
amg_slabs.setup_slabs (global Galerkin, slab-respecting aggregates), i.e. that nothing
but these exchanges is needed."""
import os
import sys

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import amg_v2  # noqa: E402


class RankLevel:
    """rows [lo, hi) of one level on one rank"""

    def __init__(self, lo, hi, A_rows):
        self.lo, self.hi = int(lo), int(hi)
        self.A = A_rows.tocsr()            # (hi-lo) x n_global, global column ids
        self.needs = None                  # sorted non-owned columns referenced by the rows
        self.agg = None                    # local row -> GLOBAL coarse id (or -1)


def split_rows(A, offsets):
    A = A.tocsr()
    return [RankLevel(a, b, A[a:b]) for a, b in zip(offsets[:-1], offsets[1:])]


def coarsen_once(ranks, n_global):
    """one pairwise pass on every rank + the exchanges; returns the next level's ranks,
    its offsets and the number of map entries every rank had to receive"""
    counts, local_aggs = [], []
    for R in ranks:
        own = R.A[:, R.lo:R.hi].tocsr()                 # matching sees owned columns only
        a = amg_v2.pair_pass(own)
        local_aggs.append(a)
        counts.append(int(a.max()) + 1 if (a >= 0).any() else 0)
    coff = np.concatenate([[0], np.cumsum(counts)])     # the all-gather + exclusive scan
    agg_global = -np.ones(n_global, dtype=np.int64)     # (only a stand-in for the owners' memory)
    for R, a, c in zip(ranks, local_aggs, coff[:-1]):
        R.agg = np.where(a >= 0, a + c, -1)
        agg_global[R.lo:R.hi] = R.agg
    nc = int(coff[-1])
    traffic, next_ranks = [], []
    for r, R in enumerate(ranks):
        cols = R.A.indices
        halo = np.unique(cols[(cols < R.lo) | (cols >= R.hi)])
        R.needs = halo
        halo_map = agg_global[halo]                     # <- received from the owners of `halo`
        traffic.append(int(halo.size))
        # Galerkin rows of this rank: local from here on
        colmap = np.empty(cols.size, dtype=np.int64)
        inside = (cols >= R.lo) & (cols < R.hi)
        colmap[inside] = R.agg[cols[inside] - R.lo]
        colmap[~inside] = halo_map[np.searchsorted(halo, cols[~inside])]
        rows = np.repeat(np.arange(R.hi - R.lo), np.diff(R.A.indptr))
        rowmap = R.agg[rows]
        keep = (rowmap >= 0) & (colmap >= 0)
        C = sp.coo_matrix((R.A.data[keep], (rowmap[keep] - coff[r], colmap[keep])),
                          shape=(counts[r], nc)).tocsr()
        C.sum_duplicates()
        next_ranks.append(RankLevel(coff[r], coff[r + 1], C))
    return next_ranks, coff, traffic


def build(A, offsets, passes=3, passes0=4, coarsest=2000, max_levels=10):
    """hierarchy as a list of levels, each a list of RankLevel; plus per-level traffic"""
    levels = [split_rows(A, offsets)]
    maps, traffic = [], []
    n = A.shape[0]
    while n > coarsest and len(levels) < max_levels:
        ranks, offs, comp = levels[-1], None, None
        cur = ranks
        t_level = []
        for p in range(passes0 if len(levels) == 1 else passes):
            nxt, coff, t = coarsen_once(cur, n if p == 0 else cur_n)
            # composite map fine row -> coarse id of this pass, kept per fine rank
            step = np.concatenate([R.agg for R in cur])
            comp = step if comp is None else np.where(comp >= 0, step[np.maximum(comp, 0)], -1)
            cur, cur_n = nxt, int(coff[-1])
            t_level.append(t)
        if cur_n > 0.7 * n:
            break
        maps.append(comp)
        traffic.append(t_level)
        levels.append(cur)
        n = cur_n
    return levels, maps, traffic


def assemble_global(level):
    return sp.vstack([R.A for R in level]).tocsr()
