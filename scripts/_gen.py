"""Synthetic-input helpers for the timing scripts (inputs only; nothing here
solves anything)."""
import numpy as np


def delaunay_conns(pts):
    """Upper-triangular, de-duplicated Delaunay edges of `pts` (scipy/Qhull)."""
    from scipy.spatial import Delaunay
    indptr, indices = Delaunay(pts).vertex_neighbor_vertices
    rows = np.repeat(np.arange(len(pts), dtype=np.int64), np.diff(indptr))
    cols = indices.astype(np.int64)
    m = rows < cols
    c = np.stack([rows[m], cols[m]], axis=1)
    return c[np.lexsort((c[:, 1], c[:, 0]))]
