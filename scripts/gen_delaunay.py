"""Generate the Delaunay network of BASELINE config 4 once, on the host:
`pts = default_rng(seed).random((N, 3))`, pores renumbered by x (the north star's "pore
ordering by coordinate"), edges from `scipy.spatial.Delaunay(...).vertex_neighbor_vertices`
(upper triangle; the same edge set as the reference's `tri_to_am`,
openpnm/_skgraph/tools/_funcs.py:646-669, without its Python loop).

    python scripts/gen_delaunay.py N [seed] -> $B200_DATA_DIR/delaunay_<N>.npz (default /tmp/b200_data)  (conns int32, coords f32)

The 10 M-pore file (~0.75 GB) is git-ignored; it travels to the GPU box with the
snapshot for the config-4 runs and is deleted afterwards."""
import os
import sys
import time

import numpy as np


def generate(N, seed=0):
    from scipy.spatial import Delaunay
    t0 = time.perf_counter()
    pts = np.random.default_rng(seed).random((N, 3))
    pts = pts[np.argsort(pts[:, 0], kind="stable")]
    tri = Delaunay(pts)
    t1 = time.perf_counter()
    indptr, indices = tri.vertex_neighbor_vertices
    rows = np.repeat(np.arange(N, dtype=np.int32), np.diff(indptr))
    up = indices > rows
    conns = np.empty((int(up.sum()), 2), dtype=np.int32)
    conns[:, 0] = rows[up]
    conns[:, 1] = indices[up]
    t2 = time.perf_counter()
    print(f"N={N}: Delaunay {t1 - t0:.1f} s, edges {t2 - t1:.1f} s, Nt={conns.shape[0]} "
          f"({2 * conns.shape[0] / N:.2f} per pore), bandwidth {int((conns[:,1]-conns[:,0]).max())}",
          flush=True)
    return conns, pts.astype(np.float32)


if __name__ == "__main__":
    N = int(sys.argv[1])
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    conns, coords = generate(N, seed)
    out = os.path.join(os.environ.get("B200_DATA_DIR", "/tmp/b200_data"),
                       f"delaunay_{N}.npz")
    np.savez(out, conns=conns, coords=coords)
    print("wrote", out, os.path.getsize(out) / 1e6, "MB")
