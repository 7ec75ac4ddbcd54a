"""Pack the host-generated Delaunay network (scripts/gen_delaunay.py) for the trip to the
GPU box (snapshot limit 512 MiB): throats longer than `cut` mean spacings are dropped here
(the hull slivers of a Delaunay triangulation of points in a box join pores that are far
apart -- OpenPNM users trim them too; this is what leaves isolated pores / small clusters
behind, which the device health check + trim then remove), the rest is stored as CSR-like
(indptr int32, c1 - c0 as uint32) + the x coordinate, compressed.

    python scripts/pack_delaunay.py N [cut=3.0] [blocks=0] -> ship/delaunay_<N>[_mB].npz

blocks > 0 renumbers the pores with openpnm_b200.dist.spatial_order (equal-count bins
along x, Morton curve inside a bin) instead of the plain x-sort."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def unpack(path):
    """-> conns (Nt,2) int64, x (Np,) float32"""
    d = np.load(path)
    indptr, delta = d["indptr"], d["delta"]
    c0 = np.repeat(np.arange(indptr.size - 1, dtype=np.int64), np.diff(indptr))
    conns = np.empty((c0.size, 2), dtype=np.int64)
    conns[:, 0] = c0
    conns[:, 1] = c0 + delta
    return conns, d["x"]


if __name__ == "__main__":
    N = int(sys.argv[1])
    cut = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
    src = os.path.join(os.environ.get("B200_DATA_DIR", "/tmp/b200_data"), f"delaunay_{N}.npz")
    t0 = time.time()
    d = np.load(src)
    conns, coords = d["conns"], d["coords"]
    L = np.linalg.norm(coords[conns[:, 0]].astype(np.float64) - coords[conns[:, 1]], axis=1)
    keep = L < cut * N ** (-1.0 / 3.0)
    c = conns[keep]
    blocks = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    if blocks > 0:
        sys.path.insert(0, ROOT)
        from openpnm_b200 import dist as bdist
        perm, inv = bdist.spatial_order(coords, 0, blocks)
        c = np.sort(inv[c], axis=1)
        coords = coords[perm]
    order = np.lexsort((c[:, 1], c[:, 0]))
    c = c[order]
    indptr = np.zeros(N + 1, dtype=np.int64)
    np.add.at(indptr, c[:, 0].astype(np.int64) + 1, 1)
    indptr = np.cumsum(indptr).astype(np.int32)
    delta = (c[:, 1].astype(np.int64) - c[:, 0]).astype(np.uint32)
    out = os.path.join(ROOT, "ship", f"delaunay_{N}" + (f"_m{blocks}" if blocks else "") + ".npz")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    np.savez_compressed(out, indptr=indptr, delta=delta, x=coords[:, 0].copy(), cut=cut)
    print(f"kept {keep.sum()} of {keep.size} throats (cut {cut} spacings), bandwidth "
          f"{int(delta.max())}, {os.path.getsize(out) / 1e6:.0f} MB, {time.time() - t0:.0f} s")
    c2, x2 = unpack(out)
    assert np.array_equal(c2, c.astype(np.int64))
    assert np.array_equal(x2, coords[:, 0])
