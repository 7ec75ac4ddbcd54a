"""CPU worker (gloo, world_size 2+): host-side slab logic + a numpy emulation of
the device algorithm's communication pattern (contiguous halo windows with the
two neighbours, all-reduced dot products), checked against the single-process
oracle.  The arithmetic here is test scaffolding (oracle side); what is under
test is openpnm_b200.dist's partition plan and the halo-window design."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def halo_exchange(dist, torch, v, pad, h, n, rank, world):
    """v is [pad | n | pad]; same windows as b200_halo_exchange (dist.cu)."""
    reqs = []
    own = pad
    if rank > 0:
        reqs.append(dist.isend(torch.from_numpy(v[own:own + h].copy()), rank - 1))
        lo = torch.empty(h, dtype=torch.float64)
        reqs.append(dist.irecv(lo, rank - 1))
    if rank < world - 1:
        reqs.append(dist.isend(torch.from_numpy(v[own + n - h:own + n].copy()), rank + 1))
        hi = torch.empty(h, dtype=torch.float64)
        reqs.append(dist.irecv(hi, rank + 1))
    for r in reqs:
        r.wait()
    if rank > 0:
        v[own - h:own] = lo.numpy()
    if rank < world - 1:
        v[own + n:own + n + h] = hi.numpy()


def allsum(dist, torch, val):
    t = torch.tensor([val], dtype=torch.float64)
    dist.all_reduce(t)
    return float(t[0])


def main():
    import torch
    import torch.distributed as dist
    from openpnm_b200 import dist as bdist
    from oracle import oracle

    rank, world, _ = bdist.init_process_group("gloo")
    case = os.environ.get("B200_CASE", "cubic")
    if case == "cubic":
        shape = (9, 4, 5)
        net = oracle.cubic_network(shape)
        conns, Np = net["conns"], net["Np"]
        ranges = bdist.cubic_slab_ranges(shape, world)
        rb, re = ranges[rank]
        ids, cl = bdist.cubic_slab_throats(shape, rb // 20, re // 20)
        assert np.array_equal(np.sort(ids), np.flatnonzero(bdist.slab_throat_mask(conns, rb, re)))
        assert np.array_equal(cl, conns[ids])
        left, right = net["labels"]["left"], net["labels"]["right"]
    else:
        # irregular network renumbered by x-coordinate (north star: "pore ordering by coordinate")
        pts = np.random.default_rng(0).random((400, 3))
        c0 = oracle.delaunay_network(pts)
        # drop the long hull edges a plain Delaunay adds (the reference trims them
        # with reflect=True, network/_delaunay.py:60): keeps the band local
        c0 = c0[np.linalg.norm(pts[c0[:, 0]] - pts[c0[:, 1]], axis=1) < 0.25]
        perm, inv = bdist.coordinate_order(pts, axis=0)
        conns = np.sort(inv[c0], axis=1)
        coords = pts[perm]
        Np = 400
        assert bdist.bandwidth(conns) < bdist.bandwidth(c0)       # sorting narrows the band
        ranges = bdist.slab_ranges(Np, world)
        rb, re = ranges[rank]
        ids = np.flatnonzero(bdist.slab_throat_mask(conns, rb, re))
        left = np.flatnonzero(coords[:, 0] < 0.1)
        right = np.flatnonzero(coords[:, 0] > 0.9)
    Nt = conns.shape[0]
    g = 10.0 ** np.random.default_rng(1).uniform(-3, 0, Nt)
    bcv = np.full(Np, np.nan)
    bcv[left] = 1.0
    bcv[right] = 0.0
    # ---- reference: single-process oracle
    A, b, f = oracle.apply_bcs(oracle.build_A(conns, g, Np), Np, bcv, None)
    M = oracle.to_csr(A, Np)
    xr, _ = oracle.spsolve(M, b)
    # ---- this rank: rows [rb, re) from the throats touching them only
    n = re - rb
    row, col, dat = oracle.build_A(conns[ids], g[ids], Np)
    keep = (row >= rb) & (row < re)
    import scipy.sparse as sprs
    Lloc = sprs.coo_matrix((dat[keep], (row[keep] - rb, col[keep])), shape=(n, Np)).tocsr()
    # the local pure diagonal must equal the global one on owned rows (every
    # incident throat is present), and f is the all-reduced mean
    Lfull = oracle.to_csr(oracle.build_A(conns, g, Np), Np)
    assert np.allclose(Lloc.toarray(), Lfull[rb:re].toarray(), rtol=1e-14, atol=0)
    fsum = allsum(dist, torch, Lloc[np.arange(n), np.arange(rb, re)].sum())
    assert np.isclose(fsum / Np, f, rtol=1e-13)
    Mloc = M[rb:re]
    h = bdist.bandwidth(conns)
    assert all(e - b_ >= h for b_, e in ranges), "slab thinner than the bandwidth"
    pad = (h + 31) // 32 * 32
    cols = Mloc.indices - rb                       # signed local columns into [pad|n|pad]
    assert cols.min() >= -h and cols.max() < n + h
    d = Mloc[np.arange(n), np.arange(rb, re)].A1 if hasattr(Mloc[0, 0], "A1") else \
        np.asarray(Mloc[np.arange(n), np.arange(rb, re)]).ravel()
    s = np.zeros(n + 2 * pad)
    s[pad:pad + n] = 1.0 / np.sqrt(d)
    halo_exchange(dist, torch, s, pad, h, n, rank, world)
    rows = np.repeat(np.arange(n), np.diff(Mloc.indptr))
    vals = Mloc.data * s[pad + rows] * s[pad + cols]

    def spmv(vec):
        halo_exchange(dist, torch, vec, pad, h, n, rank, world)
        out = np.zeros(n)
        np.add.at(out, rows, vals * vec[pad + cols])
        return out

    bt = s[pad:pad + n] * b[rb:re]
    x = np.zeros(n + 2 * pad)
    single = (np.diff(Mloc.indptr) == 1)
    x[pad:pad + n][single] = bt[single]
    r = bt - spmv(x)
    p = np.zeros(n + 2 * pad)
    p[pad:pad + n] = r
    rr = allsum(dist, torch, r @ r)
    bnorm2 = allsum(dist, torch, b[rb:re] @ b[rb:re])
    its = 0
    while True:
        true2 = allsum(dist, torch, float(np.sum(d * r * r)))
        if true2 <= 1e-26 * bnorm2 or its > 5000:
            break
        ap = spmv(p)
        pap = allsum(dist, torch, p[pad:pad + n] @ ap)
        alpha = rr / pap
        x[pad:pad + n] += alpha * p[pad:pad + n]
        r -= alpha * ap
        rr_new = allsum(dist, torch, r @ r)
        p[pad:pad + n] = r + (rr_new / rr) * p[pad:pad + n]
        rr = rr_new
        its += 1
    xloc = s[pad:pad + n] * x[pad:pad + n]
    parts = [None] * world
    dist.all_gather_object(parts, xloc)
    xg = np.concatenate(parts)
    err = np.max(np.abs(xg - xr)) / np.max(np.abs(xr))
    assert err < 1e-8, err
    if rank == 0:
        print(f"cpu dist ok case={case} world={world} iters={its} err={err:.2e}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
