"""Slab (row-block) partition of a network over the GPUs of one box.

One process per GPU.  torch.distributed is the plumbing (rendezvous, the
broadcast of the NCCL unique id, the gather of the result); the per-iteration
halo exchange and dot-product all-reduce happen inside libb200pnm.so on its own
NCCL communicator, enqueued on the same stream as the kernels.

Partitioning: contiguous pore-index ranges.  For
`Cubic` the index is x-major (openpnm/_skgraph/generators/_cubic.py:33-36), so
rank r owns whole y-z planes and its halo is one plane per side.  Any network
whose pores are numbered so that throats only join nearby indices (e.g. a
Delaunay network sorted by x) works the same way: the halo is the contiguous
window of `bandwidth` indices on either side.

The reference's statement of this algorithm is its PETSc wrapper
(openpnm/solvers/_petsc.py:117-128 row-block ownership, :160 cg + jacobi).
"""
import ctypes as C
import os

import numpy as np

__all__ = ["slab_ranges", "cubic_slab_ranges", "slab_throat_mask", "cubic_slab_throats",
           "cubic_slab_throat_ranges",
           "bandwidth", "coordinate_order", "spatial_order", "SlabSolver", "init_process_group",
           "slab_topology_check", "join_slab_clusters"]


# ------------------------------------------------------------ host-side plan
def slab_ranges(Np, nranks, align=1):
    """Balanced contiguous row ranges [(begin, end)] with ends on multiples of
    `align` (a plane size) whenever that leaves every rank non-empty."""
    units = Np // align
    if Np % align or units < nranks:
        align, units = 1, Np
    base, extra = divmod(units, nranks)
    out, b = [], 0
    for r in range(nranks):
        e = b + (base + (1 if r < extra else 0)) * align
        out.append((b, e))
        b = e
    assert b == Np
    return out


def cubic_slab_ranges(shape, nranks):
    nx, ny, nz = (int(s) for s in shape)
    return slab_ranges(nx * ny * nz, nranks, align=ny * nz)


def slab_throat_mask(conns, row_begin, row_end):
    """Throats with at least one end in [row_begin, row_end): exactly the
    throats that contribute to the rows a rank owns."""
    c = np.asarray(conns)
    a = (c[:, 0] >= row_begin) & (c[:, 0] < row_end)
    b = (c[:, 1] >= row_begin) & (c[:, 1] < row_end)
    return a | b


def cubic_slab_throats(shape, x0, x1):
    """Global throat ids and conns of a Cubic network restricted to the throats
    touching the planes x0 <= x < x1, without building the full network.
    Throat numbering follows the reference generator: all z-neighbour pairs,
    then all y-, then all x- (openpnm/_skgraph/generators/_cubic.py:40-42,69-73),
    each block in pore-index order."""
    nx, ny, nz = (int(s) for s in shape)
    ids, tails, heads = [], [], []
    xs = np.arange(x0, x1, dtype=np.int64)
    # z-throats: pore (x,y,z) -- (x,y,z+1), id = x*ny*(nz-1) + y*(nz-1) + z
    if nz > 1:
        X, Y, Z = np.meshgrid(xs, np.arange(ny, dtype=np.int64), np.arange(nz - 1, dtype=np.int64),
                              indexing="ij")
        t = (X * ny + Y) * nz + Z
        ids.append(((X * ny + Y) * (nz - 1) + Z).ravel())
        tails.append(t.ravel())
        heads.append(t.ravel() + 1)
    off = nx * ny * (nz - 1)
    # y-throats: (x,y,z) -- (x,y+1,z), id = off + x*(ny-1)*nz + y*nz + z
    if ny > 1:
        X, Y, Z = np.meshgrid(xs, np.arange(ny - 1, dtype=np.int64), np.arange(nz, dtype=np.int64),
                              indexing="ij")
        t = (X * ny + Y) * nz + Z
        ids.append((off + (X * (ny - 1) + Y) * nz + Z).ravel())
        tails.append(t.ravel())
        heads.append(t.ravel() + nz)
    off += nx * (ny - 1) * nz
    # x-throats: (x,y,z) -- (x+1,y,z) for x0-1 <= x < x1 (both cut planes included)
    if nx > 1:
        xa = np.arange(max(x0 - 1, 0), min(x1, nx - 1), dtype=np.int64)
        X, Y, Z = np.meshgrid(xa, np.arange(ny, dtype=np.int64), np.arange(nz, dtype=np.int64),
                              indexing="ij")
        t = (X * ny + Y) * nz + Z
        ids.append((off + t).ravel())
        tails.append(t.ravel())
        heads.append(t.ravel() + ny * nz)
    ids = np.concatenate(ids) if ids else np.zeros(0, dtype=np.int64)
    conns = np.empty((ids.size, 2), dtype=np.int64)
    if ids.size:
        conns[:, 0] = np.concatenate(tails)
        conns[:, 1] = np.concatenate(heads)
    return ids, conns


def cubic_slab_throat_ranges(shape, x0, x1):
    """The same throats as `cubic_slab_throats` as three contiguous id ranges
    [(begin, end)] (z-, y-, x-throats of the planes): per-throat arrays of a slab are
    three slices of the global arrays, not a gather."""
    nx, ny, nz = (int(s) for s in shape)
    out = []
    if nz > 1:
        out.append((x0 * ny * (nz - 1), x1 * ny * (nz - 1)))
    off = nx * ny * (nz - 1)
    if ny > 1:
        out.append((off + x0 * (ny - 1) * nz, off + x1 * (ny - 1) * nz))
    off += nx * (ny - 1) * nz
    if nx > 1:
        out.append((off + max(x0 - 1, 0) * ny * nz, off + min(x1, nx - 1) * ny * nz))
    return out


def bandwidth(conns):
    """max |c1 - c0|: the halo width a slab partition of this numbering needs."""
    c = np.asarray(conns)
    return int(np.abs(c[:, 1] - c[:, 0]).max()) if c.size else 0


def coordinate_order(coords, axis=0):
    """Permutation that renumbers pores by a coordinate (north star: "pore
    ordering by coordinate").  Returns (perm, inverse): new index k holds old
    pore perm[k]; inverse[old] = new."""
    perm = np.argsort(np.asarray(coords)[:, axis], kind="stable")
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.size)
    return perm, inv


def _spread3(v):
    """10 bits -> every third bit (Morton interleave helper)"""
    v = v.astype(np.uint64) & np.uint64(0x3ff)
    v = (v | (v << np.uint64(16))) & np.uint64(0x30000ff)
    v = (v | (v << np.uint64(8))) & np.uint64(0x300f00f)
    v = (v | (v << np.uint64(4))) & np.uint64(0x30c30c3)
    v = (v | (v << np.uint64(2))) & np.uint64(0x9249249)
    return v


def spatial_order(coords, axis=0, blocks=64):
    """Renumbering for irregular networks that keeps BOTH properties the slab solver
    wants: pores are ordered by `blocks` equal-count bins along `axis` first (so the
    index bandwidth stays ~2/blocks of Np and contiguous index ranges are still
    coordinate slabs), and along a Morton (Z-order) curve inside each bin, so that
    the neighbours of a pore have nearby indices and the gathers of the SpMV hit
    L1/L2 lines that were just used.  Returns (perm, inverse) like coordinate_order."""
    c = np.asarray(coords, dtype=np.float64)
    n = c.shape[0]
    by_axis = np.argsort(c[:, axis], kind="stable")
    bin_of = np.empty(n, dtype=np.int64)
    bin_of[by_axis] = (np.arange(n, dtype=np.int64) * blocks) // max(n, 1)
    lo, hi = c.min(axis=0), c.max(axis=0)
    q = ((c - lo) / np.where(hi > lo, hi - lo, 1.0) * 1023.0).astype(np.uint64)
    # the binned axis varies least inside a bin: give it the low interleave position
    others = [a for a in range(c.shape[1]) if a != axis]
    code = _spread3(q[:, axis])
    for k, a in enumerate(others[:2]):
        code |= _spread3(q[:, a]) << np.uint64(k + 1)
    perm = np.lexsort((code, bin_of))
    inv = np.empty_like(perm)
    inv[perm] = np.arange(n)
    return perm, inv


def _runs(*cols):
    """first index of every run of equal rows in the parallel integer arrays `cols`"""
    n = cols[0].size
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    chg = np.zeros(n, dtype=bool)
    chg[0] = True
    for c in cols:
        chg[1:] |= c[1:] != c[:-1]
    return np.flatnonzero(chg)


def join_slab_clusters(parts, Np):
    """Host half of the slab topology check.  parts[r] = (labels, has_bc, interior_bad)
    of rank r (b200_slab_clusters): labels / has-BC marks of the pores of rank r's two
    cut regions (2*halo entries each, index-aligned with the neighbour's: rank r's high
    region and rank r+1's low region are the same pores).  Two local clusters that
    label the same pore are one cluster; returns the number of clusters (complete ones
    + joined ones) that contain no BC pore -- what `is_fully_connected` tests
    (openpnm/topotools/_topotools.py:993-1024).  The label arrays are run-length
    compressed first (a healthy network is one run per region), so the graph that is
    joined here has a handful of nodes."""
    import scipy.sparse as sprs
    from scipy.sparse import csgraph
    bad = int(sum(int(p[2]) for p in parts))
    labs = [np.asarray(p[0]).astype(np.int64) for p in parts]
    keys, flags, ea, eb = [], [], [], []
    for r, lab in enumerate(labs):
        hb = np.asarray(parts[r][1])
        first = _runs(lab, hb)
        first = first[lab[first] >= 0]
        keys.append(r * np.int64(Np) + lab[first])
        flags.append(hb[first].astype(bool))
        if r + 1 < len(parts):
            h2 = lab.size // 2
            hi, lo = lab[h2:], labs[r + 1][:h2]
            f2 = _runs(hi, lo)
            f2 = f2[(hi[f2] >= 0) & (lo[f2] >= 0)]
            ea.append(r * np.int64(Np) + hi[f2])
            eb.append((r + 1) * np.int64(Np) + lo[f2])
    keys = np.concatenate(keys) if keys else np.zeros(0, dtype=np.int64)
    if keys.size == 0:
        return bad
    flags = np.concatenate(flags)
    nodes, inv = np.unique(keys, return_inverse=True)
    node_bc = np.zeros(nodes.size, dtype=bool)
    np.logical_or.at(node_bc, inv, flags)
    if ea:
        a = np.searchsorted(nodes, np.concatenate(ea))
        b = np.searchsorted(nodes, np.concatenate(eb))
    else:
        a = b = np.zeros(0, dtype=np.int64)
    g = sprs.coo_matrix((np.ones(a.size, dtype=np.int8), (a, b)), shape=(nodes.size, nodes.size))
    ncomp, comp = csgraph.connected_components(g, directed=False)
    comp_bc = np.zeros(ncomp, dtype=bool)
    np.logical_or.at(comp_bc, comp, node_bc)
    return bad + int((~comp_bc).sum())


def slab_topology_check(ctx, conns_local, Np, row_begin, row_end, halo, bc_value, bc_rate):
    """`_validate_topology_health` of a slab run (collective): every rank labels the
    clusters of the throats it holds on its GPU, the labels of the cut regions are
    all-gathered (device tensors over NCCL) and joined on the host.  Returns the number
    of clusters without a BC pore."""
    import torch
    import torch.distributed as dist
    lab, hb, bad = ctx.slab_clusters(conns_local, Np, row_begin, row_end, halo, bc_value, bc_rate)
    ctx.synchronize()
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return join_slab_clusters([(lab.cpu().numpy(), hb.cpu().numpy(), bad)], Np)
    world = dist.get_world_size()
    on_dev = dist.get_backend() == "nccl"
    mine = torch.cat([lab.to(torch.int32), hb.to(torch.int32),
                      torch.tensor([bad], dtype=torch.int32, device=lab.device)])
    if not on_dev:
        mine = mine.cpu()
    allp = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allp, mine)
    n = lab.numel()
    parts = []
    for t in allp:
        h = t.cpu().numpy()
        parts.append((h[:n], h[n:2 * n].astype(np.uint8), int(h[2 * n])))
    return join_slab_clusters(parts, Np)


# --------------------------------------------------------------- processes
def init_process_group(backend=None):
    """Join the torch.distributed job described by RANK / WORLD_SIZE /
    MASTER_ADDR / MASTER_PORT (torchrun).  Returns (rank, world, local_rank)."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


class SlabSolver:
    """Per-rank driver: owns a device Context for rows [row_begin, row_end).

    Usage on every rank:
        ss = SlabSolver(Np, row_begin, row_end, device=local_rank)
        ss.assemble(conns_local, g_local, bc_value, bc_rate)   # global-length bc arrays
        x_local, iters, relres, info = ss.solve(rtol)
    """

    def __init__(self, Np, row_begin, row_end, device=0, fmt="auto"):
        import torch
        import torch.distributed as dist
        from .device import Context
        from . import _lib
        self.Np, self.row_begin, self.row_end = int(Np), int(row_begin), int(row_end)
        self.ctx = Context(device)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self._fmt = {"auto": 0, "csr": _lib.FMT_CSR, "dia": _lib.FMT_DIA}[fmt]
        self.ctx.set_format(self._fmt)
        if self.world > 1:
            buf = (C.c_ubyte * 128)()
            if self.rank == 0:
                rc = self.ctx.lib.b200_nccl_unique_id(C.cast(buf, C.c_void_p))
                if rc != 0:
                    raise RuntimeError("b200_nccl_unique_id failed")
            t = torch.tensor(list(bytes(buf)), dtype=torch.uint8)
            if dist.get_backend() == "nccl":
                t = t.cuda(device)
            dist.broadcast(t, src=0)
            raw = bytes(t.cpu().tolist())
            idbuf = (C.c_ubyte * 128).from_buffer_copy(raw)
            self.ctx._ck(self.ctx.lib.b200_comm_init(self.ctx._h, self.world, self.rank,
                                                     C.cast(idbuf, C.c_void_p)))

    def assemble(self, conns_local, g_local, bc_value=None, bc_rate=None, keep_zero_diag=False,
                 reuse_laplacian=False):
        """reuse_laplacian: the pure Laplacian (and the global diagonal mean) built by the
        previous call are still valid -- only the BCs are applied again (the reference's
        `settings.cache`, _transport.py:108-114)."""
        import torch
        import torch.distributed as dist
        ctx = self.ctx
        if reuse_laplacian and getattr(self, "_nnz_pure", None) is not None:
            f, nnz_sys = ctx.apply_bcs(bc_value, bc_rate, keep_zero_diag=keep_zero_diag)
            return f, self._nnz_pure, nnz_sys
        two_way = (g_local.dim() if hasattr(g_local, "dim") else np.ndim(g_local)) == 2
        if two_way:      # (Nt,2) conductance of AdvectionDiffusion: non-symmetric rows
            nnz = ctx.build_laplacian2(conns_local, g_local, self.Np, self.row_begin, self.row_end)
        else:
            nnz = ctx.build_laplacian(conns_local, g_local, self.Np, self.row_begin, self.row_end)
        self._nnz_pure = nnz
        if self.world > 1:
            s, c = ctx.pure_diag_sum()
            t = torch.tensor([s, float(c)], dtype=torch.float64)
            if dist.get_backend() == "nccl":
                t = t.cuda(ctx.device)
            dist.all_reduce(t)          # f = global mean of the pure diagonal
            t = t.cpu()
            ctx.set_diag_mean(float(t[0] / t[1]))
        f, nnz_sys = ctx.apply_bcs(bc_value, bc_rate, keep_zero_diag=keep_zero_diag)
        return f, nnz, nnz_sys

    def enable_p2p(self):
        """Map the peers' mailboxes and direction vectors (CUDA IPC over NVLink) so
        the PCG iteration communicates from inside its kernels (csrc/p2p.cu).
        Call after assemble(); returns False when not applicable (single rank,
        irregular operator)."""
        import torch
        import torch.distributed as dist
        from . import _lib
        ctx = self.ctx
        if self.world < 2 or self.world > 8:
            return False
        ctx.pcg_prepare(self._fmt)
        fmt, _ = ctx.pcg_format()
        ok = torch.tensor([1 if fmt == _lib.FMT_DIA else 0], dtype=torch.int64)
        dev = ctx.tdev if dist.get_backend() == "nccl" else torch.device("cpu")
        ok = ok.to(dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            return False
        def all_ok(flag):
            t = torch.tensor([1 if flag else 0], dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            return bool(int(t.item()))

        buf = (C.c_ubyte * 128)()
        ok = ctx.lib.b200_p2p_export(ctx._h, C.cast(buf, C.c_void_p)) == 0
        if not all_ok(ok):            # every rank takes the same decision
            return False
        mine = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device=dev)
        allh = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(allh, mine)
        rows = torch.tensor([self.row_end - self.row_begin], dtype=torch.int64, device=dev)
        allr = [torch.empty_like(rows) for _ in range(self.world)]
        dist.all_gather(allr, rows)
        raw = b"".join(bytes(t.cpu().tolist()) for t in allh)
        hbuf = (C.c_ubyte * len(raw)).from_buffer_copy(raw)
        rbuf = (C.c_int64 * self.world)(*[int(t.item()) for t in allr])
        ok = ctx.lib.b200_p2p_attach(ctx._h, C.cast(hbuf, C.c_void_p),
                                     C.cast(rbuf, C.c_void_p)) == 0
        if not all_ok(ok):            # e.g. IPC not permitted: stay on the NCCL back end
            self.disable_p2p()
            return False
        return True

    def disable_p2p(self):
        self.ctx._ck(self.ctx.lib.b200_p2p_disable(self.ctx._h))

    def solve(self, rtol=1e-10, maxiter=0, x0=None):
        return self.ctx.pcg(x0=x0, rtol=rtol, maxiter=maxiter)

    def rate(self, x_global_dev, pores):
        """Sum of the net rates of the pores this rank owns (caller all-reduces)."""
        pores = np.asarray(pores, dtype=np.int64)
        mine = pores[(pores >= self.row_begin) & (pores < self.row_end)]
        if mine.size == 0:
            return 0.0
        return float(self.ctx.rate_pores(x_global_dev, mine).sum().item())

    def gather(self, x_local):
        """All ranks receive the full solution (like PETSc's Scatter().toAll,
        openpnm/solvers/_petsc.py:225)."""
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return x_local
        n = torch.tensor([x_local.numel()], dtype=torch.int64, device=x_local.device)
        sizes = [torch.zeros_like(n) for _ in range(self.world)]
        self.ctx.synchronize()
        dist.all_gather(sizes, n)
        sizes = [int(s.item()) for s in sizes]
        m = max(sizes)
        mine = torch.zeros(m, dtype=x_local.dtype, device=x_local.device)
        mine[:x_local.numel()] = x_local
        parts = [torch.empty(m, dtype=x_local.dtype, device=x_local.device) for _ in sizes]
        dist.all_gather(parts, mine)
        return torch.cat([p[:k] for p, k in zip(parts, sizes)])
