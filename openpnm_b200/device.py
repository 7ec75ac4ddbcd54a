"""Device-side session: a `b200_ctx` plus the torch tensors handed to it.

PyTorch is used for exactly three things here: owning device buffers
(`torch.Tensor.data_ptr()` is what crosses the C-ABI), providing the CUDA
stream the library enqueues on, and -- in `dist.py` -- torch.distributed
for the rendezvous.  All arithmetic happens inside libb200pnm.so.
"""
import ctypes as C

import numpy as np

from . import _lib


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("openpnm_b200 needs a CUDA device (no CPU fallback)")
    return torch


class Context:
    """One GPU, one stream, one assembled transport problem."""

    def __init__(self, device=0):
        torch = _torch()
        self.lib = _lib.load()
        self.device = int(device)
        self.tdev = torch.device("cuda", self.device)
        self._h = C.c_void_p()
        rc = self.lib.b200_create(self.device, C.byref(self._h))
        if rc != 0:
            raise _lib.B200Error(rc, "b200_create failed (is a CUDA device visible?)")
        with torch.cuda.device(self.tdev):
            self.stream = torch.cuda.Stream(device=self.tdev)
        self._ck(self.lib.b200_set_stream(self._h, C.c_void_p(self.stream.cuda_stream)))
        self._keep = {}          # tensors the library holds raw pointers to
        self.n = 0
        self.row_begin = 0

    # ------------------------------------------------------------ plumbing
    def _ck(self, rc):
        if rc != 0:
            msg = self.lib.b200_last_error(self._h)
            raise _lib.B200Error(rc, msg.decode() if msg else "")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.b200_destroy(self._h)
            self._h = C.c_void_p()
        self._keep = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def to_device(self, a, dtype=None):
        """numpy / tensor -> contiguous device tensor on this context's stream."""
        torch = _torch()
        if isinstance(a, torch.Tensor):
            t = a
            if t.is_cuda:
                # produced on the caller's stream: order this context's stream after it
                cur = torch.cuda.current_stream(t.device)
                if cur != self.stream:
                    self.stream.wait_stream(cur)
                    t.record_stream(self.stream)
            with torch.cuda.stream(self.stream):
                if dtype is not None and t.dtype != dtype:
                    t = t.to(dtype)
                if t.device != self.tdev:
                    t = t.to(self.tdev, non_blocking=True)
                return t.contiguous()
        arr = np.ascontiguousarray(a)
        t = torch.from_numpy(arr)
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
        with torch.cuda.stream(self.stream):
            return t.to(self.tdev, non_blocking=True)

    def empty(self, n, dtype):
        torch = _torch()
        with torch.cuda.stream(self.stream):
            return torch.empty(int(n), dtype=dtype, device=self.tdev)

    def synchronize(self):
        self._ck(self.lib.b200_synchronize(self._h))

    def _publish(self, *ts):
        """Results are written on this context's stream; make the caller's current
        torch stream wait for them so plain torch code can use them right away."""
        torch = _torch()
        cur = torch.cuda.current_stream(self.tdev)
        if cur != self.stream:
            cur.wait_stream(self.stream)
            for t in ts:
                if t is not None:
                    t.record_stream(cur)
        return ts[0] if len(ts) == 1 else ts

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p()

    # ------------------------------------------------------------ assembly
    def build_laplacian(self, conns, g, Np, row_begin=0, row_end=None):
        torch = _torch()
        conns = self.to_device(conns, torch.int64)
        g = self.to_device(g, torch.float64)
        Nt = int(g.numel())
        if conns.numel() != 2 * Nt:
            raise ValueError("conns must be (Nt, 2)")
        row_end = int(Np) if row_end is None else int(row_end)
        nnz = C.c_int64()
        self._ck(self.lib.b200_build_laplacian(self._h, self._p(conns), self._p(g), int(Np), Nt,
                                               int(row_begin), row_end, C.byref(nnz)))
        self._keep["conns"], self._keep["g"] = conns, g
        self.n = row_end - int(row_begin)
        self.row_begin = int(row_begin)
        return nnz.value

    def build_laplacian2(self, conns, g2, Np, row_begin=0, row_end=None):
        """(Nt,2) weights -> non-symmetric Laplacian (AdvectionDiffusion); rows
        [row_begin, row_end) from the throats touching them in a slab run."""
        torch = _torch()
        conns = self.to_device(conns, torch.int64)
        g2 = self.to_device(np.ascontiguousarray(g2, dtype=np.float64), torch.float64)
        Nt = int(g2.numel() // 2)
        if conns.numel() != 2 * Nt:
            raise ValueError("conns must be (Nt, 2) and g2 (Nt, 2)")
        row_end = int(Np) if row_end is None else int(row_end)
        nnz = C.c_int64()
        self._ck(self.lib.b200_build_laplacian2_rows(self._h, self._p(conns), self._p(g2), int(Np),
                                                     Nt, int(row_begin), row_end, C.byref(nnz)))
        self._keep["conns"], self._keep["g"] = conns, g2
        self.n = row_end - int(row_begin)
        self.row_begin = int(row_begin)
        return nnz.value

    def add_to_diagonal(self, add):
        torch = _torch()
        add = self.to_device(add, torch.float64)
        self._ck(self.lib.b200_add_to_diagonal(self._h, self._p(add)))

    def check_topology(self, conns, Np, bc_value=None, bc_rate=None):
        """(number of clusters, number of clusters without any BC pore)"""
        torch = _torch()
        conns = self.to_device(conns, torch.int64)
        bv = self.to_device(bc_value, torch.float64) if bc_value is not None else None
        br = self.to_device(bc_rate, torch.float64) if bc_rate is not None else None
        nc, nb = C.c_int64(), C.c_int64()
        self._ck(self.lib.b200_check_topology(self._h, self._p(conns), int(conns.numel() // 2),
                                              int(Np), self._p(bv), self._p(br), C.byref(nc),
                                              C.byref(nb)))
        return nc.value, nb.value

    def cluster_labels(self, conns, Np):
        """(labels int32 device tensor: smallest pore index of each pore's cluster, n_clusters)"""
        torch = _torch()
        conns = self.to_device(conns, torch.int64)
        lab = self.empty(Np, torch.int32)
        nc = C.c_int64()
        self._ck(self.lib.b200_cluster_labels(self._h, self._p(conns), int(conns.numel() // 2),
                                              int(Np), self._p(lab), C.byref(nc)))
        return self._publish(lab), nc.value

    def slab_clusters(self, conns_local, Np, row_begin, row_end, halo, bc_value=None, bc_rate=None):
        """Per-rank part of the slab topology check: (labels of the two cut regions,
        their has-BC marks, number of complete clusters without a BC pore)."""
        torch = _torch()
        conns = self.to_device(conns_local, torch.int64)
        bv = self.to_device(bc_value, torch.float64) if bc_value is not None else None
        br = self.to_device(bc_rate, torch.float64) if bc_rate is not None else None
        lab = self.empty(max(4 * int(halo), 1), torch.int32)
        hb = self.empty(max(4 * int(halo), 1), torch.uint8)
        bad = C.c_int64()
        self._ck(self.lib.b200_slab_clusters(self._h, self._p(conns), int(conns.numel() // 2),
                                             int(Np), int(row_begin), int(row_end), int(halo),
                                             self._p(bv), self._p(br), self._p(lab), self._p(hb),
                                             C.byref(bad)))
        self._publish(lab, hb)
        return lab[:4 * int(halo)], hb[:4 * int(halo)], bad.value

    def network_health(self, conns, Np, want_arrays=True):
        """check_network_health on the device: dict of counts + device arrays
        (cluster_size int32, isolated uint8, looped uint8, labels int32)."""
        torch = _torch()
        conns = self.to_device(conns, torch.int64)
        Nt = int(conns.numel() // 2)
        cs = self.empty(Np, torch.int32) if want_arrays else None
        iso = self.empty(Np, torch.uint8) if want_arrays else None
        lp = self.empty(max(Nt, 1), torch.uint8) if want_arrays else None
        lab = self.empty(Np, torch.int32) if want_arrays else None
        cnt = (C.c_int64 * 6)()
        self._ck(self.lib.b200_network_health(self._h, self._p(conns), Nt, int(Np), self._p(cs),
                                              self._p(iso), self._p(lp), self._p(lab), cnt))
        if want_arrays:
            self._publish(cs, iso, lp, lab)
        return dict(n_clusters=int(cnt[0]), n_isolated=int(cnt[1]), n_disconnected=int(cnt[2]),
                    n_looped=int(cnt[3]), largest_cluster=int(cnt[5]), cluster_size=cs,
                    isolated=iso, looped=None if lp is None else lp[:Nt], labels=lab)

    def trim(self, conns, Np, pore_keep, throat_keep=None):
        """topotools.trim on the device: (conns_new (Nt_new,2) int64 device, pore_map,
        throat_map, Np_new)."""
        torch = _torch()
        conns = self.to_device(conns, torch.int64)
        Nt = int(conns.numel() // 2)
        pk = self.to_device(pore_keep, torch.uint8)
        tk = self.to_device(throat_keep, torch.uint8) if throat_keep is not None else None
        pmap = self.empty(Np, torch.int64)
        tmap = self.empty(max(Nt, 1), torch.int64)
        out = self.empty(max(2 * Nt, 2), torch.int64)
        npn, ntn = C.c_int64(), C.c_int64()
        self._ck(self.lib.b200_trim(self._h, self._p(conns), Nt, int(Np), self._p(pk), self._p(tk),
                                    self._p(pmap), self._p(tmap), self._p(out), C.byref(npn),
                                    C.byref(ntn)))
        self._publish(pmap, tmap, out)
        return out[:2 * ntn.value].view(-1, 2), pmap, tmap[:Nt], npn.value

    def apply_bcs(self, bc_value=None, bc_rate=None, keep_zero_diag=False):
        torch = _torch()
        bv = self.to_device(bc_value, torch.float64) if bc_value is not None else None
        br = self.to_device(bc_rate, torch.float64) if bc_rate is not None else None
        f, nnz = C.c_double(), C.c_int64()
        self._ck(self.lib.b200_apply_bcs(self._h, self._p(bv), self._p(br),
                                         1 if keep_zero_diag else 0, C.byref(f), C.byref(nnz)))
        return f.value, nnz.value

    def pure_diag_sum(self):
        s, c = C.c_double(), C.c_int64()
        self._ck(self.lib.b200_pure_diag_sum(self._h, C.byref(s), C.byref(c)))
        return s.value, c.value

    def set_diag_mean(self, f):
        self._ck(self.lib.b200_set_diag_mean(self._h, float(f)))

    def load_coo(self, row, col, data, n):
        torch = _torch()
        row = self.to_device(row, torch.int32)
        col = self.to_device(col, torch.int32)
        data = self.to_device(data, torch.float64)
        nnz = C.c_int64()
        self._ck(self.lib.b200_load_coo(self._h, self._p(row), self._p(col), self._p(data),
                                        int(data.numel()), int(n), C.byref(nnz)))
        self.n = int(n)
        return nnz.value

    def load_csr(self, indptr, indices, data, n):
        torch = _torch()
        indptr = self.to_device(indptr, torch.int32)
        indices = self.to_device(indices, torch.int32)
        data = self.to_device(data, torch.float64)
        self._ck(self.lib.b200_load_csr(self._h, self._p(indptr), self._p(indices), self._p(data),
                                        int(n), int(data.numel())))
        self.n = int(n)

    def set_b(self, b):
        torch = _torch()
        b = self.to_device(b, torch.float64)
        self._ck(self.lib.b200_set_b(self._h, self._p(b)))

    def get_b(self):
        torch = _torch()
        out = self.empty(self.n, torch.float64)
        self._ck(self.lib.b200_get_b(self._h, self._p(out)))
        return self._publish(out)

    def csr_shape(self, which):
        n, nnz = C.c_int64(), C.c_int64()
        self._ck(self.lib.b200_csr_shape(self._h, which, C.byref(n), C.byref(nnz)))
        return n.value, nnz.value

    def export_csr(self, which=_lib.MAT_SYSTEM):
        """(indptr, indices, data) as device tensors."""
        torch = _torch()
        n, nnz = self.csr_shape(which)
        ip = self.empty(n + 1, torch.int32)
        ix = self.empty(max(nnz, 1), torch.int32)
        dt = self.empty(max(nnz, 1), torch.float64)
        self._ck(self.lib.b200_export_csr(self._h, which, self._p(ip), self._p(ix), self._p(dt)))
        self._publish(ip, ix, dt)
        return ip, ix[:nnz], dt[:nnz]

    def export_csr_scipy(self, which=_lib.MAT_SYSTEM, ncols=None):
        import scipy.sparse as sprs
        ip, ix, dt = self.export_csr(which)
        n = ip.numel() - 1
        return sprs.csr_matrix((dt.cpu().numpy(), ix.cpu().numpy(), ip.cpu().numpy()),
                               shape=(n, ncols or n))

    # -------------------------------------------------------------- solve
    def spmv(self, which, x):
        torch = _torch()
        x = self.to_device(x, torch.float64)
        y = self.empty(self.n, torch.float64)
        self._ck(self.lib.b200_spmv(self._h, which, self._p(x), self._p(y)))
        return self._publish(y)

    def set_format(self, fmt=0):
        self._ck(self.lib.b200_set_format(self._h, int(fmt)))

    def pcg_prepare(self, fmt=0):
        self._ck(self.lib.b200_pcg_prepare(self._h, int(fmt)))

    def set_precond(self, kind="jacobi"):
        self._ck(self.lib.b200_set_precond(self._h, _lib.PRECOND_CODES[kind]))

    def amg_setup(self):
        """Build the multigrid hierarchy now; returns [(n, nnz), ...] per level."""
        nl = C.c_int()
        self._ck(self.lib.b200_amg_setup(self._h, C.byref(nl)))
        out = []
        for l in range(nl.value):
            n, nnz = C.c_int64(), C.c_int64()
            self._ck(self.lib.b200_amg_level_shape(self._h, l, C.byref(n), C.byref(nnz)))
            out.append((n.value, nnz.value))
        return out

    def amg_level_info(self, level):
        """dict(n, nnz, n_global, first_row, partitioned, halo_lo, halo_hi, coarse_offset)
        of one level of the hierarchy (slab runs: what this rank holds)."""
        out = (C.c_int64 * 8)()
        self._ck(self.lib.b200_amg_level_info(self._h, int(level), out))
        keys = ("n", "nnz", "n_global", "first_row", "partitioned", "halo_lo", "halo_hi",
                "coarse_offset")
        return dict(zip(keys, (int(v) for v in out)))

    def amg_export_level(self, level, with_agg=True):
        """(indptr, indices, data, agg) device tensors of one level (agg None if
        the level is the last one)."""
        torch = _torch()
        n, nnz = C.c_int64(), C.c_int64()
        self._ck(self.lib.b200_amg_level_shape(self._h, level, C.byref(n), C.byref(nnz)))
        ip = self.empty(n.value + 1, torch.int32)
        ix = self.empty(max(nnz.value, 1), torch.int32)
        dt = self.empty(max(nnz.value, 1), torch.float64)
        ag = self.empty(n.value, torch.int32) if with_agg else None
        self._ck(self.lib.b200_amg_export_level(self._h, level, self._p(ip), self._p(ix),
                                                self._p(dt), self._p(ag)))
        self._publish(ip, ix, dt, ag)
        return ip, ix[:nnz.value], dt[:nnz.value], ag

    def pcg_format(self):
        f, nd = C.c_int(), C.c_int()
        self._ck(self.lib.b200_pcg_format(self._h, C.byref(f), C.byref(nd)))
        return f.value, nd.value

    def pcg(self, x0=None, rtol=1e-8, atol=0.0, maxiter=0, out=None):
        torch = _torch()
        x0 = self.to_device(x0, torch.float64) if x0 is not None else None
        x = out if out is not None else self.empty(self.n, torch.float64)
        it, rel, info = C.c_int64(), C.c_double(), C.c_int()
        self._ck(self.lib.b200_pcg(self._h, self._p(x0), float(rtol), float(atol), int(maxiter),
                                   self._p(x), C.byref(it), C.byref(rel), C.byref(info)))
        return self._publish(x), it.value, rel.value, info.value

    def set_conductance_model(self, kind, conns, Np, poly, size_factors, throat_prop=None,
                              interp="mean", bc_value=None, bc_rate=None, keep_zero_diag=False):
        """Conductance that follows the quantity, evaluated on the device inside the Picard
        loop (b200_set_conductance_model).  kind: 'hydraulic' (generic_hydraulic) or
        'poisson' (_poisson_conductance); poly: coefficients a_k of the pore property."""
        torch = _torch()
        conns = self.to_device(conns, torch.int64)
        Nt = int(conns.numel() // 2)
        sf = np.ascontiguousarray(size_factors, dtype=np.float64)
        cols = 3 if sf.ndim == 2 else 1
        if sf.shape[0] != Nt or (sf.ndim == 2 and sf.shape[1] != 3):
            raise ValueError("size factors must be (Nt,) or (Nt, 3)")
        sf_d = self.to_device(sf.reshape(-1), torch.float64)
        tp = self.to_device(np.ascontiguousarray(throat_prop, dtype=np.float64),
                            torch.float64) if throat_prop is not None else None
        bv = self.to_device(bc_value, torch.float64) if bc_value is not None else None
        br = self.to_device(bc_rate, torch.float64) if bc_rate is not None else None
        a = np.ascontiguousarray(poly, dtype=np.float64)
        self._ck(self.lib.b200_set_conductance_model(
            self._h, {"hydraulic": 0, "poisson": 1}[kind], self._p(conns), Nt, int(Np),
            _lib.np_ptr(a), int(a.size), {"mean": 0, "min": 1, "max": 2}[interp], self._p(tp),
            self._p(sf_d), cols, self._p(bv), self._p(br), 1 if keep_zero_diag else 0))
        self._keep["gmodel"] = (conns, sf_d, tp, bv, br)
        self._keep["conns"] = conns
        self._gmodel_Nt = Nt

    def clear_conductance_model(self):
        self._ck(self.lib.b200_clear_conductance_model(self._h))
        self._keep.pop("gmodel", None)

    def update_conductance(self, x):
        """g(x) -> K1 -> K2 with the model that is set; returns nothing (the system on the
        device is the re-assembled one)."""
        torch = _torch()
        x = self.to_device(x, torch.float64)
        self._ck(self.lib.b200_update_conductance(self._h, self._p(x)))
        self.n = int(x.numel())

    def conductance_values(self):
        torch = _torch()
        g = self.empty(self._gmodel_Nt, torch.float64)
        self._ck(self.lib.b200_conductance_values(self._h, self._p(g)))
        self._keep["g"] = g
        return self._publish(g)

    def clear_sources(self):
        self._ck(self.lib.b200_clear_sources(self._h))
        self._keep["src"] = []

    def add_source(self, kind, mask, p1, p2, p3=None):
        torch = _torch()
        n = self.n

        def vec(v):
            if v is None:
                return None
            a = np.asarray(v, dtype=np.float64)
            if a.ndim == 0:
                a = np.full(n, float(a))
            return self.to_device(a, torch.float64)

        m = self.to_device(np.ascontiguousarray(mask, dtype=np.uint8), torch.uint8)
        t1, t2, t3 = vec(p1), vec(p2), vec(p3)
        self._ck(self.lib.b200_add_source(self._h, int(kind), self._p(m), self._p(t1),
                                          self._p(t2), self._p(t3)))
        self._keep.setdefault("src", []).append((m, t1, t2, t3))

    def picard(self, x0=None, relaxation=1.0, newton_maxiter=5000, f_rtol=1e-6, x_rtol=1e-6,
               cg_rtol=1e-10, cg_maxiter=0, update=None):
        """b200_picard; with `update(x_dev_tensor)` the loop calls back before every
        linearisation (b200_picard_cb) so host-side models can be regenerated."""
        torch = _torch()
        x0 = self.to_device(x0, torch.float64) if x0 is not None else None
        x = self.empty(self.n, torch.float64)
        ni, conv, cgt, info = C.c_int64(), C.c_int(), C.c_int64(), C.c_int()
        if update is None:
            self._ck(self.lib.b200_picard(self._h, self._p(x0), float(relaxation),
                                          int(newton_maxiter), float(f_rtol), float(x_rtol),
                                          float(cg_rtol), int(cg_maxiter), self._p(x),
                                          C.byref(ni), C.byref(conv), C.byref(cgt),
                                          C.byref(info)))
            return self._publish(x), ni.value, bool(conv.value), cgt.value, info.value
        raised = []

        def hook(_user, _xptr):
            try:
                update(x)           # the library iterates in place on this tensor
                return 0
            except BaseException as e:      # noqa: BLE001 - re-raised below
                raised.append(e)
                return 1
        cb = _lib.UPDATE_FN(hook)
        rc = self.lib.b200_picard_cb(self._h, self._p(x0), float(relaxation),
                                     int(newton_maxiter), float(f_rtol), float(x_rtol),
                                     float(cg_rtol), int(cg_maxiter), self._p(x), C.byref(ni),
                                     C.byref(conv), C.byref(cgt), C.byref(info),
                                     C.cast(cb, C.c_void_p), None)
        if raised:
            raise raised[0]
        self._ck(rc)
        return self._publish(x), ni.value, bool(conv.value), cgt.value, info.value

    def transient_rk45(self, x0, volume, t0, t1, rtol=1e-6, atol=1e-6, saveat=None,
                       capacity=None):
        """(times ndarray, states device tensor [len(times), n], nsteps, nfev, status)"""
        torch = _torch()
        x0 = self.to_device(x0, torch.float64)
        vol = self.to_device(volume, torch.float64)
        if saveat is None:
            sv = np.zeros(0)
            cap = int(capacity or 4096)
        else:
            sv = np.ascontiguousarray(saveat, dtype=np.float64)
            cap = int(sv.size)
        with torch.cuda.stream(self.stream):
            out = torch.empty((max(cap, 1), self.n), dtype=torch.float64, device=self.tdev)
        times = np.zeros(max(cap, 1))
        nout, nsteps, nfev, status = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int()
        self._ck(self.lib.b200_transient_rk45(
            self._h, self._p(x0), self._p(vol), float(t0), float(t1), float(rtol), float(atol),
            _lib.np_ptr(sv) if sv.size else C.c_void_p(), int(sv.size), max(cap, 1),
            self._p(out), _lib.np_ptr(times), C.byref(nout), C.byref(nsteps), C.byref(nfev),
            C.byref(status)))
        k = nout.value
        self._publish(out)
        return times[:k].copy(), out[:k], nsteps.value, nfev.value, status.value

    def rate_pores(self, x, pores):
        torch = _torch()
        x = self.to_device(x, torch.float64)
        pores = self.to_device(np.ascontiguousarray(pores, dtype=np.int64), torch.int64)
        out = self.empty(pores.numel(), torch.float64)
        self._ck(self.lib.b200_rate_pores(self._h, self._p(x), self._p(pores),
                                          int(pores.numel()), self._p(out)))
        return self._publish(out)

    def rate_throats(self, x, throats):
        torch = _torch()
        x = self.to_device(x, torch.float64)
        throats = self.to_device(np.ascontiguousarray(throats, dtype=np.int64), torch.int64)
        out = self.empty(throats.numel(), torch.float64)
        self._ck(self.lib.b200_rate_throats(self._h, self._p(self._keep["conns"]),
                                            self._p(self._keep["g"]), self._p(x),
                                            self._p(throats), int(throats.numel()), self._p(out)))
        return self._publish(out)

    def rate_throats_explicit(self, x, conns_sel, g_sel):
        """|g (x[c1] - x[c0])| (g (k,)) or |g0 x[c1] - g1 x[c0]| (g (k,2)) for an explicit
        list of throats given by their conns and conductances (slab runs, where the context
        holds the rank's throats only; AdvectionDiffusion's two-way conductance)."""
        torch = _torch()
        x = self.to_device(x, torch.float64)
        conns = self.to_device(np.ascontiguousarray(conns_sel, dtype=np.int64), torch.int64)
        g = np.ascontiguousarray(g_sel, dtype=np.float64)
        k = int(conns.numel() // 2)
        gd = self.to_device(g.reshape(-1), torch.float64)
        idx = self.to_device(np.arange(k, dtype=np.int64), torch.int64)
        out = self.empty(k, torch.float64)
        fn = self.lib.b200_rate_throats2 if g.ndim == 2 else self.lib.b200_rate_throats
        self._ck(fn(self._h, self._p(conns), self._p(gd), self._p(x), self._p(idx), k,
                    self._p(out)))
        return self._publish(out)

    def source_values(self, index, x):
        """(S1, S2, rate) device tensors of registered source number `index` at x"""
        torch = _torch()
        x = self.to_device(x, torch.float64)
        S1, S2, r = (self.empty(self.n, torch.float64) for _ in range(3))
        self._ck(self.lib.b200_source_values(self._h, int(index), self._p(x), self._p(S1),
                                             self._p(S2), self._p(r)))
        return self._publish(S1, S2, r)

    def time_kernel(self, which, reps=20):
        ms = C.c_double()
        self._ck(self.lib.b200_time_kernel(self._h, int(which), int(reps), C.byref(ms)))
        return ms.value

    def operator_bytes(self):
        """(bytes one K_A launch moves in the format in use, stored entries)"""
        b, e = C.c_int64(), C.c_int64()
        self._ck(self.lib.b200_operator_bytes(self._h, C.byref(b), C.byref(e)))
        return b.value, e.value

    def stats(self):
        s = _lib.Stats()
        self._ck(self.lib.b200_get_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in _lib.Stats._fields_}

    # ------------------------------------------------- host-buffer entry
    def solve_coo_host(self, row, col, data, n, b, x0, rtol, atol, maxiter):
        row = _lib.as_c(row, np.int32)
        col = _lib.as_c(col, np.int32)
        data = _lib.as_c(data, np.float64)
        b = _lib.as_c(b, np.float64)
        x0 = _lib.as_c(x0, np.float64) if x0 is not None else None
        x = np.empty(int(n), dtype=np.float64)
        it, rel, info = C.c_int64(), C.c_double(), C.c_int()
        self._ck(self.lib.b200_solve_coo_h(
            self._h, _lib.np_ptr(row), _lib.np_ptr(col), _lib.np_ptr(data), int(data.size),
            int(n), _lib.np_ptr(b), _lib.np_ptr(x0) if x0 is not None else C.c_void_p(),
            float(rtol), float(atol), int(maxiter), _lib.np_ptr(x), C.byref(it), C.byref(rel),
            C.byref(info)))
        self.n = int(n)
        return x, it.value, rel.value, info.value


def pinned_empty(shape, dtype):
    """numpy array backed by page-locked host memory (for async H2D/D2H)."""
    torch = _torch()
    tdt = {np.dtype(np.float64): torch.float64, np.dtype(np.int64): torch.int64,
           np.dtype(np.int32): torch.int32, np.dtype(np.uint8): torch.uint8}[np.dtype(dtype)]
    t = torch.empty(tuple(np.atleast_1d(shape).tolist()), dtype=tdt, pin_memory=True)
    return t.numpy()


def pinned_copy(a):
    a = np.asarray(a)
    out = pinned_empty(a.shape, a.dtype)
    out[...] = a
    return out
