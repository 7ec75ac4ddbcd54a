"""Host-side mirror of `openpnm.algorithms.{Transport, ReactiveTransport,
StokesFlow, FickianDiffusion, OhmicConduction, FourierConduction}` whose
assembly, boundary conditions, source linearisation, solve, Picard loop and
`rate()` all run on the GPU through libb200pnm.so.

Same method names, argument meaning and error behaviour as the reference
(file:line under /root/reference):
    openpnm/algorithms/_transport.py:45-74      settings, __init__
    openpnm/algorithms/_transport.py:171-207    run
    openpnm/algorithms/_transport.py:229-257    validators
    openpnm/algorithms/_transport.py:259-379    rate, set_*_BC, clear_*_BCs
    openpnm/algorithms/_algorithm.py:105-225    set_BC bookkeeping
    openpnm/algorithms/_reactive_transport.py:43-123,246-256  sources
    openpnm/algorithms/_stokes_flow.py:21-32 etc.  settings-only subclasses
`network` / `phase` may be real OpenPNM objects or the containers in
`openpnm_b200.network`.
"""
import logging

import numpy as np

from . import _lib
from .solvers import B200PCG

__all__ = ["Transport", "ReactiveTransport", "StokesFlow", "FickianDiffusion",
           "OhmicConduction", "FourierConduction", "SolutionContainer", "SteadyStateSolution",
           "AdvectionDiffusion", "TransientReactiveTransport", "TransientFickianDiffusion",
           "TransientOhmicConduction", "TransientFourierConduction",
           "TransientAdvectionDiffusion", "TransientSolution"]

logger = logging.getLogger(__name__)

_SRC_KINDS = {
    "linear": (_lib.SRC_LINEAR, ("A1", "A2", None), (0.0, 0.0, 0.0)),
    "power_law": (_lib.SRC_POWER_LAW, ("A1", "A2", "A3"), (0.0, 0.0, 0.0)),
    "standard_kinetics": (_lib.SRC_STANDARD_KINETICS, ("prefactor", "exponent", None),
                          (None, None, 0.0)),
}


# conduit-conductance models with a device kernel: name -> (kind, argument names of the
# pore property, the throat property and the size factors, their defaults)
_COND_KINDS = {
    "generic_hydraulic": ("hydraulic", ("pore_viscosity", "throat_viscosity", "size_factors"),
                          ("pore.viscosity", "throat.viscosity", "throat.hydraulic_size_factors")),
    "generic_diffusive": ("poisson", ("pore_diffusivity", "throat_diffusivity", "size_factors"),
                          ("pore.diffusivity", "throat.diffusivity",
                           "throat.diffusive_size_factors")),
    "generic_thermal": ("poisson", ("pore_conductivity", "throat_conductivity", "size_factors"),
                        ("pore.thermal_conductivity", "throat.thermal_conductivity",
                         "throat.diffusive_size_factors")),
    "generic_electrical": ("poisson", ("pore_conductivity", "throat_conductivity", "size_factors"),
                           ("pore.electrical_conductivity", "throat.electrical_conductivity",
                            "throat.diffusive_size_factors")),
}

_COPY_POOL = None


def _host_copy(src, min_chunk=1 << 21):
    """A fresh array with the contents of `src` (the page-locked staging buffer is
    reused by the next run, the caller keeps the result): large arrays are copied by a
    few threads -- numpy releases the GIL inside copyto -- because one core moves only
    ~5 GB/s (0.11 s for the 64 M doubles of a 400^3 solution)."""
    global _COPY_POOL
    n = src.size
    dst = np.empty_like(src)
    parts = min(8, n // min_chunk)
    if parts < 2:
        np.copyto(dst, src)
        return dst
    if _COPY_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _COPY_POOL = ThreadPoolExecutor(max_workers=8)
    edges = np.linspace(0, n, parts + 1).astype(np.int64)
    list(_COPY_POOL.map(lambda k: np.copyto(dst[edges[k]:edges[k + 1]], src[edges[k]:edges[k + 1]]),
                        range(parts)))
    return dst


class _Stages:
    """Host-side stage timer of run(): active only with B200_PROFILE=1 (it synchronises
    the stream at every mark, which the normal path must not do)."""

    def __init__(self, ctx):
        import os
        self.on = os.environ.get("B200_PROFILE") == "1"
        self.ctx, self.t, self.out = ctx, None, {}
        if self.on:
            import time
            self.clock = time.perf_counter
            ctx.synchronize()
            self.t = self.clock()

    def mark(self, name):
        if self.on:
            self.ctx.synchronize()
            now = self.clock()
            self.out[name] = self.out.get(name, 0.0) + 1e3 * (now - self.t)
            self.t = now


class SolutionContainer(dict):
    """openpnm/algorithms/_solution.py:10"""
    pass


class SteadyStateSolution(np.ndarray):
    """openpnm/algorithms/_solution.py:15-20"""

    def __new__(cls, x):
        return np.asarray(x).view(cls)


class TransientSolution(np.ndarray):
    """openpnm/algorithms/_solution.py:23-84: the (Np, Nt) array of stored states
    with the stored times in `.t`; calling it interpolates linearly in time."""

    def __new__(cls, t, y):
        obj = np.asarray(y).view(cls)
        obj._x = np.asarray(t)
        return obj

    def __array_finalize__(self, obj):
        if obj is not None:
            self._x = getattr(obj, "_x", None)

    @property
    def t(self):
        return self._x

    def interpolate(self, t):
        from scipy.interpolate import interp1d
        if not hasattr(self, "_interpolant") or self._interpolant is None:
            self._interpolant = interp1d(self._x, np.asarray(self), bounds_error=True)
        return self._interpolant(t)

    __call__ = interpolate


class Settings(dict):
    """attribute + item access, `_update` like openpnm.utils.SettingsAttr"""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def _update(self, d):
        self.update(dict(d))


class Transport(dict):
    """Steady-state linear transport on the GPU (reference: Transport)."""

    _defaults = dict(quantity="", conductance="", cache=True, variable_props=set(),
                     validate_topology=True, cg_rtol=None)

    def __init__(self, network, phase, name="trans_?", **kwargs):
        super().__init__()
        self.network = network
        self._phase = phase
        self.name = name
        self.settings = Settings(self._defaults)
        self.settings["variable_props"] = set()
        self.settings["phase"] = getattr(phase, "name", "phase")
        self.settings._update(kwargs)
        self.Np = int(network.Np)
        self.Nt = int(network.Nt)
        dict.__setitem__(self, "pore.all", np.ones(self.Np, dtype=bool))
        dict.__setitem__(self, "throat.all", np.ones(self.Nt, dtype=bool))
        dict.__setitem__(self, "pore.bc.rate", np.full(self.Np, np.nan))
        dict.__setitem__(self, "pore.bc.value", np.full(self.Np, np.nan))
        self.soln = {}
        self._ctx = None
        self._x_dev = None
        self._cache_key = None
        self.stats = {}

    # ------------------------------------------------------------ dict rules
    def __setitem__(self, key, value):
        n = self.Np if key.startswith("pore.") else self.Nt if key.startswith("throat.") else None
        if n is None:
            raise KeyError("keys must start with 'pore.' or 'throat.'")
        a = np.asarray(value)
        if a.ndim == 0:
            a = np.full(n, a[()])
        elif a.shape[0] != n:
            raise ValueError(f"{key}: expected length {n}")
        dict.__setitem__(self, key, np.array(a))

    def __getitem__(self, key):
        try:
            return dict.__getitem__(self, key)
        except KeyError:
            pre = key + "."
            sub = {k[len(pre):]: v for k, v in self.items() if k.startswith(pre)}
            if sub:
                return sub
            if key == "pore.source":
                return {}          # _transport.py:76-83
            raise KeyError(key)

    _missing = object()

    def pop(self, key, default=_missing):
        """dict.pop that also removes a nested prefix ('pore.source' drops every
        'pore.source.*'), as the reference's Base2 does (_base2.py:166-177)."""
        if dict.__contains__(self, key):
            return dict.pop(self, key)
        pre = key + "."
        hits = {k[len(pre):]: dict.pop(self, k) for k in list(self.keys())
                if k.startswith(pre)}
        if hits:
            return hits
        if default is Transport._missing:
            raise KeyError(key)
        return default

    @property
    def Ps(self):
        return np.arange(self.Np)

    @property
    def Ts(self):
        return np.arange(self.Nt)

    @property
    def x(self):
        """Shortcut to the solution currently stored on the algorithm."""
        return self[self.settings["quantity"]]

    @x.setter
    def x(self, value):
        self[self.settings["quantity"]] = value

    def _parse_indices(self, indices):
        if indices is None:
            return np.array([], dtype=np.int64)
        a = np.asarray(indices)
        if a.dtype == bool:
            a = np.flatnonzero(a)
        return np.atleast_1d(a).astype(np.int64).ravel()

    # --------------------------------------------------------------- BCs
    def set_value_BC(self, pores=None, values=[], mode="add"):
        self.set_BC(pores=pores, bctype="value", bcvalues=values, mode=mode)

    def set_rate_BC(self, pores=None, rates=[], mode="add"):
        self.set_BC(pores=pores, bctype="rate", bcvalues=rates, mode=mode)

    def clear_value_BCs(self):
        self.set_BC(pores=None, bctype="value", mode="remove")

    def clear_rate_BCs(self):
        self.set_BC(pores=None, bctype="rate", mode="remove")

    def clear_BCs(self, bctype=[]):
        self.set_BC(pores=None, bctype=bctype, mode="remove")

    def set_BC(self, pores=None, bctype=[], bcvalues=[], mode="add"):
        """openpnm/algorithms/_algorithm.py:105-225 (same modes and errors)."""
        if not isinstance(mode, str):
            for item in mode:
                self.set_BC(pores=pores, bctype=bctype, bcvalues=bcvalues, mode=item)
            return
        if len(bctype) == 0:
            bctype = list(self["pore.bc"].keys())
        if not isinstance(bctype, str):
            for item in bctype:
                self.set_BC(pores=pores, bctype=item, bcvalues=bcvalues, mode=mode)
            return
        if mode not in ("add", "overwrite", "remove"):
            raise Exception(f"Unsupported mode {mode}")
        bc_types = list(self["pore.bc"].keys())
        other_types = [t for t in bc_types if t != bctype]
        pores = self.Ps if pores is None else self._parse_indices(pores)
        values = np.array(bcvalues, dtype=float)
        if values.size == 1:
            values = np.ones(pores.size) * values.ravel()[0]
        if values.size > 1 and values.size != pores.size:
            raise Exception("The number of values must match the number of locations")
        if mode == "add":
            mask = np.ones(pores.size, dtype=bool)
            for item in bc_types:
                mask[~np.isnan(self[f"pore.bc.{item}"][pores])] = False
            if not np.all(mask):
                raise Exception("Some of the given locations already have BCs, either use "
                                "mode='remove' first or use mode='overwrite' instead")
            self[f"pore.bc.{bctype}"][pores[mask]] = values[mask]
        elif mode == "overwrite":
            self[f"pore.bc.{bctype}"][pores] = values
            for item in other_types:
                self[f"pore.bc.{item}"][pores] = np.nan
        elif mode == "remove":
            self[f"pore.bc.{bctype}"][pores] = np.nan

    # -------------------------------------------------------- validators
    def _validate_settings(self):
        for k in ("quantity", "conductance", "phase"):
            if self.settings.get(k) is None:
                raise Exception(f"'{k}' hasn't been defined on this algorithm")

    def _validate_topology_health(self, ctx=None):
        """Every cluster must touch a BC pore (_transport.py:243-251 ->
        topotools.is_fully_connected, topotools/_topotools.py:993-1024).
        With a device context the clusters are found on the GPU
        (b200_check_topology: lock-free union-find over the throat list);
        without one, scipy's C connected_components does the same on the host
        (a check only -- nothing of the solve runs on the CPU)."""
        msg = "Your network is clustered, making Ax = b ill-conditioned"
        if ctx is not None:
            conns = self._device_conns(ctx)
            bv, br = self._device_bcs(ctx)
            _, nbad = ctx.check_topology(conns, self.Np, bv, br)
            if nbad:
                raise Exception(msg)
            return
        import scipy.sparse as sprs
        from scipy.sparse import csgraph
        conns = np.asarray(self.network["throat.conns"])
        bc = ~np.isnan(self["pore.bc.rate"]) | ~np.isnan(self["pore.bc.value"])
        am = sprs.coo_matrix((np.ones(conns.shape[0], dtype=np.int8),
                              (conns[:, 0], conns[:, 1])), shape=(self.Np, self.Np))
        ncomp, labels = csgraph.connected_components(am, directed=False)
        if ncomp == 1:
            return
        touched = np.zeros(ncomp, dtype=bool)
        touched[labels[bc]] = True
        if not touched.all():
            raise Exception(msg)

    def _device_conns(self, ctx):
        """throat.conns on the device, uploaded once per (context, network)
        unless caching is off."""
        import torch
        conns = np.asarray(self.network["throat.conns"], dtype=np.int64)
        key = (id(ctx), conns.__array_interface__["data"][0], conns.shape[0])
        memo = getattr(self, "_run_memo", None)       # one upload per run() even with cache off
        if memo is not None and memo.get("conns_key") == key:
            return memo["conns"]
        if not self.settings["cache"] or getattr(self, "_conns_key", None) != key:
            self._conns_dev = ctx.to_device(conns, torch.int64)
            self._conns_key = key
        if memo is not None:
            memo["conns_key"], memo["conns"] = key, self._conns_dev
        return self._conns_dev

    def _device_bcs(self, ctx):
        """'pore.bc.value' / 'pore.bc.rate' on the device, uploaded once per run() (the
        topology check and the BC kernel both read them)."""
        import torch
        memo = getattr(self, "_run_memo", None)
        if memo is not None and "bcs" in memo:
            return memo["bcs"]
        out = (ctx.to_device(self["pore.bc.value"], torch.float64),
               ctx.to_device(self["pore.bc.rate"], torch.float64))
        if memo is not None:
            memo["bcs"] = out
        return out

    # ------------------------------------------------------------- solve
    def _sources(self):
        return []

    def _n_sources(self):
        return 0

    def _conductance(self):
        g = np.asarray(self._phase[self.settings["conductance"]], dtype=np.float64)
        if g.shape != (self.Nt,):
            raise Exception("Received weights are of incorrect length")
        return g

    def _assemble(self, ctx):
        """_build_A + _build_b + _apply_BCs on the device."""
        g = self._conductance()
        key = (id(ctx), id(self.network), g.__array_interface__["data"][0], g.size)
        if not (self.settings["cache"] and self._cache_key == key):
            ctx.build_laplacian(self._device_conns(ctx), g, self.Np)
            self._cache_key = key
        bv, br = self._device_bcs(ctx)
        f, nnz = ctx.apply_bcs(bv, br, keep_zero_diag=self._n_sources() > 0)
        return f, nnz

    def run(self, solver=None, x0=None, verbose=False):
        """Builds A and b on the GPU, solves, stores the result in `alg.x` and
        `alg.soln` (reference: Transport.run, _transport.py:171-207)."""
        if solver is None:
            # one default solver per algorithm, so that the Laplacian cached on its
            # device context survives between runs (settings.cache, _transport.py:108-114)
            solver = getattr(self, "_default_solver", None)
            if solver is None:
                solver = self._default_solver = B200PCG()
        if not getattr(solver, "_b200_native", False):
            raise TypeError("openpnm_b200 algorithms need a B200PCG solver "
                            "(there is no CPU path)")
        self._validate_settings()
        x0_given = x0 is not None
        x0 = np.zeros(self.Np) if x0 is None else np.array(x0, dtype=np.float64)
        if x0_given and not np.isfinite(x0).all():
            raise Exception("x0 contains inf/nan values")
        dict.__setitem__(self, "pore.initial_guess", x0)
        self.soln = SolutionContainer()
        self.soln.is_converged = False
        if self._world_size() > 1 and getattr(solver, "distributed", True):
            return self._run_slabs(solver, x0, x0_given)
        ctx = solver.ctx
        self._ctx = ctx
        ctx.set_format(solver._fmt_code())
        self._run_memo = {}
        st = _Stages(ctx)
        if self.settings["validate_topology"]:
            self._validate_topology_health(ctx)
        st.mark("topology")
        try:
            self._assemble(ctx)
            st.mark("upload+assemble")
            ctx.clear_sources()
            for mask, kind, p1, p2, p3 in self._sources():
                ctx.add_source(kind, mask, p1, p2, p3)
            s = self.settings
            cg_rtol = s["cg_rtol"] if s.get("cg_rtol") else solver.tol
            x_dev, num_iter, conv, cg_iters, info = ctx.picard(
                x0=x0 if x0_given else None,
                relaxation=s.get("relaxation_factor", 1.0),
                newton_maxiter=s.get("newton_maxiter", 5000),
                f_rtol=s.get("f_rtol", 1e-6), x_rtol=s.get("x_rtol", 1e-6),
                cg_rtol=cg_rtol, cg_maxiter=0 if solver.maxiter is None else solver.maxiter,
                update=self._update_hook(ctx))
        except _lib.B200Error as e:
            if getattr(self, "_gplan", None) is not None:
                self._gplan = None
                ctx.clear_conductance_model()
            if e.code == -4:
                raise Exception("A or b contains inf/nan values") from e
            raise
        st.mark("picard+solve")
        self._x_dev = x_dev
        x = self._download(ctx, x_dev)
        st.mark("download")
        q = self.settings["quantity"]
        if getattr(self, "_gplan", None) is not None:
            # leave the phase as regenerate_models would at the converged x
            plan, self._gplan = self._gplan, None
            try:
                self._phase[self.settings["conductance"]] = ctx.conductance_values().cpu().numpy()
                self._phase[plan["pore_prop"]] = sum(a * x ** k for k, a in enumerate(plan["poly"]))
            finally:
                ctx.clear_conductance_model()
        dict.__setitem__(self, q, x)
        self.soln[q] = SteadyStateSolution(x)      # a view, as in _solution.py:17-20
        self.soln.is_converged = bool(conv)
        self.soln.num_iter = int(num_iter)
        self.soln.cg_iters = int(cg_iters)
        self.soln.exit_code = int(info)
        self.stats = ctx.stats()
        self.soln.relres = self.stats["relres"]
        if not conv:
            logger.warning(f"{self.name} didn't converge after {num_iter} iterations")
        self._post_run(x)
        st.mark("post")
        self._run_memo = None
        if st.on:
            self.stats["host_stage_ms"] = st.out

    def _post_run(self, x):
        pass

    def _phase_model(self, prop):
        """(name, arguments) of the phase model that produces `prop`, or None"""
        models = getattr(self._phase, "models", None) or {}
        key = prop if prop in models else next((k for k in models if k.split("@")[0] == prop), None)
        if key is None:
            return None
        m = models[key]
        fn = m["model"]
        return (fn if isinstance(fn, str) else getattr(fn, "__name__", str(fn))), m

    def _device_conductance_plan(self, iterative):
        """The chain quantity -> polynomial pore property -> (neighbour-interpolated or
        fixed) throat property -> generic_hydraulic / generic_diffusive / ... conductance
        has device kernels (b200_set_conductance_model); returns its description, or None
        when the conductance follows the quantity through anything else (then the host
        hook regenerates the models every iteration)."""
        gkey, q = self.settings["conductance"], self.settings["quantity"]
        gm = self._phase_model(gkey)
        if gm is None or gm[0] not in _COND_KINDS:
            return None
        kind, names, defaults = _COND_KINDS[gm[0]]
        pkey, tkey, skey = (gm[1].get(n, d) for n, d in zip(names, defaults))
        pm = self._phase_model(pkey)
        if pm is None or pm[0] != "polynomial" or pm[1].get("prop") != q:
            return None
        poly = np.asarray(pm[1].get("a"), dtype=np.float64).ravel()
        if poly.size < 1 or poly.size > 8:
            return None
        tm, throat_prop, interp = self._phase_model(tkey), None, "mean"
        if tm is None:
            if tkey in iterative:
                return None
            throat_prop = np.asarray(self._phase[tkey], dtype=np.float64)
        elif tm[0] == "from_neighbor_pores" and tm[1].get("prop") == pkey and \
                tm[1].get("mode", "min") in ("mean", "min", "max"):
            interp = tm[1].get("mode", "min")
        else:
            return None
        if set(iterative) - {pkey, tkey, gkey} - self._device_source_props(iterative):
            return None          # something else follows the quantity through Python
        try:
            sf = self.network[skey]
        except KeyError:
            return None
        if isinstance(sf, dict):
            return None
        return dict(kind=kind, poly=poly, pore_prop=pkey, throat_prop_key=tkey,
                    throat_prop=throat_prop, interp=interp, size_factors=np.asarray(sf))

    def _device_source_props(self, iterative):
        """iterative properties that are the outputs of device-evaluated source terms"""
        out = set()
        for it in self["pore.source"].keys():
            if not self._source_on_host(it, iterative):
                out |= {f"pore.{it}", f"pore.{it}.S1", f"pore.{it}.S2", f"pore.{it}.rate"}
        return out

    def _update_hook(self, ctx):
        """Callback for b200_picard_cb, or None when everything that depends on
        the quantity has a device kernel."""
        return None

    @staticmethod
    def _world_size():
        try:
            import torch.distributed as dist
            return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        except Exception:
            return 1

    def _run_slabs(self, solver, x0, x0_given):
        """The same `run()` under torchrun (one process per GPU): every rank holds the
        network as OpenPNM does, assembles the rows of its slab of the index-ordered
        pores from the throats touching them, the PCG runs slab-parallel
        (csrc/dist.cu, csrc/p2p.cu) and every rank receives the full solution
        (the algorithm of openpnm/solvers/_petsc.py:117-128,160,225)."""
        import torch
        import torch.distributed as dist
        from . import dist as bdist
        world, rank = dist.get_world_size(), dist.get_rank()
        g = self._conductance()
        rb, re, band, sel, conns_loc = self._slab_plan(world, rank)
        key = (id(self.network), world, rank, solver.device, solver.fmt)
        ss = getattr(solver, "_slab", None)
        if ss is None or getattr(solver, "_slab_key", None) != key:
            ss = bdist.SlabSolver(self.Np, rb, re, device=solver.device, fmt=solver.fmt)
            solver._slab, solver._slab_key, solver._slab_p2p = ss, key, None
        ctx = ss.ctx
        ctx.set_precond(getattr(solver, "precond", "jacobi"))
        self._ctx, self._slab = ctx, ss
        # the Laplacian of an unchanged (network, conductance array) is kept on the device
        # between runs (settings.cache, _transport.py:108-114); the verdict is collective
        ckey = (id(ctx), id(self.network), g.__array_interface__["data"][0], g.size)
        reuse = bool(self.settings["cache"]) and getattr(self, "_slab_cache_key", None) == ckey
        flag = torch.tensor([1 if reuse else 0],
                            device=ctx.tdev if dist.get_backend() == "nccl" else "cpu")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        reuse = bool(int(flag.item()))
        st = _Stages(ctx)
        conns_dev = self._slab_conns_dev if reuse else ctx.to_device(conns_loc, torch.int64)
        self._slab_conns_dev = conns_dev                        # this rank's throats only
        bv_dev = self._slab_bc_dev(ctx, "pore.bc.value", rb, re, band)
        br_dev = self._slab_bc_dev(ctx, "pore.bc.rate", rb, re, band)
        st.mark("upload conns+bcs")
        if self.settings["validate_topology"]:
            # every rank labels the clusters of its own throats on its GPU; the cut
            # regions are joined on the host (dist.slab_topology_check)
            if bdist.slab_topology_check(ctx, conns_dev, self.Np, rb, re, band, bv_dev, br_dev):
                raise Exception("Your network is clustered, making Ax = b ill-conditioned")
        st.mark("topology")
        self._update_hook(None)        # raises if a model needs the host (single-GPU only)
        try:
            srcs = self._sources()
            f, nnz0, nnz = ss.assemble(conns_dev, None if reuse else self._slab_take(ctx, g, sel),
                                       bv_dev, br_dev,
                                       keep_zero_diag=len(srcs) > 0 or self._slab_keep_diag(),
                                       reuse_laplacian=reuse)
            self._slab_after_assemble(ctx)
            self._slab_cache_key = ckey
            st.mark("upload g+assemble")
            ctx.clear_sources()
            if solver._slab_p2p is None:              # maps the peers once (source-free operator)
                solver._slab_p2p = ss.enable_p2p()
            for smask, kind, p1, p2, p3 in srcs:      # per-pore arrays of the owned rows
                loc = [None if p is None else
                       (p[rb:re] if np.ndim(p) else np.full(re - rb, float(p))) for p in (p1, p2, p3)]
                ctx.add_source(kind, smask[rb:re], *loc)
            s = self.settings
            cg_rtol = s["cg_rtol"] if s.get("cg_rtol") else solver.tol
            x_loc, num_iter, conv, cg_iters, info = ctx.picard(
                x0=x0[rb:re] if x0_given else None,
                relaxation=s.get("relaxation_factor", 1.0),
                newton_maxiter=s.get("newton_maxiter", 5000),
                f_rtol=s.get("f_rtol", 1e-6), x_rtol=s.get("x_rtol", 1e-6),
                cg_rtol=cg_rtol, cg_maxiter=0 if solver.maxiter is None else solver.maxiter)
        except _lib.B200Error as e:
            if e.code == -4:
                raise Exception("A or b contains inf/nan values") from e
            raise
        st.mark("picard+solve")
        x_dev = ss.gather(x_loc)
        self._x_dev = x_dev
        st.mark("gather")
        x = self._download(ctx, x_dev)
        st.mark("download")
        q = self.settings["quantity"]
        dict.__setitem__(self, q, x)
        self.soln[q] = SteadyStateSolution(x)
        self.soln.is_converged = bool(conv)
        self.soln.num_iter = int(num_iter)
        self.soln.cg_iters = int(cg_iters)
        self.soln.exit_code = int(info)
        self.stats = ctx.stats()
        self.soln.relres = self.stats["relres"]
        self._post_run(x)
        st.mark("post")
        if st.on:
            self.stats["host_stage_ms"] = st.out

    def _slab_plan(self, world, rank):
        """(row_begin, row_end, index bandwidth, selector of this rank's throats, their
        conns), cached per (network, world, rank).  A lattice made by this package's
        `Cubic` gets its slab's throats from index arithmetic (dist.cubic_slab_throats):
        no rank ever touches the full 191 M-throat list of a 400^3 network; any other
        network is masked once."""
        from . import dist as bdist
        from .network import Cubic
        net = self.network
        key = (id(net), net.Nt, world, rank)
        plan = getattr(self, "_slab_plan_cache", None)
        if plan is not None and plan[0] == key:
            return plan[1]
        shape = getattr(net, "shape", None)
        lattice = shape is not None and len(shape) == 3 and int(np.prod(shape)) == self.Np
        ranges = bdist.cubic_slab_ranges(shape, world) if lattice else \
            bdist.slab_ranges(self.Np, world)
        rb, re = ranges[rank]
        nx, ny, nz = (int(v) for v in shape) if lattice else (0, 0, 0)
        analytic = type(net) is Cubic and lattice and getattr(net, "connectivity", 6) == 6 and \
            net.Nt == (nx - 1) * ny * nz + nx * (ny - 1) * nz + nx * ny * (nz - 1) and \
            not getattr(net, "_conns_touched", False)
        if analytic:
            band = ny * nz if nx > 1 else (nz if ny > 1 else 1)
            _, conns_loc = bdist.cubic_slab_throats(shape, rb // (ny * nz), re // (ny * nz))
            sel = bdist.cubic_slab_throat_ranges(shape, rb // (ny * nz), re // (ny * nz))
        else:
            conns = np.asarray(net["throat.conns"], dtype=np.int64)
            band = bdist.bandwidth(conns)
            sel = np.flatnonzero(bdist.slab_throat_mask(conns, rb, re))
            conns_loc = conns[sel]
        if min(e - b for b, e in ranges) < band:
            raise Exception(f"a slab would be thinner than the matrix half-bandwidth ({band}); "
                            "renumber the pores by coordinate (openpnm_b200.dist.coordinate_order) "
                            "or use fewer ranks")
        out = (rb, re, band, sel, conns_loc)
        self._slab_plan_cache = (key, out)
        return out

    def _slab_keep_diag(self):
        return False

    def _slab_after_assemble(self, ctx):
        pass

    def _slab_bc_dev(self, ctx, key, rb, re, band):
        """Global-length BC array on this rank's device of which only the part the rank
        reads -- its rows and the halo columns, [rb - band, re + band) -- is uploaded;
        the rest of the (persistent) buffer stays NaN = "no BC" and is never looked at."""
        import torch
        cache = self.__dict__.setdefault("_slab_bc_cache", {})
        ck = (id(ctx), key, self.Np)
        buf = cache.get(ck)
        if buf is None:
            with torch.cuda.stream(ctx.stream):
                buf = torch.full((self.Np,), float("nan"), dtype=torch.float64, device=ctx.tdev)
            cache[ck] = buf
        lo, hi = max(rb - band, 0), min(re + band, self.Np)
        src = np.ascontiguousarray(self[key][lo:hi], dtype=np.float64)
        with torch.cuda.stream(ctx.stream):
            buf[lo:hi].copy_(torch.from_numpy(src), non_blocking=True)
        return buf

    def _slab_take(self, ctx, a, sel):
        """Per-throat array of this rank's throats on the device: `sel` is either an
        index array or (lattice) a list of contiguous (begin, end) ranges, which are
        sliced into a page-locked staging buffer and uploaded from there."""
        import torch
        if not isinstance(sel, list):
            return a[sel]
        if np.ndim(a) != 1:
            return np.concatenate([a[b:e] for b, e in sel])
        n = sum(e - b for b, e in sel)
        stage = getattr(self, "_g_stage", None)
        if stage is None or stage.numel() != n:
            stage = self._g_stage = torch.empty(n, dtype=torch.float64, pin_memory=True)
        out, o = stage.numpy(), 0
        for b, e in sel:
            out[o:o + e - b] = a[b:e]
            o += e - b
        return ctx.to_device(stage, torch.float64)

    def _download(self, ctx, x_dev):
        """Device -> host through a page-locked staging buffer (kept per size)."""
        import torch
        n = int(x_dev.numel())
        stage = getattr(self, "_x_stage", None)
        if stage is None or stage.numel() != n:
            stage = torch.empty(n, dtype=torch.float64, pin_memory=True)
            self._x_stage = stage
        # x_dev may have been produced on another stream (a slab run gathers it with
        # torch.distributed on torch's current stream): order this context's stream after
        # it, or the copy below could read the gathered tensor before it is complete --
        # recycled allocator memory, e.g. an old NaN-filled BC buffer, instead of x
        cur = torch.cuda.current_stream(x_dev.device)
        if cur != ctx.stream:
            ctx.stream.wait_stream(cur)
            x_dev.record_stream(ctx.stream)
        with torch.cuda.stream(ctx.stream):
            stage.copy_(x_dev, non_blocking=True)
        ctx.synchronize()
        return _host_copy(stage.numpy())

    @property
    def A(self):
        """The coefficient matrix as assembled on the device (scipy CSR)."""
        if self._ctx is None:
            raise Exception("run() first")
        return self._ctx.export_csr_scipy(_lib.MAT_SYSTEM)

    @property
    def b(self):
        if self._ctx is None:
            raise Exception("run() first")
        return self._ctx.get_b().cpu().numpy()

    def rate(self, pores=[], throats=[], mode="group"):
        """Net rate leaving `pores` / magnitude through `throats`
        (reference: Transport.rate, _transport.py:259-330)."""
        pores = self._parse_indices(pores)
        throats = self._parse_indices(throats)
        if throats.size > 0 and pores.size > 0:
            raise Exception("Must specify either pores or throats, not both")
        if throats.size == 0 and pores.size == 0:
            raise Exception("Must specify either pores or throats")
        if self._ctx is None:
            raise Exception("run() first")
        x = self._x_dev if self._x_dev is not None else self.x
        if getattr(self, "_slab", None) is not None and self._world_size() > 1:
            return self._rate_slabs(x, pores, throats, mode)
        if throats.size:
            R = self._ctx.rate_throats(x, throats).cpu().numpy()
        else:
            R = self._ctx.rate_pores(x, pores).cpu().numpy()
        if mode == "group":
            R = np.sum(R)
        return np.array(R, ndmin=1)


def _rate_slabs(self, x, pores, throats, mode):
    """rate() in a slab-parallel run: every rank evaluates the pores it owns, the
    per-pore values are summed over ranks (each pore is owned exactly once)."""
    import torch
    import torch.distributed as dist
    ss = self._slab
    if throats.size:
        # every rank holds the full x on its device; the requested throats are named by
        # their own conns / conductances (the context only knows the slab's throats)
        conns = np.asarray(self.network["throat.conns"], dtype=np.int64)[throats]
        R = ss.ctx.rate_throats_explicit(x, conns, self._conductance()[throats]).cpu().numpy()
        return np.array(np.sum(R) if mode == "group" else R, ndmin=1)
    out = torch.zeros(pores.size, dtype=torch.float64, device=ss.ctx.tdev)
    mine = np.flatnonzero((pores >= ss.row_begin) & (pores < ss.row_end))
    if mine.size:
        out[torch.from_numpy(mine).to(ss.ctx.tdev)] = ss.ctx.rate_pores(x, pores[mine])
    ss.ctx.synchronize()
    dist.all_reduce(out)
    R = out.cpu().numpy()
    return np.array(np.sum(R) if mode == "group" else R, ndmin=1)


Transport._rate_slabs = _rate_slabs


class ReactiveTransport(Transport):
    """Steady transport with (optional) nonlinear source terms; the Picard /
    Newton-linearised outer loop runs on the device (reference:
    ReactiveTransport, _reactive_transport.py)."""

    def __init__(self, network, phase, name="react_trans_?", **kwargs):
        super().__init__(network=network, phase=phase, name=name, **kwargs)
        for k, v in dict(relaxation_factor=1.0, newton_maxiter=5000, f_rtol=1e-6,
                         x_rtol=1e-6).items():
            self.settings.setdefault(k, v)

    def _dependency_edges(self):
        """(upstream, downstream) property pairs of the phase's models: OpenPNM's
        own graph when the phase is a real one (core/_models.py dependency_graph),
        else the string arguments of the stand-in Phase's models."""
        models = getattr(self._phase, "models", None)
        if models is None:
            return []
        if hasattr(models, "dependency_graph"):
            return [(str(a), str(b)) for a, b in models.dependency_graph(deep=True).edges]
        edges = []
        for prop, m in models.items():
            for k, v in m.items():
                if k != "model" and isinstance(v, str) and \
                        v.split(".", 1)[0] in ("pore", "throat"):
                    edges.append((v, prop.split("@")[0]))
        return edges

    @property
    def iterative_props(self):
        """Properties downstream of the quantity (and of settings.variable_props)
        in the model dependency graph, in lexicographic topological order; they
        are regenerated every outer iteration (reference: _algorithm.py:38-67)."""
        import heapq
        q = self.settings["quantity"]
        base = set(self.settings.get("variable_props", ())) | {q}
        down = {}
        for a, b in self._dependency_edges():
            down.setdefault(a, set()).add(b)
        nodes, stack = set(), [n for n in base if n in down]
        edges = set()
        while stack:
            a = stack.pop()
            if a in nodes:
                continue
            nodes.add(a)
            for b in down.get(a, ()):
                edges.add((a, b))
                stack.append(b)
        indeg = {n: 0 for n in nodes}
        for a, b in edges:
            indeg[b] += 1
        heap = [n for n, d in indeg.items() if d == 0]
        heapq.heapify(heap)
        order = []
        while heap:
            a = heapq.heappop(heap)
            order.append(a)
            for b in sorted(down.get(a, ())):
                if (a, b) in edges:
                    indeg[b] -= 1
                    if indeg[b] == 0:
                        heapq.heappush(heap, b)
        return [n for n in order if n != q]

    def _source_on_host(self, item, iterative):
        """True when source `item` cannot be evaluated by a device kernel: its
        model is not one of linear / power_law / standard_kinetics, or one of its
        parameter arrays itself depends on the quantity."""
        try:
            _, m = self._model_of(item)
        except NotImplementedError:
            return True
        return any(isinstance(v, str) and v in iterative
                   for k, v in m.items() if k not in ("model", "X", "regen_mode"))

    def _update_hook(self, ctx):
        """_update_iterative_props + _update_A_and_b + _apply_sources for models
        that live in Python (_algorithm.py:69-91, _reactive_transport.py:125-148,
        226-232): the device loop hands x back, the phase regenerates the
        iterative properties, and the new conductance / S1 / S2 go to the device.
        Returns None when nothing needs the host (the usual case)."""
        iterative = self.iterative_props
        items = list(self["pore.source"].keys())
        host_items = [it for it in items if self._source_on_host(it, iterative)]
        g_variable = self.settings["conductance"] in iterative
        self._gplan = None
        if not host_items and not g_variable:
            return None
        if self._world_size() > 1:
            raise NotImplementedError("host-evaluated models run on one GPU")
        if g_variable and not host_items and self.settings.get("device_conductance", True):
            plan = self._device_conductance_plan(iterative)
            if plan is not None and ctx is not None:
                # the whole chain has device kernels: the Picard loop re-evaluates g,
                # re-assembles and re-applies the BCs itself, nothing comes back to the host
                ctx.set_conductance_model(
                    plan["kind"], self._device_conns(ctx), self.Np, plan["poly"],
                    plan["size_factors"], throat_prop=plan["throat_prop"], interp=plan["interp"],
                    bc_value=self["pore.bc.value"], bc_rate=self["pore.bc.rate"],
                    keep_zero_diag=self._n_sources() > 0)
                self._gplan = plan
                self._cache_key = None
                return None
        phase, q = self._phase, self.settings["quantity"]
        dev_items = [it for it in items if it not in host_items]
        device_terms = dict(zip(dev_items, self._sources_for(dev_items)))

        def update(x_dev):
            x = self._download(ctx, x_dev)
            dict.__setitem__(self, q, x)
            phase[q] = x
            phase.regenerate_models(propnames=iterative)
            if g_variable:
                ctx.build_laplacian(self._device_conns(ctx), self._conductance(), self.Np)
                self._cache_key = None
                ctx.apply_bcs(self["pore.bc.value"], self["pore.bc.rate"],
                              keep_zero_diag=self._n_sources() > 0)
            ctx.clear_sources()
            for it in items:
                mask = np.asarray(self["pore.source." + it], dtype=bool)
                if it in host_items:
                    try:
                        S1 = np.asarray(phase[f"pore.{it}.S1"], dtype=np.float64)
                        S2 = np.asarray(phase[f"pore.{it}.S2"], dtype=np.float64)
                    except KeyError:
                        # the reference stops applying sources at the first model
                        # that left no S1/S2 (_reactive_transport.py:136-148)
                        break
                    ctx.add_source(_lib.SRC_LINEAR, mask, S1, S2, None)
                else:
                    _, kind, p1, p2, p3 = device_terms[it]
                    ctx.add_source(kind, mask, p1, p2, p3)
        return update

    def set_source(self, pores, propname, mode="add"):
        """_reactive_transport.py:70-123"""
        if isinstance(mode, list):
            for item in mode:
                self.set_source(pores=pores, propname=propname, mode=item)
            return
        if not propname.startswith("pore."):
            propname = "pore." + propname
        pores = self._parse_indices(pores)
        locs_BC = np.zeros(self.Np, dtype=bool)
        for item in self["pore.bc"].keys():
            locs_BC |= np.isfinite(self[f"pore.bc.{item}"])
        if np.any(locs_BC[pores]):
            raise Exception("BCs present in given pores, can't assign source term")
        prop = "pore.source." + propname.split(".", 1)[1]
        if mode == "add":
            if prop not in self.keys():
                self[prop] = False
            dict.__getitem__(self, prop)[pores] = True
        elif mode == "remove":
            dict.__getitem__(self, prop)[pores] = False
        elif mode == "clear":
            self[prop] = False
        else:
            raise Exception(f"Unsupported mode {mode}")

    def set_BC(self, pores=None, bctype=[], bcvalues=[], mode="add"):
        """_reactive_transport.py:246-256: BCs may not land on source pores."""
        if isinstance(mode, str) and mode != "remove" and pores is not None:
            ps = self._parse_indices(pores)
            for item in self["pore.source"].keys():
                if np.any(self["pore.source." + item][ps]):
                    raise Exception("Source term already present in given pores, "
                                    "can't assign BCs")
        super().set_BC(pores=pores, bctype=bctype, bcvalues=bcvalues, mode=mode)

    def _model_of(self, item):
        models = getattr(self._phase, "models", None)
        key = "pore." + item
        if models is None or key not in models:
            # OpenPNM stores 'propname@domain' keys too
            cands = [k for k in (models or {}) if k.split("@")[0] == key]
            if not cands:
                raise KeyError(f"no source model '{key}' on the phase")
            key = cands[0]
        m = models[key]
        fn = m["model"]
        name = fn if isinstance(fn, str) else getattr(fn, "__name__", str(fn))
        if name not in _SRC_KINDS:
            raise NotImplementedError(
                f"source model '{name}' has no device kernel (linear, power_law, "
                "standard_kinetics are implemented)")
        return name, m

    def _n_sources(self):
        return len(self["pore.source"])

    def _sources(self):
        """Source terms with a device kernel; the others are handed over by the
        update hook (`_update_hook`) every outer iteration."""
        iterative = self.iterative_props
        return self._sources_for([it for it in self["pore.source"].keys()
                                  if not self._source_on_host(it, iterative)])

    def _sources_for(self, items):
        out = []
        for item in items:
            mask = np.asarray(self["pore.source." + item], dtype=bool)
            name, m = self._model_of(item)
            if m.get("X", self.settings["quantity"]) != self.settings["quantity"]:
                raise Exception("source term X must be the algorithm's quantity")
            kind, keys, defaults = _SRC_KINDS[name]
            ps = []
            for k, d in zip(keys, defaults):
                if k is None:
                    ps.append(None)
                    continue
                v = m.get(k, d)
                if isinstance(v, str):
                    v = np.asarray(self._phase[v], dtype=np.float64)
                ps.append(v)
            out.append((mask, kind, ps[0], ps[1], ps[2]))
        return out

    def _post_run(self, x):
        """Leave S1 / S2 / rate of each source on the phase at the converged x,
        as `regenerate_models` does in the reference (_algorithm.py:89-91); like
        there, nothing is written when no property depends on the quantity."""
        if not self["pore.source"]:
            return
        try:
            self._phase[self.settings["quantity"]] = x
        except Exception:
            return
        iterative = self.iterative_props
        ctx, slab = self._ctx, getattr(self, "_slab", None)
        slab = slab if (slab is not None and self._world_size() > 1) else None
        # index = registration order on the context: run() registers the device-evaluated
        # items in this order, the host hook re-registers ALL items in this order
        for k, item in enumerate(self["pore.source"].keys()):
            if self._source_on_host(item, iterative):
                continue            # regenerated on the phase by the update hook
            # S1 / S2 / rate at the converged x from the device kernel that linearised
            # them during the run (b200_source_values), not re-derived in numpy
            if slab is not None:
                xl = self._x_dev[slab.row_begin:slab.row_end]
                vals = [slab.gather(v) for v in ctx.source_values(k, xl)]
            else:
                vals = ctx.source_values(k, self._x_dev)
            S1, S2, r = (v.cpu().numpy() for v in vals)
            try:
                self._phase[f"pore.{item}.S1"] = S1
                self._phase[f"pore.{item}.S2"] = S2
                self._phase[f"pore.{item}.rate"] = r
            except Exception:
                pass


class AdvectionDiffusion(ReactiveTransport):
    """Advection-diffusion with an (Nt,2) conductance: non-symmetric A, solved on
    the device by Jacobi-scaled BiCGStab (reference: AdvectionDiffusion,
    openpnm/algorithms/_advection_diffusion.py:43-105)."""

    def __init__(self, network, phase, name="ad_?", **kwargs):
        super().__init__(network=network, phase=phase, name=name, **kwargs)
        self.settings._update(dict(quantity="pore.concentration",
                                   conductance="throat.ad_dif_conductance",
                                   diffusive_conductance="throat.diffusive_conductance",
                                   hydraulic_conductance="throat.hydraulic_conductance",
                                   pressure="pore.pressure", cache=False))
        self.settings._update(kwargs)
        dict.__setitem__(self, "pore.bc.outflow", np.full(self.Np, np.nan))

    def set_outflow_BC(self, pores, mode="add"):
        """Zero-gradient outflow: stores the net hydraulic inflow of each pore as its
        'outflow' BC value (_advection_diffusion.py:56-84)."""
        pores = self._parse_indices(pores)
        conns = np.asarray(self.network["throat.conns"])
        sel = np.isin(conns[:, 0], pores) | np.isin(conns[:, 1], pores)   # find_neighbor_throats
        C12 = conns[sel]
        P12 = np.asarray(self._phase[self.settings["pressure"]])[C12]
        gh = np.asarray(self._phase[self.settings["hydraulic_conductance"]])[sel]
        Q12 = -gh * np.diff(P12, axis=1).squeeze(axis=1)
        Qp = np.zeros(self.Np)
        np.add.at(Qp, C12[:, 0], -Q12)
        np.add.at(Qp, C12[:, 1], Q12)
        self.set_BC(pores=pores, bcvalues=Qp[pores], bctype="outflow", mode=mode)

    def _conductance(self):
        g = np.asarray(self._phase[self.settings["conductance"]], dtype=np.float64)
        if g.shape not in ((self.Nt,), (self.Nt, 2)):
            raise Exception("Received weights are of incorrect length")
        return g

    def _assemble(self, ctx):
        g = self._conductance()
        if g.ndim == 1:                       # a symmetric conductance: the parent's path
            f, nnz = super()._assemble(ctx)
        else:
            ctx.build_laplacian2(self._device_conns(ctx), g, self.Np)
            self._cache_key = None
            # the outflow step re-writes the whole diagonal (COO setdiag), so every
            # diagonal entry is explicit in the reference's matrix
            f, nnz = ctx.apply_bcs(self["pore.bc.value"], self["pore.bc.rate"],
                                   keep_zero_diag=True)
        if g.ndim == 1 and np.isfinite(self["pore.bc.outflow"]).any():
            f, nnz = ctx.apply_bcs(self["pore.bc.value"], self["pore.bc.rate"],
                                   keep_zero_diag=True)
        ctx.add_to_diagonal(self["pore.bc.outflow"])
        return f, nnz

    # slab runs (torchrun): the rows of a slab are assembled from its throats with the
    # two-way weights (b200_build_laplacian2_rows), every diagonal entry is kept explicit
    # and the outflow values are added to the owned rows, as `_assemble` does on one GPU;
    # the solve is the slab-parallel BiCGStab
    def _slab_keep_diag(self):
        return True

    def _slab_after_assemble(self, ctx):
        ctx.add_to_diagonal(self["pore.bc.outflow"])

    def rate(self, pores=[], throats=[], mode="group"):
        throats_i = self._parse_indices(throats)
        g = self._conductance()
        if throats_i.size and g.ndim == 2:
            # |g[:,0] x[c1] - g[:,1] x[c0]| (_transport.py:306-319 with the flipped (Nt,2) g)
            if self._parse_indices(pores).size:
                raise Exception("Must specify either pores or throats, not both")
            if self._ctx is None:
                raise Exception("run() first")
            c = np.asarray(self.network["throat.conns"])[throats_i]
            x = self._x_dev if self._x_dev is not None else self.x
            R = self._ctx.rate_throats_explicit(x, c, g[throats_i]).cpu().numpy()
            return np.array(np.sum(R) if mode == "group" else R, ndmin=1)
        return super().rate(pores=pores, throats=throats, mode=mode)


class TransientReactiveTransport(ReactiveTransport):
    """dy/dt = (-A y + b) / V integrated on the device (reference:
    TransientReactiveTransport, _transient_reactive_transport.py:30-123)."""

    def __init__(self, network, phase, name="trans_react_?", **kwargs):
        super().__init__(network=network, phase=phase, name=name, **kwargs)
        self.settings.setdefault("pore_volume", "pore.volume")
        dict.__setitem__(self, "pore.ic", np.full(self.Np, np.nan))

    def run(self, x0, tspan, saveat=None, integrator=None):
        """Same signature and `saveat` handling as the reference (:51-99)."""
        from .integrators import B200RK45
        if np.isscalar(saveat):
            saveat = np.arange(*tspan, saveat)
        if (saveat is not None) and (tspan[1] not in saveat):
            saveat = np.hstack((saveat, [tspan[1]]))
        integrator = B200RK45() if integrator is None else integrator
        if not getattr(integrator, "_b200_native", False):
            raise TypeError("openpnm_b200 transient algorithms need a B200RK45 integrator")
        self._validate_settings()
        ctx = integrator.ctx
        self._ctx = ctx
        self._run_memo = {}
        ctx.set_format(integrator._fmt_code())
        if self.settings["validate_topology"]:
            self._validate_topology_health(ctx)
        x0 = np.ones(self.Np, dtype=float) * x0
        dict.__setitem__(self, "pore.ic", x0)
        bc = ~np.isnan(self["pore.bc.value"])            # _merge_inital_and_boundary_values
        x0[bc] = self["pore.bc.value"][bc]
        q = self.settings["quantity"]
        dict.__setitem__(self, q, x0)
        V = np.asarray(self.network[self.settings["pore_volume"]], dtype=np.float64)
        V = V * np.ones(self.Np)
        if self._update_hook(None) is not None:
            raise NotImplementedError("transient runs need device-evaluated models "
                                      "(linear, power_law, standard_kinetics sources; "
                                      "constant conductance)")
        if self._world_size() > 1 and getattr(integrator, "distributed", True):
            return self._run_transient_slabs(integrator, x0, V, tspan, saveat)
        try:
            self._assemble(ctx)
            ctx.clear_sources()
            for mask, kind, p1, p2, p3 in self._sources():
                ctx.add_source(kind, mask, p1, p2, p3)
            cap = None
            if saveat is None:
                cap = integrator.max_saved or max(16, min(20000, int(2e9 // (8 * self.Np))))
            t, Y, nsteps, nfev, status = ctx.transient_rk45(
                x0, V, float(tspan[0]), float(tspan[1]), rtol=integrator.rtol,
                atol=integrator.atol, saveat=saveat, capacity=cap)
        except _lib.B200Error as e:
            if e.code == -4:
                raise Exception("A or b contains inf/nan values") from e
            raise
        if status == -1:
            raise Exception("Required step size is less than spacing between numbers.")
        if status == -2:
            raise Exception("too many accepted steps to store them all: pass `saveat`")
        Yh = np.ascontiguousarray(Y.T.cpu().numpy())
        self.soln = SolutionContainer()
        self.soln[q] = TransientSolution(t, Yh)
        self.soln.nsteps, self.soln.nfev = int(nsteps), int(nfev)
        if Yh.shape[1]:
            dict.__setitem__(self, q, Yh[:, -1].copy())
            self._x_dev = Y[-1]
        self.stats = ctx.stats()
        self._post_run(self[q])


def _run_transient_slabs(self, integrator, x0, V, tspan, saveat):
    """The transient run under torchrun: slabs of the index-ordered pores as in
    `Transport._run_slabs`; `b200_transient_rk45` exchanges the halo of every right-hand
    side's operand and all-reduces the step controller's norms, so every rank takes the same
    steps; the saved states are gathered and every rank ends with the full solution."""
    import torch
    import torch.distributed as dist
    from . import dist as bdist
    world, rank = dist.get_world_size(), dist.get_rank()
    g = self._conductance()
    if g.ndim != 1:
        raise NotImplementedError("TransientAdvectionDiffusion is not slab-parallel: pass "
                                  "B200RK45(distributed=False)")
    rb, re, band, sel, conns_loc = self._slab_plan(world, rank)
    key = (id(self.network), world, rank, integrator.device, integrator.fmt)
    ss = getattr(integrator, "_slab", None)
    if ss is None or getattr(integrator, "_slab_key", None) != key:
        ss = bdist.SlabSolver(self.Np, rb, re, device=integrator.device, fmt=integrator.fmt)
        integrator._slab, integrator._slab_key = ss, key
    ctx = ss.ctx
    self._ctx, self._slab = ctx, ss
    conns_dev = ctx.to_device(conns_loc, torch.int64)
    bv_dev = self._slab_bc_dev(ctx, "pore.bc.value", rb, re, band)
    br_dev = self._slab_bc_dev(ctx, "pore.bc.rate", rb, re, band)
    if self.settings["validate_topology"]:
        if bdist.slab_topology_check(ctx, conns_dev, self.Np, rb, re, band, bv_dev, br_dev):
            raise Exception("Your network is clustered, making Ax = b ill-conditioned")
    q = self.settings["quantity"]
    try:
        srcs = self._sources()
        ss.assemble(conns_dev, self._slab_take(ctx, g, sel), bv_dev, br_dev,
                    keep_zero_diag=len(srcs) > 0)
        self._slab_cache_key = None
        ctx.clear_sources()
        for smask, kind, p1, p2, p3 in srcs:
            loc = [None if p is None else
                   (p[rb:re] if np.ndim(p) else np.full(re - rb, float(p))) for p in (p1, p2, p3)]
            ctx.add_source(kind, smask[rb:re], *loc)
        cap = None
        if saveat is None:
            cap = integrator.max_saved or max(16, min(20000, int(2e9 // (8 * (re - rb)))))
        t, Y, nsteps, nfev, status = ctx.transient_rk45(
            x0[rb:re], V[rb:re], float(tspan[0]), float(tspan[1]), rtol=integrator.rtol,
            atol=integrator.atol, saveat=saveat, capacity=cap)
    except _lib.B200Error as e:
        if e.code == -4:
            raise Exception("A or b contains inf/nan values") from e
        raise
    if status == -1:
        raise Exception("Required step size is less than spacing between numbers.")
    if status == -2:
        raise Exception("too many accepted steps to store them all: pass `saveat`")
    nout = int(Y.shape[0])
    with torch.cuda.stream(ctx.stream):
        mine = Y.T.contiguous().reshape(-1)                  # (n_local, nout), row-major
    Yfull = ss.gather(mine).reshape(self.Np, nout)           # rank blocks stack along the pores
    Yh = Yfull.cpu().numpy()
    self.soln = SolutionContainer()
    self.soln[q] = TransientSolution(t, Yh)
    self.soln.nsteps, self.soln.nfev = int(nsteps), int(nfev)
    if nout:
        dict.__setitem__(self, q, Yh[:, -1].copy())
        self._x_dev = Yfull[:, -1].contiguous()     # on the stream that gathered Yfull
    self.stats = ctx.stats()
    self._post_run(self[q])


TransientReactiveTransport._run_transient_slabs = _run_transient_slabs


class TransientFickianDiffusion(TransientReactiveTransport):
    """_transient_fickian_diffusion.py"""

    def __init__(self, network, phase, name="trans_fick_?", **kwargs):
        super().__init__(network=network, phase=phase, name=name, **kwargs)
        self.settings._update(dict(quantity="pore.concentration",
                                   conductance="throat.diffusive_conductance"))
        self.settings._update(kwargs)


class TransientOhmicConduction(TransientReactiveTransport):
    """_transient_ohmic_conduction.py"""

    def __init__(self, network, phase, name="trans_ohmic_?", **kwargs):
        super().__init__(network=network, phase=phase, name=name, **kwargs)
        self.settings._update(dict(quantity="pore.voltage",
                                   conductance="throat.electrical_conductance"))
        self.settings._update(kwargs)


class TransientFourierConduction(TransientReactiveTransport):
    """_transient_fourier_conduction.py"""

    def __init__(self, network, phase, name="trans_fourier_?", **kwargs):
        super().__init__(network=network, phase=phase, name=name, **kwargs)
        self.settings._update(dict(quantity="pore.temperature",
                                   conductance="throat.thermal_conductance"))
        self.settings._update(kwargs)


class TransientAdvectionDiffusion(TransientReactiveTransport, AdvectionDiffusion):
    """_transient_advection_diffusion.py: the transient driver on the
    non-symmetric AdvectionDiffusion assembly."""

    def __init__(self, network, phase, name="trans_ad_?", **kwargs):
        super().__init__(network=network, phase=phase, name=name, **kwargs)


class StokesFlow(ReactiveTransport):
    """_stokes_flow.py:21-32"""

    def __init__(self, network, phase, name="stokes_?", **kwargs):
        super().__init__(network=network, phase=phase, name=name, **kwargs)
        self.settings._update(dict(quantity="pore.pressure",
                                   conductance="throat.hydraulic_conductance"))
        self.settings._update(kwargs)


class FickianDiffusion(ReactiveTransport):
    """_fickian_diffusion.py:22-55"""

    def __init__(self, network, phase, name="fick_?", **kwargs):
        super().__init__(network=network, phase=phase, name=name, **kwargs)
        self.settings._update(dict(quantity="pore.concentration",
                                   conductance="throat.diffusive_conductance"))
        self.settings._update(kwargs)


class OhmicConduction(ReactiveTransport):
    """_ohmic_conduction.py:22-23"""

    def __init__(self, network, phase, name="ohmic_?", **kwargs):
        super().__init__(network=network, phase=phase, name=name, **kwargs)
        self.settings._update(dict(quantity="pore.voltage",
                                   conductance="throat.electrical_conductance"))
        self.settings._update(kwargs)


class FourierConduction(ReactiveTransport):
    """_fourier_conduction.py:22-23"""

    def __init__(self, network, phase, name="fourier_?", **kwargs):
        super().__init__(network=network, phase=phase, name=name, **kwargs)
        self.settings._update(dict(quantity="pore.temperature",
                                   conductance="throat.thermal_conductance"))
        self.settings._update(kwargs)
