#!/usr/bin/env python
"""bench.py -- StokesFlow time-to-solve and DOF*iter/s on a synthetic Cubic
network (BASELINE.json: Cubic 400^3, 64M pores, Jacobi-PCG rtol 1e-10), on
N B200s of one box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--size 400]
    python bench.py --impl reference ...      # CPU arm: unmodified OpenPNM + SciPy

One "step" = one full pass of the hot path on the resident network:
assemble the Laplacian from conns + g, apply the inlet/outlet Dirichlet BCs,
Jacobi-PCG to rtol 1e-10.  `value` has the inputs already in HBM; `e2e` goes
through the public API (`StokesFlow.run` + `rate`, under torchrun for N > 1) from
pinned HOST arrays, topology check included, and reads x back every step.
Next to it, in the same line:
  amg     the same step with B200PCG(precond='amg') at every N (time-to-solve),
  parity  the 1e-8 claim at this size: Jacobi-PCG and the multigrid solve at rtol
          1e-12 and 1e-13 compared with each other (x, inlet/outlet rates, residual on
          the canonical CSR) and, for N > 1, with a single-GPU solve made by rank 0,
  matched the GPU arm on the reference arm's own (small) problem, for a same-problem
          time ratio next to the rate ratio.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RTOL = 1e-10
P_IN, P_OUT = 2e5, 1e5


def conductance(Nt, ids=None, seed=0):
    """g = 10^U(-15,-12), default_rng(seed)."""
    g = 10.0 ** np.random.default_rng(seed).uniform(-15, -12, Nt)
    return g if ids is None else g[ids]


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200", "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            busy = sorted(sm)[len(sm) // 2:]          # upper half = samples under load
            out = {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)),
                   "reasons": sorted(reasons), "samples": len(sm)}
        return out


def pardiso_note():
    try:
        import pypardiso  # noqa: F401
        return "pypardiso present"
    except Exception:
        return "Pardiso unavailable (pypardiso is not installed); SuperLU/ScipyCG used"


# ------------------------------------------------------------------ CPU arm
def cpu_solve_sample(n, rtol=RTOL):
    """Fallback when the reference cannot be imported: the oracle port of its path
    (numpy assembly as `_build_A`/`_apply_BCs` + SciPy CSR SpMV Jacobi-PCG)."""
    from oracle import oracle
    net = oracle.cubic_network([n, n, n], spacing=1e-4)
    g = conductance(net["Nt"])
    bcv = np.full(net["Np"], np.nan)
    bcv[net["labels"]["left"]] = P_IN
    bcv[net["labels"]["right"]] = P_OUT
    t1 = time.perf_counter()
    A, b, f = oracle.apply_bcs(oracle.build_A(net["conns"], g, net["Np"]), net["Np"], bcv, None)
    M = oracle.to_csr(A, net["Np"])
    x, info, its = oracle.jacobi_pcg(M, b, rtol=rtol)
    q = oracle.rate(net["conns"], g, x, pores=net["labels"]["left"])[0]
    t4 = time.perf_counter()
    return dict(Np=net["Np"], iters=its, info=info, t_step=t4 - t1, rate=q)


def reference_openpnm_sample(n, rtol=RTOL):
    """The REAL reference (unmodified OpenPNM from baseline/_ref or /root/reference,
    imported through tests/ref_shim.py): `StokesFlow.run(solver=ScipyCG(tol))` + `rate()`
    on a Cubic n^3 sample -- its own assembly (x3), topology check and SciPy CG.
    Returns None when the reference cannot be imported."""
    import contextlib
    import logging
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    logging.disable(logging.CRITICAL)      # the reference logs through rich, to stdout
    try:
        with contextlib.redirect_stdout(sys.stderr):   # stdout carries the one JSON line only
            import ref_shim
            op = ref_shim.import_reference()
    except Exception:
        op = None
    if op is None:
        return None
    pn = op.network.Cubic(shape=[n, n, n], spacing=1e-4)
    ph = op.phase.Phase(network=pn)
    ph["throat.hydraulic_conductance"] = conductance(pn.Nt)
    sf = op.algorithms.StokesFlow(network=pn, phase=ph)
    sf.set_value_BC(pores=pn.pores("left"), values=P_IN)
    sf.set_value_BC(pores=pn.pores("right"), values=P_OUT)
    its = [0]

    class CountingCG(op.solvers.ScipyCG):
        def solve(self, A, b, **kw):
            def cb(xk):
                its[0] += 1
            return super().solve(A, b, callback=cb, **kw)

    t0 = time.perf_counter()
    sf.run(solver=CountingCG(tol=rtol))
    q = sf.rate(pores=pn.pores("left"))[0]
    dt = time.perf_counter() - t0
    out = dict(Np=pn.Np, iters=its[0], t_step=dt, rate=float(q),
               converged=bool(sf.soln.is_converged))
    op.Workspace().clear()
    return out


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    try:
        import threadpoolctl
        threads = max(d.get("num_threads", 1) for d in threadpoolctl.threadpool_info()) or 1
    except Exception:
        threads = 1
    n = args.ref_size
    real = reference_openpnm_sample(min(n, 24))        # also the import probe / warm-up
    rates = []
    if real is not None:
        for _ in range(max(args.warmup - 1, 0)):
            reference_openpnm_sample(min(n, 24))
        t, its, Np = [], 0, 0
        for _ in range(args.steps):
            r = reference_openpnm_sample(n)
            t.append(r["t_step"])
            rates.append(r["rate"])
            its, Np = r["iters"], r["Np"]
        kind, solver = "reference", f"ScipyCG(tol={RTOL:g}) (SciPy's unpreconditioned CG)"
        sample = (f"unmodified OpenPNM StokesFlow.run(solver=ScipyCG(tol={RTOL:g})) + rate() on "
                  f"Cubic {n}^3 -- a bounded sample of the GPU arm's Cubic {args.size}^3 workload "
                  f"(same g distribution, BCs and rtol; the full size would take the CPU hours): "
                  f"its own topology check, 3 assemblies and SciPy's serial CG, {its} iterations. "
                  + pardiso_note())
    else:
        n = args.cpu_size
        for _ in range(args.warmup):
            cpu_solve_sample(min(n, 40))
        t, its, Np = [], 0, 0
        for _ in range(args.steps):
            r = cpu_solve_sample(n)
            t.append(r["t_step"])
            rates.append(r["rate"])
            its, Np = r["iters"], r["Np"]
        kind, solver = "port", "oracle port: SciPy CSR SpMV Jacobi-PCG"
        sample = (f"oracle port on Cubic {n}^3 -- a bounded sample of the GPU arm's Cubic "
                  f"{args.size}^3 workload (same g distribution/BCs/rtol): numpy assembly as the "
                  f"reference + SciPy CSR SpMV Jacobi-PCG, {its} iterations; serial. "
                  + pardiso_note())
    ms = 1e3 * float(np.mean(t))
    val = Np * its / (ms * 1e-3)
    cfg = workload_config(n, Np, 3 * n * n * (n - 1), {"solver": solver})
    cfg["sample_of"] = f"Cubic {args.size}x{args.size}x{args.size} (the GPU arm's workload)"
    cfg["size_solved"] = n
    line = {
        "impl": "reference", "metric": "stokesflow_dof_iter_per_s", "value": val,
        "unit": "DOF*iter/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "same_config": False,
        "time_to_solve_s": ms * 1e-3, "cg_iters": its, "rate_inlet": float(np.mean(rates)),
        "cpu_baseline": {"value": val, "unit": "DOF*iter/s", "cores": 1, "kind": kind,
                         "blas_threads": threads, "sample": sample},
        "e2e": {"value": val, "unit": "DOF*iter/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def workload_config(size, Np, Nt, extra, solver="Jacobi-PCG"):
    cfg = {"workload": f"Cubic {size}x{size}x{size} StokesFlow (g=10^U(-15,-12) seed 0, "
                       f"left={P_IN:g}/right={P_OUT:g} Dirichlet), rtol {RTOL:g}",
           "rtol": RTOL, "l2": "inputs larger than L2 (every PCG vector >= 4x the 126 MB L2 at "
                               "400^3); no flush needed"}
    if Np:
        cfg.update(Np=int(Np), Nt=int(Nt))
    if extra:
        cfg.update(extra)
    cfg.setdefault("solver", solver)
    return cfg


# ------------------------------------------------------------------ GPU arm
def matched_leg(b2, n, steps):
    """The GPU arm on the reference arm's own problem (Cubic n^3 through the public API,
    host arrays in, x and rate out): one same-problem time next to the CPU's."""
    import torch
    pn = b2.Cubic([n, n, n], spacing=1e-4, coords=False)
    ph = b2.Phase(pn)
    ph["throat.hydraulic_conductance"] = conductance(pn.Nt)
    out = {}
    for precond in ("jacobi", "amg"):
        sf = b2.StokesFlow(network=pn, phase=ph)
        sf.set_value_BC(pores=pn.pores("left"), values=P_IN)
        sf.set_value_BC(pores=pn.pores("right"), values=P_OUT)
        sf.settings["cache"] = False
        solver = b2.B200PCG(tol=RTOL, precond=precond)
        sf.run(solver=solver)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            sf.run(solver=solver)
            q = sf.rate(pores=pn.pores("left"))[0]
        torch.cuda.synchronize()
        out[precond] = {"time_s": (time.perf_counter() - t0) / steps,
                        "iters": int(sf.soln.cg_iters), "rate_inlet": float(q)}
        solver.ctx.close()
    return out


def gpu_arm(args):
    import torch
    import openpnm_b200 as b2
    from openpnm_b200 import dist as bdist
    from openpnm_b200.device import Context, pinned_copy

    rank, world, local = bdist.init_process_group()
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            # plain `python bench.py --gpus N`: re-launch under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                   f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
                   "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)]
            cmd += sys.argv[1:]
            return subprocess.call(cmd)
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    import torch.distributed as dist

    n = args.size
    shape = (n, n, n)
    Np = n ** 3
    Nt = 3 * n * n * (n - 1)
    ranges = bdist.cubic_slab_ranges(shape, world)
    rb, re = ranges[rank]
    x0p, x1p = rb // (n * n), re // (n * n)

    # ---- host inputs (outside every timed region): the network as the API user holds
    # it (lazy lattice: no rank materialises the full throat list unless it needs it),
    # the full conductance array, the BC array
    t_host = time.perf_counter()
    pn = b2.Cubic(shape, spacing=1e-4, coords=False, lazy=world > 1)
    g_full = pinned_copy(conductance(Nt))
    if world == 1:
        conns_h, g_h = pn.conns, g_full
    else:
        ids, conns_h = bdist.cubic_slab_throats(shape, x0p, x1p)
        g_h = g_full[ids]
        del ids
    left = np.arange(0, n * n, dtype=np.int64)
    right = np.arange(Np - n * n, Np, dtype=np.int64)
    bcv_h = np.full(Np, np.nan)
    bcv_h[left] = P_IN
    bcv_h[right] = P_OUT
    conns_h, bcv_h = pinned_copy(conns_h), pinned_copy(bcv_h)
    if world > 1:
        g_h = pinned_copy(g_h)
    t_host = time.perf_counter() - t_host

    ss = bdist.SlabSolver(Np, rb, re, device=local, fmt=args.fmt)
    ctx = ss.ctx
    p2p = False
    conns_d, g_d, bcv_d = ctx.to_device(conns_h), ctx.to_device(g_h), ctx.to_device(bcv_h)
    left_d = left[(left >= rb) & (left < re)]
    right_d = right[(right >= rb) & (right < re)]
    ctx.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device=ctx.tdev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(vs):
        t = torch.tensor([float(v) for v in vs], dtype=torch.float64, device=ctx.tdev)
        if world > 1:
            dist.all_reduce(t)
        return [float(v) for v in t.tolist()]

    def device_step(rtol=RTOL):
        f, nnz0, nnz = ss.assemble(conns_d, g_d, bcv_d, None)
        x, it, rel, info = ss.solve(rtol=rtol)
        return x, it, rel, info, nnz

    def timed(fn, reps):
        """max over ranks of the CUDA-event time of `reps` calls, per call [ms]"""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ctx.stream):
            e0.record()
            for _ in range(reps):
                out = fn()
            e1.record()
        barrier()
        return allmax(e0.elapsed_time(e1)) / reps, out

    def rates(x_loc):
        xg = ss.gather(x_loc)
        q = allsum([ss.rate(xg, left_d), ss.rate(xg, right_d)])
        return q[0], q[1], xg

    if world > 1 and not args.no_p2p:
        ss.assemble(conns_d, g_d, bcv_d, None)
        p2p = ss.enable_p2p()        # CUDA-IPC mapping of the peers: collectives fused in-kernel
    for _ in range(args.warmup):
        x, it, rel, info, nnz = device_step()
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    launches0 = ctx.stats()["kernel_launches"]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(ctx.stream):
        ev0.record()
        iters = []
        t_solve_ms = []
        for _ in range(args.steps):
            x, it, rel, info, nnz = device_step()
            iters.append(it)
            t_solve_ms.append(ctx.stats()["t_solve_ms"])
        ev1.record()
    barrier()
    launches = ctx.stats()["kernel_launches"] - launches0
    clk = clocks.stop() if clocks else None
    ms_step = allmax(ev0.elapsed_time(ev1)) / args.steps
    nnz_global = int(allsum([nnz])[0])
    cg_iters = int(np.mean(iters))
    value = Np * cg_iters / (ms_step * 1e-3)

    # ---- parity-side checks on the timed result: conservation
    q_in, q_out, _xg = rates(x)
    del _xg

    # ---- dominant-kernel timing (live, CUDA events on the launching stream)
    fmt_code, nd = ctx.pcg_format()
    t_ka = ctx.time_kernel(0, 30)
    t_kb = ctx.time_kernel(1, 30)
    t_kc = ctx.time_kernel(2, 30)
    n_loc = re - rb
    nnz_loc = nnz
    b_spmv = 12 * nnz_loc + 4 * (n_loc + 1) + 16 * n_loc          # SURVEY 8d CSR model
    b_iter = b_spmv + 88 * n_loc
    if fmt_code == 2:          # nd value arrays + p + r read, Ap written
        fmt_spmv = (nd + 3) * 8 * n_loc
    else:                      # off-diagonal (val, col) pairs + row pointers + p, r, Ap
        fmt_spmv = 12 * (nnz_loc - n_loc) + 4 * n_loc + 24 * n_loc
    fmt_iter = fmt_spmv + 56 * n_loc          # K_BC: read x,r,p,Ap, write x,r,p
    peak, peak_src = peaks()
    traffic, traffic_src = None, None
    try:     # DRAM bytes per launch from the committed `ncu --set full` capture of K_A
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tk = json.load(f)["cubic400_1gpu_dia"]["k_dia_spmv_dot"]
        per_row = (tk["dram_read_bytes"] + tk["dram_write_bytes"]) / 400 ** 3
        if fmt_code == 2 and nd == 3:
            traffic = per_row * n_loc
            traffic_src = ("ncu --set full capture of this kernel at Cubic 400^3 on 1 GPU "
                           "(profiles/r01_ncu_full_kA_kBC_v2.csv)" if world == 1 and n == 400 else
                           f"{per_row:.1f} B/row from that capture x the {n_loc} rows of one launch "
                           "here (ncu replays cannot run under torchrun)")
    except Exception:
        traffic = None
    ms_iter = float(np.mean(t_solve_ms)) / max(cg_iters, 1)

    # ---- the same step with the multigrid preconditioner, at every N: assemble, build
    # the hierarchy, solve to the same rtol -- time-to-solve is the metric here, an AMG
    # iteration is not comparable to a PCG iteration
    amg = None
    if not args.no_amg:
        ctx.set_precond("amg")
        try:
            device_step()                               # warm-up (allocates the hierarchy)
            l0 = ctx.stats()["kernel_launches"]
            ms_amg, (xa, ita, rela, infoa, _nnz) = timed(device_step, args.steps)
            l_amg = (ctx.stats()["kernel_launches"] - l0) // args.steps
            sta = ctx.stats()
            dx = allmax(float((xa - x).abs().max())) / allmax(float(x.abs().max()))
            qa_in, qa_out, _xg = rates(xa)
            del _xg
            nlev = len(ctx.amg_setup())
            lv = [ctx.amg_level_info(l) for l in range(nlev)]
            amg = {"time_to_solve_s": ms_amg * 1e-3, "iters": int(ita), "relres": rela,
                   "info": int(infoa), "hierarchy_setup_ms": sta["t_setup_ms"],
                   "solve_ms": sta["t_solve_ms"], "gpu_launches_per_step": int(l_amg),
                   "levels": [[i["n_global"], "rows" if i["partitioned"] else "full"] for i in lv],
                   "max_rel_diff_vs_jacobi_pcg": dx,
                   "rate_inlet": qa_in, "rate_outlet": qa_out,
                   "speedup_vs_jacobi_pcg": ms_step / ms_amg,
                   "what": "B200PCG(precond='amg'): aggregation-AMG K-cycle + flexible CG "
                           "(csrc/amg.cu); step = assembly + hierarchy build + solve"
                           + ("" if world == 1 else "; slabs: ONE hierarchy, levels marked 'rows' "
                              "are row-partitioned by the slabs, 'full' ones replicated")}
            del xa
        finally:
            ctx.set_precond("jacobi")

    # ---- parity at this size (north star: x and rate within 1e-8 at CG rtol 1e-12):
    # the two independent solvers at rtol 1e-12 and 1e-13 against each other, residuals on
    # the canonical CSR (b200_spmv(B200_MAT_SYSTEM)), and -- for N > 1 -- against a
    # single-GPU solve of the whole problem made by rank 0 in a context of its own
    parity = None
    if not args.no_parity:
        parity = {"what": "max |x_a - x_b| / max |x_b| and relative inlet-rate differences between "
                          "Jacobi-PCG and multigrid-FCG solves of this very system; residual = "
                          "|b - A x|_2/|b|_2 on the canonical CSR", "runs": []}
        sols = {}
        b_loc = ctx.get_b()
        bn2 = allsum([float((b_loc * b_loc).sum())])[0]
        for rtol in (1e-12, 1e-13):
            for precond in (("jacobi", "amg") if not args.no_amg else ("jacobi",)):
                ctx.set_precond(precond)
                ms_p, (xp, itp, relp, infop, _nnz) = timed(lambda: device_step(rtol), 1)
                qi, qo, xg = rates(xp)
                r = ctx.spmv(1, xg) - ctx.get_b()
                res = (allsum([float((r * r).sum())])[0] / bn2) ** 0.5
                del r, xg
                sols[(rtol, precond)] = (xp.clone(), qi, qo)
                parity["runs"].append({"rtol": rtol, "precond": precond, "iters": int(itp),
                                       "info": int(infop), "time_to_solve_s": ms_p * 1e-3,
                                       "residual_canonical_csr": res, "rate_inlet": qi,
                                       "rate_outlet": qo,
                                       "conservation": abs(qi + qo) / abs(qi) if qi else None})
        ctx.set_precond("jacobi")

        def diff(a, b):
            xa_, qa_, _ = sols[a]
            xb_, qb_, _ = sols[b]
            return {"x": allmax(float((xa_ - xb_).abs().max())) / allmax(float(xb_.abs().max())),
                    "rate_inlet": abs(qa_ - qb_) / abs(qb_)}
        pairs = {}
        if not args.no_amg:
            pairs["amg_vs_jacobi@1e-12"] = diff((1e-12, "amg"), (1e-12, "jacobi"))
            pairs["amg_vs_jacobi@1e-13"] = diff((1e-13, "amg"), (1e-13, "jacobi"))
            pairs["amg@1e-12_vs_amg@1e-13"] = diff((1e-12, "amg"), (1e-13, "amg"))
        pairs["jacobi@1e-12_vs_jacobi@1e-13"] = diff((1e-12, "jacobi"), (1e-13, "jacobi"))
        # the benchmarked tolerance against the tight solve: what rtol 1e-10 really delivers
        sols[(RTOL, "jacobi")] = (x, q_in, q_out)
        pairs["jacobi@1e-10_vs_jacobi@1e-13"] = diff((RTOL, "jacobi"), (1e-13, "jacobi"))
        if world > 1 and not args.no_cross_check:
            # rank 0 alone solves the whole problem on its GPU (multigrid, rtol 1e-13) and
            # compares the gathered slab solution with it: N GPUs against 1 GPU in one run
            ref_key = (1e-13, "amg") if not args.no_amg else (1e-13, "jacobi")
            xg = ss.gather(sols[ref_key][0])
            cross = None
            if rank == 0:
                ctx1 = Context(local)
                ctx1.set_precond("amg" if not args.no_amg else "jacobi")
                ctx1.build_laplacian(pn.conns, g_full, Np)
                ctx1.apply_bcs(bcv_h, None)
                x1, it1, rel1, info1 = ctx1.pcg(rtol=1e-13)
                q1 = float(ctx1.rate_pores(x1, left).sum().item())
                cross = {"x": float((xg - x1).abs().max() / x1.abs().max()),
                         "rate_inlet": abs(sols[ref_key][1] - q1) / abs(q1),
                         "single_gpu_iters": int(it1), "single_gpu_info": int(info1)}
                del x1
                ctx1.close()
            del xg
            barrier()
            if cross is not None:
                pairs[f"{world}_gpus_vs_1_gpu@1e-13"] = cross
        parity["pairs"] = pairs
        if rank == 0:
            tight = {k: v for k, v in pairs.items() if "1e-10" not in k}
            parity["max_pairwise_x_at_rtol<=1e-12"] = max(v["x"] for v in tight.values())
            parity["max_pairwise_rate_at_rtol<=1e-12"] = max(v["rate_inlet"] for v in tight.values())
            parity["within_1e-8"] = bool(parity["max_pairwise_x_at_rtol<=1e-12"] <= 1e-8 and
                                         parity["max_pairwise_rate_at_rtol<=1e-12"] <= 1e-8)
        sols.clear()

    # ---- e2e through the public API with HOST buffers: every step copies its inputs
    # from pinned host memory, checks the topology on the device, solves and reads the
    # solution + the inlet rate back.  The same call on 1 GPU and under torchrun.
    ph = b2.Phase(pn)
    dict.__setitem__(ph, "throat.hydraulic_conductance", g_full)      # the pinned buffer
    sf = b2.StokesFlow(network=pn, phase=ph)
    if world == 1:
        dict.__setitem__(pn, "throat.conns", conns_h)
    dict.__setitem__(sf, "pore.bc.value", bcv_h)
    dict.__setitem__(sf, "pore.bc.rate", pinned_copy(np.full(Np, np.nan)))
    sf.settings["cache"] = False                # re-upload + re-assemble every step

    def e2e_leg(precond):
        solver = b2.B200PCG(tol=RTOL, fmt=args.fmt, precond=precond, device=local)
        if world == 1:
            solver._ctx = ctx
            ctx.set_precond(precond)
        else:
            solver._slab, solver._slab_p2p = ss, (p2p if precond == "jacobi" else False)
            solver._slab_key = (id(pn), world, rank, solver.device, solver.fmt)
        sf.run(solver=solver)                        # warm-up
        barrier()
        t0 = time.perf_counter()
        its_ = []
        reps = min(args.steps, 5)                    # host-buffer legs: a few steps suffice
        for _ in range(reps):
            sf.run(solver=solver)
            qq = sf.rate(pores=left)[0]
            its_.append(sf.soln.cg_iters)
        barrier()
        return allmax((time.perf_counter() - t0) / reps), float(np.mean(its_)), float(qq)

    t_e2e, it_e2e, qq = e2e_leg("jacobi")
    if world == 1:
        h2d = conns_h.nbytes + g_h.nbytes + 2 * bcv_h.nbytes
    else:     # per rank: its slab's throats + conductances, and the window of the two BC arrays
        # it reads (its rows + one halo plane on either side), summed over the ranks
        win = min(re + n * n, Np) - max(rb - n * n, 0)
        h2d = int(allsum([conns_h.nbytes + g_h.nbytes + 2 * 8 * win])[0])
    e2e = {"value": Np * it_e2e / t_e2e, "unit": "DOF*iter/s",
           "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(world * (8 * Np + 8 * left.size)),
           "ms_per_step": 1e3 * t_e2e, "time_to_solve_s": t_e2e,
           "api": "openpnm_b200.StokesFlow.run(solver=B200PCG()) + rate()"
                  + ("" if world == 1 else " under torchrun (every rank: its slab's inputs up, "
                     "the full x down)"),
           "validate_topology": bool(sf.settings["validate_topology"]),
           "rate_inlet": qq}
    if amg is not None:
        ta, ia, qa = e2e_leg("amg")
        amg["e2e_time_to_solve_s"] = ta
        amg["e2e_rate_inlet"] = qa
        amg["e2e_iters"] = ia
        ctx.set_precond("jacobi")

    cpu = None
    matched = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r = reference_openpnm_sample(args.ref_size)
        if r is not None:
            cpu = {"value": r["Np"] * r["iters"] / r["t_step"], "unit": "DOF*iter/s", "cores": 1,
                   "kind": "reference", "time_to_solve_s": r["t_step"], "iters": r["iters"],
                   "rate_inlet": r["rate"],
                   "sample": f"unmodified OpenPNM StokesFlow.run(solver=ScipyCG(tol={RTOL:g})) + "
                             f"rate() on Cubic {args.ref_size}^3, same distribution/BCs: "
                             f"{r['iters']} iterations, {r['t_step']:.1f} s; SciPy CG is serial. "
                             + pardiso_note()}
        else:
            r = cpu_solve_sample(args.cpu_size)
            cpu = {"value": r["Np"] * r["iters"] / r["t_step"], "unit": "DOF*iter/s", "cores": 1,
                   "kind": "port", "time_to_solve_s": r["t_step"], "iters": r["iters"],
                   "rate_inlet": float(r["rate"]),
                   "sample": f"oracle port on Cubic {args.cpu_size}^3 same distribution/BCs/rtol, "
                             f"{r['iters']} iterations, {r['t_step']:.1f} s; numpy assembly + SciPy "
                             f"CSR SpMV Jacobi-PCG, serial. " + pardiso_note()}
        # the GPU arm on that very problem, through the same public call
        size_m = args.ref_size if cpu["kind"] == "reference" else args.cpu_size
        m = matched_leg(b2, size_m, max(args.steps, 3))
        matched = {"workload": f"Cubic {size_m}^3 StokesFlow, the CPU arm's problem, through "
                               "StokesFlow.run + rate from host arrays",
                   "gpu": m, "cpu_time_s": cpu["time_to_solve_s"], "cpu_iters": cpu["iters"],
                   "same_config_time_ratio": cpu["time_to_solve_s"] / m["jacobi"]["time_s"],
                   "same_config_time_ratio_amg": cpu["time_to_solve_s"] / m["amg"]["time_s"],
                   "rate_inlet_rel_diff_vs_cpu":
                       abs(m["jacobi"]["rate_inlet"] - cpu["rate_inlet"]) / abs(cpu["rate_inlet"])}

    if rank == 0:
        line = {
            "metric": "stokesflow_dof_iter_per_s", "value": value, "unit": "DOF*iter/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": workload_config(n, Np, Nt, {
                "nnz": nnz_global, "cg_iters": cg_iters, "relres": rel, "info": info,
                "format": {1: "csr", 2: f"dia{nd}-sym"}.get(fmt_code, "?"),
                "parallelism": f"slab{world}", "host_prep_s": round(t_host, 1),
                "comm": ("none" if world == 1 else
                         "in-kernel peer stores over NVLink (CUDA IPC)" if p2p else "NCCL")}),
            "time_to_solve_s": ms_step * 1e-3,
            "ms_per_cg_iter": ms_iter,
            "rate_inlet": q_in, "rate_outlet": q_out,
            "conservation": abs(q_in + q_out) / abs(q_in) if q_in else None,
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": {"bound": "hbm", "kernel": "K_A SpMV + 3 dots (k_dia_spmv_dot / "
                         "k_csr_stream_spmv_dot)", "achieved": b_spmv / (t_ka * 1e-3) / 1e9,
                         "peak": peak, "unit": "GB/s",
                         "frac": b_spmv / (t_ka * 1e-3) / 1e9 / peak, "traffic": traffic,
                         "traffic_source": traffic_src,
                         "peak_source": peak_src, "ms": t_ka,
                         "algorithmic_bytes": int(b_spmv), "format_bytes": int(fmt_spmv),
                         "frac_format_bytes": fmt_spmv / (t_ka * 1e-3) / 1e9 / peak},
            "roofline_iter": {"bound": "hbm", "kernel": "one PCG iteration (K_A SpMV + dots, K_BC fused x/r/p update + r.r)",
                              "achieved": b_iter / (ms_iter * 1e-3) / 1e9, "peak": peak,
                              "unit": "GB/s", "frac": b_iter / (ms_iter * 1e-3) / 1e9 / peak,
                              "algorithmic_bytes": int(b_iter), "format_bytes": int(fmt_iter),
                              "frac_format_bytes": fmt_iter / (ms_iter * 1e-3) / 1e9 / peak,
                              "ms": ms_iter, "kernel_ms": {"K_A": t_ka, "K_BC": t_kb, "K_BC_with_true_norm": t_kc}},
        }
        if amg:
            line["amg"] = amg
        if parity:
            line["parity"] = parity
        line["e2e"] = e2e
        if cpu:
            line["cpu_baseline"] = cpu
        if matched:
            line["matched"] = matched
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=int(os.environ.get("B200_BENCH_SIZE", "400")))
    ap.add_argument("--cpu-size", type=int, default=100)
    ap.add_argument("--ref-size", type=int, default=64)
    ap.add_argument("--fmt", default="auto", choices=["auto", "csr", "dia"])
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-p2p", action="store_true")
    ap.add_argument("--no-amg", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-cross-check", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    return gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
