#!/usr/bin/env python
"""bench.py -- StokesFlow time-to-solve and DOF*iter/s on a synthetic Cubic
network (BASELINE.json: Cubic 400^3, 64M pores, Jacobi-PCG rtol 1e-10), on
N B200s of one box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--size 400]
    python bench.py --impl reference ...      # CPU arm: oracle port + SciPy

One "step" = one full pass of the hot path on the resident network:
assemble the Laplacian from conns + g, apply the inlet/outlet Dirichlet BCs,
Jacobi-PCG to rtol 1e-10, rate() through the inlet and outlet.  `value` has
the inputs already in HBM; `e2e` goes through the public API
(`StokesFlow.run`) from pinned HOST arrays and reads x back every step.
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


# ------------------------------------------------------------------ CPU arm
def cpu_solve_sample(n, rtol=RTOL):
    """The reference's CPU path on a Cubic n^3 sample: numpy assembly exactly as
    `_build_A`/`_apply_BCs` (oracle restatement) + SciPy CSR SpMV inside a
    Jacobi-PCG with ScipyCG's stopping rule.  Returns dict with timings."""
    from oracle import oracle
    t0 = time.perf_counter()
    net = oracle.cubic_network([n, n, n], spacing=1e-4)
    g = conductance(net["Nt"])
    bcv = np.full(net["Np"], np.nan)
    bcv[net["labels"]["left"]] = P_IN
    bcv[net["labels"]["right"]] = P_OUT
    t1 = time.perf_counter()
    A, b, f = oracle.apply_bcs(oracle.build_A(net["conns"], g, net["Np"]), net["Np"], bcv, None)
    M = oracle.to_csr(A, net["Np"])
    t2 = time.perf_counter()
    x, info, its = oracle.jacobi_pcg(M, b, rtol=rtol)
    t3 = time.perf_counter()
    q = oracle.rate(net["conns"], g, x, pores=net["labels"]["left"])[0]
    t4 = time.perf_counter()
    return dict(Np=net["Np"], iters=its, info=info, t_gen=t1 - t0, t_assemble=t2 - t1,
                t_solve=t3 - t2, t_rate=t4 - t3, t_step=t4 - t1, rate=q)


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
    if real is not None:
        for _ in range(max(args.warmup - 1, 0)):
            reference_openpnm_sample(min(n, 24))
        t, its, Np = [], 0, 0
        for _ in range(args.steps):
            r = reference_openpnm_sample(n)
            t.append(r["t_step"])
            its, Np = r["iters"], r["Np"]
        kind = "reference"
        sample = (f"unmodified OpenPNM StokesFlow.run(solver=ScipyCG(tol={RTOL:g})) + rate() on "
                  f"Cubic {n}^3 (same g distribution/BCs as the GPU arm's Cubic {args.size}^3): "
                  f"its own topology check, 3 assemblies and SciPy's serial CG, {its} iterations")
    else:
        n = args.cpu_size
        for _ in range(args.warmup):
            cpu_solve_sample(min(n, 40))
        t, its, Np = [], 0, 0
        for _ in range(args.steps):
            r = cpu_solve_sample(n)
            t.append(r["t_step"])
            its, Np = r["iters"], r["Np"]
        kind = "port"
        sample = (f"Cubic {n}^3 StokesFlow (same g distribution/BCs/rtol as the GPU arm's Cubic "
                  f"{args.size}^3): numpy assembly as the reference + SciPy CSR SpMV Jacobi-PCG, "
                  f"{its} iterations; SciPy's SpMV and CG are serial")
    ms = 1e3 * float(np.mean(t))
    val = Np * its / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": "stokesflow_dof_iter_per_s", "value": val,
        "unit": "DOF*iter/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.size, None, None, None),
        "time_to_solve_s": ms * 1e-3,
        "cpu_baseline": {"value": val, "unit": "DOF*iter/s", "cores": 1, "kind": kind,
                         "blas_threads": threads, "sample": sample},
        "e2e": {"value": val, "unit": "DOF*iter/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def workload_config(size, Np, Nt, extra):
    cfg = {"workload": f"Cubic {size}x{size}x{size} StokesFlow (g=10^U(-15,-12) seed 0, "
                       f"left={P_IN:g}/right={P_OUT:g} Dirichlet), Jacobi-PCG rtol {RTOL:g}",
           "rtol": RTOL, "l2": "inputs larger than L2 (every PCG vector >= 4x the 126 MB L2 at "
                               "400^3); no flush needed"}
    if Np:
        cfg.update(Np=int(Np), Nt=int(Nt))
    if extra:
        cfg.update(extra)
    return cfg


# ------------------------------------------------------------------ GPU arm
def gpu_arm(args):
    import torch
    import openpnm_b200 as b2
    from openpnm_b200 import dist as bdist
    from openpnm_b200.device import pinned_copy

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

    # ---- host inputs (outside every timed region)
    t_host = time.perf_counter()
    if world == 1:
        pn = b2.Cubic(shape, spacing=1e-4, coords=False)
        conns_h = pn.conns
        g_h = conductance(Nt)
        left, right = pn.pores("left"), pn.pores("right")
    else:
        ids, conns_h = bdist.cubic_slab_throats(shape, x0p, x1p)
        g_h = conductance(Nt, ids)
        del ids
        left = np.arange(0, n * n, dtype=np.int64)
        right = np.arange(Np - n * n, Np, dtype=np.int64)
    bcv_h = np.full(Np, np.nan)
    bcv_h[left] = P_IN
    bcv_h[right] = P_OUT
    conns_h, g_h, bcv_h = pinned_copy(conns_h), pinned_copy(g_h), pinned_copy(bcv_h)
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

    def device_step():
        f, nnz0, nnz = ss.assemble(conns_d, g_d, bcv_d, None)
        x, it, rel, info = ss.solve(rtol=RTOL)
        return x, it, rel, info, nnz

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
    ms_total = ev0.elapsed_time(ev1)
    launches = ctx.stats()["kernel_launches"] - launches0
    clk = clocks.stop() if clocks else None
    tmax = torch.tensor([ms_total], dtype=torch.float64, device=ctx.tdev)
    nnz_t = torch.tensor([float(nnz)], dtype=torch.float64, device=ctx.tdev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(nnz_t)
    ms_step = float(tmax.item()) / args.steps
    nnz_global = int(nnz_t.item())
    cg_iters = int(np.mean(iters))
    value = Np * cg_iters / (ms_step * 1e-3)

    # ---- parity-side checks on the timed result: residual + conservation
    xg = ss.gather(x)
    q_in = torch.tensor([ss.rate(xg, left_d), ss.rate(xg, right_d)], dtype=torch.float64,
                        device=ctx.tdev)
    if world > 1:
        dist.all_reduce(q_in)
    q_in, q_out = float(q_in[0]), float(q_in[1])

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
    traffic = None
    try:     # DRAM bytes per launch from the committed ncu capture of this very configuration
        if world == 1 and n == 400 and fmt_code == 2:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                tk = json.load(f)["cubic400_1gpu_dia"]["k_dia_spmv_dot"]
            traffic = tk["dram_read_bytes"] + tk["dram_write_bytes"]
    except Exception:
        traffic = None
    ms_iter = float(np.mean(t_solve_ms)) / max(cg_iters, 1)

    # ---- the same step with the multigrid preconditioner (1 GPU): assemble, build
    # the hierarchy, solve to the same rtol -- time-to-solve is the metric here,
    # an AMG iteration is not comparable to a PCG iteration
    amg = None
    if not args.no_amg:
        ctx.set_precond("amg")
        try:
            device_step()                               # warm-up (allocates the hierarchy)
            barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(ctx.stream):
                a0.record()
                for _ in range(args.steps):
                    xa, ita, rela, infoa, _nnz = device_step()
                    sta = ctx.stats()
                a1.record()
            barrier()
            t_amg = torch.tensor([a0.elapsed_time(a1) / args.steps], dtype=torch.float64,
                                 device=ctx.tdev)
            if world > 1:
                dist.all_reduce(t_amg, op=dist.ReduceOp.MAX)
            ms_amg = float(t_amg.item())
            dx = float((xa - x).abs().max() / x.abs().max())
            amg = {"time_to_solve_s": ms_amg * 1e-3, "iters": int(ita), "relres": rela,
                   "info": int(infoa), "hierarchy_setup_ms": sta["t_setup_ms"],
                   "solve_ms": sta["t_solve_ms"],
                   "levels": ctx.amg_setup(),
                   "max_rel_diff_vs_jacobi_pcg": dx,
                   "speedup_vs_jacobi_pcg": ms_step / ms_amg,
                   "what": "B200PCG(precond='amg'): aggregation-AMG K-cycle + flexible CG "
                           "(csrc/amg.cu); step = assembly + hierarchy build + solve"
                           + ("" if world == 1 else "; slabs: one hierarchy, levels row-partitioned "
                              "by the slabs, small levels replicated")}
            del xa
        finally:
            ctx.set_precond("jacobi")
            ss.assemble(conns_d, g_d, bcv_d, None)

    # ---- e2e through the public API with HOST buffers: every step copies its
    # inputs from pinned host memory and reads the solution back
    e2e = None
    if world == 1:
        ph = b2.Phase(pn)
        dict.__setitem__(ph, "throat.hydraulic_conductance", g_h)      # the pinned buffers
        sf = b2.StokesFlow(network=pn, phase=ph)
        dict.__setitem__(pn, "throat.conns", conns_h)
        dict.__setitem__(sf, "pore.bc.value", bcv_h)
        dict.__setitem__(sf, "pore.bc.rate", pinned_copy(np.full(Np, np.nan)))
        sf.settings["validate_topology"] = False    # reference's check is O(minutes) at 64M pores
        sf.settings["cache"] = False                # re-upload + re-assemble every step
        solver = b2.B200PCG(tol=RTOL, fmt=args.fmt)
        solver._ctx = ctx
        sf.run(solver=solver)                        # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_iters = []
        for _ in range(args.steps):
            sf.run(solver=solver)
            qq = sf.rate(pores=left)[0]
            e2e_iters.append(sf.soln.cg_iters)
        torch.cuda.synchronize()
        t_e2e = (time.perf_counter() - t0) / args.steps
        e2e = {"value": Np * float(np.mean(e2e_iters)) / t_e2e, "unit": "DOF*iter/s",
               "h2d_bytes_per_step": int(conns_h.nbytes + g_h.nbytes + 2 * bcv_h.nbytes),
               "d2h_bytes_per_step": int(8 * Np + 8 * left.size),
               "ms_per_step": 1e3 * t_e2e, "api": "openpnm_b200.StokesFlow.run + rate",
               "rate_inlet": float(qq)}
        if amg is not None:
            # the same public call with the multigrid preconditioner (host buffers in and out)
            ctx.set_precond("amg")
            try:
                sf.run(solver=solver)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(args.steps):
                    sf.run(solver=solver)
                    qa = sf.rate(pores=left)[0]
                torch.cuda.synchronize()
                amg["e2e_time_to_solve_s"] = (time.perf_counter() - t0) / args.steps
                amg["e2e_rate_inlet"] = float(qa)
            finally:
                ctx.set_precond("jacobi")
    else:
        # slab API (openpnm_b200.dist.SlabSolver): each rank uploads the throats of its
        # slab + the BC array, assembles, solves and downloads its part of x
        x_h = pinned_copy(np.zeros(re - rb))

        def e2e_step():
            ss.assemble(conns_h, g_h, bcv_h, None)
            xl, it_, rel_, info_ = ss.solve(rtol=RTOL)
            with torch.cuda.stream(ctx.stream):
                torch.from_numpy(x_h).copy_(xl, non_blocking=True)
            ctx.synchronize()
            return it_

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        e2e_iters = [e2e_step() for _ in range(args.steps)]
        barrier()
        t_e2e = torch.tensor([(time.perf_counter() - t0) / args.steps], dtype=torch.float64,
                             device=ctx.tdev)
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        t_e2e = float(t_e2e.item())
        nbytes = torch.tensor([float(conns_h.nbytes + g_h.nbytes + bcv_h.nbytes),
                               float(x_h.nbytes)], dtype=torch.float64, device=ctx.tdev)
        dist.all_reduce(nbytes)
        e2e = {"value": Np * float(np.mean(e2e_iters)) / t_e2e, "unit": "DOF*iter/s",
               "h2d_bytes_per_step": int(nbytes[0].item()),
               "d2h_bytes_per_step": int(nbytes[1].item()), "ms_per_step": 1e3 * t_e2e,
               "api": "openpnm_b200.dist.SlabSolver.assemble + solve, x to host (all ranks)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r = reference_openpnm_sample(args.ref_size)
        if r is not None:
            cpu = {"value": r["Np"] * r["iters"] / r["t_step"], "unit": "DOF*iter/s", "cores": 1,
                   "kind": "reference",
                   "sample": f"unmodified OpenPNM StokesFlow.run(solver=ScipyCG(tol={RTOL:g})) + "
                             f"rate() on Cubic {args.ref_size}^3, same distribution/BCs: "
                             f"{r['iters']} iterations, {r['t_step']:.1f} s; SciPy CG is serial"}
        else:
            r = cpu_solve_sample(args.cpu_size)
            cpu = {"value": r["Np"] * r["iters"] / r["t_step"], "unit": "DOF*iter/s", "cores": 1,
                   "kind": "port",
                   "sample": f"Cubic {args.cpu_size}^3 same distribution/BCs/rtol, {r['iters']} "
                             f"iterations, {r['t_step']:.1f} s (assembly {r['t_assemble']:.1f} s); "
                             f"numpy assembly + SciPy CSR SpMV Jacobi-PCG, serial"}

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
        if e2e:
            line["e2e"] = e2e
        if cpu:
            line["cpu_baseline"] = cpu
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
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    return gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
