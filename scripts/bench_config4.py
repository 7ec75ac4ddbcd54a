"""BASELINE config 4: Delaunay network (irregular coordination), FickianDiffusion, isolated
pores trimmed -- on 1 GPU or, under torchrun, slab-partitioned over N GPUs.

    python scripts/bench_config4.py ship/delaunay_<N>[_mB].npz [--reps 3]
    python -m torch.distributed.run --nproc-per-node 8 ... scripts/bench_config4.py <file>

The network comes from scripts/gen_delaunay.py + pack_delaunay.py (scipy's Delaunay of
default_rng(0) points, pores numbered by coordinate, over-long hull throats cut).  Here:
device health check -> drop isolated pores / minor clusters on the device (topotools) ->
`FickianDiffusion.run(solver=B200PCG(...))` with Jacobi and with the multigrid, x < 0.01
-> 1, x > 0.99 -> 0.  Reported: conservation, independent residual on the canonical CSR,
iterations, times, and the K_A kernel (SELL-32 SpMV + dots) against the HBM peak.
Prints one JSON line (rank 0)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))


def main():
    import torch
    import torch.distributed as dist
    import openpnm_b200 as b2
    from openpnm_b200 import dist as bdist, topotools
    import pack_delaunay

    path = sys.argv[1]
    reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 3
    rtol = float(sys.argv[sys.argv.index("--rtol") + 1]) if "--rtol" in sys.argv else 1e-10
    rank, world, local = bdist.init_process_group()
    t0 = time.perf_counter()
    conns, x = pack_delaunay.unpack(path)
    Np0 = x.size
    t_load = time.perf_counter() - t0

    # ---- health + trim on the device (every rank does the same, deterministic)
    hctx = b2.B200PCG(device=local).ctx
    conns_d = hctx.to_device(conns, torch.int64)
    hctx.synchronize()
    t0 = time.perf_counter()
    conns_new, pmap, tmap, Np, counts = topotools.trim_to_largest_cluster(hctx, conns_d, Np0)
    hctx.synchronize()
    t_trim = time.perf_counter() - t0
    keep = (pmap >= 0).cpu().numpy()
    xk = x[keep].astype(np.float64)
    conns_h = conns_new.cpu().numpy()
    tkeep = (tmap >= 0).cpu().numpy()
    del conns_d, conns_new, pmap, tmap, conns
    hctx.close()

    pn = b2.Network(conns=conns_h, Np=Np)
    Nt = pn.Nt
    g = (10.0 ** np.random.default_rng(0).uniform(-15, -12, tkeep.size))[tkeep]
    ph = b2.Phase(pn)
    dict.__setitem__(ph, "throat.diffusive_conductance", g)
    bcv = np.full(Np, np.nan)
    bcv[xk < 0.01] = 1.0
    bcv[xk > 0.99] = 0.0
    inlet, outlet = np.flatnonzero(bcv == 1.0), np.flatnonzero(bcv == 0.0)
    band = bdist.bandwidth(conns_h)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device=torch.device("cuda", local))
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if "--ka-only" in sys.argv:
        # kernel A/B runs: assemble, prepare the operator, time K_A, nothing else
        ctx = b2.B200PCG(device=local).ctx
        ctx.build_laplacian(conns_h, g, Np)
        ctx.apply_bcs(bcv, None)
        ctx.pcg(rtol=1e-2)
        t = ctx.time_kernel(0, 50)
        ob, oe = ctx.operator_bytes()
        print(json.dumps({"K_A_ms": t, "GBps": ob / t / 1e6, "frac": ob / t / 1e6 / 6535.4,
                          "variant": os.environ.get("B200_SELL_VARIANT", "0"),
                          "kernel": os.environ.get("B200_CSR_KERNEL", "sell-staged")}), flush=True)
        return

    out = {"config": f"Delaunay {Np0} pores ({os.path.basename(path)}), FickianDiffusion, "
                     f"g=10^U(-15,-12), x<0.01 -> 1 / x>0.99 -> 0, rtol {rtol:g}",
           "n_gpus": world, "Np_before_trim": int(Np0), "Np": int(Np), "Nt": int(Nt),
           "health": counts, "trim_ms": 1e3 * t_trim, "load_s": round(t_load, 1),
           "index_bandwidth": int(band), "runs": {}}
    xs = {}
    for precond in ("jacobi", "amg"):
        fd = b2.FickianDiffusion(network=pn, phase=ph)
        dict.__setitem__(fd, "pore.bc.value", bcv)
        solver = b2.B200PCG(tol=rtol, precond=precond, device=local)
        fd.run(solver=solver)                  # warm-up: allocations, p2p / NCCL setup
        ts = []
        for _ in range(reps):
            barrier()
            t0 = time.perf_counter()
            fd.run(solver=solver)
            barrier()
            ts.append(time.perf_counter() - t0)
        qi = fd.rate(pores=inlet)[0]
        qo = fd.rate(pores=outlet)[0]
        ctx = fd._ctx
        st = ctx.stats()
        # independent residual on the canonical CSR of the rows this rank holds
        xg = fd._x_dev
        r = ctx.spmv(1, xg) - ctx.get_b()
        b_loc = ctx.get_b()
        rb = torch.stack([(r * r).sum(), (b_loc * b_loc).sum()])
        if world > 1:
            dist.all_reduce(rb)
        res = float((rb[0] / rb[1]).sqrt().item())
        fmt, nd = ctx.pcg_format()
        t_ka = ctx.time_kernel(0, 30)
        ob, oe = ctx.operator_bytes()
        peak = 6535.4
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peak = float(json.load(f)["hbm_gbs"])
        except Exception:
            pass
        out["runs"][precond] = {
            "run_s_api": allmax(float(np.mean(ts))), "solve_ms_device": allmax(st["t_solve_ms"]),
            "setup_ms_device": allmax(st["t_setup_ms"]), "iters": int(fd.soln.cg_iters),
            "converged": bool(fd.soln.is_converged), "rate_in": float(qi), "rate_out": float(qo),
            "conservation": abs(qi + qo) / abs(qi), "residual_canonical_csr": res,
            "fmt": {1: "sell32 (irregular)", 2: f"dia{nd}"}.get(fmt, str(fmt)),
            "K_A_ms": allmax(t_ka), "K_A_format_bytes_per_rank": int(ob),
            "K_A_stored_entries_per_rank": int(oe),
            "K_A_GBps_per_rank": ob / (t_ka * 1e-3) / 1e9,
            "K_A_frac_of_hbm_peak": ob / (t_ka * 1e-3) / 1e9 / peak}
        xs[precond] = fd._x_dev.clone()
        if precond == "amg":
            nlev = len(ctx.amg_setup())
            out["runs"][precond]["levels"] = [
                [i["n_global"], "rows" if i["partitioned"] else "full"]
                for i in (ctx.amg_level_info(l) for l in range(nlev))]
    out["max_rel_diff_amg_vs_jacobi"] = float(
        (xs["amg"] - xs["jacobi"]).abs().max() / xs["jacobi"].abs().max())
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
