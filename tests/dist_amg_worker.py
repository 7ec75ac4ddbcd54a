"""Worker for the multi-GPU multigrid test (torchrun, one rank per GPU): the slab
solve with `precond='amg'` -- flexible CG on the global operator, every rank
preconditioning with the hierarchy of its own block -- against the CPU oracle and
against the slab Jacobi-PCG."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["B200_AMG_SLABS"] = "1"      # the block-wise slab multigrid is opt-in (csrc/amg.cu)


def main():
    import torch
    import torch.distributed as dist
    from openpnm_b200 import dist as bdist
    from oracle import oracle

    rank, world, local = bdist.init_process_group()
    for shape, fmt in (((48, 24, 24), "auto"), ((48, 24, 24), "csr")):
        nx, ny, nz = shape
        net = oracle.cubic_network(shape)
        Np, Nt = net["Np"], net["Nt"]
        g_all = 10.0 ** np.random.default_rng(3).uniform(-15, -12, Nt)
        bcv = np.full(Np, np.nan)
        bcv[net["labels"]["left"]] = 2.0
        bcv[net["labels"]["right"]] = 1.0
        rb, re = bdist.cubic_slab_ranges(shape, world)[rank]
        ids, conns = bdist.cubic_slab_throats(shape, rb // (ny * nz), re // (ny * nz))
        A, b, _ = oracle.apply_bcs(oracle.build_A(net["conns"], g_all, Np), Np, bcv, None)
        xr, _ = oracle.spsolve(oracle.to_csr(A, Np), b)
        ss = bdist.SlabSolver(Np, rb, re, device=local, fmt=fmt)
        out = {}
        for precond in ("jacobi", "amg"):
            ss.ctx.set_precond(precond)
            ss.assemble(conns, g_all[ids], bcv, None)
            ss.ctx.synchronize()
            t0 = time.perf_counter()
            x_loc, it, rel, info = ss.solve(rtol=1e-12)
            ss.ctx.synchronize()
            dt = time.perf_counter() - t0
            assert info == 0, (precond, info, it, rel)
            x = ss.gather(x_loc).cpu().numpy()
            err = np.max(np.abs(x - xr)) / np.max(np.abs(xr))
            assert err < 1e-8, (precond, shape, fmt, err)
            out[precond] = (it, dt)
        # one rank: ~35 iterations; block-Jacobi coupling of the slabs costs iterations
        # (numpy model with exact block solves: 77 / 101 / 126 for 2 / 4 / 8 slabs here)
        assert out["amg"][0] < out["jacobi"][0] / 2 and out["amg"][0] <= (60 if world == 1 else 260), \
            {k: v[0] for k, v in out.items()}
        if rank == 0:
            print(f"dist amg ok shape={shape} fmt={fmt} world={world} jacobi={out['jacobi'][0]} it "
                  f"amg={out['amg'][0]} it ({out['amg'][1]*1e3:.1f} ms) "
                  f"fmtcode={ss.ctx.pcg_format()}", flush=True)
        ss.ctx.close()
    if "--big" in sys.argv:
        # timing at a size the oracle does not solve: the two slab solves must agree
        n = 200
        shape = (n, n, n)
        Np, Nt = n ** 3, 3 * n * n * (n - 1)
        rb, re = bdist.cubic_slab_ranges(shape, world)[rank]
        ids, conns = bdist.cubic_slab_throats(shape, rb // (n * n), re // (n * n))
        g_all = 10.0 ** np.random.default_rng(0).uniform(-15, -12, Nt)
        bcv = np.full(Np, np.nan)
        bcv[:n * n] = 2e5
        bcv[Np - n * n:] = 1e5
        ss = bdist.SlabSolver(Np, rb, re, device=local, fmt="auto")
        res = {}
        for precond in ("jacobi", "amg"):
            ss.ctx.set_precond(precond)
            for rep in range(2):
                ss.assemble(conns, g_all[ids], bcv, None)
                if world > 1:
                    dist.barrier()
                ss.ctx.synchronize()
                t0 = time.perf_counter()
                x_loc, it, rel, info = ss.solve(rtol=1e-10)
                ss.ctx.synchronize()
                dt = time.perf_counter() - t0
            assert info == 0, (precond, info, it, rel)
            res[precond] = (x_loc.clone(), it, dt, ss.ctx.stats())
        dx = float((res["amg"][0] - res["jacobi"][0]).abs().max() / res["jacobi"][0].abs().max())
        assert dx < 1e-5, dx          # both are rtol-1e-10 solutions of a kappa ~ 1e5 system
        if rank == 0:
            sj, sa = res["jacobi"][3], res["amg"][3]
            print(f"dist amg big Cubic {n}^3 world={world}: jacobi {res['jacobi'][1]} it "
                  f"{res['jacobi'][2]*1e3:.1f} ms | amg {res['amg'][1]} it {res['amg'][2]*1e3:.1f} ms "
                  f"(setup {sa['t_setup_ms']:.1f} + solve {sa['t_solve_ms']:.1f}) | max rel diff "
                  f"{dx:.1e}", flush=True)
        ss.ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except BaseException:
        import traceback
        msg = traceback.format_exc().strip().splitlines()
        print(f"WORKER-ERR rank={os.environ.get('RANK')}: " + " | ".join(msg[-6:]), flush=True)
        os._exit(1)
