"""Worker for the multi-GPU multigrid test (torchrun, one rank per GPU): the slab
solve with `precond='amg'` -- flexible CG on the global operator preconditioned by ONE
multigrid hierarchy whose levels are row-partitioned by the slabs (csrc/amg.cu) --
against the CPU oracle and against the slab Jacobi-PCG.

Also checks the hierarchy itself, gathered from the ranks: every level is the Galerkin
product of the one above it for the exported aggregation (A_{l+1} = P^T A_l P, couplings
between ranks included), aggregates never cross a rank boundary, partitioned levels
stack to the global matrix, replicated levels are identical on all ranks.

`--big` adds a timing run at 200^3 (no oracle), `--huge N` one at N^3."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def gathered_levels(ctx, world):
    """[(A_l global scipy CSR, agg_l global or None, info of this rank)] from all ranks"""
    import scipy.sparse as sprs
    import torch.distributed as dist
    levels = ctx.amg_setup()
    out = []
    for l in range(len(levels)):
        info = ctx.amg_level_info(l)
        last = l + 1 == len(levels)
        ip, ix, dt, ag = ctx.amg_export_level(l, with_agg=not last)
        part = (info, ip.cpu().numpy(), ix.cpu().numpy(), dt.cpu().numpy(),
                None if ag is None else ag.cpu().numpy())
        parts = [part]
        if world > 1:
            parts = [None] * world
            dist.all_gather_object(parts, part)
        ng = info["n_global"]
        if info["partitioned"]:
            assert [p[0]["first_row"] for p in parts] == \
                list(np.cumsum([0] + [p[0]["n"] for p in parts[:-1]])), "rows are not contiguous"
            assert sum(p[0]["n"] for p in parts) == ng
            A = sprs.vstack([sprs.csr_matrix((p[3], p[2], p[1]), shape=(p[0]["n"], ng))
                             for p in parts]).tocsr()
            agg = None if last else np.concatenate([p[4] for p in parts])
            owner = np.concatenate([np.full(p[0]["n"], r) for r, p in enumerate(parts)])
        else:
            A = sprs.csr_matrix((part[3], part[2], part[1]), shape=(ng, ng))
            for p in parts:        # a replicated level is the same on every rank
                assert np.array_equal(p[1], part[1]) and np.array_equal(p[2], part[2])
                assert np.array_equal(p[3].view(np.uint64), part[3].view(np.uint64))
            agg = None if last else part[4]
            owner = None
        out.append((A, agg, info, owner))
    return out


def check_hierarchy(ctx, world, expect_partitioned_coarse=False):
    import scipy.sparse as sprs
    lv = gathered_levels(ctx, world)
    assert len(lv) >= 2 and lv[-1][0].shape[0] <= 2048 and not lv[-1][2]["partitioned"]
    if expect_partitioned_coarse:
        assert lv[1][2]["partitioned"], "level 1 should be row-partitioned in this case"
    for l in range(len(lv) - 1):
        A, agg, info, owner = lv[l]
        Ac = lv[l + 1][0].copy()
        n, nc = A.shape[0], Ac.shape[0]
        keep = agg >= 0
        assert agg.max() == nc - 1 and np.unique(agg[keep]).size == nc
        if owner is not None:
            # aggregates never cross a rank boundary; coarse ids follow the rank order
            lo = np.full(nc, world, dtype=np.int64)
            hi = np.full(nc, -1, dtype=np.int64)
            np.minimum.at(lo, agg[keep], owner[keep])
            np.maximum.at(hi, agg[keep], owner[keep])
            assert np.array_equal(lo, hi), "an aggregate spans two ranks"
            assert np.all(np.diff(lo) >= 0), "coarse numbering does not follow the rank order"
        P = sprs.csr_matrix((np.ones(keep.sum()), (np.flatnonzero(keep), agg[keep])), shape=(n, nc))
        G = (P.T @ A @ P).tocsr()
        G.sum_duplicates()
        G.sort_indices()
        Ac.sum_duplicates()
        Ac.sort_indices()
        assert np.array_equal(G.indptr, Ac.indptr) and np.array_equal(G.indices, Ac.indices), \
            f"level {l + 1}: structure differs from P^T A P"
        np.testing.assert_allclose(Ac.data, G.data, rtol=1e-11, atol=0)
        assert abs(Ac - Ac.T).max() <= 1e-12 * abs(Ac).max()
    return [(a.shape[0], bool(i["partitioned"])) for a, _, i, _ in lv]


def lattice_case(bdist, oracle, shape, rank, world, seed=3):
    nx, ny, nz = shape
    Np, Nt = nx * ny * nz, (nx - 1) * ny * nz + nx * (ny - 1) * nz + nx * ny * (nz - 1)
    rb, re = bdist.cubic_slab_ranges(shape, world)[rank]
    ids, conns = bdist.cubic_slab_throats(shape, rb // (ny * nz), re // (ny * nz))
    g_all = 10.0 ** np.random.default_rng(seed).uniform(-15, -12, Nt)
    bcv = np.full(Np, np.nan)
    bcv[:ny * nz] = 2.0
    bcv[Np - ny * nz:] = 1.0
    return Np, Nt, rb, re, ids, conns, g_all, bcv


def main():
    import torch
    import torch.distributed as dist
    from openpnm_b200 import dist as bdist
    from oracle import oracle

    rank, world, local = bdist.init_process_group()
    small = "--skip-small" not in sys.argv
    # ---- 1: against the direct solve; the coarse levels are replicated at this size
    for shape, fmt in ((((48, 24, 24), "auto"), ((48, 24, 24), "csr")) if small else ()):
        Np, Nt, rb, re, ids, conns, g_all, bcv = lattice_case(bdist, oracle, shape, rank, world)
        net = oracle.cubic_network(shape)
        assert np.array_equal(conns, net["conns"][ids])
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
        shape_of = check_hierarchy(ss.ctx, world)
        # one multigrid for the whole problem: the count does not depend on the number of
        # slabs (numpy model tests/models/amg_slabs.py: 38 / 38 / 38 / 39 on 1 / 2 / 4 / 8)
        assert out["amg"][0] < out["jacobi"][0] / 3, {k: v[0] for k, v in out.items()}
        if rank == 0:
            print(f"dist amg ok shape={shape} fmt={fmt} world={world} jacobi={out['jacobi'][0]} it "
                  f"amg={out['amg'][0]} it ({out['amg'][1]*1e3:.1f} ms) levels={shape_of} "
                  f"fmtcode={ss.ctx.pcg_format()}", flush=True)
        ss.ctx.close()

    # ---- 2: row-partitioned coarse levels (replication threshold forced down), against
    # the slab Jacobi-PCG at rtol 1e-13 and the explicitly recomputed residual
    os.environ["B200_AMG_REPL"] = "64"
    try:
        if not small:
            raise StopIteration
        shape = (64, 40, 36)
        Np, Nt, rb, re, ids, conns, g_all, bcv = lattice_case(bdist, oracle, shape, rank, world, 5)
        ss = bdist.SlabSolver(Np, rb, re, device=local, fmt="auto")
        res = {}
        for precond in ("jacobi", "amg"):
            ss.ctx.set_precond(precond)
            ss.assemble(conns, g_all[ids], bcv, None)
            x_loc, it, rel, info = ss.solve(rtol=1e-13)
            assert info == 0, (precond, info, it, rel)
            res[precond] = (ss.gather(x_loc).cpu().numpy(), it)
        dx = np.max(np.abs(res["amg"][0] - res["jacobi"][0])) / np.max(np.abs(res["jacobi"][0]))
        assert dx < 1e-8, dx
        assert res["amg"][1] < res["jacobi"][1] / 3, (res["amg"][1], res["jacobi"][1])
        shape_of = check_hierarchy(ss.ctx, world, expect_partitioned_coarse=world > 1)
        # bit-repeatable: the same solve again gives the same bits
        ss.assemble(conns, g_all[ids], bcv, None)
        x2, it2, _, _ = ss.solve(rtol=1e-13)
        assert it2 == res["amg"][1]
        assert np.array_equal(ss.gather(x2).cpu().numpy().view(np.uint64),
                              res["amg"][0].view(np.uint64)), "slab multigrid solve is not repeatable"
        if rank == 0:
            print(f"dist amg ok shape={shape} partitioned coarse levels world={world} "
                  f"jacobi={res['jacobi'][1]} it amg={res['amg'][1]} it levels={shape_of} "
                  f"max rel diff {dx:.1e}", flush=True)
        ss.ctx.close()
    except StopIteration:
        pass
    finally:
        del os.environ["B200_AMG_REPL"]

    sizes = []
    if "--big" in sys.argv:
        sizes.append(200)
    if "--huge" in sys.argv:
        sizes.append(int(sys.argv[sys.argv.index("--huge") + 1]))
    for n in sizes:
        # timing at a size the oracle does not solve: the two slab solves must agree
        shape = (n, n, n)
        Np, Nt = n ** 3, 3 * n * n * (n - 1)
        rb, re = bdist.cubic_slab_ranges(shape, world)[rank]
        ids, conns = bdist.cubic_slab_throats(shape, rb // (n * n), re // (n * n))
        g_all = 10.0 ** np.random.default_rng(0).uniform(-15, -12, Nt)
        g_loc = g_all[ids]
        del g_all, ids
        bcv = np.full(Np, np.nan)
        bcv[:n * n] = 2e5
        bcv[Np - n * n:] = 1e5
        ss = bdist.SlabSolver(Np, rb, re, device=local, fmt="auto")
        res = {}
        if "--p2p" in sys.argv and world > 1:
            ss.assemble(conns, g_loc, bcv, None)
            ss.enable_p2p()
        for precond in (("jacobi",) if "--jacobi-only" in sys.argv else ("jacobi", "amg")):
            ss.ctx.set_precond(precond)
            for rep in range(3):
                ss.assemble(conns, g_loc, bcv, None)
                if world > 1:
                    dist.barrier()
                ss.ctx.synchronize()
                t0 = time.perf_counter()
                x_loc, it, rel, info = ss.solve(rtol=1e-10)
                ss.ctx.synchronize()
                dt = time.perf_counter() - t0
            assert info == 0, (precond, info, it, rel)
            res[precond] = (x_loc.clone(), it, dt, ss.ctx.stats())
        if "--jacobi-only" in sys.argv:
            if rank == 0:
                sj = res["jacobi"][3]
                print(f"dist jacobi-only Cubic {n}^3 world={world} p2p={'--p2p' in sys.argv} "
                      f"order={os.environ.get('B200_P2P_ORDER', 'new')}: {res['jacobi'][1]} it "
                      f"{res['jacobi'][2]*1e3:.1f} ms wall, device solve {sj['t_solve_ms']:.1f} ms "
                      f"= {sj['t_solve_ms']/res['jacobi'][1]:.4f} ms/it", flush=True)
            ss.ctx.close()
            continue
        dx = float((res["amg"][0] - res["jacobi"][0]).abs().max() / res["jacobi"][0].abs().max())
        assert dx < 1e-5, dx          # both are rtol-1e-10 solutions of a kappa ~ 1e5 system
        if rank == 0:
            sj, sa = res["jacobi"][3], res["amg"][3]
            lv = [ss.ctx.amg_level_info(l) for l in range(len(ss.ctx.amg_setup()))]
            print(f"dist amg big Cubic {n}^3 world={world}: jacobi {res['jacobi'][1]} it "
                  f"{res['jacobi'][2]*1e3:.1f} ms | amg {res['amg'][1]} it {res['amg'][2]*1e3:.1f} ms "
                  f"(setup {sa['t_setup_ms']:.1f} + solve {sa['t_solve_ms']:.1f}) | max rel diff "
                  f"{dx:.1e} | levels "
                  + " ".join(f"{i['n_global']}{'p' if i['partitioned'] else 'r'}" for i in lv),
                  flush=True)
        else:
            ss.ctx.amg_setup()        # collective
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
        print(f"WORKER-ERR rank={os.environ.get('RANK')}: " + " | ".join(msg[-8:]), flush=True)
        os._exit(1)
