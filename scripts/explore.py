"""Exploration: iteration counts, solve times and per-kernel bandwidth."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import openpnm_b200 as b2
from openpnm_b200 import _lib


def run(n, fmt, rtol=1e-10):
    t0 = time.time()
    pn = b2.Cubic([n, n, n], spacing=1e-4, coords=False)
    g = 10.0 ** np.random.default_rng(0).uniform(-15, -12, pn.Nt)
    bcv = np.full(pn.Np, np.nan)
    bcv[pn.pores("left")] = 2e5
    bcv[pn.pores("right")] = 1e5
    t_host = time.time() - t0
    s = b2.B200PCG(tol=rtol, fmt=fmt)
    ctx = s.ctx
    conns_d = ctx.to_device(pn.conns)
    g_d = ctx.to_device(g)
    bcv_d = ctx.to_device(bcv)
    ctx.synchronize()
    ctx.set_format({"auto": 0, "csr": 1, "dia": 2}[fmt])
    for rep in range(2):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        with torch.cuda.stream(ctx.stream):
            e[0].record()
            nnz0 = ctx.build_laplacian(conns_d, g_d, pn.Np)
            e[1].record()
            f, nnz = ctx.apply_bcs(bcv_d, None)
            e[2].record()
            x, it, rel, info = ctx.pcg(rtol=rtol)
            e[3].record()
        torch.cuda.synchronize()
        st = ctx.stats()
    tb, tbc, ts = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3])
    Np = pn.Np
    model_iter = 12 * nnz + 4 * (Np + 1) + 16 * Np + 88 * Np
    model_spmv = 12 * nnz + 4 * (Np + 1) + 16 * Np
    print(f"n={n} fmt={fmt} Np={Np} nnz={nnz} host_gen={t_host:.1f}s build={tb:.1f}ms bc={tbc:.1f}ms "
          f"pcg={ts:.1f}ms (setup {st['t_setup_ms']:.1f} solve {st['t_solve_ms']:.1f}) iters={it} "
          f"info={info} relres={rel:.2e} ms/iter={st['t_solve_ms']/max(it,1):.4f} "
          f"DOFit/s={Np*it/(ts*1e-3):.3e} model GB/s iter={model_iter*it/(st['t_solve_ms']*1e-3)/1e9:.0f}")
    for which, name, bytes_model in ((0, "K_A spmv+dot", model_spmv), (1, "K_BC", 56 * Np),
                                     (2, "K_BC exact", 64 * Np), (3, "K3 csr spmv", model_spmv)):
        ms = ctx.time_kernel(which, 20)
        print(f"    {name:14s} {ms*1e3:9.1f} us   model {bytes_model/ (ms*1e-3)/1e9:8.0f} GB/s")
    rl = ctx.rate_pores(x, pn.pores("left")).sum().item()
    rr = ctx.rate_pores(x, pn.pores("right")).sum().item()
    print(f"    rate left {rl:.9e} right {rr:.9e} imbalance {abs(rl+rr)/abs(rl):.2e}")
    ctx.close()


if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:]] or [100, 200]
    for n in sizes:
        for fmt in ("dia", "csr"):
            run(n, fmt)
