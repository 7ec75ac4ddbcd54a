"""Iteration-count (performance) bounds of the multigrid preconditioner.

Kept apart from the parity files and collected LAST (the file name sorts after every
other test file) so that a convergence-rate regression can never stop `pytest -x`
before the golden-vector parity tests have run (round-1 verdict).  Every bound is
relative to the Jacobi-PCG count on the same system, not a magic constant."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _lattice(b2, shape, seed):
    pn = b2.Cubic(list(shape), coords=False)
    g = 10.0 ** np.random.default_rng(seed).uniform(-15, -12, pn.Nt)
    bcv = np.full(pn.Np, np.nan)
    bcv[pn.pores("left")] = 2e5
    bcv[pn.pores("right")] = 1e5
    return pn, g, bcv


@pytest.mark.parametrize("n", [24, 48, 64])
def test_multigrid_needs_far_fewer_iterations_than_jacobi(b2, n):
    pn, g, bcv = _lattice(b2, (n, n, n), seed=1)
    its = {}
    for precond in ("jacobi", "amg"):
        ctx = b2.B200PCG(precond=precond).ctx
        ctx.build_laplacian(pn.conns, g, pn.Np)
        ctx.apply_bcs(bcv, None)
        x, it, rel, info = ctx.pcg(rtol=1e-10, maxiter=20000)
        assert info == 0
        its[precond] = it
    assert its["amg"] <= max(60, its["jacobi"] // 4), its


def test_multigrid_count_is_nearly_size_independent(b2):
    its = []
    for n in (32, 64, 96):
        pn, g, bcv = _lattice(b2, (n, n, n), seed=2)
        ctx = b2.B200PCG(precond="amg").ctx
        ctx.build_laplacian(pn.conns, g, pn.Np)
        ctx.apply_bcs(bcv, None)
        x, it, rel, info = ctx.pcg(rtol=1e-10)
        assert info == 0
        its.append(it)
    assert its[-1] <= 2 * its[0], its
