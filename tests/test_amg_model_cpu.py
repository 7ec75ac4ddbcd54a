"""CPU: the numpy model of the multigrid preconditioner (tests/models/amg_v2.py,
amg_slabs.py) -- the design evidence behind csrc/amg.cu, kept executable:

* ~30 outer iterations whatever the lattice size (Jacobi-PCG grows with N);
* the hierarchy must be built on the UNSCALED operator: a piecewise-constant P on
  the Jacobi-scaled one (near-null-space D^1/2 1, not 1) loses the coarse correction;
* aggregates that never cross a slab boundary cost nothing when the Galerkin
  operators stay global (the round-2 multi-GPU design), whereas block-wise
  hierarchies (additive Schwarz, the experimental B200_AMG_SLABS mode) do.

Nothing here is product code and nothing under openpnm_b200/ imports it."""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "models"))
import amg_slabs  # noqa: E402
import amg_v2  # noqa: E402
from oracle import oracle  # noqa: E402

PRM = dict(deg=2, ratio=8.0, klev=2, crtol=0.1, cmax=30)


def problem(shape, seed=3):
    net = oracle.cubic_network(shape)
    Np = net["Np"]
    g = 10.0 ** np.random.default_rng(seed).uniform(-15, -12, net["Nt"])
    bcv = np.full(Np, np.nan)
    bcv[net["labels"]["left"]] = 2.0
    bcv[net["labels"]["right"]] = 1.0
    A, b, _ = oracle.apply_bcs(oracle.build_A(net["conns"], g, Np), Np, bcv, None)
    return oracle.to_csr(A, Np), b


def solve(A, b, levels):
    x, it = amg_v2.fcg(A, b, lambda r: amg_v2.cycle(levels, 0, r, PRM), rtol=1e-10)
    assert np.linalg.norm(b - A @ x) <= 1e-10 * np.linalg.norm(b)
    return it


def test_iterations_do_not_grow_with_the_lattice():
    its = {}
    for n in (16, 24, 32):
        A, b = problem((n, n, n))
        its[n] = solve(A, b, amg_v2.setup(A, passes=3, passes0=4, coarsest=500))
    xj, _info, itj = oracle.jacobi_pcg(*problem((32, 32, 32)), rtol=1e-10, maxiter=5000)
    assert max(its.values()) <= 45 and its[32] <= its[16] + 8, its
    assert itj > 4 * its[32]


def test_hierarchy_must_be_built_on_the_unscaled_operator():
    A, b = problem((24, 24, 24))
    good = solve(A, b, amg_v2.setup(A, passes=3, passes0=4, coarsest=500))
    s = 1.0 / np.sqrt(A.diagonal())
    At = (sp.diags(s) @ A @ sp.diags(s)).tocsr()
    bad = solve(At, s * b, amg_v2.setup(At, passes=3, passes0=4, coarsest=500))
    assert bad > 1.5 * good, (good, bad)


def test_slab_respecting_aggregates_cost_nothing_but_blockwise_hierarchies_do():
    shape = (32, 12, 12)
    A, b = problem(shape)
    Np = A.shape[0]
    plane = shape[1] * shape[2]
    base = solve(A, b, amg_v2.setup(A, passes=3, passes0=4, coarsest=500))
    for world in (2, 4):
        owner = (np.arange(Np) // plane) * world // shape[0]
        it = solve(A, b, amg_slabs.setup_slabs(A, owner, coarsest=500))
        assert it <= base + 4, (world, base, it)
    # block-wise: every slab its own hierarchy, no coupling between them in M
    cuts = np.linspace(0, shape[0], 3).astype(int) * plane
    hier = [amg_v2.setup(A[a:c, a:c].tocsr(), passes=3, passes0=4, coarsest=500)
            for a, c in zip(cuts[:-1], cuts[1:])]

    def M(r):
        z = np.empty_like(r)
        for (a, c), lv in zip(zip(cuts[:-1], cuts[1:]), hier):
            z[a:c] = amg_v2.cycle(lv, 0, r[a:c], PRM)
        return z
    x, it_block = amg_v2.fcg(A, b, M, rtol=1e-10)
    assert it_block > 2 * base, (base, it_block)


def test_distributed_setup_needs_only_the_halo_of_the_aggregation_map():
    """The rank-by-rank construction of tests/models/amg_dist_model.py (owned-column
    matching, one integer all-gather for the numbering, an index-list exchange of the
    aggregation map at the slab interfaces, local Galerkin rows) reproduces the
    global slab-respecting hierarchy level by level -- and so inherits its
    iteration count."""
    import amg_dist_model
    shape = (32, 12, 12)
    A, b = problem(shape)
    Np = A.shape[0]
    plane = shape[1] * shape[2]
    for world in (2, 4):
        offsets = (np.linspace(0, shape[0], world + 1).astype(int) * plane)
        owner = np.searchsorted(offsets, np.arange(Np), side="right") - 1
        ref = amg_slabs.setup_slabs(A, owner, coarsest=500)
        levels, maps, traffic = amg_dist_model.build(A, offsets, coarsest=500)
        assert len(levels) == len(ref)
        for lv, L in zip(levels, ref):
            G = amg_dist_model.assemble_global(lv)
            assert G.shape == L.A.shape
            D = (G - L.A).tocsr()
            assert abs(D).max() <= 1e-12 * abs(L.A).max()
        # the exchanged map entries are interface-sized, not volume-sized
        assert max(max(t) for t in traffic[0]) <= 2 * plane
        # the composite map of the fine level is the prolongation the global model uses
        P = ref[0].P.tocsr()
        k = maps[0] >= 0
        assert np.array_equal(P.indices, maps[0][k]) and P.nnz == k.sum()


def test_box_aggregation_keeps_every_level_a_lattice_operator():
    """tests/models/amg_boxes.py (what DESIGN section 10 names as the next step for lattices):
    2x2x2 boxes of the index lattice as aggregates -> every Galerkin operator is a 7-point
    lattice operator (3 super-diagonals at (1, cz, cy*cz): the DIA-sym format of the fine
    level serves all levels), still symmetric with non-positive off-diagonals, and the
    outer iteration count is not worse than with pairwise matching on the benchmark's
    conductance field"""
    import amg_boxes
    shape = (24, 20, 18)
    A, b = problem(shape)
    levels = amg_boxes.setup(A, shape, coarsest=400)
    assert [L.shape for L in levels[:3]] == [(24, 20, 18), (12, 10, 9), (6, 5, 5)]
    for L in levels:
        cy, cz = L.shape[1], L.shape[2]
        assert set(amg_boxes.super_diagonals(L.A)) <= {1, cz, cy * cz}
        assert abs(L.A - L.A.T).max() <= 1e-12 * abs(L.A).max()
        off = L.A - sp.diags(L.A.diagonal())
        assert off.max() <= 0.0
    # coarse couplings = the conductance crossing the box face (structured reduction)
    L0, L1 = levels[0], levels[1]
    agg, cs = amg_boxes.box_aggregates(L0.A, shape)
    C = L0.A.tocoo()
    m = (agg[C.row] == 0) & (agg[C.col] == 1)
    assert np.isclose(L1.A[0, 1], C.data[m].sum(), rtol=1e-14)
    its_box = solve(A, b, levels)
    its_match = solve(A, b, amg_v2.setup(A, passes=3, passes0=4, coarsest=400))
    assert its_box <= its_match + 2 and its_box <= 40, (its_box, its_match)
