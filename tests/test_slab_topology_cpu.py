"""CPU: the host half of the slab topology check (openpnm_b200/dist.py:
join_slab_clusters) against the global answer.  The per-rank device part
(b200_slab_clusters, csrc/topology.cu) is emulated here with scipy's
connected_components on exactly the throats a rank holds; the GPU test
tests/dist_worker.py runs the real thing on 2/4/8 ranks."""
import numpy as np
import pytest
import scipy.sparse as sprs
from scipy.sparse import csgraph

from openpnm_b200 import dist as bdist


def emulate_rank(conns, Np, rb, re, halo, bc):
    c = conns[bdist.slab_throat_mask(conns, rb, re)]
    am = sprs.coo_matrix((np.ones(len(c), dtype=np.int8), (c[:, 0], c[:, 1])), shape=(Np, Np))
    ncomp, comp = csgraph.connected_components(am, directed=False)
    root = np.full(ncomp, Np, dtype=np.int64)
    np.minimum.at(root, comp, np.arange(Np))
    labels = root[comp]
    has_bc = np.zeros(Np, dtype=bool)
    has_bc[labels[bc]] = True
    clip = lambda v: min(max(v, 0), Np)
    w0a, w0b, w1a, w1b = clip(rb - halo), clip(rb + halo), clip(re - halo), clip(re + halo)
    owned = np.zeros(Np, dtype=bool)
    window = np.zeros(Np, dtype=bool)
    owned[labels[rb:re]] = True
    window[labels[w0a:w0b]] = True
    window[labels[w1a:w1b]] = True
    roots = np.flatnonzero(labels == np.arange(Np))
    interior_bad = int(np.sum(owned[roots] & ~window[roots] & ~has_bc[roots]))
    lab = np.full(4 * halo, -1, dtype=np.int32)
    hb = np.zeros(4 * halo, dtype=np.uint8)
    lab[w0a - (rb - halo):w0b - (rb - halo)] = labels[w0a:w0b]
    hb[w0a - (rb - halo):w0b - (rb - halo)] = has_bc[labels[w0a:w0b]]
    lab[2 * halo + w1a - (re - halo):2 * halo + w1b - (re - halo)] = labels[w1a:w1b]
    hb[2 * halo + w1a - (re - halo):2 * halo + w1b - (re - halo)] = has_bc[labels[w1a:w1b]]
    return lab, hb, interior_bad


def global_bad(conns, Np, bc):
    am = sprs.coo_matrix((np.ones(len(conns), dtype=np.int8), (conns[:, 0], conns[:, 1])),
                         shape=(Np, Np))
    ncomp, comp = csgraph.connected_components(am, directed=False)
    touched = np.zeros(ncomp, dtype=bool)
    touched[comp[bc]] = True
    return int((~touched).sum())


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("world", [1, 2, 3, 5])
def test_joined_slab_clusters_match_the_global_count(seed, world):
    rng = np.random.default_rng(seed)
    Np, halo = 600, 17
    a = rng.integers(0, Np, 900)
    b = a + rng.integers(1, halo + 1, a.size)
    ok = b < Np
    conns = np.stack([a[ok], b[ok]], axis=1)
    conns = conns[rng.random(len(conns)) < (0.35 + 0.1 * seed)]     # fragmented on purpose
    bc = np.zeros(Np, dtype=bool)
    bc[rng.integers(0, Np, 12 + 10 * seed)] = True
    ranges = bdist.slab_ranges(Np, world)
    parts = [emulate_rank(conns, Np, rb, re, halo, bc) for rb, re in ranges]
    assert bdist.join_slab_clusters(parts, Np) == global_bad(conns, Np, bc)


def test_connected_lattice_passes_and_a_cut_off_block_fails():
    from oracle import oracle
    net = oracle.cubic_network((8, 4, 3))
    Np, conns = net["Np"], net["conns"]
    bc = np.zeros(Np, dtype=bool)
    bc[net["labels"]["left"]] = True
    for world in (1, 2, 4, 8):
        ranges = bdist.cubic_slab_ranges((8, 4, 3), world)
        parts = [emulate_rank(conns, Np, rb, re, 12, bc) for rb, re in ranges]
        assert bdist.join_slab_clusters(parts, Np) == 0
        # cut the lattice between planes 4 and 5: the right block has no BC pore
        keep = ~((conns[:, 0] // 12 == 4) & (conns[:, 1] // 12 == 5))
        parts = [emulate_rank(conns[keep], Np, rb, re, 12, bc) for rb, re in ranges]
        assert bdist.join_slab_clusters(parts, Np) == 1
