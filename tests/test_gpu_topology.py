"""GPU: network health and trimming on the device (csrc/topology.cu)
against the oracle's restatement of check_network_health / topotools.trim, and, when
the reference is importable, against OpenPNM itself."""
import os
import sys

import numpy as np
import pytest

from oracle import oracle

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


def broken_delaunay(n=3000, cutoff=0.09, seed=2):
    """Delaunay network with its long (hull) throats cut: leaves isolated pores and
    small clusters, like a trimmed OpenPNM Delaunay network."""
    pts = np.random.default_rng(seed).random((n, 3))
    pts = pts[np.argsort(pts[:, 0], kind="stable")]
    conns = oracle.delaunay_network(pts)
    L = np.linalg.norm(pts[conns[:, 0]] - pts[conns[:, 1]], axis=1)
    return pts, conns[L < cutoff]


def test_cluster_labels_and_health_match_scipy(b2):
    pts, conns = broken_delaunay()
    Np = pts.shape[0]
    ctx = b2.B200PCG().ctx
    ref = oracle.network_health(conns, Np)
    assert ref["n_clusters"] > 3 and ref["isolated"].sum() > 0      # the case bites
    lab, nc = ctx.cluster_labels(conns, Np)
    lab = lab.cpu().numpy()
    assert nc == ref["n_clusters"]
    # same partition; the device label is the smallest pore index of the cluster
    root = np.full(ref["n_clusters"], Np)
    np.minimum.at(root, ref["cluster_num"], np.arange(Np))
    assert np.array_equal(lab, root[ref["cluster_num"]])
    h = ctx.network_health(conns, Np)
    assert h["n_clusters"] == ref["n_clusters"]
    assert np.array_equal(h["cluster_size"].cpu().numpy(), ref["cluster_size"])
    assert np.array_equal(h["isolated"].cpu().numpy().astype(bool), ref["isolated"])
    assert h["n_isolated"] == int(ref["isolated"].sum())
    assert h["n_disconnected"] == ref["disconnected"].size
    assert h["largest_cluster"] == int(ref["cluster_size"].max()) and h["n_looped"] == 0


def test_trim_matches_reference_semantics(b2):
    pts, conns = broken_delaunay(seed=3)
    Np = pts.shape[0]
    ctx = b2.B200PCG().ctx
    rng = np.random.default_rng(0)
    pores = rng.choice(Np, 200, replace=False)
    throats = rng.choice(conns.shape[0], 300, replace=False)
    new_ref, Pk, Tk = oracle.trim(conns, Np, pores, throats)
    pk = np.ones(Np, dtype=np.uint8)
    pk[pores] = 0
    tk = np.ones(conns.shape[0], dtype=np.uint8)
    tk[throats] = 0
    new, pmap, tmap, Np_new = ctx.trim(conns, Np, pk, tk)
    assert Np_new == Pk.size
    assert np.array_equal(new.cpu().numpy(), new_ref)
    assert np.array_equal(np.flatnonzero(pmap.cpu().numpy() >= 0), Pk)
    assert np.array_equal(np.flatnonzero(tmap.cpu().numpy() >= 0), Tk)
    with pytest.raises(Exception, match="Cannot delete ALL pores"):
        ctx.trim(conns, Np, np.zeros(Np, dtype=np.uint8))


def test_health_then_trim_then_solve(b2):
    """config 4 in small: cut Delaunay -> drop isolated pores and minor clusters on the
    device -> assemble -> solve; x against the direct solve of the trimmed network."""
    from openpnm_b200 import topotools
    pts, conns = broken_delaunay(n=5000, cutoff=0.07, seed=5)
    Np = pts.shape[0]
    pn = b2.Network(coords=pts, conns=conns)
    health = topotools.check_network_health(pn)
    ref = oracle.network_health(conns, Np)
    assert health["isolated_pores"] == np.flatnonzero(ref["isolated"]).tolist()
    assert health["disconnected_pores"] == ref["disconnected"].tolist()
    # a clustered network is refused, exactly as the reference does
    g0 = np.ones(pn.Nt)
    ph = b2.Phase(pn)
    ph["throat.diffusive_conductance"] = g0
    fd = b2.FickianDiffusion(network=pn, phase=ph)
    bcv = np.full(Np, np.nan)
    bcv[pts[:, 0] < 0.1] = 1.0
    bcv[pts[:, 0] > 0.9] = 0.0
    fd["pore.bc.value"] = bcv
    with pytest.raises(Exception, match="clustered"):
        fd.run(solver=b2.B200PCG(tol=1e-12))
    Ps, Ts = topotools.trim(pn, pores=health["disconnected_pores"])
    new_ref, Pk, Tk = oracle.trim(conns, Np, ref["disconnected"])
    assert np.array_equal(Ps, Pk) and np.array_equal(Ts, Tk)
    assert np.array_equal(pn.conns, new_ref) and pn.Np == Pk.size and pn.Nt == Tk.size
    assert np.array_equal(pn["pore.coords"], pts[Pk])
    assert topotools.check_network_health(pn)["disconnected_pores"] == []
    g = 10.0 ** np.random.default_rng(1).uniform(-15, -12, pn.Nt)
    ph = b2.Phase(pn)
    ph["throat.diffusive_conductance"] = g
    fd = b2.FickianDiffusion(network=pn, phase=ph)
    bcv = np.full(pn.Np, np.nan)
    bcv[pn["pore.coords"][:, 0] < 0.1] = 1.0
    bcv[pn["pore.coords"][:, 0] > 0.9] = 0.0
    fd["pore.bc.value"] = bcv
    fd.run(solver=b2.B200PCG(tol=1e-13))
    refx = oracle.transport_run(pn.conns, g, pn.Np, bc_value=bcv)["x"]
    assert np.max(np.abs(fd.x - refx)) <= 1e-8


def test_against_the_reference_itself(b2):
    import ref_shim
    if ref_shim.find_reference() is None:
        pytest.skip("reference OpenPNM not available")
    op = ref_shim.import_reference()
    from openpnm_b200 import topotools
    pts, conns = broken_delaunay(n=1500, cutoff=0.11, seed=8)
    pn = op.network.Network(coords=pts, conns=conns)
    h_ref = op.utils.check_network_health(pn)
    mine = b2.Network(coords=pts, conns=conns)
    h = topotools.check_network_health(mine)
    for key in ("isolated_pores", "disconnected_pores", "looped_throats"):
        assert h[key] == list(h_ref[key]), key
    op.topotools.trim(network=pn, pores=h_ref["disconnected_pores"])
    topotools.trim(mine, pores=h["disconnected_pores"])
    assert np.array_equal(mine.conns, pn["throat.conns"])
    assert np.allclose(mine["pore.coords"], pn["pore.coords"])
    op.Workspace().clear()
