"""Network health and trimming on the device.

Mirrors, for the part the transport path needs before it can assemble:
`openpnm.utils.check_network_health` (openpnm/utils/_health.py:40-98; the models
openpnm/models/network/_health.py:11-76 behind it) and `openpnm.topotools.trim`
(openpnm/topotools/_topotools.py:157-229).  All arithmetic happens in
libb200pnm.so (csrc/topology.cu): lock-free union-find for the clusters, integer
histograms for sizes and degrees, flag -> scan -> scatter compaction for the trim.
"""
import numpy as np

from .device import Context

__all__ = ["check_network_health", "trim", "trim_to_largest_cluster"]


def _ctx(ctx):
    return ctx if ctx is not None else Context()


def check_network_health(network, ctx=None):
    """dict with the reference's keys that the path depends on: 'isolated_pores',
    'disconnected_pores' (pores outside the largest cluster), 'looped_throats'
    -- index lists like the reference's HealthDict -- plus the counts.
    Duplicate / bidirectional throats are not searched for: the assembly (K1) sums
    duplicated throats exactly as SciPy's `tocsr` does, so they do not harm the solve."""
    ctx = _ctx(ctx)
    h = ctx.network_health(network["throat.conns"], network.Np)
    mx = h["largest_cluster"]
    out = {
        "isolated_pores": np.flatnonzero(h["isolated"].cpu().numpy()).tolist(),
        "disconnected_pores": np.flatnonzero(h["cluster_size"].cpu().numpy() < mx).tolist(),
        "looped_throats": np.flatnonzero(h["looped"].cpu().numpy()).tolist(),
        "n_clusters": h["n_clusters"],
        "largest_cluster": mx,
    }
    return out


def trim(network, pores=(), throats=(), ctx=None):
    """Remove pores (with their throats) and/or throats from one of this package's
    `Network` stand-ins, in place: every 'pore.*' / 'throat.*' array is compacted and
    'throat.conns' renumbered, as topotools.trim does for a real OpenPNM project."""
    ctx = _ctx(ctx)
    Np, Nt = network.Np, network.Nt
    pk = np.ones(Np, dtype=np.uint8)
    pk[np.asarray(pores, dtype=np.int64)] = 0
    tk = None
    if np.size(throats):
        tk = np.ones(Nt, dtype=np.uint8)
        tk[np.asarray(throats, dtype=np.int64)] = 0
    conns_new, pmap, tmap, Np_new = ctx.trim(network["throat.conns"], Np, pk, tk)
    Ps = np.flatnonzero(pmap.cpu().numpy() >= 0)
    Ts = np.flatnonzero(tmap.cpu().numpy() >= 0)
    for key in list(network.keys()):
        arr = dict.__getitem__(network, key)
        if key == "throat.conns":
            continue
        dict.__setitem__(network, key, arr[Ts] if key.startswith("throat.") else arr[Ps])
    dict.__setitem__(network, "throat.conns", conns_new.cpu().numpy())
    network._Np, network._Nt = int(Np_new), int(Ts.size)
    return Ps, Ts


def trim_to_largest_cluster(ctx, conns_dev, Np):
    """Device-resident form for large networks (config 4): drop isolated pores and
    every cluster but the largest (the reference's suggested fix,
    openpnm/utils/_health.py:60-64,85-88).  Returns (conns_new device (Nt',2) int64,
    pore_map device int64 (new index or -1), throat_map, Np_new, health counts)."""
    import torch
    h = ctx.network_health(conns_dev, Np)
    with torch.cuda.stream(ctx.stream):
        keep = (h["cluster_size"] == h["largest_cluster"]).to(torch.uint8)
    conns_new, pmap, tmap, Np_new = ctx.trim(conns_dev, Np, keep)
    counts = {k: h[k] for k in ("n_clusters", "n_isolated", "n_disconnected", "n_looped",
                                "largest_cluster")}
    return conns_new, pmap, tmap, Np_new, counts
