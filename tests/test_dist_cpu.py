"""CPU (gloo, world_size 2 and 3): slab partition plan + halo-window algorithm."""
import os
import subprocess
import sys

import numpy as np
import pytest

from openpnm_b200 import dist as bdist
from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_ranges():
    assert bdist.slab_ranges(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert bdist.cubic_slab_ranges((8, 5, 6), 8) == [(30 * i, 30 * (i + 1)) for i in range(8)]
    r = bdist.cubic_slab_ranges((13, 4, 7), 4)
    assert r[0] == (0, 4 * 28) and r[-1][1] == 13 * 28 and all((b % 28) == 0 for b, _ in r)
    assert bdist.cubic_slab_ranges((2, 3, 3), 4) == bdist.slab_ranges(18, 4)   # fewer planes than ranks


@pytest.mark.parametrize("shape", [(6, 3, 4), (5, 1, 7), (4, 6, 1), (7, 1, 1)])
def test_cubic_slab_throats_cover_the_network(shape):
    net = oracle.cubic_network(shape)
    seen = np.zeros(net["Nt"], dtype=int)
    for x0, x1 in [(0, 2), (2, 3), (3, shape[0])]:
        if x0 >= shape[0]:
            continue
        x1 = min(x1, shape[0])
        ids, conns = bdist.cubic_slab_throats(shape, x0, x1)
        assert np.array_equal(conns, net["conns"][ids])
        plane = shape[1] * shape[2]
        mask = bdist.slab_throat_mask(net["conns"], x0 * plane, x1 * plane)
        assert np.array_equal(np.sort(ids), np.flatnonzero(mask))
        seen[ids] += 1
    assert seen.min() >= 1 and seen.max() <= 2        # cut throats belong to both sides


def test_coordinate_order_narrows_bandwidth():
    pts = np.random.default_rng(0).random((300, 3))
    c = oracle.delaunay_network(pts)
    perm, inv = bdist.coordinate_order(pts)
    assert np.array_equal(inv[perm], np.arange(300))
    assert np.all(np.diff(pts[perm][:, 0]) >= 0)
    assert bdist.bandwidth(np.sort(inv[c], axis=1)) < bdist.bandwidth(c)


@pytest.mark.parametrize("world,case", [(2, "cubic"), (3, "cubic"), (2, "delaunay")])
def test_gloo_slab_pcg(world, case):
    env = dict(os.environ, B200_CASE=case, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port",
           str(29700 + world + (10 if case != "cubic" else 0)),
           os.path.join(ROOT, "tests", "dist_cpu_worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "cpu dist ok" in r.stdout
