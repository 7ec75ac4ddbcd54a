"""GPU: slab-partitioned PCG over 2 (or more) B200s of one box, through NCCL."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


def test_single_rank_worker():
    """the slab driver with world_size 1 (no NCCL) on one GPU"""
    env = dict(os.environ, WORLD_SIZE="1", RANK="0", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dist_worker.py")],
                       env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "dist ok" in r.stdout


@pytest.mark.parametrize("world", [2, 4, 8])
def test_slab_solve_multi_gpu(world):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port",
           str(29600 + world), os.path.join(ROOT, "tests", "dist_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    errs = [ln for ln in r.stdout.splitlines() if "WORKER-ERR" in ln]
    assert r.returncode == 0, "\n".join(errs[:4]) + r.stderr[-1500:]
    assert r.stdout.count("dist ok") >= 11


def test_multigrid_single_rank_worker():
    env = dict(os.environ, WORLD_SIZE="1", RANK="0", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dist_amg_worker.py")],
                       env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("dist amg ok") == 3


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multigrid_slab_solve_multi_gpu(world):
    """precond='amg' on slabs: one row-partitioned multigrid hierarchy for the whole problem"""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port",
           str(29700 + world), os.path.join(ROOT, "tests", "dist_amg_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    errs = [ln for ln in r.stdout.splitlines() if "WORKER-ERR" in ln]
    assert r.returncode == 0, "\n".join(errs[:4]) + r.stderr[-1500:]
    assert r.stdout.count("dist amg ok") == 3
