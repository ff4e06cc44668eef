"""Multi-GPU parity: N bricks over NCCL vs the single-domain oracle.  Needs >= 2 GPUs on the box; the
worker (tests/mgpu_worker.py) is launched with torch.distributed.run exactly like bench.py is."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _run(n, *extra, env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + (os.getpid() % 400)), str(ROOT / "tests" / "mgpu_worker.py"), *extra]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "mgpu parity OK" in r.stdout
    return r.stdout


@pytest.mark.parametrize("rebuild_every", [10, 0])
def test_two_gpu_spheres(rebuild_every):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, "--mode", "sphere", "--steps", "60", "--rebuild-every", str(rebuild_every))


@pytest.mark.parametrize("env", [{"EDB_NO_FAN": "1"}, {"EDB_NO_P2P": "1"}], ids=["push-kernel", "nccl"])
def test_two_gpu_spheres_other_ghost_paths(env):
    """The per-step ghost refresh has three implementations: stores fused into the gather (default), a separate push kernel over
    peer memory (EDB_NO_FAN), and NCCL send/recv + all-reduce (EDB_NO_P2P, also the path when CUDA IPC is unavailable)."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, "--mode", "sphere", "--steps", "60", "--rebuild-every", "0", env=env)


def test_two_gpu_ghost_slots_current_after_run_verlet(tmp_path):
    """Inside edb_run_verlet the remote ghosts are read in the landing zone and the ghost slots are only brought home at the end of the
    call.  The ContactState view taken right after the call recomputes contact points from the ghost SLOTS: it must be bit-identical to
    the one of a run that copies the landing zone into the slots at every step (EDB_LAND_COPY=1), and operator-by-operator steps that
    follow must match the single-domain oracle."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    import numpy as np
    a, b = str(tmp_path / "inplace"), str(tmp_path / "copy")
    _run(2, "--mode", "sphere", "--steps", "25", "--rebuild-every", "0", "--then-ops", "3", "--dump", a)
    _run(2, "--mode", "sphere", "--steps", "25", "--rebuild-every", "0", "--then-ops", "3", "--dump", b, env={"EDB_LAND_COPY": "1"})
    for r in range(2):
        x, y = np.load(f"{a}.rank{r}.npz"), np.load(f"{b}.rank{r}.npz")
        for k in ("dn", "cp", "fn", "ft"):
            assert x[k].shape == y[k].shape and len(x[k]) > 100 and np.array_equal(x[k], y[k]), (r, k)


def test_two_gpu_drum_slabs():
    """The N>1 bench workload (drum cut into slabs along its periodic axis, cylinder driver on every rank), parity vs the oracle."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, "--mode", "drum", "--steps", "60", "--rebuild-every", "0")


@pytest.mark.parametrize("mode", ["sphere", "poly"])
def test_two_gpu_cost_balanced_cuts(mode):
    """Bricks cut where the dem_cost_model field balances (edb_set_brick_cuts) instead of uniformly: a box with a heap on one side, the
    two bricks hold different cell grids (3 + 5 cells along x), particles migrate across the cut, polyhedron twins cross it — same parity
    bar against the oracle, whose virtual bricks are cut at the same cells."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, "--mode", mode, "--steps", "60", "--rebuild-every", "10", "--balanced")


@pytest.mark.parametrize("env", [{}, {"EDB_DRIVERS_UNFUSED": "1"}], ids=["fused-loop", "operator-loop"])
def test_two_gpu_moving_driver(env):
    """A stress-controlled floor under a two-brick box: the reaction of the contacts is all-reduced (NCCL) before force_to_accel_driver,
    every rank integrates the same driver; drivers, particles and interactions against the single-domain oracle."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, "--mode", "sphere", "--steps", "60", "--rebuild-every", "0", "--moving", env=env)


@pytest.mark.parametrize("rebuild_every", [10, 0])
def test_two_gpu_polyhedra(rebuild_every):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, "--mode", "poly", "--steps", "30", "--rebuild-every", str(rebuild_every))


def test_eight_gpu_spheres():
    if _ngpu() < 8:
        pytest.skip("needs 8 GPUs")
    _run(8, "--mode", "sphere", "--steps", "40", "--rebuild-every", "10")
