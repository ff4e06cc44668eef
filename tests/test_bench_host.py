"""Host-side pieces of bench.py and of the scaling workloads that run without a GPU: the CPU arm (`--impl reference`) on the workload
of N GPUs as one domain, the strong-scaling brick dimensions, the global-workload builder (same particles as the N per-rank
generators), and the oracle run on that domain against the oracle run brick by brick through the virtual-brick emulation."""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from exadem_b200 import parallel

ROOT = Path(__file__).resolve().parent.parent


def test_strong_brick_dims_divide_the_fixed_box():
    assert parallel.strong_brick_dims(200, 1) == (200, 200, 200)
    assert parallel.strong_brick_dims(200, 2) == (100, 200, 200)
    assert parallel.strong_brick_dims(200, 4) == (100, 100, 200)
    assert parallel.strong_brick_dims(200, 8) == (100, 100, 100)
    with pytest.raises(ValueError):
        parallel.strong_brick_dims(25, 2)


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_global_workload_is_the_union_of_the_rank_workloads(nranks):
    make = lambda r, n: parallel.brick_workload((10, 10, 10), r, n)
    wl = parallel.global_workload(make, nranks)
    pg = parallel.process_grid(nranks)
    assert wl.particles.n == 1000 * nranks and len(np.unique(wl.particles.id)) == wl.particles.n
    assert wl.lo == (0.0, 0.0, 0.0) and wl.hi == tuple(10.0 * pg[d] for d in range(3))
    owner = parallel.owner_rank(wl.particles.vec("r"), wl.lo, wl.hi, pg)
    for r in range(nranks):
        part = make(r, nranks).particles
        mine = owner == r
        assert mine.sum() == part.n and np.array_equal(np.sort(wl.particles.id[mine]), np.sort(part.id))
    # drum slabs: N slabs make one drum N times longer
    mk = lambda r, n: parallel.drum_slab_workload(3000, r, n)
    wd = parallel.global_workload(mk, nranks)
    one = mk(0, nranks)
    assert wd.particles.n == nranks * one.particles.n and abs((wd.hi[1] - wd.lo[1]) - nranks * (one.hi[1] - one.lo[1])) < 1e-12


def run_bench(*args):
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def test_reference_arm_line_on_one_and_on_two_gpus_worth_of_work():
    keys = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}
    a = run_bench("--impl", "reference", "--particles", "8000", "--steps", "4", "--warmup", "3")
    b = run_bench("--impl", "reference", "--gpus", "2", "--particles", "8000", "--steps", "4", "--warmup", "3")
    for d, n in ((a, 1), (b, 2)):
        assert keys <= set(d) and d["impl"] == "reference" and d["n_gpus"] == n and d["value"] > 0 and d["gpu_launches"] == 0
        assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["e2e"]["h2d_bytes_per_step"] == 0
        assert d["run"]["threads_in_parallel_region"] == d["run"]["threads_requested"]     # the launcher's OMP / affinity settings are undone
    assert b["run"]["particles"] == 2 * a["run"]["particles"] and "2 slabs" in b["config"]["workload"]
    c = run_bench("--impl", "reference", "--gpus", "8", "--workload", "box", "--scaling", "strong", "--global-side", "20", "--steps", "3", "--warmup", "3")
    assert c["scaling"] == "strong" and c["run"]["particles"] == 8000 and "2x2x2 bricks of 10x10x10" in c["config"]["workload"]


def test_reference_arm_is_immune_to_a_pinned_launcher():
    """torchrun exports OMP_NUM_THREADS=1 and a launcher may pin the rank to one core: the CPU arm must still use every core"""
    import os
    if (os.cpu_count() or 1) < 2:
        pytest.skip("one core")
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(["taskset", "-c", "0", sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--particles", "6000", "--steps", "3",
                        "--warmup", "3"], capture_output=True, text=True, timeout=600, env=env)
    if r.returncode != 0 and "taskset" in (r.stderr or ""):
        pytest.skip("taskset unavailable")
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    assert d["run"]["threads_in_parallel_region"] == (os.cpu_count() or 1), d["run"]


@pytest.mark.gpu
def test_gpu_arm_line_carries_the_whole_contract():
    """`python bench.py` (N = 1, small workload): every key the driver reads is there and means something"""
    d = run_bench("--steps", "6", "--warmup", "3", "--particles", "30000", "--no-poly", "--cpu-seconds", "1")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["metric"] == "interaction-updates/s" and d["n_gpus"] == 1 and d["steps"] == 6 and d["higher_is_better"] is True
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None and d["scaling"] == "weak"
    # value = records processed by the contact kernels in the timed region / its duration
    assert d["value"] > 0 and abs(d["value"] * d["ms_per_step"] * 1e-3 / d["run"]["interactions_per_gpu"] - 1.0) < 0.02
    assert "workload" in d["config"] and "l2" in d["config"]
    r = d["roofline"]
    for k in ("kernel", "bound", "achieved", "peak", "unit", "frac", "traffic", "mean_launch_ms", "share_of_step", "per_step_ms"):
        assert k in r, k
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and 0 < r["frac"] < 1.5 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    other = d.get("roofline_contact") or d.get("roofline_gather")
    assert other and other["kernel"] != r["kernel"]
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and "sample" in c and c["unit"] == d["unit"]
    e = d["e2e"]
    n = d["run"]["particles_per_gpu"]
    assert e["h2d_bytes_per_step"] == 13 * 8 * n and e["d2h_bytes_per_step"] == 22 * 8 * n and 0 < e["value"] < d["value"]
    assert d["gpu_launches"] >= 3 * 6      # at least contact + gather + flag per step, counted inside the timed region
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(d["clocks"])
