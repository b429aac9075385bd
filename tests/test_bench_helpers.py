"""Host-side pieces of bench.py that decide whether a run is accepted: the per-cell hash of the N > 1 parity gate, the shared
workload description of the two arms.  CPU only."""
import importlib.util
import os
import types

import numpy as np
import pytest

import harness

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


def test_cell_hash_sees_values_order_and_membership():
    rng = np.random.default_rng(1)
    counts = np.array([3, 0, 2, 4], dtype=np.uint64)
    aos = rng.normal(size=(9, 8))
    h = bench.cell_hashes(np, counts, aos)
    assert h.shape == (4,) and h[1] == 0
    assert np.array_equal(h, bench.cell_hashes(np, counts, aos.copy()))
    flipped = aos.copy()
    flipped[4, 2] = np.nextafter(flipped[4, 2], np.inf)                 # one ulp in one particle of cell 2
    assert (bench.cell_hashes(np, counts, flipped) != h).tolist() == [False, False, True, False]
    swapped = aos.copy()
    swapped[[5, 6]] = swapped[[6, 5]]                                   # two particles of cell 3 trade places
    assert (bench.cell_hashes(np, counts, swapped) != h).tolist() == [False, False, False, True]
    moved = np.array([2, 1, 2, 4], dtype=np.uint64)                     # a particle of cell 0 counted in cell 1
    assert (bench.cell_hashes(np, moved, aos) != h).tolist()[:2] == [True, True]
    signed = aos.copy()
    signed[0, 3] = 0.0
    zero = bench.cell_hashes(np, counts, signed)
    signed[0, 3] = -0.0
    assert bench.cell_hashes(np, counts, signed)[0] != zero[0]          # bit patterns, not values


def test_both_arms_describe_the_same_workload():
    for workload in ("uniform", "medium", "large", "dipole"):
        args = types.SimpleNamespace(workload=workload, grid=32, particles_per_gpu=32_000_000)
        for world in (1, 2, 8):
            cfg = bench.workload_config(args, world)
            assert set(cfg) == {"workload", "particles", "particles_per_gpu", "grid", "dt"}
            assert cfg["particles_per_gpu"] * world <= cfg["particles"] < (cfg["particles_per_gpu"] + 1) * world
    assert bench.workload_config(types.SimpleNamespace(workload="uniform", grid=32, particles_per_gpu=32_000_000), 8)["particles"] == 256_000_000


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's own CPU loop on the box's host threads) on a small universe: one JSON line
    with the contract's keys, the b200 arm's `config`, an e2e that repeats the value with zero bytes copied, and a cpu_baseline
    describing this very run."""
    import json, subprocess, sys
    if not harness.Ref.available(True):
        pytest.skip("oracle/_ref not built")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--particles-per-gpu", "200000", "--steps", "3", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-1000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 3 and line["warmup"] == 1
    assert line["unit"] == "particle-pushes/s" and line["dtype"] == "f64" and line["vs_baseline"] is None and line["value"] > 0
    assert line["config"] == bench.workload_config(types.SimpleNamespace(workload="uniform", grid=32, particles_per_gpu=200000), 1)
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["value"] == line["value"] and line["cpu_baseline"]["cores"] >= 1
    assert line["with_deposit"]["value"] > 0
