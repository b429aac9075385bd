"""Host-side pieces of bench.py that decide whether a run is accepted: the per-cell hash of the N > 1 parity gate, the shared
workload description of the two arms.  CPU only."""
import importlib.util
import os
import types

import numpy as np

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
