"""Host-side logic of the multi-GPU path on CPU: world_size-2 (and 4) gloo runs of the z-slab input decomposition
and of the rank aggregation bench.py performs; plus the rank protocol of `bench.py --impl reference`."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.mark.parametrize("world", [2, 4])
def test_slab_decomposition_over_gloo(world):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29530 + world), os.path.join(HERE, "gloo_slab_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert "gloo slab check ok: %d ranks" % world in out.stdout


def test_reference_arm_runs_on_rank0_only():
    """Under torchrun the reference arm is run and printed by rank 0 alone; other ranks exit 0 without work."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True,
                         timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_slab_of_matches_a_plain_loop():
    """capi.slab_of (what bench.py --workload dipole uses to cut a universe into z-slabs) against a cell-by-cell loop."""
    import numpy as np
    from allscale_ipic3d_b200 import capi
    rng = np.random.default_rng(5)
    nx = ny = 3
    nz = 4
    counts = rng.integers(0, 7, nx * ny * nz).astype(np.uint64)
    counts[5] = 0
    aos = rng.normal(size=(int(counts.sum()), 8))
    first = np.concatenate([[0], np.cumsum(counts)]).astype(int)
    for world in (1, 2, 4):
        lz = nz // world
        seen = 0
        for rank in range(world):
            c, a = capi.slab_of(counts, aos, nx * ny, nz, rank, world)
            want_c, want_a = [], []
            for xy in range(nx * ny):
                for zl in range(lz):
                    cell = xy * nz + rank * lz + zl
                    want_c.append(counts[cell]); want_a.append(aos[first[cell]:first[cell + 1]])
            assert np.array_equal(c, np.array(want_c)) and np.array_equal(a, np.concatenate(want_a))
            seen += len(a)
        assert seen == len(aos)
