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


def test_balanced_planes():
    """capi.balanced_planes: slabs cut by particle count -- every slab at least one plane, planes add up, and on the
    medium.inp deck (particles only in the inner 75 % box: planes 8..55 of 64) the eight slabs hold equal shares."""
    import numpy as np
    from allscale_ipic3d_b200 import capi
    assert capi.balanced_planes(np.ones(32), 8) == [4] * 8
    assert capi.balanced_planes(np.ones(8), 8) == [1] * 8
    rng = np.random.default_rng(3)
    for nz, world in ((64, 8), (20, 3), (7, 7), (160, 8), (9, 2)):
        load = rng.integers(0, 1000, nz).astype(float)
        load[rng.integers(0, nz, nz // 3)] = 0.0
        planes = capi.balanced_planes(load, world)
        assert len(planes) == world and sum(planes) == nz and min(planes) >= 1
    with pytest.raises(ValueError):
        capi.balanced_planes(np.ones(3), 4)
    medium = np.zeros(64)
    medium[8:56] = 4096 * 1333.0
    planes = capi.balanced_planes(medium, 8)
    assert planes == [14, 6, 6, 6, 6, 6, 6, 14]
    z0 = np.concatenate([[0], np.cumsum(planes)])
    shares = [medium[z0[r]:z0[r + 1]].sum() for r in range(8)]
    assert max(shares) == min(shares)


def test_slab_of_with_explicit_ranges():
    """slab_of with (z_begin, z_count): uneven slabs put together give back the universe."""
    import numpy as np
    from allscale_ipic3d_b200 import capi
    rng = np.random.default_rng(9)
    nx = ny = 2
    nz = 7
    counts = rng.integers(0, 5, nx * ny * nz).astype(np.uint64)
    aos = rng.normal(size=(int(counts.sum()), 8))
    first = np.concatenate([[0], np.cumsum(counts)]).astype(int)
    planes = [3, 1, 3]
    got = {}
    for rank, lz in enumerate(planes):
        z0 = sum(planes[:rank])
        c, a = capi.slab_of(counts, aos, nx * ny, nz, rank, len(planes), z_begin=z0, z_count=lz)
        off = np.concatenate([[0], np.cumsum(c)]).astype(int)
        for xy in range(nx * ny):
            for zl in range(lz):
                k = xy * lz + zl
                got[xy * nz + z0 + zl] = a[off[k]:off[k + 1]]
    for cell in range(nx * ny * nz):
        assert np.array_equal(got[cell], aos[first[cell]:first[cell + 1]])
