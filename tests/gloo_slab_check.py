"""Helper of tests/test_slab_host_gloo.py: one process per slab over the gloo backend (CPU).  Each rank builds its
z-slab of the reference's :U:N universe with the host generator bench.py uses, the ranks all-gather per-slab
digests and timings exactly the way bench.py aggregates (max over ranks), and rank 0 checks the slabs against
the undecomposed universe."""
import hashlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from allscale_ipic3d_b200 import capi  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    grid, n_total = 8, 40_003
    props = capi.uniform_benchmark_props(grid)
    lz = grid // world
    counts, aos = capi.uniform_universe(props, n_total, rank * lz, lz)
    # what bench.py reports: whole-job totals and the max over ranks of a per-rank time
    local = torch.tensor([float(counts.sum()), float(rank + 1)], dtype=torch.float64)
    total = local.clone(); dist.all_reduce(total, op=dist.ReduceOp.SUM)
    worst = local.clone(); dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    assert int(total[0].item()) == n_total, "slab populations do not add up"
    assert worst[1].item() == world
    gathered = [None] * world
    dist.all_gather_object(gathered, (counts.tobytes(), hashlib.sha256(aos.tobytes()).hexdigest(), len(aos)))
    if rank == 0:
        gc, ga = capi.uniform_universe(props, n_total)                  # the undecomposed universe
        start = np.concatenate([[0], np.cumsum(gc)]).astype(np.int64)
        nxy = grid * grid
        for r, (cbytes, digest, n) in enumerate(gathered):
            idx = (np.arange(nxy)[:, None] * grid + r * lz + np.arange(lz)[None, :]).reshape(-1)
            assert np.frombuffer(cbytes, dtype=np.uint64).tolist() == gc[idx].tolist(), "rank %d counts" % r
            rows = np.concatenate([np.arange(start[i], start[i + 1]) for i in idx])
            assert hashlib.sha256(ga[rows].tobytes()).hexdigest() == digest and len(rows) == n, "rank %d particles" % r
        print("gloo slab check ok: %d ranks" % world, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
