"""Multi-process check of the z-slab path over NCCL (one process per GPU).  Not collected by pytest itself
(needs >= 2 GPUs): tests/test_gpu_multi.py launches it under torchrun when the box has them, and it can be run
by hand:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_nccl_check.py

Every rank builds the same universe on the host, keeps its own z-slab, steps it through the C ABI with NCCL
send/recv exchange, and compares its slab with the CPU oracle's result (bit-exact counts and particles).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import harness  # noqa: E402
from harness import Oracle, Props, Ref, bits_equal  # noqa: E402


def slab_of(props, counts, aos, rank, world, planes=None):
    nz = int(props.size[2])
    planes = planes or [nz // world] * world
    z0, lz = int(sum(planes[:rank])), planes[rank]
    nxy = int(props.size[0] * props.size[1])
    idx = (np.arange(nxy)[:, None] * nz + z0 + np.arange(lz)[None, :]).reshape(-1)
    start = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    rows = np.concatenate([np.arange(start[i], start[i + 1]) for i in idx]) if len(idx) else np.zeros(0, dtype=np.int64)
    return idx, counts[idx], aos[rows.astype(np.int64)]


def main():
    import torch
    import torch.distributed as dist
    import allscale_ipic3d_b200 as ipic
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")

    def check(name, props, counts, aos, steps, batches, flags=0, fields=None, planes=None):
        o = Oracle(props)
        o.set_cells(counts, aos)
        if fields is not None:
            o.field(fields[0]); o.bcfield(fields[1])
        nz = int(props.size[2])
        mine = planes or [nz // world] * world
        z0, lz = int(sum(mine[:rank])), mine[rank]
        ctx = ipic.Context(props, device=local, rank=rank, nranks=world, z_begin=z0, z_count=lz)
        uid = [ipic.Context.comm_unique_id() if rank == 0 else None]      # an NCCL id serves one communicator
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0])
        ctx.set_headroom(3.0, 256)
        idx, c0, a0 = slab_of(props, counts, aos, rank, world, planes)
        ctx.upload_cells(c0, a0)
        if fields is not None:
            ctx.upload_fields(fields[0], fields[1], None)
        for b in range(batches):
            ctx.step(steps, flags)
            o.step(steps, flags)
            _, wc, wa = slab_of(props, o.counts(), o.particles(), rank, world, planes)
            gc, ga = ctx.download_cells()
            assert np.array_equal(gc, wc), "%s: rank %d counts differ after batch %d" % (name, rank, b)
            if flags == 0 and props.externalMagneticField[2] == 0.0:
                assert bits_equal(ga, wa), "%s: rank %d particles differ after batch %d" % (name, rank, b)
            else:
                # free-running dipole steps (10 here): 1e-12 norm-relative, the bar of north_star
                ep, ev = harness.norm_rel_err(ga[:, 0:3], wa[:, 0:3]), harness.norm_rel_err(ga[:, 3:6], wa[:, 3:6])
                assert ep <= 1e-12 and ev <= 1e-12, "%s: pos %.3g vel %.3g" % (name, ep, ev)
        if flags & harness.PROJECT_CURRENT:
            f, bc, j = ctx.download_fields()
            jw = o.density()
            n1 = [int(props.size[a]) + 1 for a in range(3)]
            if not flags & (harness.SOLVER_FORWARD | harness.SOLVER_LEAPFROG):
                # projection alone: J lives on the slab's own node planes (halo exchange of one plane of corner sums)
                j = j.reshape(n1 + [3])[:, :, z0:z0 + lz + 1]
                jw = jw.reshape(n1 + [3])[:, :, z0:z0 + lz + 1]
            assert np.max(np.abs(j - jw)) <= 1e-12 * np.max(np.abs(o.density())), "%s: J differs" % name
            fw = o.field()
            assert np.max(np.abs(f - fw)) <= 1e-12 * max(np.max(np.abs(fw)), 1e-300), "%s: fields differ" % name
        ctx.close()
        dist.barrier()
        if rank == 0:
            print("ok: %s (%d ranks)" % (name, world), flush=True)

    # 1. uniform plasma, heavy migration (big dt), pipelined multi-step calls
    n = 2 * world
    props = Props.make(n, cellWidth=(10.0, 10.0, 10.0), dt=0.15 * 40)
    r = Ref.uniform_grid(props, 5000 * n)
    check("uniform heavy migration %d^3" % n, props, r.counts(), r.particles(), steps=4, batches=5)

    # 2. Earth dipole (absorption, reflection), committed golden start state
    g = np.load(os.path.join(HERE, "golden", "small_dipole.npz"))
    p = Props()
    import ctypes
    ctypes.memmove(ctypes.byref(p), g["props"].tobytes(), ctypes.sizeof(p))
    if int(p.size[2]) % world == 0:
        check("small dipole", p, g["counts0"], g["p0"], steps=2, batches=5)

    # 3. full PIC step with all-gathered current, random particles and fields (same seed on every rank)
    n = 2 * world
    props = Props.make(n, cellWidth=(0.5, 0.5, 0.5), dt=0.05)
    rng = np.random.default_rng(77)
    n_per, nc = 40, n ** 3
    cells = np.arange(nc)
    lo = np.stack([cells // (n * n), (cells // n) % n, cells % n], 1) * 0.5
    aos = np.zeros((nc * n_per, 8))
    aos[:, 0:3] = np.repeat(lo, n_per, 0) + rng.random((nc * n_per, 3)) * 0.5
    aos[:, 3:6] = rng.uniform(-1.0, 1.0, (nc * n_per, 3))
    aos[:, 6] = rng.choice([-1.0, 1.0], nc * n_per) * 0.01
    aos[:, 7] = np.where(aos[:, 6] < 0, -25.0, 1.0)
    fields = (rng.normal(0, 0.1, ((n + 3) ** 3, 9)), rng.normal(0, 0.1, ((n + 2) ** 3, 3)))
    counts = np.full(nc, n_per, dtype=np.uint64)
    fl = harness.PROJECT_CURRENT | harness.UPDATE_BOUNDARIES | harness.SOLVER_LEAPFROG
    check("full PIC step (interpolated)", props, counts, aos, steps=1, batches=3, flags=fl | harness.MOVER_INTERPOLATED, fields=fields)
    check("full PIC step (fused deposit)", props, counts, aos, steps=1, batches=3, flags=fl, fields=fields)
    # projection alone: the current's halo exchange instead of the all-gather, fused and as a separate pass
    check("projection only (fused deposit, J halo)", props, counts, aos, steps=2, batches=2, flags=harness.PROJECT_CURRENT, fields=fields)
    check("projection only (interpolated mover, J halo)", props, counts, aos, steps=1, batches=2, flags=harness.PROJECT_CURRENT | harness.MOVER_INTERPOLATED, fields=fields)

    # 4. slabs of different thickness (cut by particle count): 2 world + 3 planes, the first rank takes the extra ones
    n = 2 * world + 3
    props = Props.make(n, cellWidth=(10.0, 10.0, 10.0), dt=0.15 * 40)
    r = Ref.uniform_grid(props, 150 * n ** 3)
    check("uneven slabs, uniform heavy migration", props, r.counts(), r.particles(), steps=3, batches=3, planes=[5] + [2] * (world - 1))
    check("uneven slabs, projection only", props, r.counts(), r.particles(), steps=2, batches=2, flags=harness.PROJECT_CURRENT, planes=[2] * (world - 1) + [5])
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
