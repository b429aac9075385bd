#!/usr/bin/env python
"""bench.py — particle-pushes/s of the hot path (mover + current deposit + cell-to-cell migration).

Workload (config.workload): the reference's own synthetic `:U:N` uniform plasma (benchmark.h:57-92,
initCells cell.h:638-697) on the fixed 32^3-cell grid, N = 32 M particles per GPU (weak scaling: the
reference's scripts multiply N by the node count, scripts/ipic3d_udist_8nodes.sh:49-52), z-slab
decomposed over the ranks.  A "step" is one iteration of the simulateSteps loop body over all particles.

    python bench.py [--gpus N] [--steps K] [--warmup W]              our arm  (CUDA, C ABI)
    python bench.py --impl reference [...]                           the reference's CPU path (oracle/_ref)

Prints ONE JSON line (rank 0).  See DESIGN.md §6 for how every field is measured.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# Algorithmic bytes per push: read pos,vel,q,qom (64 B) + write pos,vel (48 B) = 112 B for a
# universe with per-particle species.  A universe that holds ONE species (every :U: run) is stored with 6 FP64 fields per
# particle (q, qom are constants of the context), so the bytes the path has to move are 48 R + 48 W = 96 B; the roofline
# is quoted on the figure of the layout actually resident, and the 112 B equivalent is reported next to it.
BYTES_PER_PUSH = {8: 112, 6: 96}
SURVEY_BYTES_PER_PUSH = 112
METRIC = "particle-pushes/sec (mover+deposit+migration)"
UNIT = "particle-pushes/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--particles-per-gpu", type=int, default=32_000_000)
    ap.add_argument("--grid", type=int, default=32)
    ap.add_argument("--no-deposit", action="store_true", help="time the loop as the reference ships it (STEP 1 off)")
    ap.add_argument("--workload", default="uniform", choices=["uniform", "dipole", "medium", "large"],
                    help="uniform = the headline :U:N plasma; dipole = a secondary data point (fixed total, z-slabs over the ranks): the reference's Earth-dipole "
                         ".inp set-up (medium.inp physics: 0.625 R_E cells, 10 MeV protons, absorption, ~17 %% migration per step) on --grid^3 "
                         "cells with 8^3 particles per cell")
    ap.add_argument("--device-init", action="store_true",
                    help="device-resident leg only: create the :U: universe in HBM with ipic_generate_uniform (bit-identical to the host "
                         "generator) instead of generating on the host and uploading; implies --no-e2e")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity-gate", action="store_true", help="skip the correctness gate that runs before the timed region")
    ap.add_argument("--cpu-particles", type=int, default=0, help="particles of the CPU runs (0 = the workload's own N, capped only by host memory)")
    ap.add_argument("--cpu-steps", type=int, default=5, help="timed steps of the cpu_baseline leg of the b200 arm (2 warm-up steps before)")
    ap.add_argument("--gate-steps", type=int, default=7, help="steps of the parity gate (the reference's benchmark protocol: 2 warm-up + 5 timed)")
    return ap.parse_args()


def workload_config(args, world):
    """The workload both arms run -- nothing arm-specific in here, so the two JSON lines carry the same `config`."""
    g = args.grid
    if args.workload == "uniform":
        return {"workload": "uniform plasma :U:N (reference generator initCells(uniform<>), benchmark.h:57-92), %dx%dx%d cells, %d particles per GPU "
                            "(N = %d), dt 0.15, weak scaling over z-slabs" % (g, g, g, args.particles_per_gpu, args.particles_per_gpu * world),
                "particles": args.particles_per_gpu * world, "particles_per_gpu": args.particles_per_gpu, "grid": [g, g, g], "dt": 0.15}
    if args.workload == "medium":
        return {"workload": "inputs/medium.inp as `ipic3d medium.inp` sets it up (Earth dipole, 64x64x64 cells, 10^3 particles per cell = 262144000, "
                            "two-phase initCells), first K steps of the 2400, strong scaling over z-slabs",
                "particles": 262144000, "particles_per_gpu": 262144000 // world, "grid": [64, 64, 64], "dt": 0.01}
    if args.workload == "large":
        return {"workload": "CAPACITY DEMO: inputs/large.inp as `ipic3d large.inp` means it (Earth dipole, 160x160x160 cells, 12^3 particles per cell = "
                            "7077888000; the reference's int particle count overflows on this deck, so there is no reference run to compare with), "
                            "universe created on the devices, first K steps, z-slabs",
                "particles": 7077888000, "particles_per_gpu": 7077888000 // world, "grid": [160, 160, 160], "dt": 0.01}
    return {"workload": "SECONDARY: Earth-dipole .inp set-up of the reference (medium.inp physics, host mirror generator), %dx%dx%d cells, "
                        "8^3 particles per cell, strong scaling over z-slabs" % (g, g, g),
            "particles": g * g * g * 512, "particles_per_gpu": g * g * g * 512 // world, "grid": [g, g, g], "dt": 0.01}


def host_memory_allows(n_particles):
    """The reference keeps every particle in a std::vector per cell (64 B each) plus transfer buffers; ask for 3x."""
    try:
        import psutil
        return psutil.virtual_memory().available > 3 * 64 * n_particles
    except Exception:
        return n_particles <= 64_000_000


def measured_peaks():
    """HBM roofline denominator: MEASURED_PEAKS.json (driver-written) else the profiling recipe's fallback."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ---------------------------------------------------------------------------------------------------
# the CPU side: the reference's own implementation (oracle/_ref, threaded pfor shim).  Used (a) as the reference arm,
# (b) as the cpu_baseline leg of our arm and (c) as the CHECKER of the parity gate -- never on the measured GPU path.
# ---------------------------------------------------------------------------------------------------

def _harness():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import harness
    return harness


def reference_universe(args, n_particles):
    """The workload's universe built by the UNMODIFIED reference headers (oracle/_ref/libipic3d_ref_mt.so)."""
    harness = _harness()
    threaded = harness.Ref.available(True)
    if not harness.Ref.available(threaded):
        return None, 0
    if args.workload == "uniform":
        if args.grid == 32:
            r = harness.Ref.uniform(n_particles, threaded=threaded)             # processUniform, benchmark.h:81
        else:
            r = harness.Ref.uniform_grid(harness.Props.make(args.grid, cellWidth=(10.0, 10.0, 10.0), dt=0.15), n_particles, threaded=threaded)
    elif args.workload == "medium":
        r = harness.Ref.from_inp(os.path.join(ROOT, "tests", "golden", "inputs", "medium.inp"), threaded=threaded)   # main(), ipic3d.cxx:32-83
    else:
        return None, 0
    return r, (r.threads() if threaded else 1)


def time_reference(args, n_particles, steps, warmup, deposit=False, keep=False, world=None):
    """`warmup` + `steps` steps of the reference's own time loop on all host threads; returns (result dict, universe or None)."""
    harness = _harness()
    r, cores = reference_universe(args, n_particles)
    if r is None:
        return None, None
    flags = harness.PROJECT_CURRENT if deposit else 0
    if warmup:
        r.step(warmup, flags)
    t = r.step(steps, flags)
    res = {"value": steps * n_particles / t, "seconds": t, "cores": cores, "kind": "reference", "unit": UNIT,
           "sample": "%d particles (%s), %d warm-up + %d timed steps of %s, %d host threads" % (
               n_particles, "the whole workload" if n_particles == workload_config(args, world or int(os.environ.get("WORLD_SIZE", "1")))["particles"] else "bounded sample of the workload",
               warmup, steps,
               "ref_step_full (projectToDensityField + mover + export/import, the order of the commented loop simulator.h:189-223)" if deposit
               else "simulateSteps as shipped (mover + export/import; STEP 1 is commented out in the reference, simulator.h:189-200)", cores)}
    if keep:
        return res, r
    r.close()
    return res, None


def run_reference(args):
    """--impl reference: the reference's CPU path on this box's host cores, same config / steps / warm-up as our arm."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    if rank != 0:
        return 0
    cfg = workload_config(args, world)
    n = args.cpu_particles or cfg["particles"]
    bounded = None
    if not host_memory_allows(n):
        # the host cannot hold the whole universe of the N-GPU workload three times over: bounded sample, stated
        while n > 1_000_000 and not host_memory_allows(n):
            n //= 2
        bounded = "host memory: sample of %d particles instead of %d" % (n, cfg["particles"])
    t0 = time.time()
    res, _ = time_reference(args, n, args.steps, args.warmup, deposit=False, world=world)
    if res is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref was not built (reference sources absent at build time)"}))
        return 0
    # second, labelled number: the same loop WITH the reference's own projectToDensityField (what our arm fuses in)
    dep_steps = max(1, min(args.steps, 5))
    dep, _ = time_reference(args, n, dep_steps, 1, deposit=True, world=world) if args.workload == "uniform" else (None, None)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["seconds"] / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak" if args.workload == "uniform" else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "arm": {"what": "stock simulateSteps of the unmodified reference headers (mover + export/import); the reference ships STEP 1 (deposit) commented "
                        "out, so `value` is its fastest shipped loop and `with_deposit` the like-for-like mover+deposit+migration",
                "particles_run": n, "bounded": bounded},
        "with_deposit": None if dep is None else {"value": dep["value"], "unit": UNIT, "steps": dep_steps, "warmup": 1, "sample": dep["sample"]},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.time() - t0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------
# parity gate: runs BEFORE the timed region, on the same code path (C ABI, NCCL slabs for N > 1)
# ---------------------------------------------------------------------------------------------------

def cell_hashes(np, counts, aos):
    """Order-sensitive 64-bit hash of every cell's particle records (bit patterns), vectorised."""
    counts = np.asarray(counts, dtype=np.int64)
    u = np.ascontiguousarray(aos, dtype=np.float64).reshape(-1, 8).view(np.uint64)
    start = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    with np.errstate(over="ignore"):
        h = np.zeros(len(u), dtype=np.uint64)
        for f in range(8):
            h = (h ^ u[:, f]) * np.uint64(0x9E3779B97F4A7C15 + 2 * f)
        h ^= h >> np.uint64(31)
        within = (np.arange(len(u), dtype=np.int64) - np.repeat(start[:-1], counts)).astype(np.uint64)
        h *= (within * np.uint64(2) + np.uint64(1))
        cs = np.concatenate([np.zeros(1, dtype=np.uint64), np.cumsum(h, dtype=np.uint64)])
        return cs[start[1:]] - cs[start[:-1]]


def parity_gate(args, np, dist, ipic, capi, rank, world, local, flags):
    """Correctness gate in front of the timed region.

    N = 1: the timed workload itself (e.g. :U:32000000) for --gate-steps steps on the GPU and on the reference
           (libipic3d_ref_mt.so): per-cell counts equal and EVERY particle bit-identical, in per-cell order.
    N > 1: through the same NCCL slab path that is timed: (a) :U:(4 M x N) for --gate-steps steps, per-cell counts and
           an order-sensitive hash of every cell's particle bits equal to the reference's; (b) the reference's test.inp
           deck (Earth dipole, 8^3 cells, absorption + reflection + 17 % migration per step) for 10 steps: counts
           bit-exact, positions / velocities within 1e-12 norm-relative (the dipole's pow differs by one ulp in 0.08 %
           of pushes).
    Returns (gate dict, cpu_baseline dict or None)."""
    harness = _harness()
    gate = {"ran": True, "ok": False, "checks": []}
    cpu = None
    steps = args.gate_steps
    lz_of = lambda nz: nz // world

    def bcast(obj):
        if world == 1:
            return obj
        box = [obj if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    def gather(obj):
        if world == 1:
            return [obj]
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    def slab_ctx(props, zr=None):
        ctx = ipic.Context(props, device=local, rank=rank, nranks=world, z_begin=zr[0] if zr else None, z_count=zr[1] if zr else None)
        if world > 1:
            ctx.comm_init(bcast(ipic.Context.comm_unique_id() if rank == 0 else None))
        return ctx

    # ---- the medium.inp deck: truncated-step parity at full size (SURVEY 8d "M verified on a truncated step range") -------
    if args.workload == "medium":
        deck = os.path.join(ROOT, "tests", "golden", "inputs", "medium.inp")
        shm = "/dev/shm/ipic_gate_%d" % os.getppid() if world > 1 else "/dev/shm/ipic_gate_%d" % os.getpid()
        shm = bcast(shm)
        have_ref, n_ref = True, 0
        if rank == 0:
            threaded = harness.Ref.available(True)
            have_ref = harness.Ref.available(threaded) and host_memory_allows(262144000)
            if have_ref:
                r = harness.Ref.from_inp(deck, threaded=threaded)
                t = r.step(steps)
                if not args.no_cpu_baseline:
                    cpu = {"value": steps * r.num_particles / t, "unit": UNIT, "cores": r.threads() if threaded else 1, "kind": "reference",
                           "sample": "the whole workload: medium.inp universe built by the reference's main() set-up, first %d steps of its simulateSteps "
                                     "(includes the first, cold step), %d host threads" % (steps, r.threads() if threaded else 1)}
                np.save(shm + "_counts.npy", r.counts())
                n_ref = r.total()
                mm = np.lib.format.open_memmap(shm + "_particles.npy", mode="w+", dtype=np.float64, shape=(n_ref, 8))
                r.lib.ref_get_particles(r.h, mm.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
                mm.flush(); del mm
                r.close()
        have_ref = bcast(have_ref)
        if not have_ref:
            return {"ran": False, "ok": None, "why": "oracle/_ref not built or not enough host memory for the reference's 16.8 GB universe"}, None
        props, counts, host, zr = capi.deck_slab_pinned(deck, rank, world, __import__("torch"), balanced=True)
        nz = int(props.size[2]); nxy = int(props.size[0] * props.size[1]); z0, lz = zr
        with slab_ctx(props, zr) as ctx:
            ctx.upload_cells_ptr(counts, host.data_ptr())
            del host
            ctx.step(steps, flags)
            got_c, got_p = ctx.download_cells()
        ref_counts = np.load(shm + "_counts.npy")
        ref_p = np.load(shm + "_particles.npy", mmap_mode="r")
        cells = (np.arange(nxy)[:, None] * nz + z0 + np.arange(lz)[None, :]).reshape(-1)
        same = bool(np.array_equal(got_c, ref_counts[cells]))
        ep = ev = 0.0
        checked = 0
        if same:
            ref_start = np.concatenate([[0], np.cumsum(ref_counts)]).astype(np.int64)
            my_start = np.concatenate([[0], np.cumsum(got_c)]).astype(np.int64)
            for k in range(0, len(cells), 61):                       # every 61st cell of the slab, all of its particles
                a = got_p[my_start[k]:my_start[k + 1]]
                b = np.asarray(ref_p[ref_start[cells[k]]:ref_start[cells[k] + 1]])
                if len(a):
                    ep = max(ep, harness.norm_rel_err(a[:, 0:3], b[:, 0:3])); ev = max(ev, harness.norm_rel_err(a[:, 3:6], b[:, 3:6]))
                    checked += len(a)
        parts = gather((same, ep, ev, checked, int(got_c.sum()), lz))
        del ref_p
        if world > 1:
            dist.barrier()
        if rank == 0:
            for f in ("_counts.npy", "_particles.npy"):
                try:
                    os.unlink(shm + f)
                except OSError:
                    pass
        ok = all(p[0] for p in parts) and max(p[1] for p in parts) <= 1e-12 and max(p[2] for p in parts) <= 1e-12
        gate["checks"].append({"case": "medium.inp (64^3 cells, 262144000 particles), first %d steps, %d GPU(s)%s vs the reference's own run of the deck" % (
                                   steps, world, " as NCCL z-slabs cut by particle count" if world > 1 else ""),
                               "planes_per_slab": [p[5] for p in parts],
                               "counts_equal_all_262144_cells": all(p[0] for p in parts), "particles_alive": int(sum(p[4] for p in parts)),
                               "particles_compared": int(sum(p[3] for p in parts)), "max_norm_rel_err_pos": max(p[1] for p in parts),
                               "max_norm_rel_err_vel": max(p[2] for p in parts), "tolerance": 1e-12})
        gate["ok"] = bool(ok)
        return gate, cpu

    # ---- (a) the uniform plasma ------------------------------------------------------------------
    if args.workload == "uniform":
        n_gate = args.particles_per_gpu * world if world == 1 else min(args.particles_per_gpu, 4_000_000) * world
        props = capi.uniform_benchmark_props(args.grid)
        lz = lz_of(args.grid)
        counts, aos = capi.uniform_universe(props, n_gate, rank * lz, lz)
        ref_counts = ref_hash = ref_aos = None
        have_ref = True
        if rank == 0:
            res, r = time_reference(args, n_gate, max(1, steps - 2), min(2, steps - 1), deposit=False, keep=True)
            have_ref = r is not None
            if have_ref:
                if world == 1 and not args.no_cpu_baseline:
                    cpu = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
                ref_counts, ref_aos = r.counts(), r.particles()
                r.close()
        have_ref = bcast(have_ref)
        if not have_ref:
            return {"ran": False, "ok": None, "why": "oracle/_ref not built"}, None
        with slab_ctx(props) as ctx:
            ctx.upload_cells(counts, aos)
            ctx.step(steps, flags)
            got_counts, got_aos = ctx.download_cells()
            misplaced = ctx.count_misplaced()
        del aos
        if world == 1:
            ok_counts = bool(np.array_equal(got_counts, ref_counts))
            ok_bits = bool(got_aos.shape == ref_aos.shape and np.array_equal(got_aos.view(np.uint64), ref_aos.view(np.uint64)))
            gate["checks"].append({"case": ":U:%d, %d steps, 1 GPU vs libipic3d_ref_mt.so" % (n_gate, steps), "cells": int(len(ref_counts)),
                                   "particles": int(len(ref_aos)), "counts_equal": ok_counts, "all_particle_bits_equal": ok_bits, "misplaced": misplaced})
            ok = ok_counts and ok_bits and misplaced == 0
        else:
            mine = (got_counts, cell_hashes(np, got_counts, got_aos), misplaced)
            parts = gather(mine)
            ok = True
            if rank == 0:
                nxy = args.grid * args.grid
                full_c = np.zeros(nxy * args.grid, dtype=np.uint64)
                full_h = np.zeros(nxy * args.grid, dtype=np.uint64)
                for r_, (c_, h_, _) in enumerate(parts):
                    idx = (np.arange(nxy)[:, None] * args.grid + r_ * lz + np.arange(lz)[None, :]).reshape(-1)
                    full_c[idx] = c_
                    full_h[idx] = h_
                ok_counts = bool(np.array_equal(full_c, ref_counts))
                ok_bits = bool(np.array_equal(full_h, cell_hashes(np, ref_counts, ref_aos)))
                bad = int(sum(p[2] for p in parts))
                gate["checks"].append({"case": ":U:%d, %d steps, %d NCCL z-slabs vs libipic3d_ref_mt.so" % (n_gate, steps, world), "cells": int(len(ref_counts)),
                                       "particles": int(len(ref_aos)), "counts_equal": ok_counts, "per_cell_ordered_bit_hash_equal": ok_bits, "misplaced": bad})
                ok = ok_counts and ok_bits and bad == 0
            ok = bcast(ok)
        del ref_aos, got_aos
        if not ok:
            return gate, cpu

    # ---- (b) the dipole deck through the slab path (N > 1; one GPU is covered by tests/ and smoke()) ---------------
    deck = os.path.join(ROOT, "tests", "golden", "inputs", "test.inp")
    if world > 1 and 8 % world == 0 and os.path.exists(deck):
        dsteps = 10
        pack = None
        if rank == 0:
            r = harness.Ref.from_inp(deck, threaded=harness.Ref.available(True))
            c0, p0 = r.counts(), r.particles()
            r.step(dsteps)
            pack = (bytes(r.props), c0, p0, r.counts(), r.particles())
            r.close()
        pack = bcast(pack)
        props = capi.Props()
        ctypes.memmove(ctypes.byref(props), pack[0], ctypes.sizeof(props))
        nz = int(props.size[2]); nxy = int(props.size[0] * props.size[1])
        c_mine, p_mine = capi.slab_of(pack[1], pack[2], nxy, nz, rank, world)
        with slab_ctx(props) as ctx:
            ctx.upload_cells(c_mine, p_mine)
            ctx.step(dsteps, flags)
            got_c, got_p = ctx.download_cells()
        want_c, want_p = capi.slab_of(pack[3], pack[4], nxy, nz, rank, world)
        same = bool(np.array_equal(got_c, want_c))
        ep = harness.norm_rel_err(got_p[:, 0:3], want_p[:, 0:3]) if same else float("inf")
        ev = harness.norm_rel_err(got_p[:, 3:6], want_p[:, 3:6]) if same else float("inf")
        parts = gather((same, ep, ev))
        ok = all(p[0] for p in parts) and max(p[1] for p in parts) <= 1e-12 and max(p[2] for p in parts) <= 1e-12
        gate["checks"].append({"case": "test.inp deck (Earth dipole, 8^3 cells, 262144 particles), %d steps, %d NCCL z-slabs vs libipic3d_ref" % (dsteps, world),
                               "counts_equal": all(p[0] for p in parts), "max_norm_rel_err_pos": max(p[1] for p in parts),
                               "max_norm_rel_err_vel": max(p[2] for p in parts), "tolerance": 1e-12})
        if not ok:
            return gate, cpu
    gate["ok"] = True
    return gate, cpu


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------

def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import allscale_ipic3d_b200 as ipic
    from allscale_ipic3d_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    cfg = workload_config(args, world)
    nz = cfg["grid"][2]
    if nz % world:
        raise SystemExit("grid z extent %d not divisible by %d ranks" % (nz, world))
    lz = nz // world
    zr = None                                             # (z_begin, z_count) when the slabs are cut by particle count
    flags = 0 if args.no_deposit else ipic.PROJECT_CURRENT

    # ---- parity gate (and, at N = 1, the cpu_baseline leg: the same reference run serves both) ------------------
    gate, cpu = {"ran": False, "ok": None, "why": "--no-parity-gate" if args.workload != "large" else
                 "the reference cannot run large.inp (its int particle count overflows, ipic3d.cxx:70-71): no oracle at this size; the same kernels "
                 "are gated on medium.inp and test.inp"}, None
    if not args.no_parity_gate and args.workload in ("uniform", "medium"):
        gate, cpu = parity_gate(args, np, dist, ipic, capi, rank, world, local, flags)
        if gate.get("ran") and not gate.get("ok"):
            if rank == 0:
                print(json.dumps({"metric": METRIC, "value": None, "unit": UNIT, "n_gpus": world, "config": cfg, "parity_gate": gate,
                                  "error": "parity gate failed: nothing was timed"}))
            if world > 1:
                dist.destroy_process_group()
            return 3

    if args.workload == "dipole":
        # every rank builds the whole universe (host mirror generator) and keeps its z-slab; the total is fixed: strong scaling
        props, counts, aos = capi.dipole_universe(os.path.join(ROOT, "tests", "golden", "inputs", "medium.inp"), args.grid, 8)
        n_total = int(counts.sum())
        if world > 1:
            counts, aos = capi.slab_of(counts, aos, nz * nz, nz, rank, world)
        n_local = int(counts.sum())
        host = torch.from_numpy(np.ascontiguousarray(aos)).pin_memory()
        del aos
    elif args.workload == "large":
        # only the device generator can build this universe: 453 GB of particle records would not fit the host
        args.no_e2e = True
        free_b, _ = torch.cuda.mem_get_info()
        if cfg["particles_per_gpu"] * 48 * 2 * 1.3 > free_b:
            raise SystemExit("large.inp: %.0f GB of particle state per GPU at %d GPUs do not fit (%.0f GB free); use --gpus 8"
                             % (cfg["particles_per_gpu"] * 48 * 2 * 1.3 / 1e9, world, free_b / 1e9))
        deck = capi.Deck(os.path.join(ROOT, "tests", "golden", "inputs", "large.inp"))
        assert deck.num_particles == cfg["particles"]
        props = capi.Props.from_other(deck.props)
        planes = capi.balanced_planes(deck.plane_population(), world)
        zr = (int(sum(planes[:rank])), planes[rank])
        counts = deck.counts(*zr)
        n_total = cfg["particles"]
        n_local = int(counts.sum())
        host = None
    elif args.workload == "medium" and args.device_init:
        # phase 1 of initCells on the host (sequential by construction, ~2 s), phase 2 on the GPU (ipic_generate_normal_speed)
        args.no_e2e = True
        deck = capi.Deck(os.path.join(ROOT, "tests", "golden", "inputs", "medium.inp"))
        props = capi.Props.from_other(deck.props)
        planes = capi.balanced_planes(deck.plane_population(), world)
        zr = (int(sum(planes[:rank])), planes[rank])
        counts = deck.counts(*zr)
        n_total = cfg["particles"]
        n_local = int(counts.sum())
        host = None
    elif args.workload == "medium":
        # the deck as `ipic3d medium.inp` sets it up; every rank generates ONLY its z-slab (phase 1 of initCells is cheap and
        # replicated, phase 2 is per cell), straight into pinned memory
        props, counts, host, zr = capi.deck_slab_pinned(os.path.join(ROOT, "tests", "golden", "inputs", "medium.inp"), rank, world, torch, balanced=True)
        n_total = cfg["particles"]
        n_local = int(counts.sum())
    elif args.device_init:
        args.no_e2e = True
        props = capi.uniform_benchmark_props(args.grid)
        n_total = args.particles_per_gpu * world
        counts = capi.uniform_counts(props, n_total, rank * lz, lz)
        n_local = int(counts.sum())
        host = None
    else:
        props = capi.uniform_benchmark_props(args.grid)
        n_total = args.particles_per_gpu * world
        counts = capi.uniform_counts(props, n_total, rank * lz, lz)
        n_local = int(counts.sum())
        # the host universe of this slab, pinned (what simulateSteps hands to the C ABI)
        host = torch.empty((n_local, 8), dtype=torch.float64).pin_memory()
        capi.uniform_particles_into(props, n_total, host.data_ptr(), rank * lz, lz)
    result = None if args.no_e2e else torch.empty((n_local + n_local // 8 + 4096, 8), dtype=torch.float64).pin_memory()

    ctx = ipic.Context(props, device=local, rank=rank, nranks=world, z_begin=zr[0] if zr else None, z_count=zr[1] if zr else None)
    if world > 1:
        uid = [ipic.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0])
    t_init = time.perf_counter()
    if host is None and args.workload in ("medium", "large"):
        ctx.generate_normal_speed(counts, *deck.distribution())
        deck.close()
    elif host is None:
        ctx.generate_uniform(n_total)
    else:
        ctx.upload_cells_ptr(counts, host.data_ptr())
    particle_fields = ctx.particle_fields()
    t_init = time.perf_counter() - t_init

    # ---- device-resident throughput: W warm-up steps, then exactly K timed steps -------------------
    sampler = ClockSampler(local)
    ctx.step(args.warmup, flags)
    ctx.profile_reset()
    alive_before = sum_over_ranks(ctx.total_particles())
    barrier()
    sampler.start()
    t_first, t_rest = ctx.step(args.steps, flags)          # CUDA events on the context's stream bracket the K steps
    barrier()
    clocks = sampler.stop()
    mover_ms, import_ms, launches = ctx.profile_get()
    phases = [p / args.steps for p in ctx.profile_phases()]
    seconds = max_over_ranks(t_first + t_rest)
    alive = sum_over_ranks(ctx.total_particles())
    misplaced = sum_over_ranks(ctx.count_misplaced())
    relayouts = sum_over_ranks(ctx.relayout_count())
    # :U: universe: nothing is absorbed (planetRadius = 0).  Dipole decks: the planet absorbs particles; the metric follows the
    # reference and counts the INITIAL population per step (benchmark.h:50 / ipic3d.cxx:96-99)
    pushes = args.steps * n_total
    value = pushes / seconds

    hbm_peak, peak_src = measured_peaks()
    mover_s = max_over_ranks(mover_ms / 1e3) / args.steps
    bpp = BYTES_PER_PUSH[particle_fields]
    n_pushed = alive_before / world if args.workload != "uniform" else n_local       # particles one mover launch actually streams
    achieved = n_pushed * bpp / mover_s / 1e9
    kernel = ("k_move_ws (warp-specialised: 4 push warps + 1 deposit warp per cell; fused mover + deposit + classify/compaction)"
              if (not args.no_deposit and args.workload == "uniform") else
              "k_move (fused mover%s + classify/compaction)" % ("" if args.no_deposit else " + deposit"))
    roofline = {"bound": "hbm", "kernel": kernel, "achieved": achieved,
                "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None, "peak_source": peak_src,
                "frac_basis": "resident layout: %d B/push (%d FP64 fields per particle in HBM); the SURVEY 8d figure of 112 B/push is alongside" % (bpp, particle_fields),
                "bytes_per_push": bpp, "pushes_per_launch": n_pushed, "launch_ms": mover_s * 1e3,
                "at_survey_112B_per_push": {"achieved": n_pushed * SURVEY_BYTES_PER_PUSH / mover_s / 1e9,
                                            "frac": n_pushed * SURVEY_BYTES_PER_PUSH / mover_s / 1e9 / hbm_peak,
                                            "note": "effective bandwidth had q,qom been stored per particle; can exceed the physical peak"},
                "import_ms": import_ms / args.steps,
                "phase_ms_rank0": {"movers": phases[0], "imports": phases[1], "wait_exchange": phases[2], "wait_current_and_fields": phases[3]},
                "whole_step_frac": value / world * bpp / 1e9 / hbm_peak}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            with open(traffic_file) as f:
                tj = json.load(f)
            ent = tj.get("fields%d" % particle_fields) or {}
            roofline["traffic"] = ent.get("bytes_per_push") * n_pushed if ent.get("bytes_per_push") else None
            roofline["traffic_source"] = ent.get("source")
        except Exception:
            pass

    # ---- end to end through the host-buffer API (what simulateSteps does: upload, K steps, download) ---
    e2e = None
    if not args.no_e2e:
        cnt_out = np.zeros(len(counts), dtype=np.uint64)
        barrier()
        t0 = time.perf_counter()
        ctx.upload_cells_ptr(counts, host.data_ptr())
        ctx.step(args.steps, flags)
        ctx.download_cells_ptr(cnt_out, result.data_ptr(), result.shape[0])
        barrier()
        t_call = max_over_ranks(time.perf_counter() - t0)
        # strict variant: the whole universe crosses PCIe both ways on EVERY step
        rt_steps = max(1, min(3, args.steps))
        barrier()
        t0 = time.perf_counter()
        for _ in range(rt_steps):
            ctx.upload_cells_ptr(counts, host.data_ptr())
            ctx.step(1, flags)
            ctx.download_cells_ptr(cnt_out, result.data_ptr(), result.shape[0])
        barrier()
        t_rt = max_over_ranks(time.perf_counter() - t0)
        # the reference's benchmark protocol (benchmark.h:38,43): simulateSteps(2) then simulateSteps(5) on the SAME universe; with
        # the universe resident between the two calls only the upload before the first and the download after the second cross PCIe
        barrier()
        t0 = time.perf_counter()
        ctx.upload_cells_ptr(counts, host.data_ptr())
        ctx.step(2, flags)
        barrier()
        t1 = time.perf_counter()
        ctx.step(5, flags)
        barrier()
        t2 = time.perf_counter()
        ctx.download_cells_ptr(cnt_out, result.data_ptr(), result.shape[0])
        barrier()
        t3 = time.perf_counter()
        e2e = {"value": pushes / t_call, "unit": UNIT,
               "h2d_bytes_per_step": n_local * 64 / args.steps, "d2h_bytes_per_step": int(cnt_out.sum()) * 64 / args.steps,
               "call": "simulateSteps(%d, universe) semantics: pinned host AoS -> H2D, %d steps, D2H of all cells" % (args.steps, args.steps),
               "seconds": t_call,
               "roundtrip_every_step": {"value": rt_steps * n_total / t_rt, "steps": rt_steps,
                                        "h2d_bytes_per_step": n_local * 64, "d2h_bytes_per_step": n_local * 64},
               "resident_protocol": {"what": "benchmark.h:38-50 protocol, universe resident in HBM between the two simulateSteps calls: upload, 2 warm-up steps | "
                                             "5 timed steps | download", "timed_5_steps_value": 5 * n_total / max_over_ranks(t2 - t1),
                                     "whole_protocol_value": 7 * n_total / max_over_ranks(t3 - t0), "upload_plus_warmup_s": max_over_ranks(t1 - t0),
                                     "download_s": max_over_ranks(t3 - t2)}}

    if cpu is None and rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload == "uniform":
        n_cpu = args.cpu_particles or n_total
        while n_cpu > 1_000_000 and not host_memory_allows(n_cpu):
            n_cpu //= 2
        res, _ = time_reference(args, n_cpu, args.cpu_steps, 2, deposit=False)
        if res:
            cpu = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}

    ctx.close()
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": seconds / args.steps * 1e3, "higher_is_better": True, "scaling": "weak" if args.workload == "uniform" else "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "arm": {"what": "B200: fused mover%s + migration through the C ABI (libipic3d_b200.so), z-slab per rank%s" % (
                        "" if args.no_deposit else " + current deposit", ", NCCL send/recv of face messages" if world > 1 else ""),
                    "deposit": not args.no_deposit, "parallelism": "zslab%d" % world, "particle_fields_in_hbm": particle_fields,
                    "z_planes_of_rank0": list(zr) if zr else [0, lz],
                    "l2": "inputs (%.2f GB of particle state per GPU) are larger than L2; no flush needed" % (n_local * 64 / 1e9),
                    "relayouts_since_upload": relayouts,
                    "universe_into_hbm_s": t_init, "universe_source": "device generator" if host is None else "host generator + upload of pinned AoS"},
            "parity_gate": gate,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "particles_alive_before": alive_before, "particles_alive_after": alive, "misplaced_after": misplaced,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    try:
        return run_b200(args)
    except BaseException as e:      # noqa: BLE001 -- a failing rank must take the job down at once: its peers would wait in NCCL for ever
        if isinstance(e, SystemExit) and not e.code:
            return 0
        import traceback
        traceback.print_exc()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(1)


if __name__ == "__main__":
    sys.exit(main())
