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
    ap.add_argument("--workload", default="uniform", choices=["uniform", "dipole"],
                    help="uniform = the headline :U:N plasma; dipole = a secondary data point (fixed total, z-slabs over the ranks): the reference's Earth-dipole "
                         ".inp set-up (medium.inp physics: 0.625 R_E cells, 10 MeV protons, absorption, ~17 %% migration per step) on --grid^3 "
                         "cells with 8^3 particles per cell")
    ap.add_argument("--device-init", action="store_true",
                    help="device-resident leg only: create the :U: universe in HBM with ipic_generate_uniform (bit-identical to the host "
                         "generator) instead of generating on the host and uploading; implies --no-e2e")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-particles", type=int, default=8_000_000, help="size of the bounded CPU sample")
    ap.add_argument("--cpu-steps", type=int, default=3)
    return ap.parse_args()


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
# reference arm: the reference's own CPU implementation (oracle/_ref, threaded pfor shim)
# ---------------------------------------------------------------------------------------------------

def time_reference(n_particles, steps, warmup, deposit):
    """Runs the UNMODIFIED reference headers (oracle/_ref/libipic3d_ref_mt.so): `:U:n` universe built by the
    reference's initCells, then its own simulateSteps (stock path), all host threads."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import harness
    threaded = harness.Ref.available(True)
    if not harness.Ref.available(threaded):
        return None
    r = harness.Ref.uniform(n_particles, threaded=threaded)
    cores = r.threads() if threaded else 1
    flags = harness.PROJECT_CURRENT if deposit else 0
    if warmup:
        r.step(warmup, flags)
    t = r.step(steps, flags)
    r.close()
    return {"value": steps * n_particles / t, "seconds": t, "cores": cores, "kind": "reference",
            "sample": ":U:%d on the 32^3 grid, %d warm-up + %d timed steps of %s, %d host threads" % (
                n_particles, warmup, steps,
                "ref_step_full (projectToDensityField + mover + export/import)" if deposit else "simulateSteps as shipped (mover + export/import)",
                cores)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # one bounded sample per step-batch: the same :U: workload at a size the host finishes in minutes
    n = args.cpu_particles
    steps, warm = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
    t0 = time.time()
    res = time_reference(n, steps, warm, deposit=False)
    if res is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref was not built (reference sources absent at build time)"}))
        return 0
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": res["seconds"] / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "uniform plasma :U:N, 32x32x32 cells, reference CPU path (stock simulateSteps: mover + export/import; "
                               "STEP 1 deposit is commented out in the reference, simulator.h:189-200)",
                   "particles": n, "grid": [32, 32, 32], "deposit": False},
        "cpu_baseline": {k: res[k] for k in ("value", "cores", "kind", "sample")} | {"unit": UNIT},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.time() - t0,
    }
    print(json.dumps(line))
    return 0


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

    nz = args.grid
    if nz % world:
        raise SystemExit("grid z extent %d not divisible by %d ranks" % (nz, world))
    lz = nz // world
    if args.workload == "dipole":
        # every rank builds the whole universe (host mirror generator) and keeps its z-slab; the total is fixed: strong scaling
        props, counts, aos = capi.dipole_universe(os.path.join(ROOT, "tests", "golden", "inputs", "medium.inp"), args.grid, 8)
        n_total = int(counts.sum())
        if world > 1:
            counts, aos = capi.slab_of(counts, aos, nz * nz, nz, rank, world)
        n_local = int(counts.sum())
        host = torch.from_numpy(np.ascontiguousarray(aos)).pin_memory()
        del aos
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

    flags = 0 if args.no_deposit else ipic.PROJECT_CURRENT
    ctx = ipic.Context(props, device=local, rank=rank, nranks=world)
    if world > 1:
        uid = [ipic.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0])
    if host is None:
        ctx.generate_uniform(n_total)
    else:
        ctx.upload_cells_ptr(counts, host.data_ptr())
    particle_fields = ctx.particle_fields()

    # ---- device-resident throughput: W warm-up steps, then exactly K timed steps -------------------
    sampler = ClockSampler(local)
    ctx.step(args.warmup, flags)
    ctx.profile_reset()
    barrier()
    sampler.start()
    t_first, t_rest = ctx.step(args.steps, flags)          # CUDA events on the context's stream bracket the K steps
    barrier()
    clocks = sampler.stop()
    mover_ms, import_ms, launches = ctx.profile_get()
    phases = [p / args.steps for p in ctx.profile_phases()]
    seconds = max_over_ranks(t_first + t_rest)
    alive = ctx.total_particles()
    # :U: universe: nothing is absorbed (planetRadius = 0).  Dipole: the planet absorbs particles; the metric follows the
    # reference and counts the INITIAL population per step (benchmark.h:50 / ipic3d.cxx:96-99)
    pushes = args.steps * n_total
    value = pushes / seconds

    hbm_peak, peak_src = measured_peaks()
    mover_s = max_over_ranks(mover_ms / 1e3) / args.steps
    bpp = BYTES_PER_PUSH[particle_fields]
    achieved = n_local * bpp / mover_s / 1e9
    kernel = ("k_move_ws (warp-specialised: 4 push warps + 1 deposit warp per cell; fused mover + deposit + classify/compaction)"
              if (not args.no_deposit and args.workload == "uniform") else
              "k_move (fused mover%s + classify/compaction)" % ("" if args.no_deposit else " + deposit"))
    roofline = {"bound": "hbm", "kernel": kernel, "achieved": achieved,
                "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None, "peak_source": peak_src,
                "bytes_per_push": bpp, "pushes_per_launch": n_local, "launch_ms": mover_s * 1e3,
                "at_survey_112B_per_push": {"achieved": n_local * SURVEY_BYTES_PER_PUSH / mover_s / 1e9,
                                            "frac": n_local * SURVEY_BYTES_PER_PUSH / mover_s / 1e9 / hbm_peak,
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
            roofline["traffic"] = ent.get("bytes_per_push") * n_local if ent.get("bytes_per_push") else None
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
        e2e = {"value": pushes / t_call, "unit": UNIT,
               "h2d_bytes_per_step": n_local * 64 / args.steps, "d2h_bytes_per_step": int(cnt_out.sum()) * 64 / args.steps,
               "call": "simulateSteps(%d, universe) semantics: pinned host AoS -> H2D, %d steps, D2H of all cells" % (args.steps, args.steps),
               "seconds": t_call,
               "roundtrip_every_step": {"value": rt_steps * n_total / t_rt, "steps": rt_steps,
                                        "h2d_bytes_per_step": n_local * 64, "d2h_bytes_per_step": n_local * 64}}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        res = time_reference(args.cpu_particles, args.cpu_steps, 1, deposit=False)
        if res:
            cpu = {k: res[k] for k in ("value", "cores", "kind", "sample")}
            cpu["unit"] = UNIT

    ctx.close()
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": seconds / args.steps * 1e3, "higher_is_better": True, "scaling": "weak" if args.workload == "uniform" else "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": ("uniform plasma :U:N (reference generator, bit-identical), %dx%dx%d cells, %d particles per GPU, z-slabs"
                                    % (args.grid, args.grid, args.grid, args.particles_per_gpu)) if args.workload == "uniform" else
                                   ("SECONDARY: Earth-dipole .inp set-up of the reference (medium.inp physics, host mirror generator), %dx%dx%d cells, "
                                    "8^3 particles per cell" % (args.grid, args.grid, args.grid)),
                       "particles": n_total, "particles_per_gpu": n_total // world, "grid": [args.grid] * 3,
                       "dt": float(props.dt), "deposit": not args.no_deposit, "parallelism": "zslab%d" % world,
                       "particle_fields_in_hbm": particle_fields,
                       "l2": "inputs (%.2f GB of particle state per GPU) are larger than L2; no flush needed" % (n_local * 64 / 1e9)},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "particles_alive_after": alive if world == 1 else None,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
