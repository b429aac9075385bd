"""GPU probe: step time of the dipole workload as ONE context vs an in-process group of z-slabs (devices may repeat).
usage: python scripts/probe_slabs.py [grid] [devices, e.g. 0,0 or 0,1]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import allscale_ipic3d_b200 as ipic
from allscale_ipic3d_b200 import capi

grid = int(sys.argv[1]) if len(sys.argv) > 1 else 32
devices = [int(d) for d in (sys.argv[2] if len(sys.argv) > 2 else "0,0").split(",")]
props, counts, aos = capi.dipole_universe(os.path.join(ROOT, "tests", "golden", "inputs", "medium.inp"), grid, 8)
print("dipole universe %d^3, %d particles" % (grid, len(aos)), flush=True)
for flags, name in ((0, "mover+migration"), (ipic.PROJECT_CURRENT, "with deposit")):
    with ipic.Context(props) as c:
        c.upload_cells(counts, aos)
        c.step(12, flags)
        t0 = time.perf_counter(); a, b = c.step(10, flags); wall = time.perf_counter() - t0
        print("%-16s one context : device %.3f ms/step, wall %.3f ms/step, relayouts %d" % (name, (a + b) * 100, wall * 100, c.relayout_count()), flush=True)
    with ipic.Group(props, devices) as g:
        g.upload_cells(counts, aos)
        g.step(12, flags)
        for m in g.members:
            m.profile_reset()
        t0 = time.perf_counter(); a, b = g.step(10, flags); wall = time.perf_counter() - t0
        print("%-16s %d slabs %s: device %.3f ms/step, wall %.3f ms/step, relayouts %d" % (name, len(devices), devices, (a + b) * 100, wall * 100, g.relayout_count()), flush=True)
        for r, m in enumerate(g.members):
            print("      slab %d phases ms/step (movers, imports, wait exchange, wait J+fields): %s" % (r, ["%.3f" % (p / 10) for p in m.profile_phases()]), flush=True)
