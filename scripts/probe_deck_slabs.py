"""GPU probe: a deck's universe as uneven z-slabs (cut by particle count) of ONE process -- all slabs on device 0 unless devices
are given -- created on the device, stepped, re-layouts counted.  usage: python scripts/probe_deck_slabs.py <deck> <slabs> [steps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import allscale_ipic3d_b200 as ipic
from allscale_ipic3d_b200 import capi

deck_path, slabs = sys.argv[1], int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 25
with capi.Deck(deck_path) as d:
    planes = capi.balanced_planes(d.plane_population(), slabs)
    with ipic.Group(d.props, [0] * slabs, planes=planes) as g:
        for m in g.members:
            d.generate_on_device(m)
        t0 = time.perf_counter()
        a, b = g.step(steps, ipic.PROJECT_CURRENT)
        print("planes %s: %d steps, device %.3f ms/step, wall %.3f ms/step, relayouts per slab %s, particles %d"
              % (planes, steps, (a + b) / steps * 1e3, (time.perf_counter() - t0) / steps * 1e3, [m.relayout_count() for m in g.members], int(g.cell_counts().sum())))
