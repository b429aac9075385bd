"""Small cases for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family on tiny universes --
the mbarrier / TMA ring of the movers, the warp-specialised mover, the chunked mover's look-back, the outbox sort,
import, face packing, deposit gather, field solver.  usage: compute-sanitizer --tool <tool> python scripts/sanitize_case.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import allscale_ipic3d_b200 as ipic
from allscale_ipic3d_b200 import capi

props = capi.Props.make(6, cellWidth=(10.0, 10.0, 10.0), dt=0.15 * 40)
FULL = ipic.PROJECT_CURRENT | ipic.UPDATE_BOUNDARIES | ipic.SOLVER_FORWARD
with ipic.Context(props) as c:
    c.generate_uniform(60000)                       # ~280 particles per cell: several tiles per cell, > 32 leavers per cell (sort pass)
    c.step(2, 0)
    c.step(2, ipic.PROJECT_CURRENT)                 # k_move_ws
    c.step(1, FULL)
    c.step(1, ipic.MOVER_INTERPOLATED | FULL)
    counts, aos = c.download_cells()
    print("single ok", int(counts.sum()), c.count_misplaced())
os.environ["IPIC_CHUNKED"] = "1"
with ipic.Context(props) as c:
    c.upload_cells(counts, aos)
    c.step(2, ipic.PROJECT_CURRENT)
    print("chunked ok", c.total_particles())
del os.environ["IPIC_CHUNKED"]
with ipic.Group(props, [0, 0, 0], planes=[1, 3, 2]) as g:
    g.upload_cells(counts, aos)
    g.step(2, ipic.PROJECT_CURRENT)
    g.step(1, FULL)
    print("slabs ok", int(g.cell_counts().sum()))
deck = os.path.join(ROOT, "tests", "golden", "inputs", "test.inp")
with capi.Deck(deck) as d, ipic.Context(d.props) as c:
    d.generate_on_device(c)
    c.step(3, ipic.PROJECT_CURRENT)                 # dipole mover (k_move + k_sort_outbox)
    print("deck ok", c.total_particles())
