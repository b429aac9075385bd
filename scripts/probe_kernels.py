"""GPU probe: launches every kernel family of the library once at a realistic size, so that one `ncu -k regex:<kernel> -c 1`
per family yields a capture of each (scripts/gpu_r2.sh `ncuall`).  usage: python scripts/probe_kernels.py [particles]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import allscale_ipic3d_b200 as ipic
from allscale_ipic3d_b200 import capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
props = capi.uniform_benchmark_props(32)
FULL_FWD = ipic.PROJECT_CURRENT | ipic.UPDATE_BOUNDARIES | ipic.SOLVER_FORWARD
FULL_LF = ipic.PROJECT_CURRENT | ipic.UPDATE_BOUNDARIES | ipic.SOLVER_LEAPFROG

with ipic.Context(props) as c:
    c.generate_uniform(n)                                   # k_init_uniform
    c.step(2, FULL_FWD)                                     # k_move_ws, k_gather_current, k_boundaries, k_solver_plane<1>, k_import, k_commit_flag
    c.step(2, FULL_LF)                                      # k_solver_plane<0>
    c.step(2, ipic.MOVER_INTERPOLATED | FULL_FWD)           # k_move<..INTERP..>, deposit-only k_move
    c.step(2, 0)                                            # k_move<6,0,0,1,1>
    print("diagnostics", c.diagnostics())                   # k_cell_energy, k_field_energy, k_sum_pairs
    print("misplaced", c.count_misplaced())                 # k_count_misplaced
    counts, aos = c.download_cells()                        # k_gather_aos
    c.set_uniform_species(False)
    c.upload_cells(counts, aos)                             # k_scatter_aos (8 fields)
    c.step(2, ipic.PROJECT_CURRENT)                         # k_move_ws<8,1>
    c.set_uniform_species(True)
    c.upload_cells(counts, aos)                             # k_check_species, k_scatter_aos

# a cell-segment overflow: re-layout (k_copy_heads) and redo of the import
with ipic.Context(props) as c:
    conv = aos.copy()
    conv[:, 3:6] = 0.0
    conv[:, 3] = np.where(conv[:, 0] < 160.0, 60.0, -60.0)     # everybody converges on the two middle x-planes
    c.set_headroom(0.05, 32)
    c.upload_cells(counts, conv)
    c.step(12, 0)
    print("relayouts", c.relayout_count())

os.environ["IPIC_CHUNKED"] = "1"
with ipic.Context(props) as c:
    c.upload_cells(counts, aos)
    c.step(2, ipic.PROJECT_CURRENT)                         # k_move_chunk
del os.environ["IPIC_CHUNKED"]

with ipic.Group(props, [0, 0]) as g:                        # two slabs on one GPU: k_face_count / scan / copy, k_import<true>
    g.upload_cells(counts, aos)
    g.step(3, ipic.PROJECT_CURRENT)

deck = os.path.join(ROOT, "tests", "golden", "inputs", "tiny.inp")
with capi.Deck(deck) as d, ipic.Context(d.props) as c:      # k_init_deck_positions, k_init_normal_speed (2 M particles), dipole k_move, k_sort_outbox
    d.generate_on_device(c)
    c.step(2, ipic.PROJECT_CURRENT)
    print("deck", c.total_particles())

tp = capi.Props.make(16, cellWidth=(7.97267e6 / 2,) * 3, dt=0.01, speedOfLight=299792458.0, planetRadius=6378137.0,
                     origin=(-8 * 7.97267e6 / 2,) * 3, externalMagneticField=(0.0, 0.0, 3.07e-5))
rng = np.random.default_rng(1)
tr = np.zeros((1 << 20, 8))
tr[:, 0:3] = rng.uniform(-3e7, 3e7, (len(tr), 3))
tr[:, 3:6] = rng.normal(0, 3e7, (len(tr), 3))
tr[:, 6] = 1.602176565e-19
tr[:, 7] = 1.602176565e-19 / 1.672621777e-27
cnt, sec = capi.trace(tp, tr, 50, 10)                       # k_trace
print("trace", cnt.sum(), sec)
