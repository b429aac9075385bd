"""GPU probe: cost of the field phase (ghost update + forward / leapfrog solver) per step on the 32^3 benchmark grid:
full PIC step vs the same step without solver, for the current IPIC_SOLVER_MODE setting (coop | planes | one | cluster)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import allscale_ipic3d_b200 as ipic
from allscale_ipic3d_b200 import capi
props = capi.uniform_benchmark_props(32)
with ipic.Context(props) as c:
    c.generate_uniform(4_000_000)
    out = []
    for name, flags in (("projection only", ipic.PROJECT_CURRENT), ("forward", ipic.PROJECT_CURRENT | ipic.UPDATE_BOUNDARIES | ipic.SOLVER_FORWARD),
                        ("leapfrog", ipic.PROJECT_CURRENT | ipic.UPDATE_BOUNDARIES | ipic.SOLVER_LEAPFROG)):
        c.step(3, flags)
        a, b = c.step(20, flags)
        out.append("%s %.3f ms/step" % (name, (a + b) / 20 * 1e3))
    print("IPIC_SOLVER_MODE=%s: %s" % (os.environ.get("IPIC_SOLVER_MODE", "unset (coop)"), "; ".join(out)))
