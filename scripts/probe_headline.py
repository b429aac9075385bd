"""GPU probe (no torch, a few seconds): device time per step of the headline :U:32000000 universe, mover + deposit + migration,
created in HBM by the device generator -- to see at a glance whether a kernel change moved the headline path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import allscale_ipic3d_b200 as ipic
from allscale_ipic3d_b200 import capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32_000_000
props = capi.uniform_benchmark_props(32)
with ipic.Context(props) as c:
    c.generate_uniform(n)
    c.step(5, ipic.PROJECT_CURRENT)
    c.profile_reset()
    a, b = c.step(20, ipic.PROJECT_CURRENT)
    mv, im, launches = c.profile_get()
    print(":U:%d  %.4f ms/step (mover %.4f, import %.4f)  %.4g pushes/s  launches %d  alive %d" % (
        n, (a + b) / 20 * 1e3, mv / 20, im / 20, 20 * n / (a + b), launches, c.total_particles()), flush=True)
