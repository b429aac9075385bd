"""GPU probe: device throughput of the reference's :C: / :E: / :B: benchmark universes (mover + import, resident state).
usage: python scripts/probe_loads.py [N]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import harness
import allscale_ipic3d_b200 as ipic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
for kind in "UCEB":
    r = harness.Ref.benchmark(kind, n, threaded=True)
    counts, aos = r.counts(), r.particles()
    with ipic.Context(r.props) as c:
        c.upload_cells(counts, aos)
        c.step(2)
        c.profile_reset()
        t0 = time.perf_counter()
        a, b = c.step(5)
        wall = time.perf_counter() - t0
        mv, im, launches = c.profile_get()
        cc = c.cell_counts()
        print("%s: N=%d  device %.3f ms/step (mover %.3f, import %.3f)  wall %.3f ms/step  %.3g pushes/s  relayouts %d  cells max %d nonempty %d launches %d" % (
            kind, n, (a + b) * 1e3 / 5, mv / 5, im / 5, wall * 1e3 / 5, 5 * n / (a + b), c.relayout_count(), cc.max(), (cc > 0).sum(), launches), flush=True)
