"""GPU probe: the big-outbox path of the import (k_big_hist / k_big_scan / k_big_copy) on the reference's :B: beam load,
where one cell holds most of the plasma and hands it to ONE neighbour every step.  The same universe is stepped twice in
this process, with the path off (IPIC_BIG_IMPORT=0: the target cell's one warp scans and copies the outbox) and on; the
two final states must be identical bit for bit (the old path is the one the parity tests pin to the reference), and the
device times per step are printed.   usage: python scripts/probe_big_import.py [N] [kinds]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import harness
import allscale_ipic3d_b200 as ipic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
kinds = sys.argv[2] if len(sys.argv) > 2 else "B"
for kind in kinds:
    t0 = time.perf_counter()
    r = harness.Ref.benchmark(kind, n, threaded=True)
    counts, aos = r.counts(), r.particles()
    print("%s: N=%d universe built in %.1f s, fullest cell %d" % (kind, n, time.perf_counter() - t0, counts.max()), flush=True)
    final = {}
    for mode in ("0", "1"):
        os.environ["IPIC_BIG_IMPORT"] = mode
        with ipic.Context(r.props) as c:
            c.upload_cells(counts, aos)
            c.step(2)
            c.profile_reset()
            a, b = c.step(5)
            mv, im, launches = c.profile_get()
            cc, pp = c.download_cells()
            final[mode] = (cc, pp)
            print("%s big-outbox path %s: device %.3f ms/step (mover %.3f, import %.3f)  %.3g pushes/s  relayouts %d  cells max %d nonempty %d launches %d" % (
                kind, "ON " if mode == "1" else "off", (a + b) * 1e3 / 5, mv / 5, im / 5, 5 * n / (a + b), c.relayout_count(), cc.max(), (cc > 0).sum(), launches), flush=True)
    same = np.array_equal(final["0"][0], final["1"][0]) and harness.bits_equal(final["0"][1], final["1"][1])
    print("%s: final states of the two paths identical bit for bit: %s" % (kind, same), flush=True)
    if not same:
        sys.exit(1)
