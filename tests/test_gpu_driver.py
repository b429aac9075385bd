"""The drop-in driver `host/bin/ipic3d_b200` (host C++ mirror + C ABI + CUDA kernels) against the reference's own
`ipic3d` built here (oracle/_ref/ipic3d_ref): same command line, same `<base>.out`, compared the way the
reference's scripts/verify_output.sh does (diff of the sorted files)."""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "host", "bin", "ipic3d_b200")
REF = os.path.join(ROOT, "oracle", "_ref", "ipic3d_ref")
DECK = os.path.join(ROOT, "tests", "golden", "inputs", "test.inp")


def run(binary, arg, cwd, env=None):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([binary, arg], cwd=cwd, capture_output=True, text=True, timeout=600, env=e)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    return out.stdout


@pytest.mark.parametrize("devices", [None, "0,0", "0,0,0,0"])
def test_test_inp_out_file_matches_reference(tmp_path, devices):
    """`ipic3d test.inp` (30 steps): sorted .out identical to the reference's, on one slab and on 2 / 4 in-process slabs."""
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref not built")
    a, b = tmp_path / "ours", tmp_path / "ref"
    a.mkdir(); b.mkdir()
    text = run(DRIVER, DECK, str(a), {"IPIC3D_DEVICES": devices} if devices else None)
    run(REF, DECK, str(b))
    ours = sorted((a / "test.out").read_text().split("\n"))
    want = sorted((b / "test.out").read_text().split("\n"))
    assert ours == want
    assert re.search(r"Throughput: [0-9.e+]+ particles/s", text)


def test_resident_and_host_universe_flows_agree(tmp_path):
    """`ipic3d test.inp` creates, steps and counts its universe in HBM by default (device generator, resident mode, `.out` from
    the device's per-cell counters); IPIC3D_RESIDENT=0 is the reference's flow (host universe, upload, steps, download of all
    particles, host writer).  Same `.out`, byte for byte."""
    a, b = tmp_path / "resident", tmp_path / "host"
    a.mkdir(); b.mkdir()
    run(DRIVER, DECK, str(a))
    run(DRIVER, DECK, str(b), {"IPIC3D_RESIDENT": "0"})
    assert (a / "test.out").read_text() == (b / "test.out").read_text()
    t1 = run(DRIVER, ":U:1000000", str(a))
    t2 = run(DRIVER, ":U:1000000", str(b), {"IPIC3D_RESIDENT": "0"})
    assert "Throughput:" in t1 and "Throughput:" in t2


def test_missing_file_and_usage(tmp_path):
    """Error behaviour of the reference's main (ipic3d.cxx:18-21) and parser (parameters.h:148-151)."""
    r = subprocess.run([DRIVER], capture_output=True, text=True)
    assert r.returncode != 0 and "Usage: ./ipic3d <config-file>" in r.stdout
    r = subprocess.run([DRIVER, str(tmp_path / "nope.inp")], capture_output=True, text=True)
    assert r.returncode != 0 and "File not found" in r.stderr


def test_uniform_benchmark_runs(tmp_path):
    """`ipic3d :U:2000000`: the benchmark driver's console protocol (scripts grep `Throughput:`)."""
    text = run(DRIVER, ":U:2000000", str(tmp_path))
    assert "particles / cell: 61 (+1)" in text
    m = re.search(r"Throughput: ([0-9.e+]+) particles/s", text)
    assert m and float(m.group(1)) > 1e7


REF_DBG = os.path.join(ROOT, "oracle", "_ref", "ipic3d_ref_dbg")


@pytest.mark.parametrize("devices", [None, "0,0"])
def test_conserved_quantities_log_and_density_snapshots(tmp_path, devices):
    """The reference's ENABLE_DEBUG_OUTPUT products (simulator.h:64-115, :160-187) for test.inp (FieldOutputCycle 1,
    ParticlesOutputCycle 10): `testConservedQuantities.out` (cycle, total momentum, E energy, B energy, total KE; the
    sums come from ipic_group_diagnostics on the device) agrees with the reference's file to its 6 printed digits, and
    the `result_<stamp>.csv.NNNNNN` per-cell density snapshots (counts only) are identical."""
    if not os.path.exists(REF_DBG):
        pytest.skip("oracle/_ref/ipic3d_ref_dbg not built")
    a, b = tmp_path / "ours", tmp_path / "ref"
    a.mkdir(); b.mkdir()
    env = {"IPIC3D_DEBUG_OUTPUT": "1"}
    if devices:
        env["IPIC3D_DEVICES"] = devices
    run(DRIVER, DECK, str(a), env)
    run(REF_DBG, DECK, str(b))
    ours = (a / "testConservedQuantities.out").read_text().splitlines()
    want = (b / "testConservedQuantities.out").read_text().splitlines()
    assert ours[0] == want[0] and len(ours) == len(want) == 31
    for lo, lw in zip(ours[1:], want[1:]):
        fo, fw = lo.split("\t"), lw.split("\t")
        assert fo[0] == fw[0] and len(fo) == len(fw) == 5
        for x, y in zip(fo[1:], fw[1:]):
            assert abs(float(x) - float(y)) <= 2e-6 * abs(float(y)), (lo, lw)      # one unit in the 6th printed digit
    snaps_o = sorted(p.name.split(".")[-1] for p in a.glob("result_*.csv.*"))
    snaps_w = sorted(p.name.split(".")[-1] for p in b.glob("result_*.csv.*"))
    assert snaps_o == snaps_w == ["000000", "000001", "000002", "000003"]
    for s in snaps_o:
        assert next(a.glob("result_*.csv." + s)).read_text() == next(b.glob("result_*.csv." + s)).read_text()
    assert sorted((a / "test.out").read_text().split("\n")) == sorted((b / "test.out").read_text().split("\n"))
