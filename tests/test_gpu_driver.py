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
