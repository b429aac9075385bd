"""The reference-side binding (host/ref_binding/ipic3d/app/simulator.h: the replacement for the reference's simulator.h that
INTEGRATION.md section 2 describes) compiled with the reference's UNMODIFIED driver, parser, initCells and .out writer
(oracle/Makefile -> oracle/_ref/ipic3d_ref_b200; built where /root/reference exists, travels with oracle/_ref).

CPU checks: the binary exists, refuses to run without a CUDA device (no CPU fallback behind the binding), and -- with a
test double of the C ABI on the CPU oracle in front of the real library -- drives the reference's own data model
correctly: `ipic3d test.inp` through the binding writes the same test.out as the reference built here."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
BOUND = os.path.join(REF, "ipic3d_ref_b200")
DECK = os.path.join(ROOT, "tests", "golden", "inputs", "test.inp")

pytestmark = pytest.mark.skipif(not os.path.exists(BOUND), reason="oracle/_ref/ipic3d_ref_b200 not built (reference sources absent at build time)")


def run(exe, cwd, env=None, args=(DECK,)):
    return subprocess.run([exe, *args], cwd=cwd, env=env, capture_output=True, text=True, timeout=900)


def build_mock(tmp_path):
    mock_dir = tmp_path / "mock"
    mock_dir.mkdir()
    subprocess.check_call(["gcc", "-O2", "-std=gnu11", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle"),
                           os.path.join(ROOT, "tests", "mock_abi", "mock_ipic3d_b200.c"), "-L" + os.path.join(ROOT, "oracle"), "-lipic3d_oracle",
                           "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-o", str(mock_dir / "libipic3d_b200.so")])
    return dict(os.environ, LD_LIBRARY_PATH=str(mock_dir) + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))


def test_binding_links_the_product_library_and_has_no_cpu_path(tmp_path):
    needed = subprocess.run(["readelf", "-d", BOUND], capture_output=True, text=True).stdout
    assert "libipic3d_b200.so" in needed and "oracle" not in needed
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    out = run(BOUND, tmp_path)
    assert out.returncode != 0 and "ipic_create failed" in out.stderr, out.stderr[-500:]
    assert not (tmp_path / "test.out").exists()


def test_binding_drives_the_reference_data_model(tmp_path):
    """Reference main + binding + test double of the C ABI (CPU oracle): same test.out as the reference itself."""
    a, b = tmp_path / "bound", tmp_path / "ref"
    a.mkdir(); b.mkdir()
    env = build_mock(tmp_path)
    got = run(BOUND, a, env)
    assert got.returncode == 0, got.stderr[-1000:]
    want = run(os.path.join(REF, "ipic3d_ref"), b)
    assert want.returncode == 0
    text = (a / "test.out").read_text()
    assert text == (b / "test.out").read_text()
    assert "Total: 257060" in text                      # the reference's result for this deck at this commit


def test_reference_simulation_tests_through_the_binding(tmp_path):
    """The reference's simulation_test.cc with the reference's own headers and the binding (oracle/_ref/ut_binding): its
    seven cases (Boris straight line, Larmor 1000 steps, ExB drift 5000 steps, migration / reflection on 3^3 and 4^3 grids, all
    through simulateSteps / simulateStep on universes edited between the calls) pass against the test double of the C ABI."""
    exe = os.path.join(REF, "ut_binding", "simulation_test")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ut_binding not built")
    out = subprocess.run([exe], cwd=ROOT, env=build_mock(tmp_path), capture_output=True, text=True, timeout=900)
    text = out.stdout + out.stderr
    assert out.returncode == 0 and "7 cases, 0 failed" in text, text[-2000:]


def test_binding_under_the_benchmark_driver(tmp_path):
    """`ipic3d :U:N` (processUniverse: two simulateSteps calls on one universe, hdr/benchmark.h:38,43) through the binding and
    the test double: the particle positions the reference dumps before and after (DUMP_PARTICLE_POSITION) are identical."""
    a, b = tmp_path / "bound", tmp_path / "ref"
    a.mkdir(); b.mkdir()
    env = dict(os.environ, DUMP_PARTICLE_POSITION="1")
    got = run(BOUND, a, dict(build_mock(tmp_path), DUMP_PARTICLE_POSITION="1"), args=(":U:60000",))
    assert got.returncode == 0 and "Throughput:" in got.stdout, got.stderr[-1000:]
    want = run(os.path.join(REF, "ipic3d_ref"), b, env, args=(":U:60000",))
    assert want.returncode == 0
    for name in ("t_begin.txt", "t_end.txt"):
        assert (a / name).read_text() == (b / name).read_text(), name
    assert (a / "t_begin.txt").read_text() != (a / "t_end.txt").read_text()
