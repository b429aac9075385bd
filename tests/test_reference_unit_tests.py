"""The reference's OWN unit-test sources (code/app/test/*.cc), unmodified, compiled against the host mirror (host/include:
the drop-in data model, parser and simulateSteps) instead of the reference's headers -- the switch a user of the reference
would make.  The binaries are built where /root/reference exists (oracle/Makefile, target mirrortests) and travel with
oracle/_ref; nothing here reads /root/reference at run time.

  init / parameters / particle / universe   pure host code: run in the CPU suite
  cell / field / transfer_buffer / ziggurat pure host code too: generators, placement check, .out writer, and the node-level
                                            KATs (projectToDensityField 1.0952, curls, interpolation, ghost update, forward and
                                            leapfrog solver) on the mirror's host-callable scalar versions
  simulation (7 cases)                      simulateSteps / simulateStep on universes the tests build and mutate by hand:
                                            Boris straight line, Larmor 1000 steps, ExB drift 5000 steps, migration and
                                            reflection on 3^3 and 4^3 grids -- every step runs on the GPU (-m gpu)
"""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UT = os.path.join(ROOT, "oracle", "_ref", "ut_mirror")


def run_cases(name, timeout=900, cwd=ROOT):
    exe = os.path.join(UT, name + "_test")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ut_mirror not built (reference sources absent at build time)")
    out = subprocess.run([exe], cwd=cwd, capture_output=True, text=True, timeout=timeout)
    text = out.stdout + out.stderr
    m = re.search(r"(\d+) cases, (\d+) failed", text)
    assert m, text[-2000:]
    assert out.returncode == 0 and int(m.group(2)) == 0, text[-3000:]
    return int(m.group(1)), text


@pytest.mark.parametrize("name,cases", [("init", 2), ("parameters", 2), ("particle", 2), ("universe", 7)])
def test_reference_host_unit_tests_pass_against_the_mirror(name, cases):
    n, _ = run_cases(name)
    assert n == cases


@pytest.mark.parametrize("name,cases", [("cell", 9), ("field", 11), ("transfer_buffer", 1), ("ziggurat_normal_distribution", 4)])
def test_reference_cell_and_field_unit_tests_pass_against_the_mirror(name, cases, tmp_path):
    """cell_test.cc writes its `.out` entry into the working directory and reads decks from PATH_TO_INPUTS
    (tests/golden/inputs, relative): run in a scratch directory that sees tests/ through a link."""
    os.symlink(os.path.join(ROOT, "tests"), tmp_path / "tests")
    n, text = run_cases(name, cwd=str(tmp_path))
    assert n == cases, text[-2000:]
    if name == "cell":
        assert "OK ] Cell.projectToDensityField" in text and "OK ] Cell.TestCellOutput" in text
    if name == "field":
        for case in ("curlB", "curlE", "interpC2N", "interpN2C", "TrilinearInterpolation", "FieldSolverStatic", "updateFieldsOnBoundaries",
                     "solveFieldForward", "solveFieldForwardTrajectory", "solveFieldLeapfrog"):
            assert "OK ] Field." + case in text, case


@pytest.mark.gpu
def test_reference_simulation_tests_pass_on_the_gpu():
    """simulation_test.cc: every TEST drives simulateSteps / simulateStep of the mirror, i.e. the CUDA kernels."""
    n, text = run_cases("simulation")
    assert n == 7, text[-2000:]
    for case in ("Simulation.SingleParticleBorisMover", "Simulation.SingleParticleBorisMoverLarmorRadius", "Simulation.SingleParticleLarmorRadius",
                 "Simulation.SingleParticleBorisMoverExBdrift", "SimulationTest.ParticleMigrationOneDirectionOneParticle",
                 "SimulationTest.ParticleMigrationTwoDirectionsOneParticle", "SimulationTest.ParticleMigrationManyParticles"):
        assert "OK ] " + case in text, case
