"""Pins the plain-C oracle bit-for-bit against the reference's own code (oracle/_ref: the unmodified reference
headers compiled against the AllScale shim).  Skipped where oracle/_ref has not been built."""
import os

import numpy as np
import pytest

import harness
from harness import Oracle, Ref, Props, bits_equal

pytestmark = pytest.mark.skipif(not Ref.available(), reason="oracle/_ref not built (needs /root/reference)")


def same_state(r, o):
    return np.array_equal(r.counts(), o.counts()) and bits_equal(r.particles(), o.particles())


def test_test_inp_every_step(golden_dir):
    """inputs/test.inp, all 30 steps: counts and every particle bit-identical; matches the committed golden run."""
    g = np.load(os.path.join(golden_dir, "test_inp.npz"))
    r = Ref.from_inp(os.path.join(golden_dir, "inputs", "test.inp"))
    assert r.ncycles == 30 and r.num_particles == 262144
    o = Oracle(r.props)
    o.set_cells(r.counts(), r.particles())
    for s in range(1, 31):
        r.step(1); o.step(1)
        assert same_state(r, o), "step %d" % s
        assert np.array_equal(o.counts(), g["counts"][s])
    assert r.total() == 257060          # what this commit of the reference produces


def test_tiny_inp_prefix(golden_dir):
    """inputs/tiny.inp: first 10 steps bit-identical (the full 100 are covered by the golden counts on the GPU side)."""
    g = np.load(os.path.join(golden_dir, "tiny_inp.npz"))
    r = Ref.from_inp(os.path.join(golden_dir, "inputs", "tiny.inp"), threaded=Ref.available(True))
    o = Oracle(r.props)
    o.set_cells(r.counts(), r.particles())
    r.step(10); o.step(10)
    assert same_state(r, o)
    assert np.array_equal(o.counts(), g["counts"][1])


def test_threaded_reference_equals_sequential():
    if not Ref.available(True):
        pytest.skip("threaded build missing")
    a, b = Ref.uniform(200000), Ref.uniform(200000, threaded=True)
    assert bits_equal(a.particles(), b.particles())
    a.step(5); b.step(5)
    assert np.array_equal(a.counts(), b.counts()) and bits_equal(a.particles(), b.particles())


@pytest.mark.parametrize("n,particles,dtscale", [(1, 200, 40), (2, 2000, 40), (3, 4000, 40), (5, 20000, 1)])
def test_uniform_generator_grids(n, particles, dtscale):
    """:U:-style plasma on small grids, including the degenerate 1- and 2-cell periodic address wrap."""
    props = Props.make(n, cellWidth=(10.0, 10.0, 10.0), dt=0.15 * dtscale)
    r = Ref.uniform_grid(props, particles)
    o = Oracle(props)
    o.set_cells(r.counts(), r.particles())
    for s in range(20):
        r.step(1); o.step(1)
        assert same_state(r, o), "step %d" % s


@pytest.mark.parametrize("flags", [
    harness.MOVER_INTERPOLATED,
    harness.PROJECT_CURRENT,
    harness.MOVER_INTERPOLATED | harness.PROJECT_CURRENT | harness.UPDATE_BOUNDARIES | harness.SOLVER_FORWARD,
    harness.MOVER_INTERPOLATED | harness.PROJECT_CURRENT | harness.UPDATE_BOUNDARIES | harness.SOLVER_LEAPFROG,
    harness.UPDATE_BOUNDARIES | harness.SOLVER_FORWARD | harness.SKIP_MIGRATION,
])
def test_full_step_modes(golden_dir, flags):
    """Deposit, ghost update, both solvers and the interpolating mover: all state bit-identical after 3 steps."""
    g = np.load(os.path.join(golden_dir, "small_fields.npz"))
    import ctypes
    props = Props(); ctypes.memmove(ctypes.byref(props), g["props"].tobytes(), ctypes.sizeof(props))
    r, o = Ref.empty(props), Oracle(props)
    for u in (r, o):
        u.set_cells(g["counts0"], g["p0"])
        u.field(g["field0"]); u.bcfield(g["bc0"])
    r.step(3, flags); o.step(3, flags)
    assert same_state(r, o)
    assert bits_equal(r.field(), o.field()) and bits_equal(r.bcfield(), o.bcfield()) and bits_equal(r.density(), o.density())


def test_field_init_bit_exact():
    Re = 6378137.0
    props = Props.make(6, cellWidth=(Re, Re, Re), planetRadius=Re, origin=(-3 * Re,) * 3, externalMagneticField=(0, 0, 3.07e-5))
    r, o = Ref.empty(props), Oracle(props)
    r.init_fields((0.02, 0.01, -0.03), (1e-5, 2e-5, 1e-4)); o.init_fields((0.02, 0.01, -0.03), (1e-5, 2e-5, 1e-4))
    assert bits_equal(r.field(), o.field()) and bits_equal(r.bcfield(), o.bcfield())


def test_diagnostics_match_reference(golden_dir):
    """The four columns of ConservedQuantities.out (writeOutputData, simulator.h:71-91): oracle == reference bit for bit,
    on a dipole universe with initialised fields (test.inp after 2 steps) and on the small full-PIC fixture."""
    r = Ref.from_inp(os.path.join(golden_dir, "inputs", "test.inp"))
    r.init_fields((0.0, 0.0, 0.0), (0.0, 0.0, 0.0))
    r.step(2)
    o = Oracle(r.props)
    o.set_cells(r.counts(), r.particles())
    o.field(r.field())
    d = r.diagnostics()
    assert bits_equal(d, o.diagnostics())
    assert d[0] > 0 and d[3] > 0 and d[2] > 0 and d[1] == 0.0
    g = np.load(os.path.join(golden_dir, "small_fields.npz"))
    import ctypes
    props = Props(); ctypes.memmove(ctypes.byref(props), g["props"].tobytes(), ctypes.sizeof(props))
    r, o = Ref.empty(props), Oracle(props)
    for u in (r, o):
        u.set_cells(g["counts0"], g["p0"])
        u.field(g["field0"]); u.bcfield(g["bc0"])
    flags = harness.MOVER_INTERPOLATED | harness.PROJECT_CURRENT | harness.UPDATE_BOUNDARIES | harness.SOLVER_FORWARD
    r.step(2, flags); o.step(2, flags)
    assert bits_equal(r.diagnostics(), o.diagnostics())
    assert all(v > 0 for v in r.diagnostics())


def test_unit_functions_random():
    rng = np.random.default_rng(7)
    for _ in range(200):
        p = rng.normal(size=8); E = rng.normal(size=3); B = rng.normal(size=3); dt = float(rng.random())
        assert bits_equal(Ref.update_velocity(p, E, B, dt), Oracle.update_velocity(p, E, B, dt))
        c = rng.normal(size=(2, 2, 2, 3)); pos = rng.random(3); vol = float(rng.random() + 0.1)
        assert bits_equal(Ref.trilinear(c, pos, vol), Oracle.trilinear(c, pos, vol))


def test_out_file_matches_reference_writer(tmp_path, golden_dir):
    """The sorted .out of the oracle state equals the reference writer's (verify_output.sh is a sorted diff)."""
    r = Ref.from_inp(os.path.join(golden_dir, "inputs", "test.inp"))
    r.step(3)
    path = str(tmp_path / "ref.out")
    r.write_counts(path)
    lines = open(path).read().splitlines()
    assert lines[0] == "[8,8,8]" and lines[-1] == "Total: %d" % r.total()
    c = r.counts().reshape(8, 8, 8)
    assert lines[1 + (3 * 8 + 4) * 8 + 5] == "3,4,5:%d" % c[3, 4, 5]
