"""Pins the plain-C oracle against the known-answer vectors of the reference's own unit tests
(/root/reference/code/app/test/*.cc, cited per case) and the committed golden files."""
import math
import os

import numpy as np

import harness
from harness import Oracle, Props, bits_equal


def test_update_position_exact():
    # particle_test.cc:7-42
    p = Oracle.update_position([0.0, 0.0, 0.0, 1.0, 2.0, 3.0, 0.0, 0.0], 1.5)
    assert list(p[:3]) == [1.5, 3.0, 4.5]


def test_boris_velocity_kat():
    # particle_test.cc:45-81: qom=4, v=(1,2,3), E=(1,-1,3), B=(0,1,2), dt=1
    p0 = [0.0, 0.0, 0.0, 1.0, 2.0, 3.0, 0.0, 4.0]
    p = Oracle.update_velocity(p0, [0.0] * 3, [0.0] * 3, 1.0)
    assert list(p[3:6]) == [1.0, 2.0, 3.0]
    p = Oracle.update_velocity(p0, [1.0, -1.0, 3.0], [0.0, 1.0, 2.0], 1.0)
    assert np.allclose(p[3:6], [-2.42857, 3.71429, 12.14286], atol=1e-4)


def test_larmor_radius_trace():
    # simulation_test.cc:187-263; last row of results/larmor_radius.txt:
    #   999 2.997e-08 -4.50134e-05 3.47983e-05 0 -63242.7 -77462 0
    q, mass = -1.602e-19, 9.109e-31
    E, B, dt = [0.0, 0.0, 0.0], [0.0, 0.0, 0.01], 3e-11
    rL = mass * 1e5 / (abs(q) * 0.01)
    assert abs(rL - 5.686e-05) < 1e-6
    p = np.array([rL, 0.0, 0.0, 0.0, 1e5, 0.0, q, q / mass])
    p = Oracle.update_velocity(p, E, B, -0.5 * dt)
    for _ in range(1000):
        p = Oracle.update_velocity(p, E, B, dt)
        p = Oracle.update_position(p, dt)
    assert abs(p[0] - -4.50134e-05) < 1e-10 and abs(p[1] - 3.47983e-05) < 1e-9 and p[2] == 0.0
    assert abs(p[3] - -63242.7) < 1e1 and abs(p[4] - -77462.0) < 1e1 and abs(p[5]) < 1e-2


def test_exb_drift_trace():
    # simulation_test.cc:340-407; last row of results/exb_drift.txt: 4999 49.99 -0.124361 -10.0701 0 0.428246 0.126505 0
    E, B, dt = [0.2, 0.0, 0.0], [0.0, 0.0, 1.0], 0.01
    p = np.array([0.0, 0.0, 0.0, 0.5, 0.0, 0.0, 1.0, 1.0])
    p = Oracle.update_velocity(p, E, B, -0.5 * dt)
    for _ in range(5000):
        p = Oracle.update_velocity(p, E, B, dt)
        p = Oracle.update_position(p, dt)
    assert abs(p[0] - -0.124361) < 1e-5 and abs(p[1] - -10.0701) < 1e-4
    assert abs(p[3] - 0.428246) < 1e-5 and abs(p[4] - 0.126505) < 1e-5


def test_trilinear_exact():
    # field_test.cc:231-273: the interpolation divides by the cell volume
    corners = np.zeros((2, 2, 2, 3))
    corners[...] = 1.0
    out = Oracle.trilinear(corners, [0.5, 0.5, 0.5], 1.0)
    assert list(out) == [1.0, 1.0, 1.0]
    out = Oracle.trilinear(corners, [0.25, 0.5, 0.75], 8.0)
    assert list(out) == [0.125, 0.125, 0.125]
    corners[1] = 3.0                                        # linear in x
    out = Oracle.trilinear(corners, [0.25, 0.1, 0.9], 1.0)
    assert np.allclose(out, 1.5, rtol=0, atol=1e-15)


def test_current_projection_kat():
    # cell_test.cc:211-291: 2^3 cells of width .5, eight particles -> J(1,1,1) = (1.0952, 1.0952, 1.0952) +- 1e-4
    props = Props.make(2, cellWidth=(0.5, 0.5, 0.5), dt=0.1)
    def P(x, y, z, v):
        return [x, y, z, v, v, v, 1.0, 1.0]
    cells = {  # flatten (x*2+y)*2+z
        (0, 0, 0): P(0.4, 0.4, 0.4, 0.6), (1, 0, 0): P(0.8, 0.4, 0.4, 0.3), (0, 1, 0): P(0.4, 0.8, 0.4, 0.1),
        (1, 1, 0): P(0.75, 0.75, 0.4, 0.8), (0, 0, 1): P(0.4, 0.4, 0.6, 0.6), (1, 0, 1): P(0.8, 0.4, 0.8, 0.3),
        (0, 1, 1): P(0.25, 0.75, 0.75, 0.8), (1, 1, 1): P(0.75, 0.75, 0.8, 0.8)}
    aos = [cells[(x, y, z)] for x in range(2) for y in range(2) for z in range(2)]
    o = Oracle(props)
    o.set_cells(np.ones(8, dtype=np.uint64), np.array(aos))
    assert o.misplaced() == 0
    o.step(1, harness.PROJECT_CURRENT | harness.SKIP_MIGRATION)
    J = o.density().reshape(3, 3, 3, 3)
    assert np.allclose(J[1, 1, 1], 1.0952, atol=1e-4)
    assert abs(J[1, 1, 1, 0] - 1.0952000000000002) < 1e-15      # hand-derived during the survey


def test_solve_field_forward_kat():
    # field_test.cc:419-463: E 0.2 -> -0.4, Bc unchanged
    props = Props.make(5, dt=1.0)
    o = Oracle(props)
    f = np.zeros((8 ** 3, 9)); f[:, 0:3] = 0.2; f[:, 3:6] = 0.4
    o.field(f)
    o.bcfield(np.full((7 ** 3, 3), 0.8))
    o.density(np.full((6 ** 3, 3), 0.6))
    o.solve_forward_node(2, 2, 2)
    node = o.field().reshape(8, 8, 8, 9)[2, 2, 2]
    assert np.allclose(node[0:3], -0.4, rtol=0, atol=1e-15)
    assert np.allclose(o.bcfield().reshape(7, 7, 7, 3)[2, 2, 2], 0.8, rtol=0, atol=1e-15)


def test_solve_field_leapfrog_zero_field():
    # field_test.cc:515-557: all-zero fields stay zero
    props = Props.make(5, dt=1.0)
    o = Oracle(props)
    o.solve_leapfrog_node(2, 2, 2)
    assert not o.field().any() and not o.bcfield().any()


def test_update_fields_on_boundaries_kat():
    # field_test.cc:316-416: 10^3 cells; faces copied periodically, edges/corners untouched
    props = Props.make(10)
    n = 13
    idx = np.stack(np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij"), -1).astype(float)
    f = np.zeros((n, n, n, 9))
    f[..., 0:3] = idx
    f[..., 3:6] = idx
    f[1:-1, 1:-1, 1:-1, 0:3] = idx[1:-1, 1:-1, 1:-1]
    f[1:-1, 1:-1, 1:-1, 3:6] = 2.0 * idx[1:-1, 1:-1, 1:-1]
    o = Oracle(props)
    o.field(f)
    o.update_boundaries()
    g = o.field().reshape(n, n, n, 9)
    assert list(g[0, 5, 11, 0:3]) == [11.0, 5.0, 11.0]      # field_test.cc:413
    assert list(g[9, 12, 2, 3:6]) == [18.0, 2.0, 4.0]       # field_test.cc:414
    assert list(g[0, 0, 5, 0:3]) == [0.0, 0.0, 5.0]         # an edge keeps its initial value


def test_dipole_field_init_kat():
    # field_test.cc:8-42 style: Bext vanishes inside the planet, B = B0 + Bext, E = -(drift x B0)
    Re = 6378137.0
    props = Props.make(4, cellWidth=(Re, Re, Re), planetRadius=Re, origin=(-2 * Re,) * 3, externalMagneticField=(0, 0, 3.07e-5))
    o = Oracle(props)
    o.init_fields((0.02, 0.0, 0.0), (0.0, 0.0, 1e-4))
    f = o.field().reshape(7, 7, 7, 9)
    assert list(f[3, 3, 3, 6:9]) == [0.0, 0.0, 0.0]          # node at the planet centre
    assert np.allclose(f[1, 1, 1, 0:3], [0.0, 0.02 * 1e-4, 0.0])
    x = np.array([-2 * Re, -2 * Re, -2 * Re])
    fac = -3.07e-5 * Re ** 3 / (x @ x) ** 2.5
    assert np.allclose(f[1, 1, 1, 6:9], [3 * x[0] * x[2] * fac, 3 * x[1] * x[2] * fac, (2 * x[2] ** 2 - x[0] ** 2 - x[1] ** 2) * fac], rtol=1e-14)
    assert bits_equal(f[1:6, 1:6, 1:6, 3:6], f[1:6, 1:6, 1:6, 6:9] + np.array([0.0, 0.0, 1e-4]))


def test_migration_single_particle_reflection():
    # simulation_test.cc:410-507 style: 3^3 cells, one particle hopping cell to cell, then the half-cell reflection
    props = Props.make(3, cellWidth=(1.0, 1.0, 1.0), dt=1.0)
    o = Oracle(props)
    counts = np.zeros(27, dtype=np.uint64); counts[13] = 1                 # cell (1,1,1)
    o.set_cells(counts, np.array([[1.5, 1.5, 1.5, 1.0, 0.0, 0.0, 1.0, 1.0]]))
    o.step(1)
    c = o.counts().reshape(3, 3, 3)
    assert c[2, 1, 1] == 1 and c.sum() == 1 and o.particles()[0, 0] == 2.5
    o.step(1)                                                              # 3.5 is outside: shift by half a cell, negate v
    p = o.particles()[0]
    assert o.counts().reshape(3, 3, 3)[2, 1, 1] == 1 and p[0] == 3.0 and p[3] == -1.0
    assert o.misplaced() == 0


def test_small_dipole_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "small_dipole.npz"))
    import ctypes
    props = Props(); ctypes.memmove(ctypes.byref(props), g["props"].tobytes(), ctypes.sizeof(props))
    o = Oracle(props)
    o.set_cells(g["counts0"], g["p0"])
    o.step(1)
    assert np.array_equal(o.counts(), g["counts1"]) and bits_equal(o.particles(), g["p1"])
    o.step(19)
    assert np.array_equal(o.counts(), g["counts20"]) and bits_equal(o.particles(), g["p20"])
    assert g["counts20"].sum() < g["counts0"].sum()          # the planet absorbed some particles


def test_full_pic_step_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "small_fields.npz"))
    import ctypes
    props = Props(); ctypes.memmove(ctypes.byref(props), g["props"].tobytes(), ctypes.sizeof(props))
    for name, solver in (("forward", harness.SOLVER_FORWARD), ("leapfrog", harness.SOLVER_LEAPFROG)):
        o = Oracle(props)
        o.set_cells(g["counts0"], g["p0"])
        o.field(g["field0"]); o.bcfield(g["bc0"])
        o.step(3, harness.MOVER_INTERPOLATED | harness.PROJECT_CURRENT | harness.UPDATE_BOUNDARIES | solver)
        assert np.array_equal(o.counts(), g[name + "_counts"])
        assert bits_equal(o.particles(), g[name + "_p"])
        assert bits_equal(o.field(), g[name + "_field"]) and bits_equal(o.bcfield(), g[name + "_bc"]) and bits_equal(o.density(), g[name + "_J"])
