"""The host C++ mirror (host/include/ipic3d/app: Parameters, UniverseProperties, initCells, initFields, ...)
against the reference's own headers (oracle/_ref): the universes a run STARTS from must be bit-identical,
otherwise no downstream parity claim means anything.  CPU only."""
import ctypes as C
import os

import numpy as np
import pytest

import harness
from harness import Props, Ref, bits_equal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_INPUTS = os.path.join(ROOT, "tests", "golden", "inputs")

pytestmark = pytest.mark.skipif(not Ref.available(), reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def mirror():
    from allscale_ipic3d_b200 import build
    lib = C.CDLL(build.build_host())
    lib.ipic_mirror_create_from_inp.restype = C.c_void_p
    lib.ipic_mirror_create_benchmark.restype = C.c_void_p
    lib.ipic_mirror_total.restype = C.c_uint64
    return lib


def mirror_state(lib, h, props):
    h = C.c_void_p(h)
    n = lib.ipic_mirror_total(h)
    counts = np.zeros(props.ncells, dtype=np.uint64)
    aos = np.zeros((n, 8))
    f, b, j = np.zeros((props.grid_len(3), 9)), np.zeros((props.grid_len(2), 3)), np.zeros((props.grid_len(1), 3))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    lib.ipic_mirror_get_counts(h, counts.ctypes.data_as(C.POINTER(C.c_uint64)))
    lib.ipic_mirror_get_particles(h, dp(aos))
    lib.ipic_mirror_get_fields(h, dp(f), dp(b), dp(j))
    lib.ipic_mirror_destroy(h)
    return counts, aos, f, b, j


def assert_same_universe(got, ref):
    counts, aos, f, b, j = got
    assert np.array_equal(counts, ref.counts())
    assert bits_equal(aos, ref.particles())
    assert bits_equal(f, ref.field()) and bits_equal(b, ref.bcfield()) and bits_equal(j, ref.density())


@pytest.mark.parametrize("deck", ["test.inp", "tiny.inp"])
def test_inp_universe_bit_identical(mirror, deck):
    """`ipic3d <deck>` set-up (parser, Earth overrides, two-phase initCells with the ziggurat, dipole fields)."""
    path = os.path.join(GOLDEN_INPUTS, deck)
    props, cycles, npart = Props(), C.c_uint64(0), C.c_uint64(0)
    h = mirror.ipic_mirror_create_from_inp(path.encode(), C.byref(props), C.byref(cycles), C.byref(npart))
    ref = Ref.from_inp(path)
    assert bytes(props) == bytes(ref.props)
    assert (cycles.value, npart.value) == (ref.ncycles, ref.num_particles)
    assert_same_universe(mirror_state(mirror, h, props), ref)


@pytest.mark.parametrize("deck,world", [("test.inp", 1), ("test.inp", 2), ("test.inp", 8), ("tiny.inp", 4)])
def test_deck_generated_slab_by_slab(deck, world):
    """capi.Deck: phase 1 of initCells once, then every z-slab on its own (what a rank of a multi-GPU run does) --
    put together, the slabs are the reference's universe bit for bit."""
    from allscale_ipic3d_b200 import capi
    path = os.path.join(GOLDEN_INPUTS, deck)
    ref = Ref.from_inp(path)
    want_counts, want_aos = ref.counts(), ref.particles()
    with capi.Deck(path) as d:
        assert bytes(d.props) == bytes(ref.props)
        assert (d.ncycles, d.num_particles) == (ref.ncycles, ref.num_particles)
        nz = int(d.props.size[2]); nxy = int(d.props.size[0] * d.props.size[1])
        for rank in range(world):
            counts, aos = d.slab(rank, world)
            wc, wa = capi.slab_of(want_counts, want_aos, nxy, nz, rank, world)
            assert np.array_equal(counts, wc) and bits_equal(aos, wa), "slab %d of %d" % (rank, world)


@pytest.mark.parametrize("kind,n", [("U", 300000), ("C", 200000), ("E", 100000), ("B", 100000)])
def test_benchmark_universes_bit_identical(mirror, kind, n):
    """`ipic3d :U:N`, `:C:N`, `:E:N`, `:B:N` (normal / spherical / beam generators, 32^3 grid)."""
    lib = Ref.load()
    lib.ref_create_benchmark.restype = C.c_void_p
    rp = Props()
    ref = Ref(lib.ref_create_benchmark(C.c_char(kind.encode()), C.c_uint64(n), C.byref(rp)), rp, lib)
    props = Props()
    h = mirror.ipic_mirror_create_benchmark(C.c_char(kind.encode()), C.c_uint64(n), C.byref(props))
    assert bytes(props) == bytes(rp)
    assert_same_universe(mirror_state(mirror, h, props), ref)


def parse_with(fn, path):
    buf = C.create_string_buffer(1 << 16)
    assert fn(path.encode(), buf, C.c_int(len(buf))) > 0
    return buf.value.decode()


def test_conserved_quantity_folds_bit_identical(mirror):
    """getTotalParticlesEnergy / getParticlesMomentum / getParticlesKineticEnergy / getFieldEnergy of the host mirror
    (cell.h:1315-1355, field.h:443-459) on the test.inp universe: bit-identical to the reference's."""
    path = os.path.join(GOLDEN_INPUTS, "test.inp")
    props, cycles, npart = Props(), C.c_uint64(0), C.c_uint64(0)
    h = C.c_void_p(mirror.ipic_mirror_create_from_inp(path.encode(), C.byref(props), C.byref(cycles), C.byref(npart)))
    got = np.zeros(4)
    mirror.ipic_mirror_diagnostics(h, got.ctypes.data_as(C.POINTER(C.c_double)))
    mirror.ipic_mirror_destroy(h)
    ref = Ref.from_inp(path)
    want = ref.diagnostics()
    assert bits_equal(got, want) and got[0] > 0 and got[2] > 0 and got[3] > 0


@pytest.mark.parametrize("solver", ["forward", "leapfrog"])
def test_host_scalar_field_and_projection_functions_bit_identical(mirror, solver):
    """The host-callable scalar versions the mirror keeps for known-answer checks: projectToDensityField,
    solveFieldForward / solveFieldLeapfrog, updateFieldsOnBoundaries, applied node by node in the order of the reference's
    (commented) loop on the test.inp universe with seeded random fields, against the reference's own functions: every
    bit of E, B, Bc and J equal.  (simulateSteps never calls them; its field phase is csrc/field_kernels.cu.)"""
    path = os.path.join(GOLDEN_INPUTS, "test.inp")
    props, cycles, npart = Props(), C.c_uint64(0), C.c_uint64(0)
    h = C.c_void_p(mirror.ipic_mirror_create_from_inp(path.encode(), C.byref(props), C.byref(cycles), C.byref(npart)))
    ref = Ref.from_inp(path)
    rng = np.random.default_rng(5)
    f = ref.field() * (1.0 + 0.25 * rng.standard_normal(ref.field().shape)) + 1e-9 * rng.standard_normal(ref.field().shape)
    b = ref.bcfield() * (1.0 + 0.25 * rng.standard_normal(ref.bcfield().shape))
    ref.field(f), ref.bcfield(b)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    mirror.ipic_mirror_set_fields(h, dp(f), dp(b), None)
    n = int(props.size[0])
    for x in range(n + 1):
        for y in range(n + 1):
            for z in range(n + 1):
                ref.project_node(x, y, z)
                assert mirror.ipic_mirror_node_op(h, 0, C.c_int64(x), C.c_int64(y), C.c_int64(z)) == 0
    op = 1 if solver == "forward" else 2
    for x in range(1, n + 2):
        for y in range(1, n + 2):
            for z in range(1, n + 2):
                (ref.solve_forward_node if solver == "forward" else ref.solve_leapfrog_node)(x, y, z)
                mirror.ipic_mirror_node_op(h, op, C.c_int64(x), C.c_int64(y), C.c_int64(z))
    ref.update_boundaries()
    mirror.ipic_mirror_node_op(h, 3, C.c_int64(0), C.c_int64(0), C.c_int64(0))
    gf, gb, gj = np.zeros_like(f), np.zeros_like(b), np.zeros((props.grid_len(1), 3))
    mirror.ipic_mirror_get_fields(h, dp(gf), dp(gb), dp(gj))
    mirror.ipic_mirror_destroy(h)
    assert np.abs(gj).max() > 0 and not bits_equal(gf, f)
    assert bits_equal(gj, ref.density()) and bits_equal(gf, ref.field()) and bits_equal(gb, ref.bcfield())


def test_host_scalar_trilinear_bit_identical(mirror):
    rng = np.random.default_rng(11)
    for _ in range(200):
        corners, pos, vol = rng.standard_normal(24), rng.random(3), float(rng.random() + 0.5)
        got = np.zeros(3)
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        mirror.ipic_mirror_trilinear(dp(corners), dp(pos), C.c_double(vol), dp(got))
        assert bits_equal(got, Ref.trilinear(corners, pos, vol))


@pytest.mark.parametrize("deck", ["test.inp", "tiny.inp", "medium.inp"])
def test_parser_matches_reference(mirror, deck):
    path = os.path.join(GOLDEN_INPUTS, deck)
    assert parse_with(mirror.ipic_mirror_parse, path) == parse_with(Ref.load().ref_parse, path)


def test_parser_quirks(mirror, tmp_path):
    """Substring key matching in table order, '#' comments, lines without '=', per-species lists, dropped last line
    without newline — the mirror and the reference must read the same values from a deliberately awkward deck."""
    base = open(os.path.join(GOLDEN_INPUTS, "test.inp")).read()
    awkward = base + "\n".join([
        "   dt   =    0.125   # trailing comment",
        "# dt = 99",
        "Ndt_like = 7",                      # contains "dt": parsed as dt (first key of the table that occurs)
        "this line has no equal sign dt 5",
        "nxc = 12 # shrink",
        "qom = -64.0 1.0 2.5",
        "Case = Other",
        "ncycles = 77",                      # last line WITHOUT newline: ignored by both
    ])
    p = tmp_path / "awkward.inp"
    p.write_text(awkward)
    a = parse_with(mirror.ipic_mirror_parse, str(p))
    b = parse_with(Ref.load().ref_parse, str(p))
    assert a == b
    assert "dt=7;" in a and "useCase=Test;" in a and "ncells=12," in a and "ncycles=77;" not in a


def test_out_file_format(tmp_path):
    """The reference writer (cell.h:1360-1388) and the driver's writer agree on a universe (checked end to end on
    the GPU by tests/test_gpu_driver.py); here: the reference's file has the documented shape."""
    r = Ref.uniform_grid(Props.make(2, cellWidth=(10.0, 10.0, 10.0)), 50)
    out = tmp_path / "x.out"
    r.write_counts(str(out))
    lines = out.read_text().split("\n")
    assert lines[0] == "[2,2,2]" and lines[1].startswith("0,0,0:") and lines[-2].startswith("Total: 50")
