"""ctypes harness around the two CPU checkers (test infrastructure, not product code).

* ``Oracle``  - oracle/libipic3d_oracle.so, the plain-C restatement (always available).
* ``Ref``     - oracle/_ref/libipic3d_ref{,_mt}.so, the UNMODIFIED reference headers compiled against the
                AllScale shim (prebuilt in the container where /root/reference exists; travels to the GPU box).

Both expose the same small surface: set/get cells (counts + AoS particles), fields, step(n, flags).
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REFERENCE_INPUTS = "/root/reference/inputs"

# step flags (include/ipic3d_b200.h)
MOVER_DIPOLE_ANALYTIC = 0
MOVER_INTERPOLATED = 1 << 0
PROJECT_CURRENT = 1 << 1
SOLVER_FORWARD = 1 << 2
SOLVER_LEAPFROG = 1 << 3
UPDATE_BOUNDARIES = 1 << 4
SKIP_MIGRATION = 1 << 5


class Props(C.Structure):
    """POD mirror of UniverseProperties == ipic_props (include/ipic3d_b200.h)."""
    _fields_ = [("size", C.c_int64 * 3), ("cellWidth", C.c_double * 3), ("dt", C.c_double),
                ("speedOfLight", C.c_double), ("planetRadius", C.c_double), ("objectCenter", C.c_double * 3),
                ("origin", C.c_double * 3), ("externalMagneticField", C.c_double * 3)]

    @staticmethod
    def make(size, cellWidth=(1.0, 1.0, 1.0), dt=1.0, speedOfLight=1.0, planetRadius=0.0,
             objectCenter=(0.0, 0.0, 0.0), origin=(0.0, 0.0, 0.0), externalMagneticField=(0.0, 0.0, 0.0)):
        p = Props()
        if np.isscalar(size):
            size = (size,) * 3
        p.size[:] = [int(s) for s in size]
        p.cellWidth[:] = cellWidth
        p.dt, p.speedOfLight, p.planetRadius = dt, speedOfLight, planetRadius
        p.objectCenter[:] = objectCenter
        p.origin[:] = origin
        p.externalMagneticField[:] = externalMagneticField
        return p

    def copy(self):
        q = Props()
        C.memmove(C.byref(q), C.byref(self), C.sizeof(Props))
        return q

    @property
    def ncells(self):
        return int(self.size[0] * self.size[1] * self.size[2])

    def grid_len(self, extra):
        return int((self.size[0] + extra) * (self.size[1] + extra) * (self.size[2] + extra))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _up(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def build_oracle():
    """Compile the C oracle (and, where /root/reference exists, oracle/_ref)."""
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "all"])


def _load(path):
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    return C.CDLL(path)


class _Universe:
    """Shared wrapper logic; subclasses bind the symbol prefix."""
    prefix = None
    lib = None

    def __init__(self, handle, props):
        self.h = C.c_void_p(handle)
        self.props = props

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    def close(self):
        if self.h:
            self._f("destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- particles --
    def counts(self):
        out = np.zeros(self.props.ncells, dtype=np.uint64)
        self._f("get_counts")(self.h, _up(out))
        return out

    def total(self):
        f = self._f("total_particles")
        f.restype = C.c_uint64
        return int(f(self.h))

    def particles(self):
        out = np.zeros((self.total(), 8), dtype=np.float64)
        if len(out):
            self._f("get_particles")(self.h, _dp(out))
        return out

    def set_cells(self, counts, aos):
        counts = np.ascontiguousarray(counts, dtype=np.uint64)
        aos = np.ascontiguousarray(aos, dtype=np.float64).reshape(-1, 8)
        assert counts.size == self.props.ncells and int(counts.sum()) == len(aos)
        self._f("set_cells")(self.h, _up(counts), _dp(aos))

    def misplaced(self):
        f = self._f("count_misplaced")
        f.restype = C.c_uint64
        return int(f(self.h))

    def diagnostics(self):
        """(total momentum, E energy, B energy, total kinetic energy) -- the columns of ConservedQuantities.out"""
        out = np.zeros(4)
        self._f("diagnostics")(self.h, _dp(out))
        return out

    # -- fields --
    def _grid(self, name, extra, width, value=None):
        n = self.props.grid_len(extra)
        if value is None:
            out = np.zeros((n, width), dtype=np.float64)
            self._f("get_" + name)(self.h, _dp(out))
            return out
        value = np.ascontiguousarray(value, dtype=np.float64).reshape(n, width)
        self._f("set_" + name)(self.h, _dp(value))

    def field(self, value=None):
        return self._grid("field", 3, 9, value)

    def bcfield(self, value=None):
        return self._grid("bcfield", 2, 3, value)

    def density(self, value=None):
        return self._grid("density", 1, 3, value)

    def init_fields(self, drift=(0.0, 0.0, 0.0), B0=(0.0, 0.0, 0.0)):
        d = np.array(drift, dtype=np.float64)
        b = np.array(B0, dtype=np.float64)
        self._f("init_fields")(self.h, _dp(d), _dp(b))

    # -- unit level --
    def project_node(self, x, y, z):
        self._f("project_node")(self.h, C.c_int64(x), C.c_int64(y), C.c_int64(z))

    def solve_forward_node(self, x, y, z):
        self._f("solve_forward_node")(self.h, C.c_int64(x), C.c_int64(y), C.c_int64(z))

    def solve_leapfrog_node(self, x, y, z):
        self._f("solve_leapfrog_node")(self.h, C.c_int64(x), C.c_int64(y), C.c_int64(z))

    def update_boundaries(self):
        self._f("update_boundaries")(self.h)


class Oracle(_Universe):
    prefix = "orc_"

    @classmethod
    def load(cls):
        if cls.lib is None:
            path = os.path.join(ORACLE_DIR, "libipic3d_oracle.so")
            if not os.path.exists(path):
                build_oracle()
            cls.lib = _load(path)
            cls.lib.orc_create.restype = C.c_void_p
        return cls.lib

    def __init__(self, props):
        self.load()
        props = props.copy()
        super().__init__(self.lib.orc_create(C.byref(props)), props)

    def step(self, n=1, flags=0):
        self.lib.orc_step(self.h, C.c_uint64(n), C.c_uint32(flags))

    @classmethod
    def update_velocity(cls, p8, E, B, dt):
        cls.load()
        p = np.array(p8, dtype=np.float64)
        cls.lib.orc_update_velocity(_dp(p), _dp(np.array(E, dtype=np.float64)), _dp(np.array(B, dtype=np.float64)), C.c_double(dt))
        return p

    @classmethod
    def update_position(cls, p8, dt):
        cls.load()
        p = np.array(p8, dtype=np.float64)
        cls.lib.orc_update_position(_dp(p), C.c_double(dt))
        return p

    @classmethod
    def trilinear(cls, corners, pos, vol):
        cls.load()
        out = np.zeros(3)
        cls.lib.orc_trilinear(_dp(np.ascontiguousarray(corners, dtype=np.float64)), _dp(np.array(pos, dtype=np.float64)), C.c_double(vol), _dp(out))
        return out

    @classmethod
    def move_particle_dipole(cls, props, p8):
        cls.load()
        p = np.array(p8, dtype=np.float64)
        n = cls.lib.orc_move_particle_dipole(C.byref(props), _dp(p))
        return p, n


class Ref(_Universe):
    """The reference itself. ``threaded=True`` selects the std::thread pfor build."""
    prefix = "ref_"
    _libs = {}

    @classmethod
    def available(cls, threaded=False):
        return os.path.exists(cls._path(threaded))

    @staticmethod
    def _path(threaded):
        return os.path.join(ORACLE_DIR, "_ref", "libipic3d_ref_mt.so" if threaded else "libipic3d_ref.so")

    @classmethod
    def load(cls, threaded=False):
        if threaded not in cls._libs:
            lib = _load(cls._path(threaded))
            for n in ("ref_create_from_inp", "ref_create_uniform", "ref_create_uniform_grid", "ref_create_empty"):
                getattr(lib, n).restype = C.c_void_p
            lib.ref_simulate_steps.restype = C.c_double
            lib.ref_step_full.restype = C.c_double
            cls._libs[threaded] = lib
        return cls._libs[threaded]

    def __init__(self, handle, props, lib):
        self.lib = lib
        super().__init__(handle, props)

    @classmethod
    def empty(cls, props, threaded=False):
        lib = cls.load(threaded)
        props = props.copy()
        return cls(lib.ref_create_empty(C.byref(props)), props, lib)

    @classmethod
    def from_inp(cls, path, threaded=False):
        """Universe exactly as the reference's main() builds it for ``ipic3d <path>``."""
        lib = cls.load(threaded)
        props = Props()
        cycles, npart = C.c_uint64(0), C.c_uint64(0)
        h = lib.ref_create_from_inp(path.encode(), C.byref(props), C.byref(cycles), C.byref(npart))
        u = cls(h, props, lib)
        u.ncycles, u.num_particles = int(cycles.value), int(npart.value)
        return u

    @classmethod
    def uniform(cls, num_particles, threaded=False):
        """Universe of the reference's ``:U:N`` benchmark (32^3 cells)."""
        lib = cls.load(threaded)
        props = Props()
        return cls(lib.ref_create_uniform(C.c_uint64(num_particles), C.byref(props)), props, lib)

    @classmethod
    def benchmark(cls, kind, num_particles, threaded=False):
        """Universe of the reference's ``:X:N`` benchmark, X in U (uniform), C (cluster), E (explosion), B (beam) (benchmark.h:81-138)."""
        lib = cls.load(threaded)
        lib.ref_create_benchmark.restype = C.c_void_p
        props = Props()
        return cls(lib.ref_create_benchmark(C.c_char(kind.encode()), C.c_uint64(num_particles), C.byref(props)), props, lib)

    @classmethod
    def uniform_grid(cls, props, num_particles, threaded=False):
        lib = cls.load(threaded)
        props = props.copy()
        return cls(lib.ref_create_uniform_grid(C.byref(props), C.c_uint64(num_particles)), props, lib)

    def simulate_steps(self, n=1):
        """simulateSteps<>(n, universe) untouched; returns seconds."""
        return float(self.lib.ref_simulate_steps(self.h, C.c_uint64(n)))

    def step(self, n=1, flags=0):
        if flags == 0:
            return self.simulate_steps(n)
        return float(self.lib.ref_step_full(self.h, C.c_uint64(n), C.c_uint32(flags)))

    def threads(self):
        return int(self.lib.ref_threads())

    def write_counts(self, filename):
        self.lib.ref_write_counts(self.h, filename.encode())

    @classmethod
    def update_velocity(cls, p8, E, B, dt):
        lib = cls.load()
        p = np.array(p8, dtype=np.float64)
        lib.ref_update_velocity(_dp(p), _dp(np.array(E, dtype=np.float64)), _dp(np.array(B, dtype=np.float64)), C.c_double(dt))
        return p

    @classmethod
    def trilinear(cls, corners, pos, vol):
        lib = cls.load()
        out = np.zeros(3)
        lib.ref_trilinear(_dp(np.ascontiguousarray(corners, dtype=np.float64)), _dp(np.array(pos, dtype=np.float64)), C.c_double(vol), _dp(out))
        return out


def bits_equal(a, b):
    """Bitwise equality of two float64 arrays (distinguishes -0.0/+0.0, equates identical NaNs)."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))


def norm_rel_err(a, b):
    """max over rows of |a-b| / |b| for (n,3) arrays (the norm-relative error of SURVEY.md §8d)."""
    a = np.asarray(a, dtype=np.float64).reshape(-1, 3)
    b = np.asarray(b, dtype=np.float64).reshape(-1, 3)
    if len(a) == 0:
        return 0.0
    num = np.linalg.norm(a - b, axis=1)
    den = np.linalg.norm(b, axis=1)
    den = np.where(den == 0, 1.0, den)
    return float(np.max(num / den))
