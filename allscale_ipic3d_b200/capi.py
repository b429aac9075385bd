"""ctypes binding of include/ipic3d_b200.h (the drop-in C ABI).  No torch types cross this boundary."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

MOVER_DIPOLE_ANALYTIC = 0
MOVER_INTERPOLATED = 1 << 0
PROJECT_CURRENT = 1 << 1
SOLVER_FORWARD = 1 << 2
SOLVER_LEAPFROG = 1 << 3
UPDATE_BOUNDARIES = 1 << 4
SKIP_MIGRATION = 1 << 5

# every symbol include/ipic3d_b200.h declares
API_SYMBOLS = [
    "ipic_last_error", "ipic_version", "ipic_device_count", "ipic_create", "ipic_create_slab", "ipic_destroy",
    "ipic_set_headroom", "ipic_comm_unique_id", "ipic_comm_init", "ipic_upload_cells", "ipic_download_cells",
    "ipic_cell_counts", "ipic_total_particles", "ipic_upload_fields", "ipic_download_fields", "ipic_step",
    "ipic_step_async", "ipic_sync", "ipic_count_misplaced", "ipic_profile_reset", "ipic_profile_get", "ipic_selftest",
    "ipic_set_exchange_capacity", "ipic_group_create", "ipic_group_step", "ipic_relayout_count",
    "ipic_alloc_pinned", "ipic_free_pinned", "ipic_set_uniform_species", "ipic_particle_fields", "ipic_profile_phases",
    "ipic_diagnostics", "ipic_group_diagnostics", "ipic_trace", "ipic_generate_uniform", "ipic_generate_normal_speed",
    "ipic_create_slab_range", "ipic_group_create_ranges",
]


class IpicError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("ipic3d_b200 error %d: %s" % (code, message))
        self.code = code


class Props(C.Structure):
    """``ipic_props``: POD mirror of ipic3d::UniverseProperties (reference universe_properties.h:52-75)."""
    _fields_ = [("size", C.c_int64 * 3), ("cellWidth", C.c_double * 3), ("dt", C.c_double),
                ("speedOfLight", C.c_double), ("planetRadius", C.c_double), ("objectCenter", C.c_double * 3),
                ("origin", C.c_double * 3), ("externalMagneticField", C.c_double * 3)]

    @classmethod
    def make(cls, size, cellWidth=(1.0, 1.0, 1.0), dt=1.0, speedOfLight=1.0, planetRadius=0.0,
             objectCenter=(0.0, 0.0, 0.0), origin=(0.0, 0.0, 0.0), externalMagneticField=(0.0, 0.0, 0.0)):
        p = cls()
        if np.isscalar(size):
            size = (size,) * 3
        p.size[:] = [int(s) for s in size]
        p.cellWidth[:] = cellWidth
        p.dt, p.speedOfLight, p.planetRadius = dt, speedOfLight, planetRadius
        p.objectCenter[:] = objectCenter
        p.origin[:] = origin
        p.externalMagneticField[:] = externalMagneticField
        return p

    @classmethod
    def from_other(cls, other):
        """Copy from any ctypes structure with the same layout (e.g. the test harness' Props)."""
        p = cls()
        assert C.sizeof(other) == C.sizeof(p)
        C.memmove(C.byref(p), C.byref(other), C.sizeof(p))
        return p

    @property
    def ncells(self):
        return int(self.size[0] * self.size[1] * self.size[2])

    def grid_len(self, extra):
        return int((self.size[0] + extra) * (self.size[1] + extra) * (self.size[2] + extra))


def lib_path():
    # IPIC_LIB: tuning builds of the same sources with other compile-time knobs (scripts/variants.py)
    return os.environ.get("IPIC_LIB") or os.path.join(_HERE, "lib", "libipic3d_b200.so")


_lib = None


def load_library():
    """Load the CUDA library; fails loudly if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise IpicError(-5, "%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                "(nvcc, sm_100a) first; there is no CPU fallback" % path)
        lib = C.CDLL(path)
        lib.ipic_last_error.restype = C.c_char_p
        lib.ipic_version.restype = C.c_char_p
        _lib = lib
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _up(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


class Context:
    """One z-slab of a universe on one GPU (``ipic_ctx``).  Single GPU: rank 0 of 1."""

    def __init__(self, props, device=0, rank=0, nranks=1, _handle=None, z_begin=None, z_count=None):
        """z_begin / z_count: the slab's planes when the slabs are not equally thick (ipic_create_slab_range)."""
        self.lib = load_library()
        self.props = Props.from_other(props)
        self.rank, self.nranks = rank, nranks
        self.h = C.c_void_p()
        if z_count is None:
            z_count = int(self.props.size[2]) // nranks
            z_begin = rank * z_count
        self.z_begin, self.z_count = int(z_begin), int(z_count)
        if _handle is not None:
            self.h = C.c_void_p(_handle)
        else:
            self._ck(self.lib.ipic_create_slab_range(C.byref(self.props), C.c_int(device), C.c_int(rank), C.c_int(nranks),
                                                     C.c_int(self.z_begin), C.c_int(self.z_count), C.byref(self.h)))
        self.owned_cells = int(self.props.size[0] * self.props.size[1]) * self.z_count

    def _ck(self, rc):
        if rc != 0:
            raise IpicError(rc, self.lib.ipic_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.ipic_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- multi-GPU wiring --
    @staticmethod
    def comm_unique_id():
        lib = load_library()
        buf = C.create_string_buffer(128)
        rc = lib.ipic_comm_unique_id(buf)
        if rc:
            raise IpicError(rc, lib.ipic_last_error().decode())
        return buf.raw

    def comm_init(self, unique_id):
        self._ck(self.lib.ipic_comm_init(self.h, C.c_char_p(unique_id)))

    # -- state --
    def set_exchange_capacity(self, particles):
        self._ck(self.lib.ipic_set_exchange_capacity(self.h, C.c_int64(particles)))

    def set_headroom(self, slack, min_pad=64):
        self._ck(self.lib.ipic_set_headroom(self.h, C.c_double(slack), C.c_int64(min_pad)))

    def relayout_count(self):
        t = C.c_uint64(0)
        self._ck(self.lib.ipic_relayout_count(self.h, C.byref(t)))
        return int(t.value)

    def set_uniform_species(self, enable):
        self._ck(self.lib.ipic_set_uniform_species(self.h, C.c_int(1 if enable else 0)))

    def particle_fields(self):
        t = C.c_int(0)
        self._ck(self.lib.ipic_particle_fields(self.h, C.byref(t)))
        return int(t.value)

    def upload_cells(self, counts, aos):
        counts = np.ascontiguousarray(counts, dtype=np.uint64)
        aos = np.ascontiguousarray(aos, dtype=np.float64).reshape(-1, 8)
        if counts.size != self.owned_cells or int(counts.sum()) != len(aos):
            raise ValueError("counts/particles shape mismatch")
        self._ck(self.lib.ipic_upload_cells(self.h, _up(counts), _dp(aos)))

    def upload_cells_ptr(self, counts, aos_ptr):
        """Upload from a raw host pointer (e.g. pinned memory owned by the caller)."""
        counts = np.ascontiguousarray(counts, dtype=np.uint64)
        self._ck(self.lib.ipic_upload_cells(self.h, _up(counts), C.c_void_p(aos_ptr)))

    def cell_counts(self):
        out = np.zeros(self.owned_cells, dtype=np.uint64)
        self._ck(self.lib.ipic_cell_counts(self.h, _up(out)))
        return out

    def total_particles(self):
        t = C.c_uint64(0)
        self._ck(self.lib.ipic_total_particles(self.h, C.byref(t)))
        return int(t.value)

    def download_cells(self):
        n = self.total_particles()
        counts = np.zeros(self.owned_cells, dtype=np.uint64)
        aos = np.zeros((n, 8), dtype=np.float64)
        self._ck(self.lib.ipic_download_cells(self.h, _up(counts), _dp(aos), C.c_uint64(n)))
        return counts, aos

    def download_cells_ptr(self, counts, aos_ptr, capacity):
        self._ck(self.lib.ipic_download_cells(self.h, _up(counts), C.c_void_p(aos_ptr), C.c_uint64(capacity)))

    def upload_fields(self, field=None, bcfield=None, density=None):
        f = None if field is None else np.ascontiguousarray(field, dtype=np.float64).reshape(self.props.grid_len(3), 9)
        b = None if bcfield is None else np.ascontiguousarray(bcfield, dtype=np.float64).reshape(self.props.grid_len(2), 3)
        j = None if density is None else np.ascontiguousarray(density, dtype=np.float64).reshape(self.props.grid_len(1), 3)
        self._ck(self.lib.ipic_upload_fields(self.h, _dp(f), _dp(b), _dp(j)))

    def download_fields(self):
        f = np.zeros((self.props.grid_len(3), 9))
        b = np.zeros((self.props.grid_len(2), 3))
        j = np.zeros((self.props.grid_len(1), 3))
        self._ck(self.lib.ipic_download_fields(self.h, _dp(f), _dp(b), _dp(j)))
        return f, b, j

    # -- the hot path --
    def step(self, n=1, flags=0):
        """Blocking: n iterations of the simulateSteps loop body; returns (firstStep, remainingSteps) seconds."""
        a, b = C.c_double(0), C.c_double(0)
        self._ck(self.lib.ipic_step(self.h, C.c_uint64(n), C.c_uint32(flags), C.byref(a), C.byref(b)))
        return a.value, b.value

    def step_async(self, n=1, flags=0):
        self._ck(self.lib.ipic_step_async(self.h, C.c_uint64(n), C.c_uint32(flags)))

    def sync(self):
        self._ck(self.lib.ipic_sync(self.h))

    def count_misplaced(self):
        t = C.c_uint64(0)
        self._ck(self.lib.ipic_count_misplaced(self.h, C.byref(t)))
        return int(t.value)

    def generate_uniform(self, num_particles):
        """The reference's `:U:N` universe (this context's cells) generated on the device, bit-identical to initCells."""
        self._ck(self.lib.ipic_generate_uniform(self.h, C.c_uint64(num_particles)))

    def generate_normal_speed(self, counts, mean, stddev, q, qom):
        """An `.inp` deck's universe (this context's cells) generated on the device: phase 2 of initCells, bit-identical to the host's."""
        counts = np.ascontiguousarray(counts, dtype=np.uint64)
        if counts.size != self.owned_cells:
            raise ValueError("one count per owned cell")
        m = (C.c_double * 3)(*mean)
        s = (C.c_double * 3)(*stddev)
        self._ck(self.lib.ipic_generate_normal_speed(self.h, _up(counts), m, s, C.c_double(q), C.c_double(qom)))

    def diagnostics(self):
        """(total momentum, E energy, B energy, total kinetic energy): the columns of ConservedQuantities.out, reduced on the device."""
        out = (C.c_double * 4)()
        self._ck(self.lib.ipic_diagnostics(self.h, out))
        return tuple(out)

    def profile_reset(self):
        self._ck(self.lib.ipic_profile_reset(self.h))

    def profile_phases(self):
        out = (C.c_double * 4)()
        self._ck(self.lib.ipic_profile_phases(self.h, out))
        return list(out)

    def profile_get(self):
        a, b, n = C.c_double(0), C.c_double(0), C.c_uint64(0)
        self._ck(self.lib.ipic_profile_get(self.h, C.byref(a), C.byref(b), C.byref(n)))
        return a.value, b.value, int(n.value)


class Group:
    """All z-slabs of a universe in ONE process (``ipic_group_create``): rank r on ``devices[r]`` (devices may
    repeat), exchange by peer copies.  Host arrays are given / returned for the whole universe."""

    def __init__(self, props, devices, planes=None):
        """planes: number of z-planes per slab (default: equal slabs)."""
        self.lib = load_library()
        self.props = Props.from_other(props)
        n = len(devices)
        self.nz = int(self.props.size[2])
        if planes is None:
            if self.nz % n:
                raise ValueError("nz=%d is not divisible by %d slabs" % (self.nz, n))
            planes = [self.nz // n] * n
        self.planes = [int(p) for p in planes]
        self.z0 = [int(sum(self.planes[:r])) for r in range(n)]
        handles = (C.c_void_p * n)()
        devs = (C.c_int * n)(*devices)
        rc = self.lib.ipic_group_create_ranges(C.byref(self.props), devs, C.c_int(n), (C.c_int * n)(*self.z0), (C.c_int * n)(*self.planes), handles)
        if rc:
            raise IpicError(rc, self.lib.ipic_last_error().decode())
        self._handles = handles
        self.members = [Context(self.props, devices[r], r, n, _handle=handles[r], z_begin=self.z0[r], z_count=self.planes[r]) for r in range(n)]

    def close(self):
        for m in self.members:
            m.close()
        self.members = []

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _slab_index(self, r):
        """Indices of rank r's cells in the global flatten, in slab order (x*ny + y)*lz + zl."""
        nxy = int(self.props.size[0] * self.props.size[1])
        return (np.arange(nxy)[:, None] * self.nz + self.z0[r] + np.arange(self.planes[r])[None, :]).reshape(-1)

    def upload_cells(self, counts, aos):
        counts = np.ascontiguousarray(counts, dtype=np.uint64)
        aos = np.ascontiguousarray(aos, dtype=np.float64).reshape(-1, 8)
        start = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        for r, m in enumerate(self.members):
            idx = self._slab_index(r)
            rows = np.concatenate([np.arange(start[i], start[i + 1]) for i in idx]) if len(idx) else np.zeros(0, dtype=np.int64)
            m.upload_cells(counts[idx], aos[rows.astype(np.int64)])

    def download_cells(self):
        ncell = self.props.ncells
        counts = np.zeros(ncell, dtype=np.uint64)
        parts = [None] * ncell
        for r, m in enumerate(self.members):
            c, a = m.download_cells()
            idx = self._slab_index(r)
            counts[idx] = c
            s = np.concatenate([[0], np.cumsum(c)]).astype(np.int64)
            for k, i in enumerate(idx):
                parts[i] = a[s[k]:s[k + 1]]
        aos = np.concatenate(parts) if ncell else np.zeros((0, 8))
        return counts, aos.reshape(-1, 8)

    def cell_counts(self):
        counts = np.zeros(self.props.ncells, dtype=np.uint64)
        for r, m in enumerate(self.members):
            counts[self._slab_index(r)] = m.cell_counts()
        return counts

    def upload_fields(self, field=None, bcfield=None, density=None):
        for m in self.members:                      # the field grids are replicated on every slab
            m.upload_fields(field, bcfield, density)

    def download_fields(self, rank=0):
        return self.members[rank].download_fields()

    def set_headroom(self, slack, min_pad=64):
        for m in self.members:
            m.set_headroom(slack, min_pad)

    def set_exchange_capacity(self, particles):
        for m in self.members:
            m.set_exchange_capacity(particles)

    def set_uniform_species(self, enable):
        for m in self.members:
            m.set_uniform_species(enable)

    def particle_fields(self):
        return [m.particle_fields() for m in self.members]

    def step(self, n=1, flags=0):
        a, b = C.c_double(0), C.c_double(0)
        rc = self.lib.ipic_group_step(self._handles, C.c_int(len(self.members)), C.c_uint64(n), C.c_uint32(flags), C.byref(a), C.byref(b))
        if rc:
            raise IpicError(rc, self.lib.ipic_last_error().decode())
        return a.value, b.value

    def count_misplaced(self):
        return sum(m.count_misplaced() for m in self.members)

    def diagnostics(self):
        out = (C.c_double * 4)()
        rc = self.lib.ipic_group_diagnostics(self._handles, C.c_int(len(self.members)), out)
        if rc:
            raise IpicError(rc, self.lib.ipic_last_error().decode())
        return tuple(out)

    def relayout_count(self):
        return sum(m.relayout_count() for m in self.members)


def trace(props, aos, timesteps, frame_interval, device=0):
    """pic_tracer's traceParticle over all particles of `aos` ((n, 8) float64): returns (counts[frames, nx, ny, nz] uint32, seconds)."""
    lib = load_library()
    aos = np.ascontiguousarray(aos, dtype=np.float64).reshape(-1, 8)
    frames = (max(timesteps, 1) - 1) // max(frame_interval, 1) + 1
    counts = np.zeros((frames, int(props.size[0]), int(props.size[1]), int(props.size[2])), dtype=np.uint32)
    sec = C.c_double(0)
    p = props if isinstance(props, Props) else Props.from_other(props)
    rc = lib.ipic_trace(C.byref(p), C.c_int(device), aos.ctypes.data_as(C.POINTER(C.c_double)), C.c_uint64(len(aos)), C.c_int(timesteps),
                        C.c_int(frame_interval), C.c_int(frames), counts.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(sec))
    if rc:
        raise IpicError(rc, lib.ipic_last_error().decode())
    return counts, sec.value


def selftest(n=1 << 22, seed=1, device=0):
    """Device self-test of the FP64 helpers (see ipic_selftest); returns the number of mismatches (must be 0)."""
    lib = load_library()
    bad = C.c_uint64(0)
    rc = lib.ipic_selftest(C.c_int(device), C.c_uint64(n), C.c_uint64(seed), C.byref(bad))
    if rc:
        raise IpicError(rc, lib.ipic_last_error().decode())
    return int(bad.value)


# -- host-side input generator (host/src/host_capi.cc; not part of the device C ABI) -----------------

_hostlib = None


def load_host_library():
    global _hostlib
    if _hostlib is None:
        path = os.path.join(_HERE, "lib", "libipic3d_host.so")
        if not os.path.exists(path):
            raise IpicError(-5, "%s is missing: run __graft_entry__.build() first" % path)
        _hostlib = C.CDLL(path)
    return _hostlib


def uniform_benchmark_props(size=32):
    """UniverseProperties of the reference's synthetic benchmarks (benchmark.h:13-25, :61-65)."""
    return Props.make(size, cellWidth=(10.0, 10.0, 10.0), dt=0.15)


def uniform_counts(props, num_particles, z_begin=0, z_count=None):
    """Per-cell populations of the reference's ``:U:N`` universe for the z-planes [z_begin, z_begin+z_count)."""
    lib = load_host_library()
    z_count = int(props.size[2]) - z_begin if z_count is None else z_count
    counts = np.zeros(int(props.size[0] * props.size[1]) * z_count, dtype=np.uint64)
    rc = lib.ipic_host_uniform_counts(C.byref(props), C.c_uint64(num_particles), C.c_int(z_begin), C.c_int(z_count), _up(counts))
    if rc:
        raise IpicError(rc, "ipic_host_uniform_counts: bad arguments")
    return counts


def uniform_particles_into(props, num_particles, aos_ptr, z_begin=0, z_count=None, threads=0):
    """Fill caller memory (e.g. pinned) with the particles of the ``:U:N`` universe, cell-major AoS."""
    lib = load_host_library()
    z_count = int(props.size[2]) - z_begin if z_count is None else z_count
    rc = lib.ipic_host_uniform_particles(C.byref(props), C.c_uint64(num_particles), C.c_int(z_begin), C.c_int(z_count),
                                         C.c_void_p(aos_ptr), C.c_int(threads))
    if rc:
        raise IpicError(rc, "ipic_host_uniform_particles: bad arguments")


def uniform_universe(props, num_particles, z_begin=0, z_count=None, threads=0):
    counts = uniform_counts(props, num_particles, z_begin, z_count)
    aos = np.empty((int(counts.sum()), 8), dtype=np.float64)
    if len(aos):
        uniform_particles_into(props, num_particles, aos.ctypes.data, z_begin, z_count, threads)
    return counts, aos


def balanced_planes(plane_load, world):
    """Cuts nz z-planes into `world` consecutive slabs of about equal load (e.g. particles per plane), at least one plane
    each: the list of planes per slab.  Greedy on the cumulative load: slab r ends where the running sum first reaches
    (r + 1) / world of the total, leaving enough planes for the slabs that follow."""
    load = np.asarray(plane_load, dtype=np.float64)
    nz = len(load)
    if world > nz:
        raise ValueError("%d slabs need at least as many z-planes (%d)" % (world, nz))
    cum = np.concatenate([[0.0], np.cumsum(load)])
    cuts = [0]
    for r in range(1, world):
        target = cum[-1] * r / world
        z = int(np.searchsorted(cum, target, side="left"))
        # the cut that leaves the smaller imbalance, within what the remaining slabs allow
        if z > 0 and abs(cum[z - 1] - target) <= abs(cum[min(z, nz)] - target):
            z -= 1
        z = max(z, cuts[-1] + 1)
        z = min(z, nz - (world - r))
        cuts.append(z)
    cuts.append(nz)
    return [cuts[r + 1] - cuts[r] for r in range(world)]


def slab_of(counts, aos, nxy, nz, rank, world, z_begin=None, z_count=None):
    """The z-slab `rank` of `world` (or the planes [z_begin, z_begin + z_count)) of a universe given as per-cell counts
    (z-fastest flatten) + cell-major AoS particles: (counts, particles) of the slab's cells in slab order (x*ny + y)*lz + zl."""
    counts = np.asarray(counts)
    lz = nz // world if z_count is None else z_count
    z0 = rank * lz if z_begin is None else z_begin
    first = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    cells = (np.arange(nxy)[:, None] * nz + z0 + np.arange(lz)[None, :]).reshape(-1)
    mine = counts[cells]
    offset = np.concatenate([[0], np.cumsum(mine)[:-1]]).astype(np.int64)
    idx = np.repeat(first[cells] - offset, mine.astype(np.int64)) + np.arange(int(mine.sum()))
    return mine, np.asarray(aos).reshape(-1, 8)[idx]


class Deck:
    """An input deck as ``ipic3d <deck>`` sets it up (host mirror's EarthDeck), generated piecewise: phase 1 of initCells
    (every cell's population) runs once at open, after which any range of z-planes can be produced on its own --
    what one rank of a multi-GPU run needs.  Bit-identical to the reference's universe (tests/test_host_mirror.py)."""

    def __init__(self, deck_path):
        lib = load_host_library()
        lib.ipic_host_deck_open.restype = C.c_void_p
        self.lib = lib
        self.props = Props()
        cycles, n = C.c_uint64(0), C.c_int64(0)
        self.h = C.c_void_p(lib.ipic_host_deck_open(deck_path.encode(), C.byref(self.props), C.byref(cycles), C.byref(n)))
        self.ncycles, self.num_particles = int(cycles.value), int(n.value)

    def close(self):
        if self.h:
            self.lib.ipic_host_deck_close(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def counts(self, z_begin=0, z_count=None):
        z_count = int(self.props.size[2]) - z_begin if z_count is None else z_count
        out = np.zeros(int(self.props.size[0] * self.props.size[1]) * z_count, dtype=np.uint64)
        if self.lib.ipic_host_deck_counts(self.h, C.c_int(z_begin), C.c_int(z_count), _up(out)):
            raise IpicError(-1, "ipic_host_deck_counts: bad arguments")
        return out

    def particles_into(self, aos_ptr, z_begin=0, z_count=None, threads=0):
        z_count = int(self.props.size[2]) - z_begin if z_count is None else z_count
        if self.lib.ipic_host_deck_particles(self.h, C.c_int(z_begin), C.c_int(z_count), C.c_void_p(aos_ptr), C.c_int(threads)):
            raise IpicError(-1, "ipic_host_deck_particles: bad arguments")

    def plane_population(self):
        """particles per z-plane of the whole universe (what balanced_planes cuts by)"""
        nz = int(self.props.size[2])
        return self.counts(0, nz).reshape(-1, nz).sum(axis=0)

    def distribution(self):
        """(mean[3], stddev[3], q, qom) of phase 2: what ipic_generate_normal_speed needs besides the populations."""
        m, s = (C.c_double * 3)(), (C.c_double * 3)()
        q, qom = C.c_double(0), C.c_double(0)
        self.lib.ipic_host_deck_distribution(self.h, m, s, C.byref(q), C.byref(qom))
        return list(m), list(s), q.value, qom.value

    def generate_on_device(self, ctx, rank=0, world=1):
        """the context's z-slab created in its HBM (phase 1 populations from the host, phase 2 on the GPU)."""
        counts = self.counts(ctx.z_begin, ctx.z_count)
        ctx.generate_normal_speed(counts, *self.distribution())
        return counts

    def slab(self, rank=0, world=1):
        """(counts, particles) of z-slab `rank` of `world` in slab order."""
        lz = int(self.props.size[2]) // world
        counts = self.counts(rank * lz, lz)
        aos = np.empty((int(counts.sum()), 8), dtype=np.float64)
        if len(aos):
            self.particles_into(aos.ctypes.data, rank * lz, lz)
        return counts, aos


def deck_slab_pinned(deck_path, rank, world, torch, balanced=False):
    """z-slab `rank` of `world` of the deck's universe generated straight into pinned host memory:
    (props, counts, tensor, (z_begin, z_count)).  balanced: slabs cut by particle count instead of equally thick."""
    with Deck(deck_path) as d:
        nz = int(d.props.size[2])
        planes = balanced_planes(d.plane_population(), world) if balanced else [nz // world] * world
        z0, lz = int(sum(planes[:rank])), planes[rank]
        counts = d.counts(z0, lz)
        host = torch.empty((max(int(counts.sum()), 1), 8), dtype=torch.float64).pin_memory()
        if counts.sum():
            d.particles_into(host.data_ptr(), z0, lz)
        return Props.from_other(d.props), counts, host[:int(counts.sum())], (z0, lz)


def dipole_universe(deck_path, grid, per_dim):
    """The reference's `ipic3d <deck>` universe (main()'s Earth set-up, two-phase initCells) built by the host C++ mirror,
    with the deck's grid replaced by grid^3 cells (box length scaled to keep the cell width) and per_dim^3 particles per cell."""
    import re
    import tempfile
    text = open(deck_path).read()
    old = int(re.search(r"^nxc = (\d+)", text, re.M).group(1))
    length = float(re.search(r"^Lx = ([0-9.]+)", text, re.M).group(1)) * grid / old
    for k in ("nxc", "nyc", "nzc"):
        text = re.sub(r"^%s = .*$" % k, "%s = %d" % (k, grid), text, flags=re.M)
    for k in ("Lx", "Ly", "Lz"):
        text = re.sub(r"^%s = .*$" % k, "%s = %r" % (k, length), text, flags=re.M)
    for k in ("npcelx", "npcely", "npcelz"):
        text = re.sub(r"^%s = .*$" % k, "%s = %d %d" % (k, per_dim, per_dim), text, flags=re.M)
    lib = load_host_library()
    lib.ipic_mirror_create_from_inp.restype = C.c_void_p
    lib.ipic_mirror_total.restype = C.c_uint64
    with tempfile.NamedTemporaryFile("w", suffix=".inp", delete=False) as f:
        f.write(text)
    props = Props()
    h = C.c_void_p(lib.ipic_mirror_create_from_inp(f.name.encode(), C.byref(props), None, None))
    os.unlink(f.name)
    n = lib.ipic_mirror_total(h)
    counts = np.zeros(props.ncells, dtype=np.uint64)
    aos = np.empty((n, 8), dtype=np.float64)
    lib.ipic_mirror_get_counts(h, _up(counts))
    lib.ipic_mirror_get_particles(h, _dp(aos))
    lib.ipic_mirror_destroy(h)
    return props, counts, aos
