"""pic_tracer (reference code/app/src/pic_tracer.cxx): the oracle's restatement of traceParticle and the host
mirror's particle generator against the reference's own code (oracle/_ref/libipic3d_ref_tracer.so), CPU only; the CUDA
kernel behind ipic_trace and the `pic_tracer_b200` executable against both, on the GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from harness import Oracle, Props, bits_equal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libipic3d_ref_tracer.so")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "pic_tracer_ref")
OUR_BIN = os.path.join(ROOT, "host", "bin", "pic_tracer_b200")

pytestmark = pytest.mark.skipif(not os.path.exists(REF_LIB), reason="oracle/_ref not built (needs /root/reference)")


def dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def up(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


@pytest.fixture(scope="module")
def ref():
    return C.CDLL(REF_LIB)


def tracer_setup(ref, R, n, seed=0):
    props = Props()
    ref.ref_tracer_config(C.c_int(R), C.byref(props))
    aos = np.zeros((n, 8))
    ref.ref_tracer_particles(C.byref(props), C.c_uint32(seed), C.c_uint64(n), dp(aos))
    return props, aos


def ref_counts(ref, props, aos, T, S):
    frames = (T - 1) // S + 1
    counts = np.zeros((frames, props.size[0], props.size[1], props.size[2]), dtype=np.uint32)
    ref.ref_trace(C.byref(props), dp(aos), C.c_uint64(len(aos)), C.c_int(T), C.c_int(S), C.c_int(frames), up(counts))
    return counts


def oracle_counts(props, aos, T, S):
    lib = Oracle.load()
    frames = (T - 1) // S + 1
    counts = np.zeros((frames, props.size[0], props.size[1], props.size[2]), dtype=np.uint32)
    lib.orc_trace(C.byref(props), dp(aos), C.c_uint64(len(aos)), C.c_int(T), C.c_int(S), up(counts))
    return counts


@pytest.mark.parametrize("R,n,T,S,seed", [(8, 20000, 31, 10, 0), (16, 5000, 51, 7, 12), (4, 3000, 101, 1, 3)])
def test_oracle_trace_equals_reference(ref, R, n, T, S, seed):
    """orc_trace == the reference's traceParticle + ParticleCount: every density frame identical."""
    props, aos = tracer_setup(ref, R, n, seed)
    want = ref_counts(ref, props, aos, T, S)
    got = oracle_counts(props, aos, T, S)
    assert np.array_equal(got, want)
    assert want[0].sum() > 0.5 * n and want[-1].sum() > 0           # most particles start outside the planet and survive


def test_mirror_generator_equals_reference(ref):
    """The host mirror's uniform_pos_normal_speed generator (what pic_tracer_b200 traces) is bit-identical to the reference's."""
    from allscale_ipic3d_b200 import build
    mirror = C.CDLL(build.build_host())
    for seed, n in ((0, 50000), (77, 1000)):
        props, want = tracer_setup(ref, 16, n, seed)
        got = np.zeros_like(want)
        mirror.ipic_mirror_tracer_particles(C.byref(props), C.c_uint32(seed), C.c_uint64(n), dp(got))
        assert bits_equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("R,n,T,S", [(16, 200000, 51, 10), (8, 30000, 151, 10), (4, 5000, 20, 3)])
def test_gpu_trace_equals_oracle(ref, R, n, T, S):
    """ipic_trace (one CUDA thread per particle, integer atomics) == the oracle's frames, cell for cell."""
    from allscale_ipic3d_b200 import capi
    props, aos = tracer_setup(ref, R, n)
    got, seconds = capi.trace(props, aos, T, S)
    want = oracle_counts(props, aos, T, S)
    assert got.shape == want.shape and seconds > 0
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_gpu_trace_argument_errors(ref):
    from allscale_ipic3d_b200 import capi
    props, aos = tracer_setup(ref, 4, 10)
    with pytest.raises(capi.IpicError):
        capi.trace(props, aos, 10, 0)


@pytest.mark.gpu
def test_pic_tracer_executable_matches_reference(tmp_path):
    """`pic_tracer N T S R`: console protocol and every result_<stamp>.csv.NNNNNN frame identical to the reference's."""
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/pic_tracer_ref not built")
    a, b = tmp_path / "ours", tmp_path / "ref"
    a.mkdir(); b.mkdir()
    args = ["50000", "40", "10", "8"]
    o = subprocess.run([OUR_BIN] + args, cwd=str(a), capture_output=True, text=True, timeout=600)
    r = subprocess.run([REF_BIN] + args, cwd=str(b), capture_output=True, text=True, timeout=600)
    assert o.returncode == 0 and r.returncode == 0, o.stdout + o.stderr
    assert o.stdout.splitlines()[:2] == r.stdout.splitlines()[:2] and "Throughput:" in o.stdout
    frames_o = sorted(p.name.split(".")[-1] for p in a.glob("result_*.csv.*"))
    frames_r = sorted(p.name.split(".")[-1] for p in b.glob("result_*.csv.*"))
    assert frames_o == frames_r == ["%06d" % t for t in range(5)]
    for f in frames_o:
        assert next(a.glob("result_*.csv." + f)).read_text() == next(b.glob("result_*.csv." + f)).read_text()
