"""Builds lib/libipic3d_b200.so with nvcc for sm_100a (in-tree; the .so travels to the GPU box)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["capi.cu", "particle_kernels.cu", "field_kernels.cu", "diag_kernels.cu", "tracer_kernels.cu", "init_kernels.cu", "comm.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-fmad=false",            # no FMA contraction: bit-faithful FP64 vs the reference's x86-64 GCC build
         "-Xcompiler", "-fPIC,-fvisibility=hidden", "-diag-suppress", "128"]


def needs_build(out):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "ipic3d_b200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


HOST_DIR = os.path.join(HERE, "..", "host")
HOST_SRC = os.path.join(HOST_DIR, "src", "host_capi.cc")
HOST_MAIN = os.path.join(HOST_DIR, "src", "ipic3d.cc")
HOST_TRACER = os.path.join(HOST_DIR, "src", "pic_tracer.cc")
HOST_INC = ["-I" + os.path.join(HOST_DIR, "include"), "-I" + os.path.join(HERE, "..", "include")]
HOST_FLAGS = ["-O2", "-std=c++17", "-pthread", "-ffp-contract=off"]      # no FMA contraction: the init must be bit-faithful


def _host_deps():
    inc = os.path.join(HOST_DIR, "include", "ipic3d", "app")
    return [os.path.join(inc, f) for f in os.listdir(inc)] + [HOST_SRC, HOST_MAIN, HOST_TRACER, os.path.join(HERE, "..", "include", "ipic3d_b200.h"), __file__]


def build_host(force=False):
    """libipic3d_host.so: host C++ mirror test hooks + the multithreaded :U:N generator used by bench.py and tests."""
    out = os.path.join(HERE, "lib", "libipic3d_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not force and os.path.exists(out) and all(os.path.getmtime(out) > os.path.getmtime(d) for d in _host_deps()):
        return out
    subprocess.check_call(["g++"] + HOST_FLAGS + HOST_INC + ["-fPIC", "-shared", "-fvisibility=hidden", "-o", out, HOST_SRC])
    return out


def build_driver(force=False):
    """host/bin/ipic3d_b200: the reference's `ipic3d <file.inp> | :U:N` driver on top of the host mirror and the C ABI."""
    out = os.path.join(HOST_DIR, "bin", "ipic3d_b200")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    lib = os.path.join(HERE, "lib", "libipic3d_b200.so")
    if not force and os.path.exists(out) and all(os.path.getmtime(out) > os.path.getmtime(d) for d in _host_deps() + [lib]):
        return out
    link = ["-L" + os.path.join(HERE, "lib"), "-lipic3d_b200", "-Wl,-rpath,$ORIGIN/../../allscale_ipic3d_b200/lib"]
    subprocess.check_call(["g++"] + HOST_FLAGS + HOST_INC + ["-o", out, HOST_MAIN] + link)
    # host/bin/pic_tracer_b200: the reference's `pic_tracer` (code/app/src/pic_tracer.cxx)
    subprocess.check_call(["g++"] + HOST_FLAGS + HOST_INC + ["-o", os.path.join(HOST_DIR, "bin", "pic_tracer_b200"), HOST_TRACER] + link)
    return out


def build(force=False, verbose=False, defines=(), name="libipic3d_b200.so"):
    out = os.path.join(HERE, "lib", name)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    build_host(force)
    if not force and not needs_build(out):
        if name == "libipic3d_b200.so":
            build_driver()
        return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-shared", "-o", out] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    subprocess.check_call(cmd)
    if name == "libipic3d_b200.so":
        build_driver(True)
    return out


if __name__ == "__main__":
    print(build(force=True, verbose=bool(os.environ.get("VERBOSE"))))
