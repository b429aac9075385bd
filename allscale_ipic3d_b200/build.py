"""Builds lib/libipic3d_b200.so with nvcc for sm_100a (in-tree; the .so travels to the GPU box)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["capi.cu", "particle_kernels.cu", "field_kernels.cu", "comm.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-fmad=false",            # no FMA contraction: bit-faithful FP64 vs the reference's x86-64 GCC build
         "-Xcompiler", "-fPIC,-fvisibility=hidden", "-diag-suppress", "128"]


def needs_build(out):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "ipic3d_b200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    out = os.path.join(HERE, "lib", "libipic3d_b200.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not force and not needs_build(out):
        return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-shared", "-o", out] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(force=True, verbose=bool(os.environ.get("VERBOSE"))))
