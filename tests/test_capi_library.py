"""CPU-side checks of the product boundary: the C-ABI library is built in-tree, exports every symbol that
include/ipic3d_b200.h declares, and fails loudly (no fallback) when no CUDA device is present."""
import ctypes
import os
import re

import pytest

import allscale_ipic3d_b200 as ipic
from allscale_ipic3d_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "ipic3d_b200.h")).read()
    return sorted(set(re.findall(r"IPIC_API\s+[\w\s\*]+?\b(ipic_\w+)\s*\(", text)))


def test_library_is_in_tree_and_loads():
    assert os.path.exists(ipic.lib_path()), "run __graft_entry__.build() first"
    lib = ipic.load_library()
    assert b"sm_100a" in lib.ipic_version()


def test_every_declared_symbol_is_exported():
    lib = ipic.load_library()
    declared = header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "%s declared in include/ipic3d_b200.h but not exported" % name
    assert sorted(capi.API_SYMBOLS) == declared


def test_props_layout_matches_header():
    # 3*int64 + 3*double + 3 doubles + 3*3 doubles
    assert ctypes.sizeof(ipic.Props) == 8 * (3 + 3 + 3 + 9)


def test_product_never_touches_the_oracle():
    """The shipped package must not import, link or dlopen anything under oracle/."""
    pkg = os.path.join(ROOT, "allscale_ipic3d_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp", ".hpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.replace("no CPU fallback", "").lower() or f == "__init__.py" and "oracle/" not in text, \
                    "%s mentions the oracle" % os.path.join(dirpath, f)


def test_no_cuda_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(ipic.IpicError) as e:
        ipic.Context(ipic.Props.make(4))
    assert e.value.code == -2
    # the context-free entry points as well: the tracer and the device self-test have no CPU path either
    import numpy as np
    from allscale_ipic3d_b200 import capi
    with pytest.raises(ipic.IpicError) as e:
        capi.trace(ipic.Props.make(4, planetRadius=0.1), np.ones((3, 8)), 5, 1)
    assert e.value.code == -2
    with pytest.raises(ipic.IpicError):
        capi.selftest(1 << 10)


def test_host_tree_never_touches_the_oracle():
    """Neither do the host C++ mirror and its drivers."""
    for sub in ("host", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, sub)):
            for f in files:
                if f.endswith((".h", ".cc", ".cpp", ".hpp")):
                    text = open(os.path.join(dirpath, f), errors="ignore").read()
                    assert "oracle" not in text.lower(), \
                        "%s refers to the oracle tree" % os.path.join(dirpath, f)


def test_time_loop_never_calls_the_host_scalar_versions():
    """The mirror keeps host-callable scalar versions of a few per-node functions for known-answer checks (cell.h:
    projectToDensityField, trilinearInterpolationF2P; field.h: curls, solvers, ghost update).  The time loop and the drivers
    must not use them -- simulateSteps steps on the GPU or fails -- and movers / export / import have no host version."""
    import re
    scalar = r"\b(projectToDensityField|trilinearInterpolationF2P|solveFieldForward|solveFieldLeapfrog|solveFieldStatically|" \
             r"updateFieldsOnBoundaries|computeCurlB|computeCurlE|interpC2N)\s*\("
    host = os.path.join(ROOT, "host")
    for rel in ("include/ipic3d/app/simulator.h", "include/ipic3d/app/benchmark.h", "include/ipic3d/app/earth_deck.h",
                "src/ipic3d.cc", "src/pic_tracer.cc"):
        text = open(os.path.join(host, rel)).read()
        code = "\n".join(line.split("//")[0] for line in text.splitlines())
        assert not re.search(scalar, code), "%s calls a host scalar version" % rel
    for dirpath, _, files in os.walk(os.path.join(host, "include")):
        for f in files:
            text = open(os.path.join(dirpath, f)).read()
            code = "\n".join(line.split("//")[0] for line in text.splitlines())
            assert not re.search(r"\b(moveParticles|exportParticles|importParticles)\s*\(", code), "%s defines a host mover/export/import" % f
