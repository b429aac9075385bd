"""B200-native particle hot path of AllScale iPIC3D (mover + current deposit + cell-to-cell migration).

The product is the C-ABI shared library ``lib/libipic3d_b200.so`` (hand-written sm_100a CUDA, see
``csrc/`` and ``include/ipic3d_b200.h``) plus the C++ host mirror of the reference's data model under
``host/``.  This Python package is only the thin ctypes binding used by tests and ``bench.py``.
There is no CPU fallback: without the built library or without a CUDA device every call fails loudly.
"""
from .capi import (Context, Group, IpicError, Props, lib_path, load_library,  # noqa: F401
                   MOVER_DIPOLE_ANALYTIC, MOVER_INTERPOLATED, PROJECT_CURRENT, SOLVER_FORWARD,
                   SOLVER_LEAPFROG, UPDATE_BOUNDARIES, SKIP_MIGRATION)

__all__ = ["Context", "Group", "IpicError", "Props", "lib_path", "load_library"]
