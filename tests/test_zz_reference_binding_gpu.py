"""GPU leg of tests/test_reference_binding.py (kept in its own, last-sorted file): the reference's unmodified driver built
with the reference-side binding (oracle/_ref/ipic3d_ref_b200) runs `ipic3d test.inp` with its time loop on the GPU through
libipic3d_b200.so and writes the same test.out as the reference built here -- per-cell counts of all 512 cells after 30
steps, in the reference's own writer's format."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
BOUND = os.path.join(REF, "ipic3d_ref_b200")
DECK = os.path.join(ROOT, "tests", "golden", "inputs", "test.inp")


@pytest.mark.gpu
def test_reference_driver_through_the_binding_on_the_gpu(tmp_path):
    if not os.path.exists(BOUND):
        pytest.skip("oracle/_ref/ipic3d_ref_b200 not built (reference sources absent at build time)")
    a, b = tmp_path / "bound", tmp_path / "ref"
    a.mkdir(); b.mkdir()
    got = subprocess.run([BOUND, DECK], cwd=a, capture_output=True, text=True, timeout=900)
    assert got.returncode == 0, got.stderr[-1000:]
    want = subprocess.run([os.path.join(REF, "ipic3d_ref"), DECK], cwd=b, capture_output=True, text=True, timeout=900)
    assert want.returncode == 0
    assert (a / "test.out").read_text() == (b / "test.out").read_text()
