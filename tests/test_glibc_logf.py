"""The restatement of glibc's logf that the device generator of the `.inp` universes uses in the tail of the ziggurat
(allscale_ipic3d_b200/csrc/glibc_logf.h) against the C library itself, over ALL positive normal floats below 2
(the generator only ever takes logarithms of values in (2e-8, 1)).  CPU only; about ten seconds on eight cores."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROGRAM = r"""
#include <atomic>
#include <cstdio>
#include <cstring>
#include <math.h>
#include <thread>
#include <vector>
#include "glibc_logf.h"
int main() {
	std::atomic<unsigned long long> bad(0);
	std::vector<std::thread> pool;
	const unsigned lo = 0x00800000u, hi = 0x40000000u;
	const int T = (int)std::max(1u, std::thread::hardware_concurrency());
	for (int t = 0; t < T; t++) pool.emplace_back([&, t] {
		unsigned long long b = 0;
		for (unsigned long long u = lo + t; u < hi; u += T) {
			unsigned uu = (unsigned)u; float x; memcpy(&x, &uu, 4);
			const float a = logf(x), c = ipic::glibc_logf(x);
			if (memcmp(&a, &c, 4) != 0) b++;
		}
		bad += b;
	});
	for (auto& th : pool) th.join();
	printf("%llu\n", bad.load());
	return 0;
}
"""


def test_logf_restatement_matches_libm_exhaustively(tmp_path):
    src = tmp_path / "check.cc"
    src.write_text(PROGRAM)
    exe = tmp_path / "check"
    # -mfma only makes fma() a single instruction; contraction stays off: the header spells every fused operation out
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-mfma", "-pthread",
                           "-I" + os.path.join(ROOT, "allscale_ipic3d_b200", "csrc"), str(src), "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600, check=True).stdout.strip()
    assert out == "0", "%s of the 1 065 353 216 floats differ from the C library's logf" % out
