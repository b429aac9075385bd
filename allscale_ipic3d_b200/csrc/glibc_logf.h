// Single-precision natural logarithm, bit for bit the one glibc >= 2.27 computes on x86-64 CPUs with FMA
// (sysdeps/ieee754/flt-32/e_logf.c = ARM optimized-routines logf, built a second time with -mfma and selected at
// run time by ifunc).  The reference's particle generator calls it from the tail of its ziggurat
// (hdr/ziggurat_normal_distribution.h:64-66: `log(uni())` on a float), so the device generator of the `.inp`
// universes (init_kernels.cu) needs exactly these bits; a library logf with a 1-ulp bound is not enough.
//
// Published algorithm: x = 2^k z with z in [OFF, 2 OFF); 16 subintervals with a centre c each, table of 1/c and log c;
// log x = log1p(z/c - 1) + log c + k ln 2 with a degree-3 polynomial, everything in double, one rounding to float.
// The products feeding a sum are contracted to fused multiply-adds the way GCC contracts them in the FMA build.
// tests/test_glibc_logf.py compares this restatement with the C library over ALL positive normal floats below 2.
// Usable from host code (fma from <cmath>) and from device code.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define IPIC_HD __host__ __device__ __forceinline__
#else
#define IPIC_HD inline
#endif

namespace ipic {

IPIC_HD float glibc_logf(float x) {
	// { 1/c, log c } of the 16 subintervals, ln 2, polynomial coefficients (logf_data.c)
	const double T[16][2] = {
		{ 0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2 }, { 0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2 },
		{ 0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2 },  { 0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3 },
		{ 0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3 }, { 0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3 },
		{ 0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4 }, { 0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4 },
		{ 0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5 }, { 0x1p+0, 0x0p+0 },
		{ 0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5 },  { 0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4 },
		{ 0x1.b2036576afce6p-1, 0x1.526e57720db08p-3 },  { 0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3 },
		{ 0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2 },  { 0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2 } };
	const double Ln2 = 0x1.62e42fefa39efp-1;
	const double A0 = -0x1.00ea348b88334p-2, A1 = 0x1.5575b0be00b6ap-2, A2 = -0x1.ffffef20a4123p-2;
	std::uint32_t ix;
	std::memcpy(&ix, &x, 4);
	if (ix == 0x3f800000u) return 0.0f;
	if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
		// zero, negative, infinite, NaN, subnormal: the generator never gets here (uni() is in (2e-8, 1)); library semantics
		if (ix * 2u == 0u) return -INFINITY;
		if (ix == 0x7f800000u) return x;
		if ((ix & 0x80000000u) || ix * 2u >= 0xff000000u) return NAN;
		const float scaled = x * 0x1p23f;
		std::memcpy(&ix, &scaled, 4);
		ix -= 23u << 23;
	}
	const std::uint32_t tmp = ix - 0x3f330000u;
	const int i = (int)((tmp >> (23 - 4)) % 16u);
	const int k = (std::int32_t)tmp >> 23;
	const std::uint32_t iz = ix - (tmp & 0xff800000u);
	float zf;
	std::memcpy(&zf, &iz, 4);
	const double z = (double)zf;
	const double r = fma(z, T[i][0], -1.0);
	const double y0 = fma((double)k, Ln2, T[i][1]);
	const double r2 = r * r;
	double y = fma(A1, r, A2);
	y = fma(A0, r2, y);
	y = fma(y, r2, y0 + r);
	return (float)y;
}

} // namespace ipic
