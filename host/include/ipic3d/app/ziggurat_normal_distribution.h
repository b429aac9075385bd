// Host C++ mirror: the ziggurat normal generator that seeds particle velocities of the `.inp` runs
// (reference hdr/ziggurat_normal_distribution.h:14-125; Marsaglia & Tsang's RNOR with a 64-bit SHR3
// state).  Every parity run starts from universes built with this generator, so the arithmetic types are
// part of the contract and are spelled out: the layer index and the integer draw are the LOW 32 BITS of
// the 64-bit SHR3 sum, the accept test compares |draw| (as double) with a 64-bit table entry, samples
// are float, `uni()` is float, tail/wedge logarithms are single precision (the reference calls the
// overloaded log on a float), the wedge test exponent is double.
#pragma once
#include <cmath>
#include <cstdint>

namespace ipic3d {

	class ziggurat_normal_distribution {
		std::uint64_t state = 123456789;
		std::uint64_t edge[128];          // kn
		float width[128];                 // wn
		float height[128];                // fn

		std::uint64_t shr3() {
			const std::uint64_t before = state;
			state ^= state << 13;
			state ^= state >> 17;
			state ^= state << 5;
			return before + state;
		}

		float uni() { return (float)(.5 + (std::int32_t)shr3() * .2328306e-9); }

	public:
		explicit ziggurat_normal_distribution(std::uint64_t seed = 0) {
			const double m1 = 2147483648.0;
			double dn = 3.442619855899, tn = dn;
			const double vn = 9.91256303526217e-3;
			state ^= seed;
			const double q = vn / std::exp(-.5 * dn * dn);
			edge[0] = (std::uint64_t)((dn / q) * m1);
			edge[1] = 0;
			width[0] = (float)(q / m1);
			width[127] = (float)(dn / m1);
			height[0] = 1.;
			height[127] = (float)std::exp(-.5 * dn * dn);
			for (int i = 126; i >= 1; --i) {
				dn = std::sqrt(-2. * std::log(vn / dn + std::exp(-.5 * dn * dn)));
				edge[i + 1] = (std::uint64_t)((dn / tn) * m1);
				tn = dn;
				height[i] = (float)std::exp(-.5 * dn * dn);
				width[i] = (float)(dn / m1);
			}
		}

		float operator()() {
			const float r = 3.442620f;
			std::int32_t hz = (std::int32_t)shr3();
			std::uint32_t iz = (std::uint32_t)hz & 127u;
			if (std::fabs((double)hz) < (double)edge[iz]) return (float)hz * width[iz];
			for (;;) {
				if (iz == 0) {                                   // base strip: sample the tail
					float x, y;
					do {
						x = (float)(-0.2904764 * std::log(uni()));   // float log, double product, rounded to float
						y = -std::log(uni());
					} while (!(x * x <= y + y));
					return hz <= 0 ? -r - x : +r + x;
				}
				const float x = (float)hz * width[iz];           // wedge
				if (height[iz] + uni() * (height[iz - 1] - height[iz]) < std::exp(-0.5 * x * x)) return x;
				hz = (std::int32_t)shr3();
				iz = (std::uint32_t)hz & 127u;
				if (std::fabs((double)hz) < (double)edge[iz]) return (float)hz * width[iz];
			}
		}
	};

} // namespace ipic3d
