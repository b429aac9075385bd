// Host C++ mirror: what the reference's `main` does with an input deck before the time loop
// (code/app/src/ipic3d.cxx:32-83).  A deck only contributes the grid, the particles per cell and the cycle count: the
// physics is overridden by hard-coded Earth-magnetosphere values:
//
//   dt 0.01 s, c 299 792 458 m/s, planet radius R = 6 378 137 m, the deck's cell width taken in units of R,
//   planet at the centre of the box, dipole moment from B = 3.07e-5 T along z,
//   10 MeV protons: positions uniform in the inner 75 % box, velocity components normal with sigma = the relativistic
//   speed of a 10 MeV proton, generator seed 0.
//
// Shared by the driver (host/src/ipic3d.cc) and the test hooks (host/src/host_capi.cc).
#pragma once
#include <cmath>
#include <cstdint>
#include <string>

#include "ipic3d/app/cell.h"
#include "ipic3d/app/init_properties.h"
#include "ipic3d/app/parameters.h"
#include "ipic3d/app/universe.h"

namespace ipic3d {

	struct EarthDeck {
		Parameters params;
		InitProperties init;
		UniverseProperties properties;
		int numParticles;                                  // an int in the reference as well (it overflows for large.inp)
		std::int64_t numParticles64;                       // the count the deck means: cells x particles per cell, exact

		explicit EarthDeck(const std::string& path, const std::string& baseName = std::string())
			: params(path), init(params), properties(params), numParticles(0), numParticles64(0) {
			constexpr double earthRadius = 6378137;        // m
			properties.outputFileBaseName = baseName;
			properties.dt = 0.01;
			properties.speedOfLight = 299792458;
			properties.planetRadius = earthRadius;
			properties.cellWidth *= properties.planetRadius;
			properties.objectCenter = { 0.0, 0.0, 0.0 };
			properties.origin = properties.objectCenter - boxSize() / 2.0;
			properties.externalMagneticField = { 0.0, 0.0, 3.07e-5 };     // T
			properties.useCase = UseCase::Dipole;
			const auto& perCell = init.particlesPerCell[0];
			numParticles64 = properties.size.x * properties.size.y * properties.size.z * (std::int64_t)perCell.x * (std::int64_t)perCell.y * (std::int64_t)perCell.z;
			// `int numParticles = size.x * size.y * size.z; numParticles = numParticles * npcelx * npcely * npcelz;` (ipic3d.cxx:70-71)
			// wraps around for large.inp (160^3 x 12^3 = 7 077 888 000): reproduced with defined (unsigned) arithmetic
			numParticles = (int)(std::uint32_t)(std::uint64_t)numParticles64;
		}

		// does the reference's 32-bit particle count still say what the deck means?  (false for large.inp: the reference cannot run it)
		bool countFitsReference() const { return numParticles64 == (std::int64_t)numParticles; }

		Vector3<double> boxSize() const { return elementwiseProduct(properties.cellWidth, properties.size); }

		// speed of a proton of 10 MeV kinetic energy
		double thermalSpeed() const {
			const double e = 1.602176565e-19, K = 1e7 * e, m = 1.672621777e-27;
			const double c = properties.speedOfLight;
			return c / std::sqrt(1.0 + (m * c * c) / K);
		}

		// 10 MeV protons in the inner 75 % box, generator seed 0 (ipic3d.cxx:72-83)
		distribution::uniform_pos_normal_speed<> protons() const {
			const double v = thermalSpeed();
			const auto low = properties.origin + 0.125 * boxSize();
			const auto high = low + 0.75 * boxSize();
			return distribution::uniform_pos_normal_speed<>(low, high, Vector3<double>{ 0, 0, 0 }, Vector3<double>{ v, v, v });
		}

		Universe createUniverse() const { return createUniverseFromDistribution(properties, init, (std::uint64_t)numParticles64, protons()); }

		// the same universe without its particles (grids only): for runs that create the particles on the device
		// (populateOnDevice in simulator.h) and never need them on the host
		Universe createUniverseWithoutParticles() const { return detail::assembleUniverse(properties, init, Cells(properties.size)); }

		// The same universe piecewise, for callers that never hold all of it (one z-slab per GPU / per process):
		// phase 1 of initCells once (cheap, sequential, identical everywhere), then any cell on its own.
		std::vector<std::uint64_t> cellPopulations() const { return detail::approximatePopulation(properties, (std::uint64_t)numParticles64, protons()); }
		void populateCell(const coordinate_type& pos, std::uint64_t n, Particle* out) const { detail::populateCell(properties, pos, n, protons(), out); }
	};

} // namespace ipic3d
