// Host C++ mirror: `InitProperties` (reference hdr/init_properties.h:14-62) — what is needed only while a
// universe is being set up.
#pragma once
#include <algorithm>
#include <cstdint>
#include <ostream>
#include <vector>

#include "ipic3d/app/parameters.h"
#include "ipic3d/app/vector.h"

namespace ipic3d {

	struct InitProperties {
		std::uint64_t numSteps = 1;
		std::vector<Vector3<unsigned>> particlesPerCell;     // per species, per dimension
		std::vector<Vector3<double>> driftVelocity;          // per species
		Vector3<double> magneticField{ 0, 0, 0 };            // B0
		double rhoInit = 1.0;                                // charge density on the nodes

		InitProperties() = default;

		InitProperties(unsigned steps, const std::vector<Vector3<unsigned>>& perCell = {}, const std::vector<Vector3<double>>& drift = {},
		               const Vector3<double>& B0 = { 0, 0, 0 }, double rho = 1.0) {
			numSteps = steps; particlesPerCell = perCell; driftVelocity = drift; magneticField = B0; rhoInit = rho;
		}

		// species-wise entries of an input deck
		explicit InitProperties(const Parameters& deck) : numSteps(deck.ncycles), magneticField{ deck.B0.x, deck.B0.y, deck.B0.z }, rhoInit(deck.rhoInit[0]) {
			const std::size_t species = (std::size_t)std::max(deck.ns, 0);
			particlesPerCell.reserve(species);
			driftVelocity.reserve(species);
			for (std::size_t s = 0; s < species; ++s) {
				particlesPerCell.push_back({ (unsigned)deck.npcelx[s], (unsigned)deck.npcely[s], (unsigned)deck.npcelz[s] });
				driftVelocity.push_back({ deck.u0[s], deck.v0[s], deck.w0[s] });
			}
		}
	};

	namespace detail {
		template<typename T>
		void printList(std::ostream& out, const std::vector<T>& values) {
			out << "[";
			for (std::size_t i = 0; i < values.size(); ++i) out << (i ? "," : "") << values[i];
			out << "]";
		}
	}

	// the console block the drivers print
	inline std::ostream& operator<<(std::ostream& out, const InitProperties& p) {
		out << "InitProperties:\n\tNumber of time steps: " << p.numSteps << "\n\tNumber of particles per cell: ";
		detail::printList(out, p.particlesPerCell);
		out << "\n\tDrift velocity: ";
		detail::printList(out, p.driftVelocity);
		return out << "\n\tMagnetic fielde: " << p.magneticField << "\n\tCharge density: " << p.rhoInit << std::endl;
	}

} // namespace ipic3d
