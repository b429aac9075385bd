// Host C++ mirror: `InitProperties` (reference hdr/init_properties.h:14-62) — what is needed only while a
// universe is being set up.
#pragma once
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
		double rhoInit = 1.0;

		InitProperties() = default;
		InitProperties(unsigned steps, const std::vector<Vector3<unsigned>>& ppc = {}, const std::vector<Vector3<double>>& drift = {},
		               const Vector3<double>& B0 = { 0, 0, 0 }, double rho = 1.0)
			: numSteps(steps), particlesPerCell(ppc), driftVelocity(drift), magneticField(B0), rhoInit(rho) {}

		explicit InitProperties(const Parameters& params) {
			numSteps = params.ncycles;
			for (int s = 0; s < params.ns; ++s) {
				driftVelocity.push_back({ params.u0[s], params.v0[s], params.w0[s] });
				particlesPerCell.push_back({ (unsigned)params.npcelx[s], (unsigned)params.npcely[s], (unsigned)params.npcelz[s] });
			}
			magneticField = { params.B0.x, params.B0.y, params.B0.z };
			rhoInit = params.rhoInit[0];
		}

		friend std::ostream& operator<<(std::ostream& out, const InitProperties& p) {
			auto list = [&out](const auto& v) { out << "["; for (std::size_t i = 0; i < v.size(); ++i) out << (i ? "," : "") << v[i]; out << "]"; };
			out << "InitProperties:\n\tNumber of time steps: " << p.numSteps << "\n\tNumber of particles per cell: ";
			list(p.particlesPerCell);
			out << "\n\tDrift velocity: ";
			list(p.driftVelocity);
			return out << "\n\tMagnetic fielde: " << p.magneticField << "\n\tCharge density: " << p.rhoInit << std::endl;
		}
	};

} // namespace ipic3d
