// Host C++ mirror: `UniverseProperties` and the cell geometry helpers (reference
// hdr/universe_properties.h:52-142).  `toAbi()` yields the POD the C ABI takes (ipic_props).
#pragma once
#include <cassert>
#include <ostream>
#include <string>

#include "ipic3d/app/parameters.h"
#include "ipic3d/app/vector.h"
#include "ipic3d_b200.h"

namespace ipic3d {

	inline std::ostream& operator<<(std::ostream& out, const UseCase& useCase) { return out << (useCase == UseCase::Dipole ? "Dipole" : "Test"); }

	struct UniverseProperties {
		UseCase useCase;
		coordinate_type size;                          // cells per dimension (cubic universes only, as in the reference)
		Vector3<double> cellWidth;
		double dt;
		double speedOfLight;
		double planetRadius;
		Vector3<double> objectCenter;
		Vector3<double> origin;                        // front lower left corner of the domain
		Vector3<double> externalMagneticField;
		int FieldOutputCycle;
		int ParticleOutputCycle;
		std::string outputFileBaseName;

		UniverseProperties(const UseCase& useCase = UseCase::Dipole, const coordinate_type& size = { 1, 1, 1 },
		                   const Vector3<double>& cellWidth = { 1.0, 1.0, 1.0 }, double dt = 1.0, double speedOfLight = 1.0,
		                   double planetRadius = 0.0, const Vector3<double>& objectCenter = { 0.0, 0.0, 0.0 },
		                   const Vector3<double>& origin = { 0.0, 0.0, 0.0 }, const Vector3<double>& externalMagneticField = { 0, 0, 0 },
		                   int FieldOutputCycle = 100, int ParticleOutputCycle = 100)
			: useCase(useCase), size(size), cellWidth(cellWidth), dt(dt), speedOfLight(speedOfLight), planetRadius(planetRadius),
			  objectCenter(objectCenter), origin(origin), externalMagneticField(externalMagneticField),
			  FieldOutputCycle(FieldOutputCycle), ParticleOutputCycle(ParticleOutputCycle) {
			assert(size.x > 0 && size.y > 0 && size.z > 0);
			assert(size.x == size.y && size.y == size.z);
			assert(cellWidth.x > 0 && cellWidth.y > 0 && cellWidth.z > 0);
			assert(dt > 0 && speedOfLight > 0 && planetRadius >= 0);
		}

		explicit UniverseProperties(const Parameters& params)
			: useCase(params.useCase), size({ params.ncells.x, params.ncells.y, params.ncells.z }),
			  cellWidth({ params.dspace.x, params.dspace.y, params.dspace.z }), dt(params.dt), speedOfLight(params.c),
			  planetRadius(params.planetRadius), objectCenter({ params.objectCenter.x, params.objectCenter.y, params.objectCenter.z }),
			  FieldOutputCycle(params.FieldOutputCycle), ParticleOutputCycle(params.ParticlesOutputCycle) {
			origin.x = params.objectCenter.x - params.ncells.x * params.dspace.x / 2.0;
			origin.y = params.objectCenter.y - params.ncells.y * params.dspace.y / 2.0;
			origin.z = params.objectCenter.z - params.ncells.z * params.dspace.z / 2.0;
			externalMagneticField = { params.B1.x, params.B1.y, params.B1.z };
		}

		// the POD mirror that crosses the C ABI
		ipic_props toAbi() const {
			ipic_props p{};
			for (int a = 0; a < 3; ++a) {
				p.size[a] = size[a]; p.cellWidth[a] = cellWidth[a]; p.objectCenter[a] = objectCenter[a];
				p.origin[a] = origin[a]; p.externalMagneticField[a] = externalMagneticField[a];
			}
			p.dt = dt; p.speedOfLight = speedOfLight; p.planetRadius = planetRadius;
			return p;
		}

		friend std::ostream& operator<<(std::ostream& out, const UniverseProperties& p) {
			return out << "Universe properties:\n\tUse Case: " << p.useCase << "\n\tSize: " << p.size << "\n\tCell width: " << p.cellWidth
			           << "\n\tTimestep: " << p.dt << "\n\tSpeed of light: " << p.speedOfLight << "\n\tPlanet radius: " << p.planetRadius
			           << "\n\tObject center: " << p.objectCenter << "\n\tOrigin of the domain: " << p.origin
			           << "\n\tExternal magnetic field: " << p.externalMagneticField << "\n\tFields output cycle: " << p.FieldOutputCycle
			           << "\n\tParticles output cycle: " << p.ParticleOutputCycle << std::endl;
		}
	};

	// origin + idx * width: position of the lower corner of a cell; the same formula locates field nodes
	inline Vector3<double> getOriginOfCell(const coordinate_type& pos, const UniverseProperties& properties) {
		const Vector3<double> idx{ (double)pos.x, (double)pos.y, (double)pos.z };
		return properties.origin + elementwiseProduct(idx, properties.cellWidth);
	}
	inline Vector3<double> getLocationForFields(const coordinate_type& pos, const UniverseProperties& properties) {
		return getOriginOfCell(pos, properties);
	}
	inline Vector3<double> getCenterOfCell(const coordinate_type& pos, const UniverseProperties& properties) {
		return getOriginOfCell(pos, properties) + properties.cellWidth / 2.0;
	}

} // namespace ipic3d
