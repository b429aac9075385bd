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

	// "Dipole" / "Test"; anything else reads as "Test", like the parser's fallback (reference hdr/universe_properties.h:15-30)
	inline std::string getUseCaseName(const UseCase& useCase) { return useCase == UseCase::Dipole ? "Dipole" : "Test"; }
	inline std::ostream& operator<<(std::ostream& out, const UseCase& useCase) { return out << getUseCaseName(useCase); }

	struct UniverseProperties {
		UseCase useCase = UseCase::Dipole;
		coordinate_type size{ 1, 1, 1 };               // cells per dimension (cubic universes only, as in the reference)
		Vector3<double> cellWidth{ 1.0, 1.0, 1.0 };
		double dt = 1.0;
		double speedOfLight = 1.0;
		double planetRadius = 0.0;
		Vector3<double> objectCenter{ 0.0, 0.0, 0.0 };
		Vector3<double> origin{ 0.0, 0.0, 0.0 };       // front lower left corner of the domain
		Vector3<double> externalMagneticField{ 0, 0, 0 };
		int FieldOutputCycle = 100;
		int ParticleOutputCycle = 100;
		std::string outputFileBaseName;

		UniverseProperties() = default;

		UniverseProperties(const UseCase& useCase, const coordinate_type& size = { 1, 1, 1 }, const Vector3<double>& cellWidth = { 1.0, 1.0, 1.0 },
		                   double dt = 1.0, double speedOfLight = 1.0, double planetRadius = 0.0, const Vector3<double>& objectCenter = { 0.0, 0.0, 0.0 },
		                   const Vector3<double>& origin = { 0.0, 0.0, 0.0 }, const Vector3<double>& externalMagneticField = { 0, 0, 0 },
		                   int FieldOutputCycle = 100, int ParticleOutputCycle = 100) {
			this->useCase = useCase; this->size = size; this->cellWidth = cellWidth;
			this->dt = dt; this->speedOfLight = speedOfLight; this->planetRadius = planetRadius;
			this->objectCenter = objectCenter; this->origin = origin; this->externalMagneticField = externalMagneticField;
			this->FieldOutputCycle = FieldOutputCycle; this->ParticleOutputCycle = ParticleOutputCycle;
			validate();
		}

		// from an input deck: the box is centred on the object (reference universe_properties.h:92-113)
		explicit UniverseProperties(const Parameters& params) {
			useCase = params.useCase;
			dt = params.dt;
			speedOfLight = params.c;
			planetRadius = params.planetRadius;
			FieldOutputCycle = params.FieldOutputCycle;
			ParticleOutputCycle = params.ParticlesOutputCycle;
			const double centre[3] = { params.objectCenter.x, params.objectCenter.y, params.objectCenter.z };
			const std::int64_t cells[3] = { params.ncells.x, params.ncells.y, params.ncells.z };
			const double width[3] = { params.dspace.x, params.dspace.y, params.dspace.z };
			const double field[3] = { params.B1.x, params.B1.y, params.B1.z };
			for (int a = 0; a < 3; ++a) {
				size[a] = cells[a];
				cellWidth[a] = width[a];
				objectCenter[a] = centre[a];
				origin[a] = centre[a] - cells[a] * width[a] / 2.0;
				externalMagneticField[a] = field[a];
			}
		}

		void validate() const {
			assert(size.x > 0 && size.y > 0 && size.z > 0);
			assert(size.x == size.y && size.y == size.z);
			assert(cellWidth.x > 0 && cellWidth.y > 0 && cellWidth.z > 0);
			assert(dt > 0 && speedOfLight > 0 && planetRadius >= 0);
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
	};

	// the console block the drivers print (one "\t<label>: <value>" line per property)
	inline std::ostream& operator<<(std::ostream& out, const UniverseProperties& p) {
		out << "Universe properties:\n";
		auto line = [&out](const char* label, const auto& value) { out << "\t" << label << ": " << value << "\n"; };
		line("Use Case", p.useCase);
		line("Size", p.size);
		line("Cell width", p.cellWidth);
		line("Timestep", p.dt);
		line("Speed of light", p.speedOfLight);
		line("Planet radius", p.planetRadius);
		line("Object center", p.objectCenter);
		line("Origin of the domain", p.origin);
		line("External magnetic field", p.externalMagneticField);
		line("Fields output cycle", p.FieldOutputCycle);
		line("Particles output cycle", p.ParticleOutputCycle);
		return out << std::flush;
	}

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
