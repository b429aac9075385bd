// Host C++ mirror: field containers and their initialisation (reference hdr/field.h:14-38, :43-160,
// :252-264).  The per-timestep field work of the reference's field.h (updateFieldsOnBoundaries,
// solveFieldForward, solveFieldLeapfrog) runs in the CUDA kernels behind `simulateSteps`
// (csrc/field_kernels.cu).
#pragma once
#include <cmath>
#include <fstream>
#include <string>

#include "ipic3d/app/init_properties.h"
#include "ipic3d/app/universe_properties.h"
#include "ipic3d/app/vector.h"

namespace ipic3d {

	struct FieldNode {
		Vector3<double> E;          // electric field on the node
		Vector3<double> B;          // magnetic field on the node
		Vector3<double> Bext;       // external (dipole) magnetic field on the node
	};
	struct BcFieldCell {
		Vector3<double> Bc;         // magnetic field at the cell centre
	};
	struct DensityNode {
		Vector3<double> J;          // current density on the node
	};
	static_assert(sizeof(FieldNode) == 72 && sizeof(BcFieldCell) == 24 && sizeof(DensityNode) == 24, "field records must match the C ABI");

	using Field = Grid3<FieldNode>;                  // (size + 3)^3: nodes plus one ghost layer on each side
	using BcField = Grid3<BcFieldCell>;              // (size + 2)^3: cell centres plus ghost layer
	using CurrentDensity = Grid3<DensityNode>;       // (size + 1)^3: nodes

	// E = -(v_drift x B0); Bext = field of a z-aligned dipole of moment -Bext.z R^3 outside the planet, 0 inside; B = B0 + Bext
	inline FieldNode getDipoleFieldAt(const Vector3<double>& location, const InitProperties& initProperties, const UniverseProperties& universeProperties) {
		FieldNode node;
		node.E = -1.0 * crossProduct(initProperties.driftVelocity[0], initProperties.magneticField);
		const double a = universeProperties.planetRadius;
		const Vector3<double> d = location - universeProperties.objectCenter;
		const double r2 = sumOfSquares(d);
		if (r2 > a * a) {
			const double fac1 = -universeProperties.externalMagneticField.z * std::pow(a, 3) / std::pow(r2, 2.5);
			node.Bext.x = 3.0 * d.x * d.z * fac1;
			node.Bext.y = 3.0 * d.y * d.z * fac1;
			node.Bext.z = (2.0 * d.z * d.z - d.x * d.x - d.y * d.y) * fac1;
		} else {
			node.Bext = { 0.0, 0.0, 0.0 };
		}
		node.B = initProperties.magneticField + node.Bext;
		return node;
	}

	// nodes [1, size+2) get the dipole field at origin + (idx-1)*width; the outermost layer stays zero
	inline Field initFields(const InitProperties& initProperties, const UniverseProperties& universeProperties) {
		const coordinate_type one(1);
		Field fields(universeProperties.size + coordinate_type(3));
		pfor(one, universeProperties.size + coordinate_type(2), [&](const coordinate_type& cur) {
			fields[cur] = getDipoleFieldAt(getLocationForFields(cur - one, universeProperties), initProperties, universeProperties);
		});
		return fields;
	}

	// Bc of a cell = mean of B over its 8 corner nodes, summed in (i,j,k) order
	inline void interpN2C(const coordinate_type& pos, const Field& fields, BcField& bcfields) {
		Vector3<double> sum(0);
		for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) for (int k = 0; k < 2; ++k)
			sum += fields[coordinate_type{ pos.x + i, pos.y + j, pos.z + k }].B;
		bcfields[pos].Bc = .125 * sum;
	}

	inline BcField initBcFields(const UniverseProperties& universeProperties, const Field& field) {
		BcField bcfield(universeProperties.size + coordinate_type(2));
		forEachCoordinate(coordinate_type(1), universeProperties.size + coordinate_type(1), [&](const coordinate_type& cur) { interpN2C(cur, field, bcfield); });
		return bcfield;
	}

	// field energy over the nodes [1, size-1) of the field grid: sum of 0.5 * cell volume * |F|^2 / (4 pi), F chosen by the
	// accessor (reference hdr/field.h:443-459); sequential fold in lexicographic order
	template<typename Accessor>
	double getFieldEnergy(const Field& field, const UniverseProperties& universeProperties, const Accessor& accessor) {
		const double vol = 0.5 * universeProperties.cellWidth.x * universeProperties.cellWidth.y * universeProperties.cellWidth.z;
		const double fourPI = 16.0 * std::atan(1.0);
		double res = 0.0;
		forEachCoordinate(coordinate_type(1), field.size() - coordinate_type(1), [&](const coordinate_type& index) {
			res += vol * sumOfSquares(accessor(field, index)) / (fourPI);
		});
		return res;
	}

	// dump of all field values: "[Nx,Ny,Nz]", then "i,j,k:E|B|Bext" per node, blank line; the same for Bc
	// (reference hdr/field.h:464-530: two consecutive dumps into the entry of that name; never called by the reference's
	// drivers or tests)
	inline void outputFieldGrids(const Field& field, const BcField& bcField, const std::string& outputFilename) {
		{
			std::ofstream out(outputFilename);
			out << field.size() << "\n";
			forEachCoordinate(coordinate_type(0), field.size(), [&](const coordinate_type& p) {
				out << p.x << "," << p.y << "," << p.z << ":" << field[p].E << "|" << field[p].B << "|" << field[p].Bext << "\n";
			});
			out << "\n";
		}
		{
			std::ofstream out(outputFilename, std::ios::app);
			out << bcField.size() << "\n";
			forEachCoordinate(coordinate_type(0), bcField.size(), [&](const coordinate_type& p) {
				out << p.x << "," << p.y << "," << p.z << ":" << bcField[p].Bc << "\n";
			});
			out << "\n";
		}
	}

	inline CurrentDensity initCurrentDensity(const UniverseProperties& universeProperties) {
		return CurrentDensity(universeProperties.size + coordinate_type(1));     // value-initialised: J = 0
	}

} // namespace ipic3d
