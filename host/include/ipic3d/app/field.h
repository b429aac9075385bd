// Host C++ mirror: field containers and their initialisation (reference hdr/field.h:14-38, :43-160,
// :252-264).  The per-timestep field work of the reference's field.h (updateFieldsOnBoundaries,
// solveFieldForward, solveFieldLeapfrog) runs in the CUDA kernels behind `simulateSteps`
// (csrc/field_kernels.cu); node-level host versions for unit-level use are at the end of this file.
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

	// ---- node-level stencils of the explicit field update, host-callable ------------------------------------------------
	// The reference exposes the pieces of its field update as free functions that act on ONE node (hdr/field.h:162-247,
	// :272-402, :407-440) and unit-tests them that way; the time loop (`simulateSteps`) runs the same arithmetic in the CUDA
	// kernels of csrc/field_kernels.cu for all nodes and never calls the functions below.  They exist so that code written
	// against the reference's node-level API (its own field_test.cc included) compiles and behaves the same here.
	// All curls are built from one operator: the mean difference quotient of a component along an axis over the four
	// parallel edges of the 2x2x2 neighbourhood that hangs below a node (cell centres) or above a cell centre (nodes).
	// The reference writes each of these as one long alternating sum; the order of the edges and the left-to-right
	// association are kept (x and y differences: edges in lexicographic order of the two other offsets; z differences:
	// the y offset varies slowest), so the results agree with the reference bit for bit.

	namespace detail {
		// corner(i,j,k), i,j,k in {0,1}: value at offset +(i,j,k) for `above` = true, at -(i,j,k) otherwise
		template<typename Corner>
		double meanDifference(const Corner& corner, bool above, int axis, int component, double width) {
			static const int slow[3] = { 1, 0, 1 }, fast[3] = { 2, 2, 0 };
			double chain = 0.0;
			for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) {
				int hi[3], lo[3];                                             // "hi" = the larger coordinate along `axis`
				hi[axis] = above ? 1 : 0; lo[axis] = above ? 0 : 1;
				hi[slow[axis]] = lo[slow[axis]] = a;
				hi[fast[axis]] = lo[fast[axis]] = b;
				const double h = corner(hi[0], hi[1], hi[2])[component], l = corner(lo[0], lo[1], lo[2])[component];
				chain = (a == 0 && b == 0) ? h - l : chain + h - l;
			}
			return .25 * chain / width;
		}
		template<typename Corner>
		Vector3<double> curlOver(const Corner& corner, bool above, const Vector3<double>& w) {
			return { meanDifference(corner, above, 1, 2, w.y) - meanDifference(corner, above, 2, 1, w.z),
			         meanDifference(corner, above, 2, 0, w.z) - meanDifference(corner, above, 0, 2, w.x),
			         meanDifference(corner, above, 0, 1, w.x) - meanDifference(corner, above, 1, 0, w.y) };
		}
	}

	// curl of the cell-centred Bc on node `pos` (the eight cells below it)
	inline void computeCurlB(const UniverseProperties& universeProperties, const coordinate_type& pos, const BcField& bcfield, Vector3<double>& curl) {
		curl = detail::curlOver([&](int i, int j, int k) -> const Vector3<double>& { return bcfield[{ pos.x - i, pos.y - j, pos.z - k }].Bc; }, false, universeProperties.cellWidth);
	}
	// curl of the nodal E on cell centre `pos` (the eight nodes above it)
	inline void computeCurlE(const UniverseProperties& universeProperties, const coordinate_type& pos, const Field& field, Vector3<double>& curl) {
		curl = detail::curlOver([&](int i, int j, int k) -> const Vector3<double>& { return field[{ pos.x + i, pos.y + j, pos.z + k }].E; }, true, universeProperties.cellWidth);
	}
	// B on a node = mean of Bc over the eight cells below it
	inline void interpC2N(const coordinate_type& pos, const BcField& bcfields, Field& fields) {
		Vector3<double> sum(0.0);
		forEachCoordinate(coordinate_type(0), coordinate_type(2), [&](const coordinate_type& d) { sum += bcfields[pos - d].Bc; });
		fields[pos].B = .125 * sum;
	}

	// the dipole case keeps its initial fields: nothing to do (hdr/field.h:272-289)
	inline void solveFieldStatically(const UniverseProperties&, const coordinate_type&, Field&) {}

	// forward (Euler) update of node `pos`: E from curl Bc and J, Bc from curl E (only where `pos` is lexicographically below
	// the last cell of the Bc grid: Vector3's `<`, as in the reference), B re-averaged
	inline void solveFieldForward(const UniverseProperties& universeProperties, const coordinate_type& pos, CurrentDensity& density, Field& field, BcField& bcfield) {
		const double c = universeProperties.speedOfLight, dt = universeProperties.dt;
		Vector3<double> curlB;
		computeCurlB(universeProperties, pos, bcfield, curlB);
		field[pos].E += (c * curlB - density[pos - coordinate_type(1)].J) * dt;       // J lives on the nodes without the ghost layer
		if (pos < bcfield.size() - coordinate_type(1)) {
			Vector3<double> curlE;
			computeCurlE(universeProperties, pos, field, curlE);
			bcfield[pos].Bc -= c * curlE * dt;
		}
		interpC2N(pos, bcfield, field);
	}

	// leapfrog update of node `pos` (the reference's active branch, hdr/field.h:353-372): one-sided differences, backward for
	// the curl of Bc, forward for the curl of E
	inline void solveFieldLeapfrog(const UniverseProperties& universeProperties, const coordinate_type& pos, CurrentDensity& density, Field& field, BcField& bcfield) {
		const double c = universeProperties.speedOfLight, dt = universeProperties.dt;
		const Vector3<double>& w = universeProperties.cellWidth;
		const auto step = [](int axis) { coordinate_type e(0); e[axis] = 1; return e; };
		const auto back = [&](int axis, int comp) { return (bcfield[pos].Bc[comp] - bcfield[pos - step(axis)].Bc[comp]) / w[axis]; };
		const Vector3<double> J = density[pos - coordinate_type(1)].J;
		Vector3<double>& E = field[pos].E;
		E.x = E.x + dt * (c * c * (back(1, 2) - back(2, 1)) - J.x);
		E.y = E.y + dt * (c * c * (back(2, 0) - back(0, 2)) - J.y);
		E.z = E.z + dt * (c * c * (back(0, 1) - back(1, 0)) - J.z);
		if (pos < bcfield.size() - coordinate_type(1)) {
			const auto fwd = [&](int axis, int comp) { return (field[pos + step(axis)].E[comp] - E[comp]) / w[axis]; };
			Vector3<double>& B = bcfield[pos].Bc;
			B.x = B.x - dt * (fwd(1, 2) - fwd(2, 1));
			B.y = B.y - dt * (fwd(2, 0) - fwd(0, 2));
			B.z = B.z - dt * (fwd(0, 1) - fwd(1, 0));
		}
	}

	// periodic ghost layers, faces only: along every axis the ghost plane 0 takes the last interior plane and the ghost plane
	// `end` the first interior one, for in-face indices 1 .. end-1 (edges and corners stay as they are; cubic grids, as in the
	// reference, hdr/field.h:407-440)
	namespace detail {
		template<typename GridT>
		void wrapFaces(GridT& grid) {
			const std::int64_t end = grid.size().x - 1;
			for (std::int64_t a = 1; a < end; ++a) for (std::int64_t b = 1; b < end; ++b) {
				grid[{ 0, a, b }] = grid[{ end - 1, a, b }];
				grid[{ end, a, b }] = grid[{ 1, a, b }];
				grid[{ a, 0, b }] = grid[{ a, end - 1, b }];
				grid[{ a, end, b }] = grid[{ a, 1, b }];
				grid[{ a, b, 0 }] = grid[{ a, b, end - 1 }];
				grid[{ a, b, end }] = grid[{ a, b, 1 }];
			}
		}
	}
	inline void updateFieldsOnBoundaries(Field& field, BcField& bcfield) {
		detail::wrapFaces(field);
		detail::wrapFaces(bcfield);
	}

} // namespace ipic3d
