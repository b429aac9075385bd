// The reference-side binding: a replacement for code/app/include/ipic3d/app/simulator.h of AllScale iPIC3D that keeps the
// reference's declarations (simulateSteps / simulateStep, DurationMeasurement, the detail:: policy functors,
// writeOutputHeader / writeOutputData) and runs the time loop through the C ABI of libipic3d_b200.so instead of the
// reference's pfor loops (hdr/simulator.h:127-242).  Everything else of the reference -- main (src/ipic3d.cxx), the .inp
// parser, initCells, the benchmark driver, the .out writer, the AllScale Grid / Vector types -- is used as it is: put this
// directory in front of the reference's include path (or copy the file over the reference's) and link libipic3d_b200.so.
//
// This is the binding INTEGRATION.md section 2 describes, written against the reference's OWN headers (AllScale API types);
// host/include/ipic3d/app/simulator.h is the same stub on top of this repository's host mirror, with the resident mode and
// the multi-GPU group added.  The repository's test build compiles the reference's unmodified driver with it
// (tests/test_reference_binding.py).
#pragma once

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <type_traits>
#include <vector>

#include "allscale/api/core/io.h"

#include "ipic3d/app/cell.h"
#include "ipic3d/app/field.h"
#include "ipic3d/app/transfer_buffer.h"
#include "ipic3d/app/universe.h"

#include "ipic3d_b200.h"

namespace ipic3d {

	template<typename T>
	using Grid = allscale::api::user::data::Grid<T,3>;

	// the policy functors select kernels: their operator() cannot run on the device, so they are tags here
	namespace detail {
		struct default_particle_to_field_projector {};
		struct default_field_solver {};
		struct forward_field_solver {};
		struct leapfrog_field_solver {};
		struct default_particle_mover {};
	}

	struct DurationMeasurement {
		double firstStep;
		double remainingSteps;
	};

	// the reference's log lines (hdr/simulator.h:62-115): unchanged host code
	template<typename StreamObject>
	void writeOutputHeader(StreamObject& streamObject) {
		streamObject << "Cycle \t Total Moment \t E energy \t B energy \t Total KE \n";
	}

	namespace b200 {

		inline void check(int rc, const char* what) {
			if (rc == IPIC_OK) return;
			std::cerr << "ipic3d_b200: " << what << " failed (" << rc << "): " << ipic_last_error() << std::endl;
			std::exit(EXIT_FAILURE);
		}

		static_assert(sizeof(Particle) == 8 * sizeof(double), "Particle must be the 8-double record of the C ABI (hdr/particle.h:9-19)");
		static_assert(sizeof(FieldNode) == 9 * sizeof(double) && sizeof(BcFieldCell) == 3 * sizeof(double) && sizeof(DensityNode) == 3 * sizeof(double),
		              "field records must be the reference's (hdr/field.h:14-38)");

		inline ipic_props toAbi(const UniverseProperties& p) {
			ipic_props abi{};
			for (int a = 0; a < 3; a++) {
				abi.size[a] = p.size[a];
				abi.cellWidth[a] = p.cellWidth[a];
				abi.origin[a] = p.origin[a];
				abi.objectCenter[a] = p.objectCenter[a];
				abi.externalMagneticField[a] = p.externalMagneticField[a];
			}
			abi.dt = p.dt;
			abi.speedOfLight = p.speedOfLight;
			abi.planetRadius = p.planetRadius;
			return abi;
		}

		// a Grid<T,3> of records of W doubles <-> one contiguous array, z fastest (the AllScale grid's storage need not be contiguous)
		template<typename GridT>
		std::vector<double> flatten(const GridT& grid, int W) {
			const auto n = grid.size();
			std::vector<double> out((std::size_t)(n.x * n.y * n.z) * W);
			double* o = out.data();
			for (std::int64_t x = 0; x < n.x; x++) for (std::int64_t y = 0; y < n.y; y++) for (std::int64_t z = 0; z < n.z; z++) {
				const double* rec = reinterpret_cast<const double*>(&grid[coordinate_type{ x, y, z }]);
				for (int k = 0; k < W; k++) *o++ = rec[k];
			}
			return out;
		}
		template<typename GridT>
		void unflatten(const std::vector<double>& in, int W, GridT& grid) {
			const auto n = grid.size();
			const double* i = in.data();
			for (std::int64_t x = 0; x < n.x; x++) for (std::int64_t y = 0; y < n.y; y++) for (std::int64_t z = 0; z < n.z; z++) {
				double* rec = reinterpret_cast<double*>(&grid[coordinate_type{ x, y, z }]);
				for (int k = 0; k < W; k++) rec[k] = *i++;
			}
		}

		template<typename Projector, typename Solver, typename Mover>
		std::uint32_t stepFlags() {
			static_assert(std::is_same<Mover, detail::default_particle_mover>::value, "the mover functor is a tag: only detail::default_particle_mover exists on the device");
			std::uint32_t flags = IPIC_MOVER_DIPOLE_ANALYTIC;
			if (std::is_same<Solver, detail::forward_field_solver>::value) flags |= IPIC_PROJECT_CURRENT | IPIC_UPDATE_BOUNDARIES | IPIC_SOLVER_FORWARD;
			if (std::is_same<Solver, detail::leapfrog_field_solver>::value) flags |= IPIC_PROJECT_CURRENT | IPIC_UPDATE_BOUNDARIES | IPIC_SOLVER_LEAPFROG;
			return flags;
		}
	}

	template<
		typename ParticleToFieldProjector 	= detail::default_particle_to_field_projector,
		typename FieldSolver 				= detail::default_field_solver,
		typename ParticleMover 				= detail::default_particle_mover
	>
	DurationMeasurement simulateSteps(std::uint64_t numSteps, Universe& universe) {
		using namespace b200;
		const std::uint32_t flags = stepFlags<ParticleToFieldProjector, FieldSolver, ParticleMover>();
		const ipic_props abi = toAbi(universe.properties);
		const auto size = universe.properties.size;

		int device = 0;
		if (const char* e = std::getenv("IPIC3D_DEVICE")) device = std::atoi(e);
		ipic_ctx* ctx = nullptr;
		check(ipic_create(&abi, device, &ctx), "ipic_create");

		// Grid<Cell,3> -> per-cell counts + one AoS array, cells in the z-fastest flatten, particles in cell order
		std::vector<std::uint64_t> counts;
		std::vector<Particle> aos;
		counts.reserve((std::size_t)(size.x * size.y * size.z));
		for (std::int64_t x = 0; x < size.x; x++) for (std::int64_t y = 0; y < size.y; y++) for (std::int64_t z = 0; z < size.z; z++) {
			const auto& ps = universe.cells[coordinate_type{ x, y, z }].particles;
			counts.push_back(ps.size());
			aos.insert(aos.end(), ps.begin(), ps.end());
		}
		check(ipic_upload_cells(ctx, counts.data(), reinterpret_cast<const double*>(aos.data())), "ipic_upload_cells");
		if (flags != IPIC_MOVER_DIPOLE_ANALYTIC) {
			const std::vector<double> f = flatten(universe.field, 9), b = flatten(universe.bcfield, 3), j = flatten(universe.currentDensity, 3);
			check(ipic_upload_fields(ctx, f.data(), b.data(), j.data()), "ipic_upload_fields");
		}

		double first = 0.0, rest = 0.0;
		check(ipic_step(ctx, numSteps, flags, &first, &rest), "ipic_step");           // the time loop, on the GPU

		std::uint64_t total = 0;
		check(ipic_total_particles(ctx, &total), "ipic_total_particles");
		aos.resize((std::size_t)total);
		check(ipic_download_cells(ctx, counts.data(), reinterpret_cast<double*>(aos.data()), total), "ipic_download_cells");
		std::size_t next = 0, cell = 0;
		for (std::int64_t x = 0; x < size.x; x++) for (std::int64_t y = 0; y < size.y; y++) for (std::int64_t z = 0; z < size.z; z++, cell++) {
			auto& ps = universe.cells[coordinate_type{ x, y, z }].particles;
			ps.assign(aos.begin() + next, aos.begin() + next + counts[cell]);
			next += counts[cell];
		}
		if (flags != IPIC_MOVER_DIPOLE_ANALYTIC) {
			std::vector<double> f(flatten(universe.field, 9).size()), b(flatten(universe.bcfield, 3).size()), j(flatten(universe.currentDensity, 3).size());
			check(ipic_download_fields(ctx, f.data(), b.data(), j.data()), "ipic_download_fields");
			unflatten(f, 9, universe.field); unflatten(b, 3, universe.bcfield); unflatten(j, 3, universe.currentDensity);
		}
		check(ipic_destroy(ctx), "ipic_destroy");
		return { first, rest };
	}

	template<
		typename ParticleToFieldProjector 	= detail::default_particle_to_field_projector,
		typename FieldSolver 				= detail::default_field_solver,
		typename ParticleMover 				= detail::default_particle_mover
	>
	void simulateStep(Universe& universe) {
		simulateSteps<ParticleToFieldProjector,FieldSolver,ParticleMover>(1, universe);
	}

} // end namespace ipic3d
