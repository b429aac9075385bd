// Host C++ mirror: the time loop entry points (reference hdr/simulator.h:42-57, :127-242, :246-275).
//
//     simulateSteps<Projector, FieldSolver, Mover>(numSteps, universe)      simulateStep<...>(universe)
//
// keep their names, template parameters and semantics; the loop body runs on the GPU(s) behind the C ABI
// (include/ipic3d_b200.h).  Per-cell functor calls cannot cross to the device, so the policy types are
// TAGS that select kernels:
//
//   all three defaults                         the loop exactly as the reference ships it: STEP 3 mover
//                                              (analytic dipole) + export, STEP 4 import; STEP 1/2 are
//                                              commented out there (simulator.h:189-209)
//   FieldSolver = forward_field_solver /       the full step the commented code describes: projection,
//                 leapfrog_field_solver        ghost update, that solver, mover, import
//   Projector = current_density_projector      (extension) STEP 1 alone on top of the shipped loop
//   Mover = interpolating_particle_mover       (extension) the commented corner-gather mover, cell.h:905-937
//
// The call uploads the universe, steps, and downloads it again, so callers may read and modify
// `universe.cells[...]` / `universe.field[...]` between calls like the reference's tests do.
// GPUs: IPIC3D_GPUS=<n> (default 1) splits the universe into n z-slabs on devices 0..n-1 of this process;
// IPIC3D_DEVICES=<ordinal,ordinal,...> names the slabs' devices explicitly.
#pragma once
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <type_traits>
#include <vector>

#include "ipic3d/app/universe.h"
#include "ipic3d_b200.h"

namespace ipic3d {

	namespace detail {
		struct default_particle_to_field_projector {};
		struct current_density_projector {};
		struct default_field_solver {};
		struct forward_field_solver {};
		struct leapfrog_field_solver {};
		struct default_particle_mover {};
		struct interpolating_particle_mover {};

		template<typename Projector, typename Solver, typename Mover>
		constexpr std::uint32_t stepFlags() {
			std::uint32_t f = 0;
			if (std::is_same<Solver, forward_field_solver>::value) f |= IPIC_PROJECT_CURRENT | IPIC_UPDATE_BOUNDARIES | IPIC_SOLVER_FORWARD;
			if (std::is_same<Solver, leapfrog_field_solver>::value) f |= IPIC_PROJECT_CURRENT | IPIC_UPDATE_BOUNDARIES | IPIC_SOLVER_LEAPFROG;
			if (std::is_same<Projector, current_density_projector>::value) f |= IPIC_PROJECT_CURRENT;
			if (std::is_same<Mover, interpolating_particle_mover>::value) f |= IPIC_MOVER_INTERPOLATED;
			return f;
		}

		// failures of the device layer end the run with a message, like a failed assertion in the reference
		inline void check(int rc, const char* what) {
			if (rc == IPIC_OK) return;
			std::cerr << "ipic3d_b200: " << what << " failed (" << rc << "): " << ipic_last_error() << std::endl;
			std::exit(EXIT_FAILURE);
		}

		// CUDA ordinals of the slabs: IPIC3D_DEVICES="0,1,2,3" (ordinals may repeat: several slabs on one GPU),
		// else the first IPIC3D_GPUS devices, else device 0.  Slabs hold whole z-planes.
		inline std::vector<int> slabDevices(std::int64_t nz) {
			std::vector<int> devices;
			int have = 0;
			check(ipic_device_count(&have), "ipic_device_count");
			if (const char* e = std::getenv("IPIC3D_DEVICES")) {
				for (const char* p = e; *p;) {
					char* end = nullptr;
					const long d = std::strtol(p, &end, 10);
					if (end == p) break;
					devices.push_back((int)d);
					p = *end ? end + 1 : end;
				}
			} else {
				int n = 1;
				if (const char* g = std::getenv("IPIC3D_GPUS")) n = std::atoi(g);
				if (n > have) n = have;
				for (int d = 0; d < n; ++d) devices.push_back(d);
			}
			if (devices.empty()) devices.push_back(0);
			while (devices.size() > 1 && nz % (std::int64_t)devices.size() != 0) devices.pop_back();
			return devices;
		}

		// pinned staging memory for the AoS transfers
		struct PinnedBuffer {
			void* p = nullptr; std::size_t bytes = 0;
			explicit PinnedBuffer(std::size_t n) : bytes(n) { if (n) check(ipic_alloc_pinned(n, &p), "ipic_alloc_pinned"); }
			~PinnedBuffer() { if (p) ipic_free_pinned(p); }
			PinnedBuffer(const PinnedBuffer&) = delete;
			PinnedBuffer& operator=(const PinnedBuffer&) = delete;
		};
	}

	struct DurationMeasurement {
		double firstStep;
		double remainingSteps;
	};

	template<
		typename ParticleToFieldProjector = detail::default_particle_to_field_projector,
		typename FieldSolver = detail::default_field_solver,
		typename ParticleMover = detail::default_particle_mover
	>
	DurationMeasurement simulateSteps(std::uint64_t numSteps, Universe& universe) {
		using clock = std::chrono::high_resolution_clock;
		const auto seconds = [](clock::duration d) { return std::chrono::duration_cast<std::chrono::microseconds>(d).count() / 1e6; };
		const std::uint32_t flags = detail::stepFlags<ParticleToFieldProjector, FieldSolver, ParticleMover>();
		const bool usesFields = flags & (IPIC_MOVER_INTERPOLATED | IPIC_SOLVER_FORWARD | IPIC_SOLVER_LEAPFROG | IPIC_UPDATE_BOUNDARIES);
		const bool writesFields = flags & (IPIC_SOLVER_FORWARD | IPIC_SOLVER_LEAPFROG | IPIC_UPDATE_BOUNDARIES);
		const bool writesCurrent = flags & IPIC_PROJECT_CURRENT;

		const auto start = clock::now();
		const UniverseProperties& props = universe.properties;
		const ipic_props abi = props.toAbi();
		const std::int64_t nx = props.size.x, ny = props.size.y, nz = props.size.z;
		const std::vector<int> devices = detail::slabDevices(nz);
		const int gpus = (int)devices.size();
		const std::int64_t lz = nz / gpus;
		std::vector<ipic_ctx*> ctx((std::size_t)gpus, nullptr);
		detail::check(ipic_group_create(&abi, devices.data(), gpus, ctx.data()), "ipic_group_create");

		// ---- upload: every slab's cells in slab order (x*ny + y)*lz + zl, particles cell-major as AoS records
		std::vector<std::vector<std::uint64_t>> counts((std::size_t)gpus);
		std::uint64_t largestSlab = 0;
		for (int r = 0; r < gpus; ++r) {
			auto& cnt = counts[(std::size_t)r];
			cnt.reserve((std::size_t)(nx * ny * lz));
			std::uint64_t total = 0;
			for (std::int64_t x = 0; x < nx; ++x) for (std::int64_t y = 0; y < ny; ++y) for (std::int64_t z = r * lz; z < (r + 1) * lz; ++z) {
				cnt.push_back(universe.cells[{ x, y, z }].particles.size());
				total += cnt.back();
			}
			largestSlab = std::max(largestSlab, total);
		}
		// room for migration into a slab while it is on the device
		const std::uint64_t stagingParticles = largestSlab + largestSlab / 4 + 65536;
		detail::PinnedBuffer staging((std::size_t)stagingParticles * sizeof(Particle));
		for (int r = 0; r < gpus; ++r) {
			char* out = static_cast<char*>(staging.p);
			for (std::int64_t x = 0; x < nx; ++x) for (std::int64_t y = 0; y < ny; ++y) for (std::int64_t z = r * lz; z < (r + 1) * lz; ++z) {
				const auto& ps = universe.cells[{ x, y, z }].particles;
				if (!ps.empty()) std::memcpy(out, ps.data(), ps.size() * sizeof(Particle));
				out += ps.size() * sizeof(Particle);
			}
			detail::check(ipic_upload_cells(ctx[(std::size_t)r], counts[(std::size_t)r].data(), static_cast<const double*>(staging.p)), "ipic_upload_cells");
			if (usesFields || writesCurrent)
				detail::check(ipic_upload_fields(ctx[(std::size_t)r], reinterpret_cast<const double*>(universe.field.data()),
				                                 reinterpret_cast<const double*>(universe.bcfield.data()),
				                                 reinterpret_cast<const double*>(universe.currentDensity.data())), "ipic_upload_fields");
		}

		// ---- the time loop: first step and the rest are timed apart, as the reference does
		double tFirst = 0, tRest = 0;
		if (numSteps > 0) detail::check(ipic_group_step(ctx.data(), gpus, 1, flags, &tFirst, nullptr), "ipic_group_step");
		const auto endFirst = clock::now();
		if (numSteps > 1) detail::check(ipic_group_step(ctx.data(), gpus, numSteps - 1, flags, &tFirst, &tRest), "ipic_group_step");

		// ---- download
		for (int r = 0; r < gpus; ++r) {
			auto& cnt = counts[(std::size_t)r];
			std::uint64_t total = 0;
			detail::check(ipic_total_particles(ctx[(std::size_t)r], &total), "ipic_total_particles");
			if (total > stagingParticles) { std::cerr << "ipic3d_b200: slab grew beyond the staging buffer" << std::endl; std::exit(EXIT_FAILURE); }
			detail::check(ipic_download_cells(ctx[(std::size_t)r], cnt.data(), static_cast<double*>(staging.p), stagingParticles), "ipic_download_cells");
			const char* in = static_cast<const char*>(staging.p);
			std::size_t c = 0;
			for (std::int64_t x = 0; x < nx; ++x) for (std::int64_t y = 0; y < ny; ++y) for (std::int64_t z = r * lz; z < (r + 1) * lz; ++z, ++c) {
				auto& ps = universe.cells[{ x, y, z }].particles;
				ps.resize((std::size_t)cnt[c]);
				if (cnt[c]) std::memcpy(ps.data(), in, (std::size_t)cnt[c] * sizeof(Particle));
				in += cnt[c] * sizeof(Particle);
			}
		}
		if (writesFields || writesCurrent)     // the grids are replicated on every slab: rank 0 has them all
			detail::check(ipic_download_fields(ctx[0], writesFields ? reinterpret_cast<double*>(universe.field.data()) : nullptr,
			                                   writesFields ? reinterpret_cast<double*>(universe.bcfield.data()) : nullptr,
			                                   writesCurrent ? reinterpret_cast<double*>(universe.currentDensity.data()) : nullptr), "ipic_download_fields");
		for (ipic_ctx* c : ctx) ipic_destroy(c);
		const auto endAll = clock::now();
		return { seconds(endFirst - start), seconds(endAll - endFirst) };
	}

	template<
		typename ParticleToFieldProjector = detail::default_particle_to_field_projector,
		typename FieldSolver = detail::default_field_solver,
		typename ParticleMover = detail::default_particle_mover
	>
	void simulateStep(Universe& universe) {
		simulateSteps<ParticleToFieldProjector, FieldSolver, ParticleMover>(1, universe);
	}

} // namespace ipic3d
