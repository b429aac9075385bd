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
// By default the call uploads the universe, steps, and downloads it again, so callers may read and modify
// `universe.cells[...]` / `universe.field[...]` between calls like the reference's tests do.  GPU contexts
// and the pinned staging buffer are kept between calls (releaseDevices() frees them).
//
// RESIDENT mode (opt-in: a `ResidentScope` object, or IPIC3D_RESIDENT=1): the universe stays in HBM between calls.
// The first simulateSteps call on a universe uploads it; later calls on the SAME universe object find it on the
// device and start stepping at once; nothing is downloaded until somebody asks for it with syncToHost(universe)
// (the scope's destructor does, so the host universe is current again when the scope ends).  A caller that changes
// the host universe inside such a scope says so with touch(universe).  This is the shape of the reference's own
// benchmark protocol (hdr/benchmark.h:38,43: simulateSteps(2) then simulateSteps(5) on the same universe) and of
// `ipic3d <deck>`, which only needs the per-cell counts afterwards (cellCounts(universe)).
// GPUs: IPIC3D_GPUS=<n> (default 1) splits the universe into n z-slabs on devices 0..n-1 of this process;
// IPIC3D_DEVICES=<ordinal,ordinal,...> names the slabs' devices explicitly.
#pragma once
#include <chrono>
#include <cstdio>
#include <fstream>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <iostream>
#include <thread>
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

		template<typename T, typename... Known> struct is_one_of : std::false_type {};
		template<typename T, typename K, typename... Rest> struct is_one_of<T, K, Rest...> : std::conditional<std::is_same<T, K>::value, std::true_type, is_one_of<T, Rest...>>::type {};

		template<typename Projector, typename Solver, typename Mover>
		constexpr std::uint32_t stepFlags() {
			// the policy parameters are TAGS for device kernels: a user-written functor cannot cross to the GPU, and silently
			// running the default in its place would be worse than not compiling
			static_assert(is_one_of<Projector, default_particle_to_field_projector, current_density_projector>::value,
			              "ParticleToFieldProjector must be one of the tags in ipic3d::detail (functors do not run on the device)");
			static_assert(is_one_of<Solver, default_field_solver, forward_field_solver, leapfrog_field_solver>::value,
			              "FieldSolver must be one of the tags in ipic3d::detail (functors do not run on the device)");
			static_assert(is_one_of<Mover, default_particle_mover, interpolating_particle_mover>::value,
			              "ParticleMover must be one of the tags in ipic3d::detail (functors do not run on the device)");
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

		// host-side packing / unpacking of the per-cell vectors is plain memory traffic: spread it over the cores
		template<typename Body>
		void parallelFor(std::size_t tasks, const Body& body) {
			unsigned workers = std::thread::hardware_concurrency();
			if (const char* e = std::getenv("NUM_WORKERS")) workers = (unsigned)std::atoi(e);     // the reference's knob
			workers = std::max(1u, std::min(workers, 64u));
			if (tasks < 256 || workers == 1) { for (std::size_t t = 0; t < tasks; ++t) body(t); return; }
			std::atomic<std::size_t> next(0);
			auto work = [&]() { for (;;) { const std::size_t first = next.fetch_add(64); if (first >= tasks) return; for (std::size_t t = first; t < std::min(tasks, first + 64); ++t) body(t); } };
			std::vector<std::thread> pool;
			for (unsigned w = 1; w < workers; ++w) pool.emplace_back(work);
			work();
			for (auto& t : pool) t.join();
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

		// Device contexts and the pinned staging buffer outlive a simulateSteps call: creating them costs far more
		// than a time step, and drivers call simulateSteps repeatedly on the same universe (benchmark.h warm-up +
		// timed run, the tests' simulateStep loops).  They are rebuilt when the universe's properties or the
		// device list change, and released by releaseDevices() (or at process exit by the driver).
		struct Session {
			ipic_props abi{};
			std::vector<int> devices;
			std::vector<ipic_ctx*> ctx;
			void* staging = nullptr;
			std::uint64_t stagingParticles = 0;
			// resident mode: which host universe the device state belongs to, whether the host copy is behind it, and what
			// the last steps changed besides the particles
			int residentDepth = 0;                   // nesting of ResidentScope objects
			const void* residentOf = nullptr;
			bool hostStale = false, fieldsStale = false, currentStale = false, fieldsOnDevice = false;

			bool residentMode() const {
				if (residentDepth > 0) return true;
				const char* e = std::getenv("IPIC3D_RESIDENT");
				return e && std::atoi(e) != 0;
			}
			void forget() { residentOf = nullptr; hostStale = fieldsStale = currentStale = fieldsOnDevice = false; }

			void release() {
				for (ipic_ctx* c : ctx) ipic_destroy(c);
				ctx.clear();
				if (staging) ipic_free_pinned(staging);
				staging = nullptr; stagingParticles = 0;
				forget();
			}
			void prepare(const ipic_props& props, const std::vector<int>& devs) {
				if (!ctx.empty() && devs == devices && std::memcmp(&props, &abi, sizeof abi) == 0) return;
				release();
				abi = props; devices = devs;
				ctx.assign(devs.size(), nullptr);
				check(ipic_group_create(&abi, devices.data(), (int)devices.size(), ctx.data()), "ipic_group_create");
			}
			void reserveStaging(std::uint64_t particles) {
				if (particles <= stagingParticles) return;
				if (staging) ipic_free_pinned(staging);
				staging = nullptr;
				stagingParticles = particles + particles / 8;
				check(ipic_alloc_pinned((std::size_t)stagingParticles * sizeof(Particle), &staging), "ipic_alloc_pinned");
			}
			static Session& instance() { static Session* s = new Session(); return *s; }   // never destroyed: CUDA may be gone at exit
		};
	}

	// ---- conserved-quantity log and density snapshots (reference hdr/simulator.h:64-115, ENABLE_DEBUG_OUTPUT) --------

	template<typename StreamObject>
	void writeOutputHeader(StreamObject& streamObject) {
		streamObject << "Cycle \t Total Moment \t E energy \t B energy \t Total KE \n";
	}

	namespace detail {
		// is the cycle one that writeOutputData logs (simulator.h:74) / dumps the density of (simulator.h:93)?
		inline bool fieldOutputDue(const UniverseProperties& p, std::uint64_t cycle, std::uint64_t numSteps) {
			return p.FieldOutputCycle > 0 && ((cycle % (std::uint64_t)p.FieldOutputCycle == 0) || (cycle + 1 == numSteps));
		}
		inline bool particleOutputDue(const UniverseProperties& p, std::uint64_t cycle) {
			return p.ParticleOutputCycle > 0 && ((cycle == 0) || ((cycle + 1) % (std::uint64_t)p.ParticleOutputCycle == 0));
		}
		template<typename StreamObject>
		void writeConservedRow(std::uint64_t cycle, const double q[4], StreamObject& streamObject) {
			streamObject << cycle << "\t" << q[0] << "\t" << q[1] << "\t" << q[2] << "\t" << q[3] << "\n";
		}
		// "t,x,y,z,density" snapshot: number of particles per cell, cells in x-major order (simulator.h:93-113)
		template<typename CountOf>
		void writeDensitySnapshot(const UniverseProperties& p, std::uint64_t cycle, const char* fileName, const CountOf& countOf) {
			const int t = (int)((cycle + 1) / (std::uint64_t)p.ParticleOutputCycle);
			std::cout << cycle << ' ' << t << '\n';
			char fileNamet[96];
			std::snprintf(fileNamet, sizeof fileNamet, "%s%06d", fileName, t);
			std::ofstream out(fileNamet);
			out << "t,x,y,z,density\n";
			for (std::int64_t x = 0; x < p.size.x; x++) for (std::int64_t y = 0; y < p.size.y; y++) for (std::int64_t z = 0; z < p.size.z; z++)
				out << t << "," << x * p.cellWidth.x << "," << y * p.cellWidth.y << "," << z * p.cellWidth.z << "," << countOf(x, y, z) << "\n";
		}
		// the reference compiles the log in with -DENABLE_DEBUG_OUTPUT; here it can also be switched on at run time
		inline bool debugOutputEnabled() {
#ifdef ENABLE_DEBUG_OUTPUT
			return true;
#else
			const char* e = std::getenv("IPIC3D_DEBUG_OUTPUT");
			return e && std::atoi(e) != 0;
#endif
		}
	}

	// write output depending on the set frequency, from a universe that lives on the host (reference simulator.h:71-115)
	template<typename StreamObject>
	void writeOutputData(const int cycle, const int numSteps, Universe& universe, StreamObject& streamObject, char* fileName) {
		const UniverseProperties& p = universe.properties;
		if (detail::fieldOutputDue(p, (std::uint64_t)cycle, (std::uint64_t)numSteps)) {
			auto getE = [](const Field& field, const coordinate_type& index) { return field[index].E; };
			auto getB = [](const Field& field, const coordinate_type& index) { return field[index].B + field[index].Bext; };
			const double q[4] = { getTotalParticlesEnergy(universe.cells, getParticlesMomentum), getFieldEnergy(universe.field, p, getE),
			                      getFieldEnergy(universe.field, p, getB), getTotalParticlesEnergy(universe.cells, getParticlesKineticEnergy) };
			detail::writeConservedRow((std::uint64_t)cycle, q, streamObject);
		}
		if (detail::particleOutputDue(p, (std::uint64_t)cycle))
			detail::writeDensitySnapshot(p, (std::uint64_t)cycle, fileName, [&](std::int64_t x, std::int64_t y, std::int64_t z) { return universe.cells[{ x, y, z }].particles.size(); });
	}

	namespace detail { inline void downloadUniverse(Session& session, Universe& universe, bool fields, bool current); }

	// frees the GPU contexts and pinned memory kept between simulateSteps calls
	inline void releaseDevices() { detail::Session::instance().release(); }

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
		// (the reference truncates to milliseconds, hdr/simulator.h:122; whole runs take a few of them here, so microseconds)
		const auto seconds = [](clock::duration d) { return std::chrono::duration_cast<std::chrono::microseconds>(d).count() / 1e6; };
		const std::uint32_t flags = detail::stepFlags<ParticleToFieldProjector, FieldSolver, ParticleMover>();
		const bool usesFields = flags & (IPIC_MOVER_INTERPOLATED | IPIC_SOLVER_FORWARD | IPIC_SOLVER_LEAPFROG | IPIC_UPDATE_BOUNDARIES);
		const bool writesFields = flags & (IPIC_SOLVER_FORWARD | IPIC_SOLVER_LEAPFROG | IPIC_UPDATE_BOUNDARIES);
		const bool writesCurrent = flags & IPIC_PROJECT_CURRENT;

		const auto start = clock::now();
		const UniverseProperties& props = universe.properties;
		const bool debugOutput = detail::debugOutputEnabled();
		const ipic_props abi = props.toAbi();
		const std::int64_t nx = props.size.x, ny = props.size.y, nz = props.size.z;
		const std::vector<int> devices = detail::slabDevices(nz);
		const int gpus = (int)devices.size();
		const std::int64_t lz = nz / gpus;
		detail::Session& session = detail::Session::instance();
		session.prepare(abi, devices);
		std::vector<ipic_ctx*>& ctx = session.ctx;

		// resident mode: the device already holds this universe (left there by an earlier call) -> no upload
		const bool resident = session.residentMode();
		const bool onDevice = resident && session.residentOf == static_cast<const void*>(&universe);
		if (!onDevice) session.forget();

		// ---- upload: every slab's cells in slab order (x*ny + y)*lz + zl, particles cell-major as AoS records
		std::vector<std::vector<std::uint64_t>> counts((std::size_t)gpus);
		std::uint64_t largestSlab = 0;
		if (!onDevice) for (int r = 0; r < gpus; ++r) {
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
		if (!onDevice) session.reserveStaging(largestSlab + largestSlab / 4 + 65536);
		struct { void* p; } staging{ session.staging };
		// slab cell c = (x*ny + y)*lz + zl  ->  the universe's cell
		const auto cellOf = [&](int r, std::size_t c) -> Cell& {
			const std::int64_t xy = (std::int64_t)c / lz, zl = (std::int64_t)c % lz;
			return universe.cells[{ xy / ny, xy % ny, r * lz + zl }];
		};
		std::vector<std::uint64_t> offset;
		const bool needFields = usesFields || writesCurrent || debugOutput;
		if (onDevice && needFields && !session.fieldsOnDevice)
			for (int r = 0; r < gpus; ++r)
				detail::check(ipic_upload_fields(ctx[(std::size_t)r], reinterpret_cast<const double*>(universe.field.data()),
				                                 reinterpret_cast<const double*>(universe.bcfield.data()),
				                                 reinterpret_cast<const double*>(universe.currentDensity.data())), "ipic_upload_fields");
		if (needFields) session.fieldsOnDevice = true;
		if (!onDevice) for (int r = 0; r < gpus; ++r) {
			const auto& cnt = counts[(std::size_t)r];
			offset.assign(cnt.size() + 1, 0);
			for (std::size_t c = 0; c < cnt.size(); ++c) offset[c + 1] = offset[c] + cnt[c];
			detail::parallelFor(cnt.size(), [&](std::size_t c) {
				const auto& ps = cellOf(r, c).particles;
				if (!ps.empty()) std::memcpy(static_cast<char*>(staging.p) + offset[c] * sizeof(Particle), ps.data(), ps.size() * sizeof(Particle));
			});
			detail::check(ipic_upload_cells(ctx[(std::size_t)r], counts[(std::size_t)r].data(), static_cast<const double*>(staging.p)), "ipic_upload_cells");
			if (usesFields || writesCurrent || debugOutput)
				detail::check(ipic_upload_fields(ctx[(std::size_t)r], reinterpret_cast<const double*>(universe.field.data()),
				                                 reinterpret_cast<const double*>(universe.bcfield.data()),
				                                 reinterpret_cast<const double*>(universe.currentDensity.data())), "ipic_upload_fields");
		}

		// ---- the time loop: first step and the rest are timed apart, as the reference does.  With the debug output on
		// (reference: -DENABLE_DEBUG_OUTPUT, simulator.h:160-187) the loop stops before every cycle that is logged; the
		// four conserved quantities are reduced on the device (ipic_group_diagnostics) and the density snapshot needs the
		// per-cell counts only, so the particles stay in HBM.
		std::ofstream logFile;
		char fileName[64] = "";
		if (debugOutput) {
			logFile.open(props.outputFileBaseName + "ConservedQuantities.out");
			writeOutputHeader(logFile);
			const auto timeStamp = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
			std::snprintf(fileName, sizeof fileName, "result_%ld.csv.", (long)timeStamp);
		}
		const auto emitOutput = [&](std::uint64_t cycle) {
			if (detail::fieldOutputDue(props, cycle, numSteps)) {
				double q[4];
				detail::check(ipic_group_diagnostics(ctx.data(), gpus, q), "ipic_group_diagnostics");
				detail::writeConservedRow(cycle, q, logFile);
			}
			if (detail::particleOutputDue(props, cycle)) {
				std::vector<std::vector<std::uint64_t>> now((std::size_t)gpus);
				for (int r = 0; r < gpus; ++r) {
					now[(std::size_t)r].resize((std::size_t)(nx * ny * lz));
					detail::check(ipic_cell_counts(ctx[(std::size_t)r], now[(std::size_t)r].data()), "ipic_cell_counts");
				}
				detail::writeDensitySnapshot(props, cycle, fileName, [&](std::int64_t x, std::int64_t y, std::int64_t z) {
					return now[(std::size_t)(z / lz)][(std::size_t)((x * ny + y) * lz + z % lz)];
				});
			}
		};
		const auto outputDue = [&](std::uint64_t cycle) { return debugOutput && (detail::fieldOutputDue(props, cycle, numSteps) || detail::particleOutputDue(props, cycle)); };
		auto endFirst = clock::now();
		for (std::uint64_t done = 0; done < numSteps;) {
			if (outputDue(done)) emitOutput(done);
			std::uint64_t next = done + 1;                       // the first step is timed on its own
			while (done > 0 && next < numSteps && !outputDue(next)) ++next;
			detail::check(ipic_group_step(ctx.data(), gpus, next - done, flags, nullptr, nullptr), "ipic_group_step");
			if (done == 0) endFirst = clock::now();
			done = next;
		}

		if (resident) {
			// the universe stays in HBM; the host copy is brought up to date by syncToHost (ResidentScope's destructor)
			session.residentOf = &universe;
			session.hostStale = true;
			session.fieldsStale = session.fieldsStale || writesFields;
			session.currentStale = session.currentStale || writesCurrent;
			const auto endAll = clock::now();
			return { seconds(endFirst - start), seconds(endAll - endFirst) };
		}
		detail::downloadUniverse(session, universe, writesFields, writesCurrent);
		const auto endAll = clock::now();
		return { seconds(endFirst - start), seconds(endAll - endFirst) };
	}

	namespace detail {
		// device -> host: every slab's particles back into the per-cell vectors, and the grids the steps changed
		inline void downloadUniverse(Session& session, Universe& universe, bool fields, bool current) {
			const UniverseProperties& props = universe.properties;
			const std::int64_t nx = props.size.x, ny = props.size.y, nz = props.size.z;
			(void)nx;
			std::vector<ipic_ctx*>& ctx = session.ctx;
			const int gpus = (int)ctx.size();
			const std::int64_t lz = nz / gpus;
			const auto cellOf = [&](int r, std::size_t c) -> Cell& {
				const std::int64_t xy = (std::int64_t)c / lz, zl = (std::int64_t)c % lz;
				return universe.cells[{ xy / ny, xy % ny, r * lz + zl }];
			};
			std::vector<std::uint64_t> cnt, offset;
			struct { void* p; } staging{ session.staging };
			for (int r = 0; r < gpus; ++r) {
				cnt.assign((std::size_t)(props.size.x * ny * lz), 0);
				std::uint64_t total = 0;
				check(ipic_total_particles(ctx[(std::size_t)r], &total), "ipic_total_particles");
				session.reserveStaging(total);
				staging.p = session.staging;
				check(ipic_download_cells(ctx[(std::size_t)r], cnt.data(), static_cast<double*>(staging.p), session.stagingParticles), "ipic_download_cells");
				offset.assign(cnt.size() + 1, 0);
				for (std::size_t c = 0; c < cnt.size(); ++c) offset[c + 1] = offset[c] + cnt[c];
				parallelFor(cnt.size(), [&](std::size_t c) {
					auto& ps = cellOf(r, c).particles;
					ps.resize((std::size_t)cnt[c]);
					if (cnt[c]) std::memcpy(ps.data(), static_cast<const char*>(staging.p) + offset[c] * sizeof(Particle), (std::size_t)cnt[c] * sizeof(Particle));
				});
			}
			if (fields || current)     // E, B, Bc are replicated on every slab: rank 0 has them all (and all of J when a solver ran)
				check(ipic_download_fields(ctx[0], fields ? reinterpret_cast<double*>(universe.field.data()) : nullptr,
				                           fields ? reinterpret_cast<double*>(universe.bcfield.data()) : nullptr,
				                           current ? reinterpret_cast<double*>(universe.currentDensity.data()) : nullptr), "ipic_download_fields");
			if (current && !fields && gpus > 1) {
				// projection without a solver: every slab holds J on its own node planes z0 .. z0 + lz only
				std::vector<DensityNode> part(universe.currentDensity.count());
				const std::int64_t jy = ny + 1, jz = nz + 1;
				for (int r = 1; r < gpus; ++r) {
					check(ipic_download_fields(ctx[(std::size_t)r], nullptr, nullptr, reinterpret_cast<double*>(part.data())), "ipic_download_fields");
					for (std::int64_t x = 0; x <= props.size.x; ++x) for (std::int64_t y = 0; y <= ny; ++y) for (std::int64_t z = r * lz + 1; z <= (r + 1) * lz; ++z)
						universe.currentDensity.data()[(x * jy + y) * jz + z] = part[(std::size_t)((x * jy + y) * jz + z)];
				}
			}
		}
	}

	// ---- resident mode (see the header comment) -----------------------------------------------------------------

	// brings the host universe up to date with the device copy left by simulateSteps in resident mode
	inline void syncToHost(Universe& universe) {
		detail::Session& session = detail::Session::instance();
		if (session.residentOf != static_cast<const void*>(&universe) || !session.hostStale) return;
		detail::downloadUniverse(session, universe, session.fieldsStale, session.currentStale);
		session.hostStale = session.fieldsStale = session.currentStale = false;
	}

	// the caller changed the host universe: the device copy (if any) is dropped, the next call uploads again
	inline void touch(Universe& universe) {
		detail::Session& session = detail::Session::instance();
		if (session.residentOf != static_cast<const void*>(&universe)) return;
		if (session.hostStale) syncToHost(universe);      // nothing the device computed is lost
		session.forget();
	}

	// Per-cell particle counts of a universe wherever it currently lives (x-major, the order of the `.out` file): from the
	// device when the host copy is stale -- 8 bytes per cell instead of 64 bytes per particle across PCIe.
	inline std::vector<std::uint64_t> cellCounts(const Universe& universe) {
		const UniverseProperties& p = universe.properties;
		std::vector<std::uint64_t> out((std::size_t)(p.size.x * p.size.y * p.size.z));
		detail::Session& session = detail::Session::instance();
		if (session.residentOf == static_cast<const void*>(&universe) && session.hostStale) {
			const int gpus = (int)session.ctx.size();
			const std::int64_t lz = p.size.z / gpus;
			std::vector<std::uint64_t> slab((std::size_t)(p.size.x * p.size.y * lz));
			for (int r = 0; r < gpus; ++r) {
				detail::check(ipic_cell_counts(session.ctx[(std::size_t)r], slab.data()), "ipic_cell_counts");
				for (std::size_t c = 0; c < slab.size(); ++c) out[(c / (std::size_t)lz) * (std::size_t)p.size.z + (std::size_t)(r * lz) + c % (std::size_t)lz] = slab[c];
			}
		} else {
			for (std::size_t i = 0; i < out.size(); ++i) out[i] = universe.cells.data()[i].particles.size();
		}
		return out;
	}

	// "<base>.out" of a universe wherever it lives (reference writer: hdr/cell.h:1360-1388, same text)
	inline void outputNumberOfParticlesPerCell(const Universe& universe, const std::string& outputFilename) {
		const UniverseProperties& p = universe.properties;
		const std::vector<std::uint64_t> counts = cellCounts(universe);
		std::ofstream out(outputFilename);
		out << p.size << "\n";
		std::uint64_t total = 0;
		std::size_t i = 0;
		for (std::int64_t x = 0; x < p.size.x; ++x) for (std::int64_t y = 0; y < p.size.y; ++y) for (std::int64_t z = 0; z < p.size.z; ++z, ++i) {
			out << x << "," << y << "," << z << ":" << counts[i] << "\n";
			total += counts[i];
		}
		out << "\nTotal: " << total << "\n";
	}

	// Creates the particles of `universe` directly in HBM (resident mode only: inside a ResidentScope): phase 2 of
	// initCells<uniform_pos_normal_speed> (hdr/cell.h:596-632) on the device(s), from the per-cell populations of phase 1
	// (z-fastest flatten, e.g. EarthDeck::cellPopulations) and the distribution's velocity parameters.  The host cells stay
	// empty until syncToHost(universe); simulateSteps finds the universe on the device.  Bit-identical to initCells.
	inline void populateOnDevice(Universe& universe, const std::vector<std::uint64_t>& population, const Vector3<double>& mean,
	                             const Vector3<double>& stddev, double q, double qom) {
		detail::Session& session = detail::Session::instance();
		if (!session.residentMode()) { std::cerr << "ipic3d_b200: populateOnDevice needs resident mode (a ResidentScope)" << std::endl; std::exit(EXIT_FAILURE); }
		const UniverseProperties& p = universe.properties;
		const ipic_props abi = p.toAbi();
		session.prepare(abi, detail::slabDevices(p.size.z));
		session.forget();
		const int gpus = (int)session.ctx.size();
		const std::int64_t lz = p.size.z / gpus;
		const double m[3] = { mean.x, mean.y, mean.z }, sd[3] = { stddev.x, stddev.y, stddev.z };
		std::vector<std::uint64_t> slab((std::size_t)(p.size.x * p.size.y * lz));
		for (int r = 0; r < gpus; ++r) {
			for (std::size_t c = 0; c < slab.size(); ++c) slab[c] = population[(c / (std::size_t)lz) * (std::size_t)p.size.z + (std::size_t)(r * lz) + c % (std::size_t)lz];
			detail::check(ipic_generate_normal_speed(session.ctx[(std::size_t)r], slab.data(), m, sd, q, qom), "ipic_generate_normal_speed");
		}
		session.residentOf = &universe;
		session.hostStale = true;
	}

	// While an object of this type lives, simulateSteps keeps universes in HBM between calls; at the end of the scope the
	// host copy of `universe` is brought up to date (unless dismissed: the caller only wanted counts / diagnostics).
	class ResidentScope {
		Universe& universe;
		bool sync = true;
	public:
		explicit ResidentScope(Universe& u) : universe(u) { detail::Session::instance().residentDepth++; }
		ResidentScope(const ResidentScope&) = delete;
		ResidentScope& operator=(const ResidentScope&) = delete;
		void dismiss() { sync = false; }
		~ResidentScope() {
			detail::Session& session = detail::Session::instance();
			if (sync) syncToHost(universe);
			if (--session.residentDepth == 0 && session.residentOf == static_cast<const void*>(&universe)) session.forget();
		}
	};

	template<
		typename ParticleToFieldProjector = detail::default_particle_to_field_projector,
		typename FieldSolver = detail::default_field_solver,
		typename ParticleMover = detail::default_particle_mover
	>
	void simulateStep(Universe& universe) {
		simulateSteps<ParticleToFieldProjector, FieldSolver, ParticleMover>(1, universe);
	}

} // namespace ipic3d
