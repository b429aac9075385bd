// Host C++ mirror: `Cell`, `Cells`, the particle generators and `initCells` (reference hdr/cell.h:31-100,
// :102-776, :1360-1388).  These run once on the host and define the INPUT of every run, so they reproduce
// the reference's random streams exactly (same libstdc++ engines, same draw order; SURVEY.md Appendix
// A.3/A.4).  The per-timestep functions of the reference's cell.h (moveParticles, exportParticles,
// importParticles, projectToDensityField) are the CUDA kernels behind `simulateSteps` (simulator.h).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <fstream>
#include <iostream>
#include <limits>
#include <random>
#include <string>
#include <vector>

#include "ipic3d/app/field.h"
#include "ipic3d/app/init_properties.h"
#include "ipic3d/app/parameters.h"
#include "ipic3d/app/particle.h"
#include "ipic3d/app/universe_properties.h"
#include "ipic3d/app/vector.h"
#include "ipic3d/app/ziggurat_normal_distribution.h"

namespace ipic3d {

	struct Cell {
		std::vector<Particle> particles;
	};

	using Cells = Grid3<Cell>;

	// ---- geometry predicates (reference hdr/cell.h:48-100) ---------------------------------------------------

	// closed on both sides: a particle exactly on a face belongs to both cells
	inline bool isInside(const UniverseProperties& properties, const coordinate_type& pos, const Particle& p) {
		const Vector3<double> rel = p.position - getCenterOfCell(pos, properties);
		const Vector3<double> half = properties.cellWidth / 2.0;
		return std::fabs(rel.x) <= half.x && std::fabs(rel.y) <= half.y && std::fabs(rel.z) <= half.z;
	}

	inline coordinate_type getCellCoordinates(const UniverseProperties& properties, const Particle& p) {
		const auto f = elementwiseDivision(p.position - properties.origin, properties.cellWidth);
		return { (std::int64_t)f.x, (std::int64_t)f.y, (std::int64_t)f.z };
	}

	inline bool isInsideUniverse(const UniverseProperties& properties, const Particle& p) {
		const Vector3<double> upper{ properties.origin.x + properties.size.x * properties.cellWidth.x,
		                             properties.origin.y + properties.size.y * properties.cellWidth.y,
		                             properties.origin.z + properties.size.z * properties.cellWidth.z };
		return properties.origin.dominatedBy(p.position) && p.position.strictlyDominatedBy(upper);
	}

	inline bool verifyCorrectParticlesPositionInCell(const UniverseProperties& properties, const Cell& cell, const coordinate_type& pos) {
		for (const Particle& p : cell.particles) if (!isInside(properties, pos, p)) return false;
		return true;
	}

	inline std::uint64_t countParticlesInDomain(const Cells& cells) {
		std::uint64_t n = 0;
		for (std::size_t i = 0; i < cells.count(); ++i) n += cells.data()[i].particles.size();
		return n;
	}

	// ---- particle distributions (reference hdr/cell.h:102-485) ------------------------------------------------

	namespace distribution {

		namespace species {
			struct electron {
				Particle operator()() const { Particle p; p.q = -1.0; p.qom = -25.0; return p; }
				void seed(std::uint32_t) {}
			};
			struct proton {
				Particle operator()() const { Particle p; p.q = 1.0; p.qom = 1.0; return p; }
				void seed(std::uint32_t) {}
			};
		}

		namespace vector {

			// one minstd engine feeding three uniform distributions, drawn x, y, z
			class uniform {
				std::uniform_real_distribution<> dx, dy, dz;
				std::minstd_rand engine;
			public:
				uniform(const Vector3<double>& lo, const Vector3<double>& hi, std::uint32_t seed = 0)
					: dx(lo.x, hi.x), dy(lo.y, hi.y), dz(lo.z, hi.z), engine(seed) {}
				Vector3<double> operator()() { const double a = dx(engine); const double b = dy(engine); const double c = dz(engine); return { a, b, c }; }
				void seed(std::uint32_t s) { engine.seed(s); }
			};

			// uniform in a spherical shell between radii R1 and R2 (per component), reference hdr/cell.h:176-220
			class uniform_r {
				Vector3<double> R1, R2;
				std::uniform_real_distribution<> u1{ 0.0, 1.0 }, u2{ 0.0, 1.0 }, u3{ 0.0, 1.0 };
				std::minstd_rand engine;
			public:
				uniform_r(const Vector3<double>& lo, const Vector3<double>& hi, std::uint32_t seed = 0) : R1(lo), R2(hi), engine(seed) {}
				Vector3<double> operator()() {
					const double a = u1(engine); const double b = u2(engine); const double c = u3(engine);
					const double nu = (1.0 - 2.0 * b);
					auto radius = [a](double r1, double r2) { return std::pow(std::pow(r1, 3) + (std::pow(r2, 3) - std::pow(r1, 3)) * a, 1.0 / 3.0); };
					return { radius(R1.x, R2.x) * std::pow(1 - nu * nu, 1.0 / 2.0) * std::cos(2.0 * M_PI * c),
					         radius(R1.y, R2.y) * std::pow(1 - nu * nu, 1.0 / 2.0) * std::sin(2.0 * M_PI * c),
					         radius(R1.z, R2.z) * nu };
				}
				void seed(std::uint32_t s) { engine.seed(s); }
			};

			// mean + stddev * ziggurat sample (float), drawn x, y, z from one generator
			class normal {
				Vector3<double> mean, stddev;
				ziggurat_normal_distribution zig;
			public:
				normal(const Vector3<double>& mean, const Vector3<double>& stddev, std::uint32_t seed = 0) : mean(mean), stddev(stddev), zig(seed) {}
				Vector3<double> operator()() {
					const double a = mean.x + stddev.x * zig();
					const double b = mean.y + stddev.y * zig();
					const double c = mean.z + stddev.z * zig();
					return { a, b, c };
				}
				void seed(std::uint32_t s) { zig = ziggurat_normal_distribution(s); }
			};
		}

		// species, then position, then velocity; re-seeding derives the sub-seeds as s -> (1023*s)&s (32-bit)
		template<typename PositionDist, typename VelocityDist, typename SpeciesDist>
		class generic_particle_generator {
			SpeciesDist speciesGen;
			PositionDist posGen;
			VelocityDist velGen;
		public:
			generic_particle_generator(const PositionDist& pos, const VelocityDist& vel, const SpeciesDist& spec) : speciesGen(spec), posGen(pos), velGen(vel) {}
			Particle operator()() {
				Particle p = speciesGen();
				p.position = posGen();
				p.velocity = velGen();
				return p;
			}
			void seed(std::uint32_t s) {
				speciesGen.seed(s);
				s = (1023 * s) & s;
				posGen.seed(s);
				s = (1023 * s) & s;
				velGen.seed(s);
			}
		};

		// position generator seeded with seed+1, velocity generator with seed+2 (reference hdr/cell.h:316, :349, :382)
		template<typename SpeciesGen = species::electron>
		struct uniform : generic_particle_generator<vector::uniform, vector::uniform, SpeciesGen> {
			using super = generic_particle_generator<vector::uniform, vector::uniform, SpeciesGen>;
			uniform(const Vector3<double>& minPos, const Vector3<double>& maxPos, const Vector3<double>& minVel, const Vector3<double>& maxVel, std::uint32_t seed = 0)
				: super({ minPos, maxPos, seed + 1 }, { minVel, maxVel, seed + 2 }, SpeciesGen()) {}
		};
		template<typename SpeciesGen = species::electron>
		struct normal : generic_particle_generator<vector::normal, vector::uniform, SpeciesGen> {
			using super = generic_particle_generator<vector::normal, vector::uniform, SpeciesGen>;
			normal(const Vector3<double>& center, const Vector3<double>& stddev, const Vector3<double>& minVel, const Vector3<double>& maxVel, std::uint32_t seed = 0)
				: super({ center, stddev, seed + 1 }, { minVel, maxVel, seed + 2 }, SpeciesGen()) {}
		};
		template<typename SpeciesGen = species::electron>
		struct uniform_pos_normal_speed : generic_particle_generator<vector::uniform, vector::normal, SpeciesGen> {
			using super = generic_particle_generator<vector::uniform, vector::normal, SpeciesGen>;
			uniform_pos_normal_speed(const Vector3<double>& minPos, const Vector3<double>& maxPos, const Vector3<double>& center, const Vector3<double>& stddev, std::uint32_t seed = 0)
				: super({ minPos, maxPos, seed + 1 }, { center, stddev, seed + 2 }, SpeciesGen()) {}
		};
		template<typename SpeciesGen = species::electron>
		struct uniform_pos_normal_speed_r : generic_particle_generator<vector::uniform_r, vector::normal, SpeciesGen> {
			using super = generic_particle_generator<vector::uniform_r, vector::normal, SpeciesGen>;
			uniform_pos_normal_speed_r(const Vector3<double>& minPos, const Vector3<double>& maxPos, const Vector3<double>& center, const Vector3<double>& stddev, std::uint32_t seed = 0)
				: super({ minPos, maxPos, seed + 1 }, { center, stddev, seed + 2 }, SpeciesGen()) {}
		};

		// rejection of everything outside a ball
		template<typename Distribution>
		class spherical {
			Distribution dist;
			Vector3<double> center;
			double radius;
		public:
			spherical(const Distribution& dist, const Vector3<double>& center, double radius) : dist(dist), center(center), radius(radius) {}
			Particle operator()() {
				for (int attempt = 0; attempt < 100000; ++attempt) {
					Particle p = dist();
					if (std::sqrt(sumOfSquares(p.position - center)) <= radius) return p;
				}
				return {};
			}
			void seed(std::uint32_t s) { dist.seed(s); }
		};
		template<typename Distribution>
		spherical<Distribution> make_spherical(const Distribution& dist, const Vector3<double>& center = 0, double radius = 1) { return { dist, center, radius }; }
	}

	// ---- initCells (reference hdr/cell.h:493-776) ---------------------------------------------------------------

	namespace detail {
		// box of a cell and its per-cell seed (x*1023 + y)*1023 + z, truncated to 32 bits
		struct CellBox { Vector3<double> low, high; std::uint32_t seed; };
		inline CellBox cellBox(const UniverseProperties& properties, const coordinate_type& pos) {
			const auto& w = properties.cellWidth;
			CellBox b;
			b.low = Vector3<double>{ w.x * pos.x, w.y * pos.y, w.z * pos.z };
			b.low += properties.origin;
			b.high = b.low + w;
			b.seed = (std::uint32_t)(((pos.x * 1023) + pos.y) * 1023 + pos.z);
			return b;
		}
	}

	namespace detail {
		// Phase 1 of the generic initCells (reference hdr/cell.h:506-594), sequential: histogram of 100 pseudo particles per
		// cell, each standing for numParticles / (100 numCells) particles; what the truncation leaves over is spread evenly,
		// the remainder over the first cells of the z-fastest flatten.  Returns the population of every cell.
		template<typename Distribution>
		std::vector<std::uint64_t> approximatePopulation(const UniverseProperties& properties, std::uint64_t numParticles, Distribution sampler) {
			const coordinate_type grid = properties.size;
			std::cout << "Approximating particle distribution ...\n";
			const std::int64_t numCells = grid.x * grid.y * grid.z;
			std::vector<std::uint64_t> population((std::size_t)numCells, 0);
			const std::int64_t pseudoParticles = numCells * 100;
			const std::uint64_t weight = numParticles / pseudoParticles;
			for (std::int64_t i = 0; i < pseudoParticles; ++i) {
				Particle p = sampler();
				while (!isInsideUniverse(properties, p)) p = sampler();
				const coordinate_type c = getCellCoordinates(properties, p);
				population[(std::size_t)((c.x * grid.y + c.y) * grid.z + c.z)] += weight;
			}
			std::uint64_t sum = 0, hi = 0, lo = std::numeric_limits<std::uint64_t>::max();
			for (std::uint64_t v : population) { sum += v; hi = std::max(hi, v); lo = std::min(lo, v); }
			const std::int64_t missing = (std::int64_t)(numParticles - sum);
			if (missing != 0) {
				const int everywhere = (int)(missing / numCells);
				const int firstCells = (int)(missing % numCells);
				for (std::int64_t lin = 0; lin < numCells; ++lin) population[(std::size_t)lin] += everywhere + (lin < firstCells ? 1 : 0);
			}
			// (the reference prints the statistics of the uncorrected histogram here, hdr/cell.h:559-591)
			std::cout << "Number of particles in cells (min/avg/max): " << lo << "/" << (sum / numCells) << "/" << hi << "\n";
			return population;
		}

		// Phase 2 for ONE cell (reference hdr/cell.h:596-632): the distribution re-seeded with the cell's seed gives species and
		// velocity, positions are uniform in the cell's box from their own engine, charge and mass are the proton's.
		// Independent per cell, so a z-slab (or one GPU's share) can be generated without the rest of the universe.
		template<typename Distribution>
		void populateCell(const UniverseProperties& properties, const coordinate_type& pos, std::uint64_t n, const Distribution& sampler, Particle* out) {
			const double e = 1.602176565e-19;      // elementary charge (C)
			const double m = 1.672621777e-27;      // proton mass (kg)
			const CellBox box = cellBox(properties, pos);
			Distribution mine = sampler;
			mine.seed(box.seed);
			distribution::vector::uniform place(box.low, box.high, box.seed);
			for (std::uint64_t i = 0; i < n; ++i) {
				Particle p = mine();
				p.position = place();
				p.q = e;
				p.qom = e / m;
				out[i] = p;
			}
		}
	}

	// Any distribution (the `.inp` runs): per-cell populations follow a coarse Monte-Carlo estimate of the
	// distribution's density, positions inside a cell are uniform, velocities and species come from the
	// re-seeded distribution, charge and mass are the proton's.
	template<typename Distribution>
	Cells initCells(const UniverseProperties& properties, std::uint64_t numParticles, const Distribution& dist) {
		const coordinate_type grid = properties.size;
		Cells cells(grid);
		const std::vector<std::uint64_t> population = detail::approximatePopulation(properties, numParticles, dist);
		std::cout << "Populating cells ...\n";
		pfor(coordinate_type(0), grid, [&](const coordinate_type& pos) {
			Cell& cell = cells[pos];
			cell.particles.resize(population[cells.flatten(pos)]);
			detail::populateCell(properties, pos, cell.particles.size(), dist, cell.particles.data());
		});
		return cells;
	}

	// The uniform benchmark load (`:U:N`): N/numCells particles per cell (+1 in the first N%numCells cells),
	// positions uniform in the cell, velocities uniform in [-0.2, 0.2)^3, electrons.  The distribution object
	// is only a tag (the reference ignores it too, hdr/cell.h:638, :670).
	inline Cells initCells(const UniverseProperties& properties, std::uint64_t numParticles, const distribution::uniform<>&) {
		const coordinate_type grid = properties.size;
		Cells cells(grid);
		std::cout << "Sorting in uniformly distributed particles ...\n";
		const std::uint64_t numCells = (std::uint64_t)(grid.x * grid.y * grid.z);
		const std::uint64_t perCell = numParticles / numCells, extra = numParticles % numCells;
		std::cout << "  particles / cell: " << perCell << " (+1)\n";
		pfor(coordinate_type(0), grid, [&](const coordinate_type& pos) {
			const detail::CellBox box = detail::cellBox(properties, pos);
			distribution::uniform<> next(box.low, box.high, Vector3<double>{ -0.2, -0.2, -0.2 }, Vector3<double>{ +0.2, +0.2, +0.2 }, box.seed);
			const std::uint64_t n = perCell + (cells.flatten(pos) < extra ? 1 : 0);
			Cell& cell = cells[pos];
			cell.particles.reserve(n);
			for (std::uint64_t i = 0; i < n; ++i) cell.particles.push_back(next());
		});
		return cells;
	}

	// The legacy lattice initialisation from a parameter file (used by the reference's tests only): a regular
	// sub-lattice of npcelx*npcely*npcelz particles per cell with Box-Muller thermal velocities.
	inline Cells initCells(const Parameters& params, const InitProperties& initProperties, const UniverseProperties& properties) {
		Cells cells(properties.size);
		const Vector3<unsigned> per = initProperties.particlesPerCell[0];
		const unsigned perCell = per.x * per.y * per.z;
		double charge = params.qom[0] / std::fabs(params.qom[0]);
		charge = charge * (properties.cellWidth.x * properties.cellWidth.y * properties.cellWidth.z) / perCell;
		const double fourPI = 16.0 * std::atan(1.0);
		charge = charge * params.rhoInit[0] / fourPI;

		forEachCoordinate(coordinate_type(0), properties.size, [&](const coordinate_type& pos) {
			Cell& cell = cells[pos];
			const Vector3<double> corner = getOriginOfCell(pos, properties);
			std::minstd_rand engine((unsigned)(pos[0] * 10000 + pos[1] * 100 + pos[2]));
			const double engineMax = std::minstd_rand::max();
			auto harvest = [&]() { return (double)engine() / engineMax; };
			for (unsigned i = 0; i < per.x; ++i) for (unsigned j = 0; j < per.y; ++j) for (unsigned k = 0; k < per.z; ++k) {
				Particle p;
				p.position.x = (i + 0.5) * (properties.cellWidth.x / per.x) + corner.x;
				p.position.y = (j + 0.5) * (properties.cellWidth.y / per.y) + corner.y;
				p.position.z = (k + 0.5) * (properties.cellWidth.z / per.z) + corner.z;
				const double prob0 = std::sqrt(-2.0 * std::log(1.0 - 0.999999 * harvest()));
				const double theta0 = 2.0 * M_PI * harvest();
				const double prob1 = std::sqrt(-2.0 * std::log(1.0 - 0.999999 * harvest()));
				const double theta1 = 2.0 * M_PI * harvest();
				p.velocity.x = params.u0[0] + params.uth[0] * (prob0 * std::cos(theta0));
				p.velocity.y = params.v0[0] + params.vth[0] * (prob0 * std::sin(theta0));
				p.velocity.z = params.w0[0] + params.wth[0] * (prob1 * std::cos(theta1));
				p.qom = params.qom[0];
				p.q = charge;
				cell.particles.push_back(p);
			}
		});
		return cells;
	}

	// ---- scalar versions of two per-timestep pieces, host-callable for known-answer checks ----------------------------
	// ("host-callable scalar versions ... stay available for the KATs").  `simulateSteps` never calls them:
	// in the time loop the interpolation is `trilinear` of the interpolating mover and the projection is the per-cell corner
	// sums of the movers plus k_gather_current (csrc/particle_kernels.cu).  The movers, export and import have no host
	// version at all.

	// value at the fractional position `pos` in [0,1]^3 between 8 corner values, divided by `vol` (reference hdr/cell.h:867-885)
	template<typename T>
	T trilinearInterpolationF2P(const T corners[2][2][2], const Vector3<double>& pos, const double vol) {
		const double along[3][2] = { { 1 - pos.x, pos.x }, { 1 - pos.y, pos.y }, { 1 - pos.z, pos.z } };
		T res = T(0);
		for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) for (int k = 0; k < 2; ++k)
			res += corners[i][j][k] * (along[0][i] * along[1][j] * along[2][k]);
		return res / vol;
	}

	// J on node `pos` = sum over the particles of the 8 cells around the node (periodic) of q v w, w the trilinear weight of
	// the node seen from the particle's cell, divided by the cell volume and by 8 (reference hdr/cell.h:787-824)
	inline void projectToDensityField(const UniverseProperties& universeProperties, const Cells& cells, const coordinate_type& pos, CurrentDensity& density) {
		const coordinate_type size = universeProperties.size;
		Vector3<double> Js(0.0);
		// a = 0: the cell below the node along that axis (the node is its upper corner: weight = relative position),
		// a = 1: the cell above (the node is its lower corner: weight = 1 - relative position)
		for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) for (int c = 0; c < 2; ++c) {
			coordinate_type cur{ pos.x + a - 1, pos.y + b - 1, pos.z + c - 1 };
			for (int d = 0; d < 3; ++d) cur[d] = cur[d] < 0 ? size[d] - 1 : (cur[d] >= size[d] ? 0 : cur[d]);
			const Vector3<double> cellOrigin = getOriginOfCell(cur, universeProperties);
			for (const Particle& p : cells[cur].particles) {
				const Vector3<double> rel = elementwiseDivision(p.position - cellOrigin, universeProperties.cellWidth);
				const double fac = (a ? 1 - rel.x : rel.x) * (b ? 1 - rel.y : rel.y) * (c ? 1 - rel.z : rel.z);
				Js += p.q * p.velocity * fac;
			}
		}
		const double vol = universeProperties.cellWidth.x * universeProperties.cellWidth.y * universeProperties.cellWidth.z;
		density[pos].J = (Js / vol) / 8.0;
	}

	// ---- conserved-quantity folds over a host universe (reference hdr/cell.h:1315-1355) --------------------------
	// Sequential, in the reference's order (particles in cell order; cells lexicographic).  simulateSteps does not use
	// them: while the universe is on the GPU the same sums come from ipic_diagnostics (device tree reductions).

	inline double getParticlesKineticEnergy(const Cell& cell) {
		double res = 0.0;
		for (const Particle& p : cell.particles) res += 0.5 * (p.q / p.qom) * sumOfSquares(p.velocity);
		return res;
	}

	inline double getParticlesMomentum(const Cell& cell) {
		double res = 0.0;
		for (const Particle& p : cell.particles) res += (p.q / p.qom) * std::sqrt(sumOfSquares(p.velocity));
		return res;
	}

	template<class T>
	double getTotalParticlesEnergy(const Cells& cells, T func) {
		double res = 0.0;
		forEachCoordinate(coordinate_type(0), cells.size(), [&](const coordinate_type& index) { res += func(cells[index]); });
		return res;
	}

	// ---- outputs (reference hdr/cell.h:1360-1429) ---------------------------------------------------------------

	// "<base>.out": "[Nx,Ny,Nz]", one "x,y,z:count" line per cell in x-major order, a blank line, "Total: n"
	inline void outputNumberOfParticlesPerCell(const Cells& cells, const std::string& outputFilename) {
		std::ofstream out(outputFilename);
		out << cells.size() << "\n";
		std::uint64_t total = 0;
		forEachCoordinate(coordinate_type(0), cells.size(), [&](const coordinate_type& p) {
			out << p.x << "," << p.y << "," << p.z << ":" << cells[p].particles.size() << "\n";
			total += cells[p].particles.size();
		});
		out << "\nTotal: " << total << "\n";
	}

	template<typename Stream>
	void outputParticlePositions(const Cells& cells, Stream& out) {
		forEachCoordinate(coordinate_type(0), cells.size(), [&](const coordinate_type& c) {
			for (const Particle& p : cells[c].particles) out << p.position.x << " " << p.position.y << " " << p.position.z << "\n";
		});
	}
	inline void outputParticlePositions(const Cells& cells, const std::string& filename) {
		std::ofstream out(filename);
		outputParticlePositions(cells, out);
	}

} // namespace ipic3d
