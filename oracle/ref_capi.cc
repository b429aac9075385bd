// TEST INFRASTRUCTURE ONLY (oracle/): never linked into, loaded by or called from the product.
//
// C handle API over the UNMODIFIED reference headers (/root/reference/code/app/include), compiled
// against the AllScale header shim in oracle/shim.  Built by oracle/Makefile into
// oracle/_ref/libipic3d_ref.so (sequential pfor) and oracle/_ref/libipic3d_ref_mt.so (std::thread pfor).
// It exists to (1) pin the plain-C oracle (oracle/ipic3d_oracle.c) against the reference's own code,
// (2) generate golden vectors, (3) serve as the timed CPU baseline (bench.py --impl reference).
//
// Nothing here restates the reference's arithmetic: every numeric step is a call into the
// reference's functions.  The only code of our own is set-up glue that mirrors
//   main()                 code/app/src/ipic3d.cxx:32-83      -> ref_create_from_inp
//   processDistribution()  hdr/benchmark.h:57-79               -> ref_create_uniform
//   the commented STEP 1/2 loop of simulateSteps hdr/simulator.h:189-209 -> ref_step_full
//   the commented corner gather of moveParticles hdr/cell.h:905-937      -> mover_interpolated
#include <chrono>
#include <cstdint>
#include <cstring>
#include <sstream>
#include <string>

#include "ipic3d/app/benchmark.h"
#include "ipic3d/app/cell.h"
#include "ipic3d/app/field.h"
#include "ipic3d/app/parameters.h"
#include "ipic3d/app/simulator.h"
#include "ipic3d/app/universe.h"

#include "../include/ipic3d_b200.h"

using namespace ipic3d;
using Coord = utils::Coordinate<3>;

namespace {

	struct Handle {
		Universe universe;
		InitProperties init;
		Handle(Universe&& u, const InitProperties& i) : universe(std::move(u)), init(i) {}
	};

	UniverseProperties toProps(const ipic_props& p) {
		UniverseProperties r;
		r.size = { p.size[0], p.size[1], p.size[2] };
		r.cellWidth = { p.cellWidth[0], p.cellWidth[1], p.cellWidth[2] };
		r.dt = p.dt;
		r.speedOfLight = p.speedOfLight;
		r.planetRadius = p.planetRadius;
		r.objectCenter = { p.objectCenter[0], p.objectCenter[1], p.objectCenter[2] };
		r.origin = { p.origin[0], p.origin[1], p.origin[2] };
		r.externalMagneticField = { p.externalMagneticField[0], p.externalMagneticField[1], p.externalMagneticField[2] };
		r.FieldOutputCycle = 0;
		r.ParticleOutputCycle = 0;
		return r;
	}

	void fromProps(const UniverseProperties& r, ipic_props* p) {
		for (int i = 0; i < 3; i++) {
			p->size[i] = r.size[i];
			p->cellWidth[i] = r.cellWidth[i];
			p->objectCenter[i] = r.objectCenter[i];
			p->origin[i] = r.origin[i];
			p->externalMagneticField[i] = r.externalMagneticField[i];
		}
		p->dt = r.dt;
		p->speedOfLight = r.speedOfLight;
		p->planetRadius = r.planetRadius;
	}

	// The interpolating mover exactly as the reference has it in comments (cell.h:905-937), followed by
	// the live sub-cycling + Boris loop (cell.h:947-960), using the reference's own functions.
	void mover_interpolated(const UniverseProperties& properties, Cell& cell, const Coord& pos, const Field& field) {
		if (cell.particles.empty()) return;
		Vector3<double> Es[2][2][2];
		Vector3<double> Bs[2][2][2];
		for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) for (int k = 0; k < 2; k++) {
			Coord cur({ pos[0] + i + 1, pos[1] + j + 1, pos[2] + k + 1 });
			Es[i][j][k] = field[cur].E;
			Bs[i][j][k] = field[cur].B;
		}
		const auto cellOrigin = getOriginOfCell(pos, properties);
		double vol = properties.cellWidth.x * properties.cellWidth.y * properties.cellWidth.z;
		for (std::size_t i = 0; i < cell.particles.size(); ++i) {
			Particle& p = cell.particles[i];
			const auto relPos = allscale::utils::elementwiseDivision((p.position - cellOrigin), (properties.cellWidth));
			auto E = trilinearInterpolationF2P(Es, relPos, vol);
			auto B = trilinearInterpolationF2P(Bs, relPos, vol);
			double B_mag = allscale::utils::sumOfSquares(B);
			double dt_sub = M_PI * properties.speedOfLight / (4.0 * fabs(p.qom) * B_mag);
			int sub_cycles = int(properties.dt / dt_sub) + 1;
			sub_cycles = std::min(sub_cycles, 100);
			dt_sub = properties.dt / double(sub_cycles);
			for (int cyc_cnt = 0; cyc_cnt < sub_cycles; cyc_cnt++) {
				p.updateVelocity(E, B, dt_sub);
				p.updatePosition(dt_sub);
			}
		}
	}
}

extern "C" {

// ---- construction -------------------------------------------------------------------------

// Mirrors main() for "<file>.inp" runs, including the hard-coded Earth overrides (ipic3d.cxx:48-83).
void* ref_create_from_inp(const char* path, ipic_props* outProps, uint64_t* outCycles, uint64_t* outParticles) {
	auto params = Parameters(std::string(path));
	InitProperties initProperties = InitProperties(params);
	UniverseProperties universeProperties = UniverseProperties(params);
	universeProperties.dt = 0.01;
	universeProperties.speedOfLight = 299792458;
	universeProperties.planetRadius = 6378137;
	universeProperties.cellWidth *= universeProperties.planetRadius;
	auto universeSize = elementwiseProduct(universeProperties.cellWidth, universeProperties.size);
	universeProperties.objectCenter = { 0.0, 0.0, 0.0 };
	universeProperties.origin = universeProperties.objectCenter - universeSize / 2.0;
	universeProperties.externalMagneticField = { 0.0, 0.0, 3.07e-5 };
	universeProperties.useCase = UseCase::Dipole;
	int numParticles = universeProperties.size.x * universeProperties.size.y * universeProperties.size.z;
	numParticles = numParticles * initProperties.particlesPerCell[0].x * initProperties.particlesPerCell[0].y * initProperties.particlesPerCell[0].z;
	double e = 1.602176565e-19;
	double K = 1e7 * e;
	double m = 1.672621777e-27;
	double v_mod = universeProperties.speedOfLight / sqrt(1.0 + (m * universeProperties.speedOfLight * universeProperties.speedOfLight) / K);
	auto low = universeProperties.origin + 0.125 * universeSize;
	auto hig = low + 0.75 * universeSize;
	auto dist = distribution::uniform_pos_normal_speed<>(low, hig, Vector3<double>{ 0, 0, 0 }, Vector3<double>{ v_mod, v_mod, v_mod });
	auto universe = createUniverseFromDistribution(universeProperties, initProperties, numParticles, dist);
	if (outProps) fromProps(universe.properties, outProps);
	if (outCycles) *outCycles = params.ncycles;
	if (outParticles) *outParticles = (uint64_t)numParticles;
	return new Handle(std::move(universe), initProperties);
}

// Mirrors processDistribution/processUniform for ":U:N" (benchmark.h:57-92).
void* ref_create_uniform(uint64_t numParticles, ipic_props* outProps) {
	UniverseProperties universeProperties;
	universeProperties.cellWidth = CELL_WIDTH;
	universeProperties.size = GRID_SIZE;
	universeProperties.dt = DELTA_T;
	universeProperties.FieldOutputCycle = 0;
	InitProperties initProperties;
	initProperties.driftVelocity.push_back(0);
	auto dist = distribution::uniform<>(Vector3<double>{ 0, 0, 0 }, UNIVERSE_SIZE, Vector3<double>{ -0.2, -0.2, -0.2 }, Vector3<double>{ +0.2, +0.2, +0.2 });
	auto universe = createUniverseFromDistribution(universeProperties, initProperties, numParticles, dist);
	if (outProps) fromProps(universe.properties, outProps);
	return new Handle(std::move(universe), initProperties);
}

// The other synthetic loads of processBenchmark (benchmark.h:94-138): 'C' cluster, 'E' explosion, 'B' beam, 'U' uniform.
void* ref_create_benchmark(char kind, uint64_t numParticles, ipic_props* outProps) {
	UniverseProperties universeProperties;
	universeProperties.cellWidth = CELL_WIDTH;
	universeProperties.size = GRID_SIZE;
	universeProperties.dt = DELTA_T;
	universeProperties.FieldOutputCycle = 0;
	InitProperties initProperties;
	initProperties.driftVelocity.push_back(0);
	const uint64_t numTimeSteps = NUM_TIME_STEPS;
	Handle* h = nullptr;
	if (kind == 'U') return ref_create_uniform(numParticles, outProps);
	if (kind == 'C') h = new Handle(createUniverseFromDistribution(universeProperties, initProperties, numParticles,
		distribution::normal<>(UNIVERSE_SIZE / 2, UNIVERSE_SIZE / 5, Vector3<double>{ -0.2, -0.2, -0.2 }, Vector3<double>{ +0.2, +0.2, +0.2 })), initProperties);
	if (kind == 'E') h = new Handle(createUniverseFromDistribution(universeProperties, initProperties, numParticles,
		distribution::make_spherical(distribution::generic_particle_generator<distribution::vector::uniform, distribution::vector::normal, distribution::species::electron>(
			{ UNIVERSE_SIZE * 0.4, UNIVERSE_SIZE * 0.6, 0 }, { Vector3<double>{ 0 }, Vector3<double>{ 1.5 }, 1 }, {}), UNIVERSE_SIZE / 2, UNIVERSE_SIZE.x / 10)), initProperties);
	if (kind == 'B') h = new Handle(createUniverseFromDistribution(universeProperties, initProperties, numParticles,
		distribution::normal<>(UNIVERSE_SIZE / 100, UNIVERSE_SIZE / 500, UNIVERSE_SIZE / numTimeSteps * 0.95 / DELTA_T, UNIVERSE_SIZE / numTimeSteps * 0.95 / DELTA_T)), initProperties);
	if (h && outProps) fromProps(h->universe.properties, outProps);
	return h;
}

// The reference's parser alone: every member it fills, as "name=value;" text (same format as the host mirror's hook).
int ref_parse(const char* path, char* out, int capacity) {
	const Parameters p{ std::string(path) };
	std::ostringstream s;
	s.precision(17);
	auto list = [&s](const char* name, const auto& v) { s << name << "="; for (std::size_t i = 0; i < v.size(); ++i) s << (i ? "," : "") << v[i]; s << ";"; };
	s << "c=" << p.c << ";dt=" << p.dt << ";ncycles=" << p.ncycles << ";L=" << p.L.x << "," << p.L.y << "," << p.L.z
	  << ";center=" << p.objectCenter.x << "," << p.objectCenter.y << "," << p.objectCenter.z << ";planetRadius=" << p.planetRadius
	  << ";delta=" << p.delta << ";ncells=" << p.ncells.x << "," << p.ncells.y << "," << p.ncells.z
	  << ";dspace=" << p.dspace.x << "," << p.dspace.y << "," << p.dspace.z << ";ns=" << p.ns
	  << ";B0=" << p.B0.x << "," << p.B0.y << "," << p.B0.z << ";B1=" << p.B1.x << "," << p.B1.y << "," << p.B1.z
	  << ";useCase=" << (p.useCase == UseCase::Dipole ? "Dipole" : "Test") << ";wmethod=" << p.wmethod << ";SimName=" << p.SimName
	  << ";PoissonCorrection=" << p.PoissonCorrection << ";SaveDirName=" << p.SaveDirName << ";FieldOutputCycle=" << p.FieldOutputCycle
	  << ";FieldOutputTag=" << p.FieldOutputTag << ";MomentsOutputTag=" << p.MomentsOutputTag << ";ParticlesOutputCycle=" << p.ParticlesOutputCycle
	  << ";ParticlesOutputTag=" << p.ParticlesOutputTag << ";";
	list("npcelx", p.npcelx); list("npcely", p.npcely); list("npcelz", p.npcelz); list("qom", p.qom); list("rhoInit", p.rhoInit);
	list("uth", p.uth); list("vth", p.vth); list("wth", p.wth); list("u0", p.u0); list("v0", p.v0); list("w0", p.w0);
	const std::string text = s.str();
	if ((int)text.size() + 1 > capacity) return -1;
	std::memcpy(out, text.c_str(), text.size() + 1);
	return (int)text.size();
}

// Same generator on an arbitrary cubic grid (for small-case parity tests).
void* ref_create_uniform_grid(const ipic_props* props, uint64_t numParticles) {
	UniverseProperties universeProperties = toProps(*props);
	InitProperties initProperties;
	initProperties.driftVelocity.push_back(0);
	auto dist = distribution::uniform<>(Vector3<double>{ 0, 0, 0 }, Vector3<double>{ 1, 1, 1 }, Vector3<double>{ -0.2, -0.2, -0.2 }, Vector3<double>{ +0.2, +0.2, +0.2 });
	auto universe = createUniverseFromDistribution(universeProperties, initProperties, numParticles, dist);
	return new Handle(std::move(universe), initProperties);
}

// Empty universe (no particles, zero fields) with the given properties.
void* ref_create_empty(const ipic_props* props) {
	InitProperties initProperties;
	initProperties.driftVelocity.push_back(0);
	return new Handle(Universe(toProps(*props)), initProperties);
}

void ref_destroy(void* h) { delete static_cast<Handle*>(h); }

void ref_get_props(void* h, ipic_props* p) { fromProps(static_cast<Handle*>(h)->universe.properties, p); }

// initFields + initBcFields + initCurrentDensity with the given drift velocity / B0 (field.h:81-157)
void ref_init_fields(void* h, const double* drift, const double* B0) {
	auto& u = static_cast<Handle*>(h)->universe;
	InitProperties ip;
	ip.driftVelocity.push_back({ drift[0], drift[1], drift[2] });
	ip.magneticField = { B0[0], B0[1], B0[2] };
	u.field = initFields(ip, u.properties);
	u.bcfield = initBcFields(u.properties, u.field);
	u.currentDensity = initCurrentDensity(u.properties);
}

// ---- state access -------------------------------------------------------------------------

uint64_t ref_total_particles(void* h) { return countParticlesInDomain(static_cast<Handle*>(h)->universe); }

void ref_get_counts(void* h, uint64_t* counts) {
	auto& u = static_cast<Handle*>(h)->universe;
	auto s = u.cells.size();
	uint64_t n = 0;
	for (int64_t x = 0; x < s.x; x++) for (int64_t y = 0; y < s.y; y++) for (int64_t z = 0; z < s.z; z++)
		counts[n++] = u.cells[{ x, y, z }].particles.size();
}

void ref_get_particles(void* h, double* aos) {
	auto& u = static_cast<Handle*>(h)->universe;
	auto s = u.cells.size();
	static_assert(sizeof(Particle) == 64, "Particle layout");
	for (int64_t x = 0; x < s.x; x++) for (int64_t y = 0; y < s.y; y++) for (int64_t z = 0; z < s.z; z++) {
		auto& ps = u.cells[{ x, y, z }].particles;
		if (!ps.empty()) std::memcpy(aos, ps.data(), ps.size() * sizeof(Particle));
		aos += ps.size() * 8;
	}
}

void ref_set_cells(void* h, const uint64_t* counts, const double* aos) {
	auto& u = static_cast<Handle*>(h)->universe;
	auto s = u.cells.size();
	uint64_t n = 0;
	for (int64_t x = 0; x < s.x; x++) for (int64_t y = 0; y < s.y; y++) for (int64_t z = 0; z < s.z; z++) {
		auto& ps = u.cells[{ x, y, z }].particles;
		ps.resize(counts[n]);
		if (counts[n]) std::memcpy(ps.data(), aos, counts[n] * sizeof(Particle));
		aos += counts[n] * 8;
		n++;
	}
}

#define REF_GRID_COPY(NAME, MEMBER, ELEM, W, DIR) \
void NAME(void* h, double* buf) { \
	auto& g = static_cast<Handle*>(h)->universe.MEMBER; auto s = g.size(); \
	static_assert(sizeof(ELEM) == W * sizeof(double), "layout"); \
	for (int64_t x = 0; x < s.x; x++) for (int64_t y = 0; y < s.y; y++) for (int64_t z = 0; z < s.z; z++) { \
		DIR; buf += W; } \
}
REF_GRID_COPY(ref_get_field, field, FieldNode, 9, std::memcpy(buf, &g[Coord(x, y, z)], sizeof(FieldNode)))
REF_GRID_COPY(ref_set_field, field, FieldNode, 9, std::memcpy(&g[Coord(x, y, z)], buf, sizeof(FieldNode)))
REF_GRID_COPY(ref_get_bcfield, bcfield, BcFieldCell, 3, std::memcpy(buf, &g[Coord(x, y, z)], sizeof(BcFieldCell)))
REF_GRID_COPY(ref_set_bcfield, bcfield, BcFieldCell, 3, std::memcpy(&g[Coord(x, y, z)], buf, sizeof(BcFieldCell)))
REF_GRID_COPY(ref_get_density, currentDensity, DensityNode, 3, std::memcpy(buf, &g[Coord(x, y, z)], sizeof(DensityNode)))
REF_GRID_COPY(ref_set_density, currentDensity, DensityNode, 3, std::memcpy(&g[Coord(x, y, z)], buf, sizeof(DensityNode)))

void ref_write_counts(void* h, const char* filename) {
	outputNumberOfParticlesPerCell(static_cast<Handle*>(h)->universe.cells, std::string(filename));
}

// the four columns writeOutputData prints (simulator.h:71-91), computed by the reference's own reductions
void ref_diagnostics(void* h, double* out4) {
	Universe& universe = static_cast<Handle*>(h)->universe;
	auto getE = [](const auto& field, const auto& index) { return field[index].E; };
	auto getB = [](const auto& field, const auto& index) { return (field[index].B + field[index].Bext); };
	out4[0] = getTotalParticlesEnergy(universe.cells, getParticlesMomentum);
	out4[1] = getFieldEnergy(universe.field, universe.properties, getE);
	out4[2] = getFieldEnergy(universe.field, universe.properties, getB);
	out4[3] = getTotalParticlesEnergy(universe.cells, getParticlesKineticEnergy);
}

uint64_t ref_count_misplaced(void* h) {
	auto& u = static_cast<Handle*>(h)->universe;
	auto s = u.cells.size();
	uint64_t bad = 0;
	for (int64_t x = 0; x < s.x; x++) for (int64_t y = 0; y < s.y; y++) for (int64_t z = 0; z < s.z; z++)
		for (auto& p : u.cells[{ x, y, z }].particles) if (!isInside(u.properties, { x, y, z }, p)) bad++;
	return bad;
}

// ---- stepping -----------------------------------------------------------------------------

// The reference's shipped time loop, untouched: simulateSteps<>(n, universe) (simulator.h:132).
double ref_simulate_steps(void* h, uint64_t n) {
	auto& u = static_cast<Handle*>(h)->universe;
	auto t0 = std::chrono::high_resolution_clock::now();
	simulateSteps(n, u);
	auto t1 = std::chrono::high_resolution_clock::now();
	return std::chrono::duration<double>(t1 - t0).count();
}

// The intended full PIC step with the reference's own functions in the order of the commented
// loop (simulator.h:189-223).  flags as in ipic3d_b200.h.
double ref_step_full(void* h, uint64_t n, uint32_t flags) {
	using allscale::api::user::algorithm::pfor;
	auto& u = static_cast<Handle*>(h)->universe;
	auto zero = Coord(0);
	auto size = u.cells.size();
	TransferBuffers transfers(size);
	auto fieldStart = Coord(1);
	auto fieldEnd = u.field.size() - Coord(1);
	auto t0 = std::chrono::high_resolution_clock::now();
	for (uint64_t i = 0; i < n; i++) {
		if (flags & IPIC_PROJECT_CURRENT) {
			ipic3d::detail::default_particle_to_field_projector proj;
			pfor(zero, u.currentDensity.size(), [&](const Coord& pos) { proj(u.properties, u.cells, pos, u.currentDensity); });
		}
		if (flags & IPIC_UPDATE_BOUNDARIES) updateFieldsOnBoundaries(u.field, u.bcfield);
		if (flags & (IPIC_SOLVER_FORWARD | IPIC_SOLVER_LEAPFROG)) {
			// sequential, lexicographic: the reference updates in place inside the loop
			allscale::api::user::algorithm::detail::forEach(fieldStart, fieldEnd, [&](const Coord& pos) {
				if (flags & IPIC_SOLVER_FORWARD) ipic3d::detail::forward_field_solver()(u.properties, pos, u.currentDensity, u.field, u.bcfield);
				else ipic3d::detail::leapfrog_field_solver()(u.properties, pos, u.currentDensity, u.field, u.bcfield);
			});
		}
		pfor(zero, size, [&](const Coord& pos) {
			if (flags & IPIC_MOVER_INTERPOLATED) mover_interpolated(u.properties, u.cells[pos], pos, u.field);
			else moveParticles(u.properties, u.cells[pos], pos, u.field);
			if (!(flags & IPIC_SKIP_MIGRATION)) exportParticles(u.properties, u.cells[pos], pos, transfers);
		});
		if (!(flags & IPIC_SKIP_MIGRATION))
			pfor(zero, size, [&](const Coord& pos) { importParticles(u.properties, u.cells[pos], pos, transfers); });
	}
	auto t1 = std::chrono::high_resolution_clock::now();
	return std::chrono::duration<double>(t1 - t0).count();
}

// ---- unit-level wrappers (for pinning the oracle function by function) -------------------------

void ref_update_velocity(double* p8, const double* E, const double* B, double dt) {
	Particle p; std::memcpy(&p, p8, 64);
	p.updateVelocity({ E[0], E[1], E[2] }, { B[0], B[1], B[2] }, dt);
	std::memcpy(p8, &p, 64);
}

void ref_update_position(double* p8, double dt) {
	Particle p; std::memcpy(&p, p8, 64);
	p.updatePosition(dt);
	std::memcpy(p8, &p, 64);
}

void ref_trilinear(const double* corners24, const double* pos, double vol, double* out3) {
	Vector3<double> c[2][2][2];
	for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) for (int k = 0; k < 2; k++)
		c[i][j][k] = { corners24[((i * 2 + j) * 2 + k) * 3], corners24[((i * 2 + j) * 2 + k) * 3 + 1], corners24[((i * 2 + j) * 2 + k) * 3 + 2] };
	auto r = trilinearInterpolationF2P(c, Vector3<double>{ pos[0], pos[1], pos[2] }, vol);
	out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}

void ref_project_node(void* h, int64_t x, int64_t y, int64_t z) {
	auto& u = static_cast<Handle*>(h)->universe;
	projectToDensityField(u.properties, u.cells, { x, y, z }, u.currentDensity);
}

void ref_solve_forward_node(void* h, int64_t x, int64_t y, int64_t z) {
	auto& u = static_cast<Handle*>(h)->universe;
	solveFieldForward(u.properties, { x, y, z }, u.currentDensity, u.field, u.bcfield);
}

void ref_solve_leapfrog_node(void* h, int64_t x, int64_t y, int64_t z) {
	auto& u = static_cast<Handle*>(h)->universe;
	solveFieldLeapfrog(u.properties, { x, y, z }, u.currentDensity, u.field, u.bcfield);
}

void ref_update_boundaries(void* h) {
	auto& u = static_cast<Handle*>(h)->universe;
	updateFieldsOnBoundaries(u.field, u.bcfield);
}

int ref_threads(void) {
#ifdef ALLSCALE_SHIM_THREADS
	return allscale::api::user::algorithm::detail::worker_count();
#else
	return 1;
#endif
}

} // extern "C"
