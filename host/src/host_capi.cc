// Host-side helpers of the B200 build exposed with C linkage for bench.py and the tests.
//
// This is NOT part of the device C ABI (include/ipic3d_b200.h); it is the part of the host C++
// mirror (host/include/ipic3d/app/) that produces the *inputs* of the hot path: the synthetic
// ":U:N" benchmark universe of the reference (benchmark.h:57-92 -> initCells(props, N, uniform<>)
// cell.h:638-697).  Bit-identical to the reference's generator because it uses the same
// libstdc++ engines in the same draw order:
//   per cell (x,y,z): seed = (x*1023 + y)*1023 + z (uint32)                       cell.h:668
//   positions : std::minstd_rand(seed+1) -> uniform_real_distribution x,y,z       cell.h:138-160, :316
//   velocities: std::minstd_rand(seed+2) -> uniform [-0.2, 0.2)^3                 cell.h:671-677
//   species   : electron, q = -1, qom = -25                                       cell.h:107-118
//   count     : N / numCells, +1 for flat cell index < N % numCells               cell.h:650-651, :684-687
#include <atomic>
#include <cstdint>
#include <random>
#include <thread>
#include <vector>

#include "../../include/ipic3d_b200.h"

namespace {

inline uint64_t cellPopulation(uint64_t n, uint64_t numCells, uint64_t lin) { return n / numCells + (lin < n % numCells ? 1 : 0); }

void fillCell(const ipic_props& p, int64_t x, int64_t y, int64_t z, uint64_t count, double* out) {
	const double lx = p.cellWidth[0] * (double)x + p.origin[0];
	const double ly = p.cellWidth[1] * (double)y + p.origin[1];
	const double lz = p.cellWidth[2] * (double)z + p.origin[2];
	const uint32_t seed = (uint32_t)(((x * 1023) + y) * 1023 + z);
	std::minstd_rand posEngine(seed + 1), velEngine(seed + 2);
	std::uniform_real_distribution<double> px(lx, lx + p.cellWidth[0]), py(ly, ly + p.cellWidth[1]), pz(lz, lz + p.cellWidth[2]);
	std::uniform_real_distribution<double> vx(-0.2, 0.2), vy(-0.2, 0.2), vz(-0.2, 0.2);
	for (uint64_t i = 0; i < count; i++, out += 8) {
		out[0] = px(posEngine); out[1] = py(posEngine); out[2] = pz(posEngine);
		out[3] = vx(velEngine); out[4] = vy(velEngine); out[5] = vz(velEngine);
		out[6] = -1.0; out[7] = -25.0;
	}
}

} // namespace

extern "C" {

// per-cell populations of the z-planes [zBegin, zBegin+zCount) in slab order (x*ny + y)*zCount + zl
IPIC_API int ipic_host_uniform_counts(const ipic_props* p, uint64_t numParticles, int zBegin, int zCount, uint64_t* counts) {
	if (!p || !counts || zBegin < 0 || zCount < 0 || zBegin + zCount > p->size[2]) return IPIC_ERR_ARG;
	const uint64_t numCells = (uint64_t)(p->size[0] * p->size[1] * p->size[2]);
	uint64_t* o = counts;
	for (int64_t x = 0; x < p->size[0]; x++) for (int64_t y = 0; y < p->size[1]; y++) for (int z = zBegin; z < zBegin + zCount; z++)
		*o++ = cellPopulation(numParticles, numCells, (uint64_t)((x * p->size[1] + y) * p->size[2] + z));
	return IPIC_OK;
}

// particles of the same cells, AoS of 8 doubles, cell-major in slab order; `threads` <= 0 = all cores
IPIC_API int ipic_host_uniform_particles(const ipic_props* p, uint64_t numParticles, int zBegin, int zCount, double* aos, int threads) {
	if (!p || !aos || zBegin < 0 || zCount < 0 || zBegin + zCount > p->size[2]) return IPIC_ERR_ARG;
	const uint64_t numCells = (uint64_t)(p->size[0] * p->size[1] * p->size[2]);
	const int64_t slabCells = p->size[0] * p->size[1] * zCount;
	std::vector<uint64_t> start((size_t)slabCells + 1, 0);
	for (int64_t c = 0; c < slabCells; c++) {
		const int64_t xy = c / zCount, z = zBegin + c % zCount, x = xy / p->size[1], y = xy % p->size[1];
		start[(size_t)c + 1] = start[(size_t)c] + cellPopulation(numParticles, numCells, (uint64_t)((x * p->size[1] + y) * p->size[2] + z));
	}
	if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
	if (threads <= 0) threads = 1;
	std::atomic<int64_t> next(0);
	auto work = [&]() {
		for (;;) {
			const int64_t c = next.fetch_add(1);
			if (c >= slabCells) return;
			const int64_t xy = c / zCount, z = zBegin + c % zCount, x = xy / p->size[1], y = xy % p->size[1];
			fillCell(*p, x, y, z, start[(size_t)c + 1] - start[(size_t)c], aos + start[(size_t)c] * 8);
		}
	};
	std::vector<std::thread> pool;
	for (int t = 1; t < threads; t++) pool.emplace_back(work);
	work();
	for (auto& t : pool) t.join();
	return IPIC_OK;
}

}

// ---- test hooks over the host C++ mirror (host/include/ipic3d/app) ----------------------------------------
// Universes built by the mirror's own Parameters / initCells / initFields, exported in the C ABI's memory
// formats so that tests/test_host_mirror.py can compare them bit for bit with the reference's.
#include <cstring>
#include <sstream>

#include "ipic3d/app/benchmark.h"
#include "ipic3d/app/earth_deck.h"
#include "ipic3d/app/universe.h"

namespace {

struct MirrorUniverse {
	ipic3d::Universe universe;
	std::uint64_t ncycles = 0, numParticles = 0;
	explicit MirrorUniverse(ipic3d::Universe&& u) : universe(std::move(u)) {}
};

// main()'s set-up for "<file>.inp" (EarthDeck, shared with host/src/ipic3d.cc), without running anything
MirrorUniverse* mirrorFromInp(const char* path) {
	const ipic3d::EarthDeck deck{ std::string(path) };
	auto* h = new MirrorUniverse(deck.createUniverse());
	h->ncycles = deck.params.ncycles;
	h->numParticles = (std::uint64_t)deck.numParticles64;
	return h;
}

} // namespace

extern "C" {

IPIC_API void* ipic_mirror_create_from_inp(const char* path, ipic_props* props, uint64_t* ncycles, uint64_t* numParticles) {
	std::ostringstream quiet;
	auto* old = std::cout.rdbuf(quiet.rdbuf());
	MirrorUniverse* h = mirrorFromInp(path);
	std::cout.rdbuf(old);
	if (props) *props = h->universe.properties.toAbi();
	if (ncycles) *ncycles = h->ncycles;
	if (numParticles) *numParticles = h->numParticles;
	return h;
}

// benchmark.h's universe for `:U:N` (kind 'U'), `:C:N` ('C'), `:E:N` ('E'), `:B:N` ('B')
IPIC_API void* ipic_mirror_create_benchmark(char kind, uint64_t numParticles, ipic_props* props) {
	if (std::string("UCEB").find(kind) == std::string::npos) return nullptr;
	std::ostringstream quiet;
	auto* old = std::cout.rdbuf(quiet.rdbuf());
	MirrorUniverse* h = new MirrorUniverse(ipic3d::benchmarkUniverse(kind, numParticles));
	std::cout.rdbuf(old);
	if (props) *props = h->universe.properties.toAbi();
	h->numParticles = numParticles;
	return h;
}

// ---- an input deck piecewise: phase 1 of initCells once, then the particles of any z-slab straight into caller memory ----
// (what a rank of a multi-GPU run needs: its own planes of `ipic3d <deck>`'s universe, never the whole of it)

struct DeckHandle {
	ipic3d::EarthDeck deck;
	std::vector<std::uint64_t> population;       // every cell of the universe, z-fastest flatten
	explicit DeckHandle(const char* path) : deck(std::string(path)) {}
};

IPIC_API void* ipic_host_deck_open(const char* path, ipic_props* props, uint64_t* ncycles, int64_t* numParticles) {
	std::ostringstream quiet;
	auto* old = std::cout.rdbuf(quiet.rdbuf());
	auto* h = new DeckHandle(path);
	h->population = h->deck.cellPopulations();
	std::cout.rdbuf(old);
	if (props) *props = h->deck.properties.toAbi();
	if (ncycles) *ncycles = h->deck.params.ncycles;
	if (numParticles) *numParticles = h->deck.numParticles64;
	return h;
}

IPIC_API void ipic_host_deck_close(void* handle) { delete static_cast<DeckHandle*>(handle); }

// what phase 2 of initCells needs besides the populations: the velocity distribution and the species (for the device generator,
// ipic_generate_normal_speed): mean 0, sigma = the 10 MeV proton's speed per component, q = e, qom = e / m_p (hdr/cell.h:503-504)
IPIC_API void ipic_host_deck_distribution(void* handle, double* mean3, double* stddev3, double* q, double* qom) {
	const DeckHandle* h = static_cast<DeckHandle*>(handle);
	const double v = h->deck.thermalSpeed();
	const double e = 1.602176565e-19, m = 1.672621777e-27;
	for (int a = 0; a < 3; ++a) { mean3[a] = 0.0; stddev3[a] = v; }
	*q = e;
	*qom = e / m;
}

// populations of the z-planes [zBegin, zBegin + zCount) in slab order (x*ny + y)*zCount + zl
IPIC_API int ipic_host_deck_counts(void* handle, int zBegin, int zCount, uint64_t* counts) {
	const DeckHandle* h = static_cast<DeckHandle*>(handle);
	const auto size = h->deck.properties.size;
	if (!counts || zBegin < 0 || zCount < 0 || zBegin + zCount > size.z) return IPIC_ERR_ARG;
	for (std::int64_t xy = 0; xy < size.x * size.y; ++xy) for (int zl = 0; zl < zCount; ++zl)
		*counts++ = h->population[(std::size_t)(xy * size.z + zBegin + zl)];
	return IPIC_OK;
}

// particles of the same cells, AoS of 8 doubles, cell-major in slab order; `threads` <= 0 = all cores
IPIC_API int ipic_host_deck_particles(void* handle, int zBegin, int zCount, double* aos, int threads) {
	const DeckHandle* h = static_cast<DeckHandle*>(handle);
	const auto size = h->deck.properties.size;
	if (!aos || zBegin < 0 || zCount < 0 || zBegin + zCount > size.z) return IPIC_ERR_ARG;
	static_assert(sizeof(ipic3d::Particle) == 64, "Particle is the 8-double record of the C ABI");
	const std::int64_t slabCells = size.x * size.y * zCount;
	std::vector<std::uint64_t> start((std::size_t)slabCells + 1, 0);
	for (std::int64_t c = 0; c < slabCells; ++c) start[(std::size_t)c + 1] = start[(std::size_t)c] + h->population[(std::size_t)((c / zCount) * size.z + zBegin + c % zCount)];
	if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
	if (threads <= 0) threads = 1;
	std::atomic<std::int64_t> next(0);
	auto work = [&]() {
		for (;;) {
			const std::int64_t c = next.fetch_add(1);
			if (c >= slabCells) return;
			const std::int64_t xy = c / zCount;
			const ipic3d::coordinate_type pos{ xy / size.y, xy % size.y, zBegin + c % zCount };
			h->deck.populateCell(pos, start[(std::size_t)c + 1] - start[(std::size_t)c], reinterpret_cast<ipic3d::Particle*>(aos) + start[(std::size_t)c]);
		}
	};
	std::vector<std::thread> pool;
	for (int t = 1; t < threads; ++t) pool.emplace_back(work);
	work();
	for (auto& t : pool) t.join();
	return IPIC_OK;
}

IPIC_API void ipic_mirror_destroy(void* handle) { delete static_cast<MirrorUniverse*>(handle); }

IPIC_API uint64_t ipic_mirror_total(void* handle) { return ipic3d::countParticlesInDomain(static_cast<MirrorUniverse*>(handle)->universe); }

IPIC_API void ipic_mirror_get_counts(void* handle, uint64_t* counts) {
	const auto& cells = static_cast<MirrorUniverse*>(handle)->universe.cells;
	for (std::size_t i = 0; i < cells.count(); ++i) counts[i] = cells.data()[i].particles.size();
}

IPIC_API void ipic_mirror_get_particles(void* handle, double* aos) {
	const auto& cells = static_cast<MirrorUniverse*>(handle)->universe.cells;
	for (std::size_t i = 0; i < cells.count(); ++i) {
		const auto& ps = cells.data()[i].particles;
		if (!ps.empty()) std::memcpy(aos, ps.data(), ps.size() * sizeof(ipic3d::Particle));
		aos += ps.size() * 8;
	}
}

IPIC_API void ipic_mirror_get_fields(void* handle, double* field, double* bc, double* J) {
	const auto& u = static_cast<MirrorUniverse*>(handle)->universe;
	std::memcpy(field, u.field.data(), u.field.count() * sizeof(ipic3d::FieldNode));
	std::memcpy(bc, u.bcfield.data(), u.bcfield.count() * sizeof(ipic3d::BcFieldCell));
	std::memcpy(J, u.currentDensity.data(), u.currentDensity.count() * sizeof(ipic3d::DensityNode));
}

IPIC_API void ipic_mirror_set_fields(void* handle, const double* field, const double* bc, const double* J) {
	auto& u = static_cast<MirrorUniverse*>(handle)->universe;
	if (field) std::memcpy(u.field.data(), field, u.field.count() * sizeof(ipic3d::FieldNode));
	if (bc) std::memcpy(u.bcfield.data(), bc, u.bcfield.count() * sizeof(ipic3d::BcFieldCell));
	if (J) std::memcpy(u.currentDensity.data(), J, u.currentDensity.count() * sizeof(ipic3d::DensityNode));
}

// the host-callable scalar versions of the per-node pieces (cell.h / field.h mirror; never used by simulateSteps), on one
// node: op 0 = projectToDensityField, 1 = solveFieldForward, 2 = solveFieldLeapfrog; op 3 = updateFieldsOnBoundaries
IPIC_API int ipic_mirror_node_op(void* handle, int op, int64_t x, int64_t y, int64_t z) {
	using namespace ipic3d;
	auto& u = static_cast<MirrorUniverse*>(handle)->universe;
	const coordinate_type pos{ x, y, z };
	switch (op) {
	case 0: projectToDensityField(u.properties, u.cells, pos, u.currentDensity); return IPIC_OK;
	case 1: solveFieldForward(u.properties, pos, u.currentDensity, u.field, u.bcfield); return IPIC_OK;
	case 2: solveFieldLeapfrog(u.properties, pos, u.currentDensity, u.field, u.bcfield); return IPIC_OK;
	case 3: updateFieldsOnBoundaries(u.field, u.bcfield); return IPIC_OK;
	}
	return IPIC_ERR_ARG;
}

// value of trilinearInterpolationF2P for 8 corner vectors [i][j][k][component] at a fractional position
IPIC_API void ipic_mirror_trilinear(const double* corners24, const double* pos3, double vol, double* out3) {
	using namespace ipic3d;
	Vector3<double> corners[2][2][2];
	for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) for (int k = 0; k < 2; ++k)
		corners[i][j][k] = { corners24[((i * 2 + j) * 2 + k) * 3], corners24[((i * 2 + j) * 2 + k) * 3 + 1], corners24[((i * 2 + j) * 2 + k) * 3 + 2] };
	const Vector3<double> r = trilinearInterpolationF2P(corners, Vector3<double>{ pos3[0], pos3[1], pos3[2] }, vol);
	out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}

// the particles pic_tracer's generator produces for one task (pic_tracer.cxx:236-262): seed = a*b of the task
IPIC_API void ipic_mirror_tracer_particles(const ipic_props* props, uint32_t seed, uint64_t n, double* aos) {
	using namespace ipic3d;
	const Vector3<double> origin{ props->origin[0], props->origin[1], props->origin[2] };
	const Vector3<double> width{ props->cellWidth[0], props->cellWidth[1], props->cellWidth[2] };
	const coordinate_type size{ props->size[0], props->size[1], props->size[2] };
	const double e = 1.602176565e-19, K = 1e7 * e, m = 1.672621777e-27;
	const double v_mod = props->speedOfLight / std::sqrt(1.0 + (m * props->speedOfLight * props->speedOfLight) / K);
	const auto low = origin + elementwiseProduct(width, size) / 8.0;
	const auto hig = low + 0.75 * elementwiseProduct(width, size);
	distribution::uniform_pos_normal_speed<> next(low, hig, Vector3<double>{ 0.0, 0.0, 0.0 }, Vector3<double>{ v_mod, v_mod, v_mod }, seed);
	for (uint64_t i = 0; i < n; i++) {
		Particle p = next();
		p.q = e;
		p.qom = e / m;
		std::memcpy(aos + i * 8, &p, sizeof p);
	}
}

// host folds of the conserved quantities (cell.h / field.h mirror): { momentum, E energy, B energy, KE }
IPIC_API void ipic_mirror_diagnostics(void* handle, double* out4) {
	using namespace ipic3d;
	const auto& u = static_cast<MirrorUniverse*>(handle)->universe;
	auto getE = [](const Field& field, const coordinate_type& index) { return field[index].E; };
	auto getB = [](const Field& field, const coordinate_type& index) { return field[index].B + field[index].Bext; };
	out4[0] = getTotalParticlesEnergy(u.cells, getParticlesMomentum);
	out4[1] = getFieldEnergy(u.field, u.properties, getE);
	out4[2] = getFieldEnergy(u.field, u.properties, getB);
	out4[3] = getTotalParticlesEnergy(u.cells, getParticlesKineticEnergy);
}

// the parser alone: values of every key the reference reads, as "name=value;" text (vectors comma separated)
IPIC_API int ipic_mirror_parse(const char* path, char* out, int capacity) {
	const ipic3d::Parameters p{ std::string(path) };
	std::ostringstream s;
	s.precision(17);
	auto list = [&s](const char* name, const auto& v) { s << name << "="; for (std::size_t i = 0; i < v.size(); ++i) s << (i ? "," : "") << v[i]; s << ";"; };
	s << "c=" << p.c << ";dt=" << p.dt << ";ncycles=" << p.ncycles << ";L=" << p.L.x << "," << p.L.y << "," << p.L.z
	  << ";center=" << p.objectCenter.x << "," << p.objectCenter.y << "," << p.objectCenter.z << ";planetRadius=" << p.planetRadius
	  << ";delta=" << p.delta << ";ncells=" << p.ncells.x << "," << p.ncells.y << "," << p.ncells.z
	  << ";dspace=" << p.dspace.x << "," << p.dspace.y << "," << p.dspace.z << ";ns=" << p.ns
	  << ";B0=" << p.B0.x << "," << p.B0.y << "," << p.B0.z << ";B1=" << p.B1.x << "," << p.B1.y << "," << p.B1.z
	  << ";useCase=" << (p.useCase == ipic3d::UseCase::Dipole ? "Dipole" : "Test") << ";wmethod=" << p.wmethod << ";SimName=" << p.SimName
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

}
