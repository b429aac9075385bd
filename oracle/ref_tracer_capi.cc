// TEST INFRASTRUCTURE ONLY (oracle/): never linked into, loaded by or called from the product.
//
// C handle over the UNMODIFIED reference tracer (/root/reference/code/app/src/pic_tracer.cxx), compiled against the
// AllScale header shim: the reference's own `traceParticle` and `ParticleCount`, plus its particle generator
// exactly as `main` sets it up (pic_tracer.cxx:206-262).  The reference's main() is renamed, not run.
#include <cstring>

#define main ref_pic_tracer_main
#include "pic_tracer.cxx"
#undef main

#include "../include/ipic3d_b200.h"

namespace {
	UniverseProperties toConfig(const ipic_props& p) {
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
		r.useCase = UseCase::Dipole;
		return r;
	}
}

extern "C" {

// the tracer's universe for resolution R (pic_tracer.cxx:206-222)
void ref_tracer_config(int R, ipic_props* out) {
	UniverseProperties config;
	config.dt = 0.01;
	config.speedOfLight = 299792458;
	config.size = { R, R, R };
	config.planetRadius = 6378137;
	config.cellWidth = (20.0 / config.size.x) * config.planetRadius;
	config.objectCenter = { 0.0, 0.0, 0.0 };
	config.origin.x = config.objectCenter.x - config.size.x * config.cellWidth.x / 2.0;
	config.origin.y = config.objectCenter.y - config.size.y * config.cellWidth.y / 2.0;
	config.origin.z = config.objectCenter.z - config.size.z * config.cellWidth.z / 2.0;
	config.externalMagneticField = { 0.0, 0.0, 3.07e-5 };
	for (int i = 0; i < 3; i++) {
		out->size[i] = config.size[i]; out->cellWidth[i] = config.cellWidth[i]; out->objectCenter[i] = config.objectCenter[i];
		out->origin[i] = config.origin[i]; out->externalMagneticField[i] = config.externalMagneticField[i];
	}
	out->dt = config.dt; out->speedOfLight = config.speedOfLight; out->planetRadius = config.planetRadius;
}

// the n particles one `map(a, b)` task generates (pic_tracer.cxx:236-262): generator seeded with a*b, q = e, qom = e/m
void ref_tracer_particles(const ipic_props* props, uint32_t seed, uint64_t n, double* aos) {
	const UniverseProperties config = toConfig(*props);
	double e = 1.602176565e-19;
	double K = 1e7 * e;
	double m = 1.672621777e-27;
	double v_mod = config.speedOfLight / sqrt(1.0 + (m * config.speedOfLight * config.speedOfLight) / K);
	auto low = config.origin + elementwiseProduct(config.cellWidth, config.size) / 8.0;
	auto hig = low + 0.75 * elementwiseProduct(config.cellWidth, config.size);
	distribution::uniform_pos_normal_speed<> next(low, hig, Vector3<double>{ 0.0, 0.0, 0.0 }, Vector3<double>{ v_mod, v_mod, v_mod }, seed);
	for (uint64_t i = 0; i < n; i++) {
		Particle p = next();
		p.q = e;
		p.qom = e / m;
		std::memcpy(aos + i * 8, &p, 64);
	}
}

// traceParticle over n particles; counts[frames][nx][ny][nz] is accumulated into
void ref_trace(const ipic_props* props, const double* aos, uint64_t n, int T, int frame_interval, int num_frames, uint32_t* counts) {
	const UniverseProperties config = toConfig(*props);
	ParticleCount res(num_frames, config.size);
	for (uint64_t i = 0; i < n; i++) {
		Particle p;
		std::memcpy(&p, aos + i * 8, 64);
		traceParticle(p, T, config, frame_interval, res);
	}
	for (int t = 0; t < num_frames; t++) for (int x = 0; x < config.size.x; x++) for (int y = 0; y < config.size.y; y++) for (int z = 0; z < config.size.z; z++)
		counts[((t * config.size.x + x) * config.size.y + y) * config.size.z + z] += res.get(t, x, y, z);
}

} // extern "C"
