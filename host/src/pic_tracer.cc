// `pic_tracer [N [T [S [R [M]]]]]` -- the reference's density tracer (code/app/src/pic_tracer.cxx:151-316) on top of
// the host mirror and the C ABI: same command line, same console protocol ("Throughput: ... particles/s"), same
// `result_<stamp>.csv.NNNNNN` frames.  The particles are generated on the host by the mirror of the reference's
// generator (bit-identical: libstdc++ minstd_rand + the in-tree ziggurat); `traceParticle` for all of them runs in
// one CUDA kernel (ipic_trace, csrc/tracer_kernels.cu).
//
// The reference hands `map(a, b)` sub-ranges of [0, N/B+1) to its runtime and seeds each task's generator with a*b,
// so WHICH particles it traces depends on how the runtime splits the range.  This driver follows the sequential
// schedule (one task covering the whole range: seed 0, particles 0..N-1), which is what the reference does when
// built against a sequential preduce.
#include <algorithm>
#include <chrono>
#include <cinttypes>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <vector>

#include "ipic3d/app/cell.h"
#include "ipic3d/app/particle.h"
#include "ipic3d/app/universe_properties.h"
#include "ipic3d_b200.h"

using namespace ipic3d;

namespace {

	// per-frame particle density (reference pic_tracer.cxx:26-70)
	class ParticleCount {
		coordinate_type size;
		std::vector<std::uint32_t> counts;
	public:
		ParticleCount(int num_frames, const coordinate_type& size) : size(size), counts((std::size_t)(num_frames * size.x * size.y * size.z), 0) {}
		std::uint32_t get(int t, int x, int y, int z) const { return counts[flatten(t, x, y, z)]; }
		std::uint32_t* data() { return counts.data(); }
		ParticleCount& operator+=(const ParticleCount& other) {
			for (std::size_t i = 0; i < counts.size(); i++) counts[i] += other.counts[i];
			return *this;
		}
	private:
		std::size_t flatten(int t, int x, int y, int z) const { return (std::size_t)(((t * size.x + x) * size.y + y) * size.z + z); }
	};

	void check(int rc, const char* what) {
		if (rc == IPIC_OK) return;
		std::cerr << "pic_tracer_b200: " << what << " failed (" << rc << "): " << ipic_last_error() << std::endl;
		std::exit(EXIT_FAILURE);
	}
}

int main(int argc, char** argv) {
	int N = 16 * 1000 * 1000;      // number of particles
	int T = 150;                   // number of time steps
	int S = 10;                    // number of time steps between frames
	int R = 64;                    // resolution of the result grid
	int M = 1;                     // number of repetitions
	if (argc > 1) N = std::atoi(argv[1]);
	if (argc > 2) T = std::atoi(argv[2]);
	if (argc > 3) S = std::atoi(argv[3]);
	if (argc > 4) R = std::atoi(argv[4]);
	if (argc > 5) M = std::atoi(argv[5]);
	const int num_frames = T / S + 1;

	std::cout << "----- particle-in-cell tracer -----\n";
	std::cout << "Tracing " << N << " particles for " << T << " time steps in a " << R << "^3 grid recording a snapshot every " << S << " time steps";
	if (M == 1) std::cout << " ...\n";
	else std::cout << M << " time(s) ...\n";

	UniverseProperties config;
	config.dt = 0.01;
	config.speedOfLight = 299792458;
	config.size = { R, R, R };
	config.planetRadius = 6378137;                                 // m (Earth radius)
	config.cellWidth = (20.0 / config.size.x) * config.planetRadius;
	config.FieldOutputCycle = 0;
	config.objectCenter = { 0.0, 0.0, 0.0 };
	config.origin.x = config.objectCenter.x - config.size.x * config.cellWidth.x / 2.0;
	config.origin.y = config.objectCenter.y - config.size.y * config.cellWidth.y / 2.0;
	config.origin.z = config.objectCenter.z - config.size.z * config.cellWidth.z / 2.0;
	config.externalMagneticField = { 0.0, 0.0, 3.07e-5 };
	config.useCase = UseCase::Dipole;
	const ipic_props abi = config.toAbi();

	const double e = 1.602176565e-19;                              // elementary charge (C)
	const double K = 1e7 * e;                                      // kinetic energy (J)
	const double m = 1.672621777e-27;                              // proton mass (kg)
	const double v_mod = config.speedOfLight / std::sqrt(1.0 + (m * config.speedOfLight * config.speedOfLight) / K);

	int device = 0;
	if (const char* d = std::getenv("IPIC3D_DEVICES")) device = std::atoi(d);
	void* pinned = nullptr;
	check(ipic_alloc_pinned((std::size_t)std::max(N, 1) * sizeof(Particle), &pinned), "ipic_alloc_pinned");
	Particle* particles = static_cast<Particle*>(pinned);

	ParticleCount res(num_frames, config.size);
	for (int i = 0; i < M; i++) {
		if (M > 1) std::cout << "Processing batch " << (i + 1) << "/" << M << " ..\n";
		const auto start = std::chrono::high_resolution_clock::now();

		// one task over the whole range: generator seeded with a*b = 0 (see the header comment)
		const auto low = config.origin + elementwiseProduct(config.cellWidth, config.size) / 8.0;
		const auto hig = low + 0.75 * elementwiseProduct(config.cellWidth, config.size);
		distribution::uniform_pos_normal_speed<> next(low, hig, Vector3<double>{ 0.0, 0.0, 0.0 }, Vector3<double>{ v_mod, v_mod, v_mod }, 0);
		for (int k = 0; k < N; k++) {
			Particle p = next();
			p.q = e;
			p.qom = e / m;
			particles[k] = p;
		}
		ParticleCount batch(num_frames, config.size);
		double deviceSeconds = 0;
		check(ipic_trace(&abi, device, reinterpret_cast<const double*>(particles), (std::uint64_t)std::max(N, 0), T + 1, S, num_frames, batch.data(), &deviceSeconds), "ipic_trace");
		res += batch;

		const auto end = std::chrono::high_resolution_clock::now();
		std::cout << "Simulation Finished" << std::endl;
		const double s = std::chrono::duration_cast<std::chrono::nanoseconds>(end - start).count() / (1e9);
		std::cout << "Simulation took " << s << "s\n";
		std::cout << "Throughput: " << ((T + 1) * double(N)) / s << " particles/s \n";
		std::cout << "Device tracing (upload + kernel): " << deviceSeconds << "s, " << ((T + 1) * double(N)) / deviceSeconds << " particles/s \n";
	}
	ipic_free_pinned(pinned);

	{
		const std::int64_t timeStamp = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
		for (int t = 0; t < num_frames; t++) {
			char fileName[64];
			std::snprintf(fileName, sizeof fileName, "result_%" PRId64 ".csv.%06d", timeStamp, t);
			std::ofstream out(fileName);
			out << "t,x,y,z,density\n";
			for (int x = 0; x < config.size.x; x++) for (int y = 0; y < config.size.y; y++) for (int z = 0; z < config.size.z; z++)
				out << t << "," << x * config.cellWidth.x << "," << y * config.cellWidth.y << "," << z * config.cellWidth.z << "," << res.get(t, x, y, z) << "\n";
		}
	}
	return EXIT_SUCCESS;
}
