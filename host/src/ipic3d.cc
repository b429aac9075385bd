// `ipic3d <file.inp>` / `ipic3d :U:N` — the reference's driver (code/app/src/ipic3d.cxx:13-114) on top of the host
// mirror: same command line, same set-up (EarthDeck: the hard-coded Earth parameters of SURVEY.md R10), same console
// lines (scripts grep "Throughput:"), same "<base>.out"; the time loop runs on the GPU(s).
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>

#include "ipic3d/app/benchmark.h"
#include "ipic3d/app/earth_deck.h"
#include "ipic3d/app/simulator.h"

using namespace ipic3d;

namespace {

	// "dir/name.inp" -> "name"
	std::string baseNameOf(const std::string& path) {
		const auto sep = path.find_last_of("/\\");
		return path.substr(sep + 1, path.find_last_of('.') - sep - 1);
	}

	int runDeck(const std::string& path) {
		std::cout << "Loading configuration file \"" << path << "\" ..." << std::endl;
		const std::string baseName = baseNameOf(path);
		std::cout << "Initializing simulation state ..." << std::endl;
		const EarthDeck deck(path, baseName);
		std::cout << deck.init << deck.properties;
		// The run only reports per-cell counts afterwards, so by default the particles live in HBM from start to end: created
		// there (phase 1 of initCells on the host, phase 2 on the device, bit-identical), stepped there, counted there.
		// IPIC3D_RESIDENT=0: the reference's flow -- host universe, upload, steps, download of every particle.
		const char* residentEnv = std::getenv("IPIC3D_RESIDENT");
		const bool resident = !(residentEnv && std::atoi(residentEnv) == 0);
		if (!deck.countFitsReference()) {
			// large.inp: the reference's `int numParticles` overflows (ipic3d.cxx:70-71) and its run is meaningless; here the deck's
			// exact count is used, which only the device-resident flow can hold (64 B per particle on the host would be 453 GB)
			std::cout << "Note: " << deck.numParticles64 << " particles exceed the reference's 32-bit count (which reads " << deck.numParticles
			          << "); using the exact 64-bit count" << std::endl;
			if (!resident) { std::cerr << "this deck needs the device-resident flow (unset IPIC3D_RESIDENT=0)" << std::endl; return EXIT_FAILURE; }
		}
		Universe universe = resident ? deck.createUniverseWithoutParticles() : deck.createUniverse();
		struct OptionalScope { ResidentScope* s; ~OptionalScope() { delete s; } } scope{ resident ? new ResidentScope(universe) : nullptr };
		if (resident) {
			scope.s->dismiss();
			const double v = deck.thermalSpeed();
			const double e = 1.602176565e-19, m = 1.672621777e-27;
			std::cout << "Populating cells ...\n";
			populateOnDevice(universe, deck.cellPopulations(), Vector3<double>{ 0, 0, 0 }, Vector3<double>{ v, v, v }, e, e / m);
		}

		std::cout << "Running simulation..." << std::endl;
		const std::uint64_t cycles = deck.params.ncycles;
		const double n = (double)deck.numParticles64;
		const DurationMeasurement t = simulateSteps(cycles, universe);

		std::cout << "Simulation measurements: " << deck.numParticles64 << " initial particles, first step " << t.firstStep << " seconds, " << (n / t.firstStep)
		          << " pps, remaining steps " << t.remainingSteps << " seconds, " << (n * (cycles - 1)) / t.remainingSteps << " pps\n"
		          << "Throughput: " << (cycles * double(n)) / t.remainingSteps << " particles/s \n";

		std::cout << "Simulation finished successfully, producing output data..." << std::endl;
		outputNumberOfParticlesPerCell(universe, baseName + ".out");       // from the device's per-cell counters when the universe lives there
		return EXIT_SUCCESS;
	}
}

int main(int argc, char** argv) {
	const bool wantsHelp = argc == 2 && (std::strcmp(argv[1], "-h") == 0 || std::strcmp(argv[1], "--help") == 0);
	if (argc != 2 || wantsHelp) {
		std::cout << "Usage: ./ipic3d <config-file>" << std::endl;
		return EXIT_FAILURE;
	}
	const std::string what = argv[1];
	const int rc = what[0] == ':' ? processBenchmark(what) : runDeck(what);
	releaseDevices();
	return rc;
}
