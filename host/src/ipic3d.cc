// `ipic3d <file.inp>` / `ipic3d :U:N` — the reference's driver (code/app/src/ipic3d.cxx:13-114) on top of the
// host mirror: same command line, same set-up incl. the hard-coded Earth parameters, same
// console lines (scripts grep "Throughput:"), same "<base>.out"; the time loop runs on the GPU(s).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>

#include "ipic3d/app/benchmark.h"
#include "ipic3d/app/cell.h"
#include "ipic3d/app/field.h"
#include "ipic3d/app/parameters.h"
#include "ipic3d/app/simulator.h"
#include "ipic3d/app/universe.h"

using namespace ipic3d;

int main(int argc, char** argv) {
	if (argc != 2 || std::strcmp(argv[1], "-h") == 0 || std::strcmp(argv[1], "--help") == 0) {
		std::cout << "Usage: ./ipic3d <config-file>" << std::endl;
		return EXIT_FAILURE;
	}
	const std::string inputFilename = argv[1];
	if (inputFilename[0] == ':') {
		const int rc = processBenchmark(inputFilename);
		releaseDevices();
		return rc;
	}

	std::cout << "Loading configuration file \"" << inputFilename << "\" ..." << std::endl;
	const Parameters params(inputFilename);

	std::cout << "Initializing simulation state ..." << std::endl;
	const auto sep = inputFilename.find_last_of("/\\");
	const std::string baseName = inputFilename.substr(sep + 1, inputFilename.find_last_of('.') - sep - 1);

	const InitProperties initProperties(params);
	std::cout << initProperties;

	// Earth magnetosphere set-up: these values override the input deck
	UniverseProperties universeProperties(params);
	universeProperties.outputFileBaseName = baseName;
	universeProperties.dt = 0.01;
	universeProperties.speedOfLight = 299792458;
	universeProperties.planetRadius = 6378137;                      // m
	universeProperties.cellWidth *= universeProperties.planetRadius;
	const auto universeSize = elementwiseProduct(universeProperties.cellWidth, universeProperties.size);
	universeProperties.objectCenter = { 0.0, 0.0, 0.0 };
	universeProperties.origin = universeProperties.objectCenter - universeSize / 2.0;
	universeProperties.externalMagneticField = { 0.0, 0.0, 3.07e-5 };   // T
	universeProperties.useCase = UseCase::Dipole;
	std::cout << universeProperties;

	int numParticles = (int)(universeProperties.size.x * universeProperties.size.y * universeProperties.size.z);
	numParticles = numParticles * initProperties.particlesPerCell[0].x * initProperties.particlesPerCell[0].y * initProperties.particlesPerCell[0].z;
	// 10 MeV protons: thermal speed from the relativistic kinetic energy
	const double e = 1.602176565e-19, K = 1e7 * e, m = 1.672621777e-27;
	const double v_mod = universeProperties.speedOfLight / std::sqrt(1.0 + (m * universeProperties.speedOfLight * universeProperties.speedOfLight) / K);
	const auto low = universeProperties.origin + 0.125 * universeSize;
	const auto hig = low + 0.75 * universeSize;
	const auto dist = distribution::uniform_pos_normal_speed<>(low, hig, Vector3<double>{ 0, 0, 0 }, Vector3<double>{ v_mod, v_mod, v_mod });
	auto universe = createUniverseFromDistribution(universeProperties, initProperties, numParticles, dist);

	std::cout << "Running simulation..." << std::endl;
	const auto duration = simulateSteps(params.ncycles, universe);

	std::cout << "Simulation measurements: " << numParticles;
	std::cout << " initial particles, first step " << duration.firstStep << " seconds, " << (numParticles / duration.firstStep);
	std::cout << " pps, remaining steps " << duration.remainingSteps << " seconds, " << (numParticles * (params.ncycles - 1)) / duration.remainingSteps << " pps\n";
	std::cout << "Throughput: " << (params.ncycles * double(numParticles)) / duration.remainingSteps << " particles/s \n";

	std::cout << "Simulation finished successfully, producing output data..." << std::endl;
	outputNumberOfParticlesPerCell(universe.cells, baseName + ".out");
	releaseDevices();
	return EXIT_SUCCESS;
}
