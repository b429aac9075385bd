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
