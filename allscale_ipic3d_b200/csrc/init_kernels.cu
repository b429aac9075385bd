// The `:U:N` benchmark universe generated on the device, bit-identical to the reference's
// initCells(properties, N, distribution::uniform<>) (hdr/cell.h:638-697, hdr/benchmark.h:81-92):
//
//   population of cell (x,y,z)   N / cells, +1 for the first N % cells cells of the z-fastest flatten
//   seed                         uint32((x*1023 + y)*1023 + z); positions from std::minstd_rand(seed + 1), velocities
//                                from std::minstd_rand(seed + 2); per particle x, y, z in that order
//   a double                     std::uniform_real_distribution<double>(a, b): generate_canonical<double, 53> takes
//                                TWO engine values u1, u2 (libstdc++ bits/random.tcc:3349-3381, range R = 2147483646):
//                                ret = ((u1 - 1) + (u2 - 1) * R) / (R * R), ret >= 1 -> nextafter(1, 0); a + ret * (b - a)
//   species                      electron: q = -1, qom = -25 (hdr/cell.h:107-118)
//
// minstd_rand is the Lehmer generator x <- 48271 x mod (2^31 - 1): particle i of a cell needs draws 6i+1 .. 6i+6, so
// every lane jumps to its particle with powers of the multiplier; one warp per cell, coalesced AoSoA stores.
// Compiled with -fmad=false: the library's floating-point steps are kept unfused and in order.
#include "init_kernels.h"
#include "glibc_logf.h"

namespace ipic {

constexpr unsigned long long kLcgA = 48271ull, kLcgM = 2147483647ull;
__device__ __forceinline__ unsigned long long lcg_mul(unsigned long long x, unsigned long long y) { return x * y % kLcgM; }

__device__ __forceinline__ double canonical(unsigned long long u1, unsigned long long u2) {
	const double R = 2147483646.0;                       // (long double)max - (long double)min + 1
	double sum = 0.0;
	double tmp = 1.0;
	sum += (double)(u1 - 1ull) * tmp;
	tmp *= R;
	sum += (double)(u2 - 1ull) * tmp;
	tmp *= R;                                            // 2147483646^2 < 2^62: the long double product is exact, one rounding to double
	double ret = sum / tmp;
	if (ret >= 1.0) ret = 0.99999999999999988897769753748434595763683319091796875;    // nextafter(1.0, 0.0)
	return ret;
}

__global__ void __launch_bounds__(128)
k_init_uniform(const Geometry g, const ParticleArrays p, const int64_t* __restrict__ cellStart, const int* __restrict__ count) {
	const int lane = threadIdx.x & 31;
	const int cell = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	if (cell >= g.ncellsLocal) return;
	const int xy = cell / g.lz, zl = cell - xy * g.lz;
	const long long cx = xy / g.ny, cy = xy - cx * g.ny, cz = g.z0 + zl;
	const int n = count[cell];
	const int64_t base = cellStart[cell];

	const unsigned seed = (unsigned)(((cx * 1023) + cy) * 1023 + cz);
	unsigned long long sp = (unsigned long long)(unsigned)(seed + 1u) % kLcgM, sv = (unsigned long long)(unsigned)(seed + 2u) % kLcgM;
	if (sp == 0) sp = 1;                                 // linear_congruential_engine::seed with c == 0
	if (sv == 0) sv = 1;
	// a^6 (one particle), then this lane's offset a^(6 lane) and the warp's stride a^(6*32)
	unsigned long long a6 = 1;
	#pragma unroll
	for (int k = 0; k < 6; k++) a6 = lcg_mul(a6, kLcgA);
	unsigned long long jump = 1, stride = 1;
	for (int k = 0; k < 32; k++) { if (k < lane) jump = lcg_mul(jump, a6); stride = lcg_mul(stride, a6); }
	sp = lcg_mul(sp, jump); sv = lcg_mul(sv, jump);

	// cell box: getOriginOfCell = origin + idx * width (universe_properties.h:125-131), upper corner = lower + width
	const double lo[3] = { g.w[0] * (double)cx + g.origin[0], g.w[1] * (double)cy + g.origin[1], g.w[2] * (double)cz + g.origin[2] };
	const double hi[3] = { lo[0] + g.w[0], lo[1] + g.w[1], lo[2] + g.w[2] };

	for (int i = lane; i < n; i += 32) {
		double v[6];
		unsigned long long t = sp;
		#pragma unroll
		for (int a = 0; a < 3; a++) {
			const unsigned long long u1 = t = lcg_mul(t, kLcgA);
			const unsigned long long u2 = t = lcg_mul(t, kLcgA);
			v[a] = (canonical(u1, u2) * (hi[a] - lo[a])) + lo[a];
		}
		t = sv;
		#pragma unroll
		for (int a = 0; a < 3; a++) {
			const unsigned long long u1 = t = lcg_mul(t, kLcgA);
			const unsigned long long u2 = t = lcg_mul(t, kLcgA);
			v[3 + a] = (canonical(u1, u2) * (0.2 - (-0.2))) + (-0.2);
		}
		double* q = p.d + slot_index(base + i, g.nf);
		#pragma unroll
		for (int f = 0; f < 6; f++) q[f * kFieldStride] = v[f];
		if (g.nf == kFields) { q[6 * kFieldStride] = -1.0; q[7 * kFieldStride] = -25.0; }
		sp = lcg_mul(sp, stride); sv = lcg_mul(sv, stride);
	}
}

void launch_init_uniform(const Geometry& g, const ParticleArrays& p, const int64_t* cellStart, const int* count, cudaStream_t st) {
	if (g.ncellsLocal <= 0) return;
	k_init_uniform<<<(g.ncellsLocal + 3) / 4, 128, 0, st>>>(g, p, cellStart, count);
}

// ---- the universes of the `.inp` decks: phase 2 of initCells<Distribution> (hdr/cell.h:596-632) --------------------------
//
// For every cell: seed = uint32((x*1023 + y)*1023 + z);
//   positions   distribution::vector::uniform(cellLow, cellHigh, seed): ONE std::minstd_rand(seed) feeding x, y, z
//               (hdr/cell.h:138-174) -> the same jump-ahead as above, one warp per cell (k_init_deck_positions);
//   velocities  the distribution re-seeded with the cell's seed (generic_particle_generator::seed, hdr/cell.h:281-287:
//               s = (1023 s) & s twice, 32-bit) -> distribution::vector::normal: mean + stddev * zig() per component with the
//               in-tree ziggurat (hdr/ziggurat_normal_distribution.h): a 64-bit xorshift whose number of draws per sample
//               depends on the sample (wedge / tail rejections), so a cell's velocities are inherently sequential:
//               one THREAD per cell (k_init_normal_speed), 262 144 cells in medium.inp keep the GPU busy;
//   species     q = e, qom = e / m_p for every particle (hdr/cell.h:624-625).
// The ziggurat's tables come from the host (libm exp / log / sqrt in double, as the reference computes them); its fast path is
// integer and float arithmetic; the tail takes logf of a float, restated bit for bit in glibc_logf.h; the wedge test
// compares a float with a double exp() and tolerates the device library's last-bit difference (an outcome could only
// flip if the float fell between the two values of exp: ~1e-14 per test).

struct ZigguratState {
	unsigned long long jsr;
	__device__ __forceinline__ unsigned long long shr3() {
		const unsigned long long before = jsr;
		jsr ^= jsr << 13;
		jsr ^= jsr >> 17;
		jsr ^= jsr << 5;
		return before + jsr;
	}
	__device__ __forceinline__ float uni() { return (float)(.5 + (double)(int)shr3() * .2328306e-9); }
};

// one sample (hdr/ziggurat_normal_distribution.h:50-103, nfix); t = the tables in shared memory
__device__ __forceinline__ float ziggurat_sample(ZigguratState& z, const ZigguratTables& t) {
	const float r = 3.442620f;
	int hz = (int)z.shr3();
	unsigned iz = (unsigned)hz & 127u;
	if (fabs((double)hz) < (double)t.kn[iz]) return (float)hz * t.wn[iz];
	for (;;) {
		if (iz == 0) {
			float x, y;
			do {
				x = (float)(-0.2904764 * (double)glibc_logf(z.uni()));
				y = -glibc_logf(z.uni());
			} while (!(x * x <= y + y));
			return hz <= 0 ? -r - x : +r + x;
		}
		const float x = (float)hz * t.wn[iz];
		const float lhs = t.fn[iz] + z.uni() * (t.fn[iz - 1] - t.fn[iz]);
		if ((double)lhs < exp(-0.5 * (double)x * (double)x)) return x;
		hz = (int)z.shr3();
		iz = (unsigned)hz & 127u;
		if (fabs((double)hz) < (double)t.kn[iz]) return (float)hz * t.wn[iz];
	}
}

__global__ void __launch_bounds__(128)
k_init_deck_positions(const Geometry g, const ParticleArrays p, const int64_t* __restrict__ cellStart, const int* __restrict__ count) {
	const int lane = threadIdx.x & 31;
	const int cell = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	if (cell >= g.ncellsLocal) return;
	const int xy = cell / g.lz, zl = cell - xy * g.lz;
	const long long cx = xy / g.ny, cy = xy - cx * g.ny, cz = g.z0 + zl;
	const int n = count[cell];
	const int64_t base = cellStart[cell];
	const unsigned seed = (unsigned)(((cx * 1023) + cy) * 1023 + cz);
	unsigned long long sp = (unsigned long long)seed % kLcgM;
	if (sp == 0) sp = 1;
	unsigned long long a6 = 1;
	#pragma unroll
	for (int k = 0; k < 6; k++) a6 = lcg_mul(a6, kLcgA);
	unsigned long long jump = 1, stride = 1;
	for (int k = 0; k < 32; k++) { if (k < lane) jump = lcg_mul(jump, a6); stride = lcg_mul(stride, a6); }
	sp = lcg_mul(sp, jump);
	// low = width * idx, then += origin; high = low + width (hdr/cell.h:603-605)
	const double lo[3] = { g.w[0] * (double)cx + g.origin[0], g.w[1] * (double)cy + g.origin[1], g.w[2] * (double)cz + g.origin[2] };
	const double hi[3] = { lo[0] + g.w[0], lo[1] + g.w[1], lo[2] + g.w[2] };
	for (int i = lane; i < n; i += 32) {
		unsigned long long t = sp;
		double* q = p.d + slot_index(base + i, g.nf);
		#pragma unroll
		for (int a = 0; a < 3; a++) {
			const unsigned long long u1 = t = lcg_mul(t, kLcgA);
			const unsigned long long u2 = t = lcg_mul(t, kLcgA);
			q[a * kFieldStride] = (canonical(u1, u2) * (hi[a] - lo[a])) + lo[a];
		}
		if (g.nf == kFields) { q[6 * kFieldStride] = g.uq; q[7 * kFieldStride] = g.uqom; }
		sp = lcg_mul(sp, stride);
	}
}

__global__ void __launch_bounds__(128)
k_init_normal_speed(const Geometry g, const ParticleArrays p, const int64_t* __restrict__ cellStart, const int* __restrict__ count,
                    const ZigguratTables* __restrict__ tables, double m0, double m1, double m2, double s0, double s1, double s2) {
	__shared__ ZigguratTables t;
	{
		const unsigned* src = reinterpret_cast<const unsigned*>(tables);
		unsigned* dst = reinterpret_cast<unsigned*>(&t);
		for (int i = threadIdx.x; i < (int)(sizeof(ZigguratTables) / 4); i += blockDim.x) dst[i] = src[i];
	}
	__syncthreads();
	const int cell = blockIdx.x * blockDim.x + threadIdx.x;
	if (cell >= g.ncellsLocal) return;
	const int xy = cell / g.lz, zl = cell - xy * g.lz;
	const long long cx = xy / g.ny, cy = xy - cx * g.ny, cz = g.z0 + zl;
	const int n = count[cell];
	const int64_t base = cellStart[cell];
	unsigned seed = (unsigned)(((cx * 1023) + cy) * 1023 + cz);
	seed = (1023u * seed) & seed;                        // position generator of the distribution (its draws are overwritten)
	seed = (1023u * seed) & seed;                        // velocity generator
	ZigguratState z{ 123456789ull ^ (unsigned long long)seed };
	const double mean[3] = { m0, m1, m2 }, stddev[3] = { s0, s1, s2 };
	for (int i = 0; i < n; i++) {
		double* q = p.d + slot_index(base + i, g.nf);
		#pragma unroll
		for (int a = 0; a < 3; a++) q[(3 + a) * kFieldStride] = mean[a] + stddev[a] * (double)ziggurat_sample(z, t);
	}
}

void launch_init_deck(const Geometry& g, const ParticleArrays& p, const int64_t* cellStart, const int* count, const ZigguratTables* tables,
                      const double mean[3], const double stddev[3], cudaStream_t st) {
	if (g.ncellsLocal <= 0) return;
	k_init_deck_positions<<<(g.ncellsLocal + 3) / 4, 128, 0, st>>>(g, p, cellStart, count);
	k_init_normal_speed<<<(g.ncellsLocal + 127) / 128, 128, 0, st>>>(g, p, cellStart, count, tables, mean[0], mean[1], mean[2], stddev[0], stddev[1], stddev[2]);
}

} // namespace ipic
