// Conserved-quantity diagnostics on the device (reference writeOutputData, hdr/simulator.h:71-91):
//   total particle "momentum"   sum (q/qom) * sqrt(|v|^2)        getParticlesMomentum       hdr/cell.h:1330-1340
//   total kinetic energy        sum 0.5 * (q/qom) * |v|^2        getParticlesKineticEnergy  hdr/cell.h:1315-1325
//   E and B field energy        sum vol * |F|^2 / (4 pi) over the nodes [1, size-1) of the field grid, vol = 0.5 * cell
//                               volume, F = E or B + Bext                                   getFieldEnergy  hdr/field.h:443-459
// so that a run with FieldOutputCycle > 0 does not have to bring 64 B per particle back to the host for four
// numbers.  The reference folds sequentially (cells in lexicographic order, particles in cell order); here every
// level is a fixed-shape tree (thread-strided partials, shared-memory tree, fixed-order final pass), so the
// result is reproducible run to run and independent of the grid of CTAs, and agrees with the sequential fold
// to rounding (tests: <= 1e-12 relative).
#include "diag_kernels.h"

namespace ipic {

constexpr int kDiagBlock = 128;

// deterministic block sum of two values per thread; result valid in thread 0
__device__ __forceinline__ void block_sum2(double& a, double& b, double* sh /* 2 * kDiagBlock */) {
	const int tid = threadIdx.x;
	sh[tid] = a; sh[kDiagBlock + tid] = b;
	__syncthreads();
	#pragma unroll
	for (int s = kDiagBlock / 2; s > 0; s >>= 1) {
		if (tid < s) { sh[tid] += sh[tid + s]; sh[kDiagBlock + tid] += sh[kDiagBlock + tid + s]; }
		__syncthreads();
	}
	a = sh[0]; b = sh[kDiagBlock];
}

// one CTA per owned cell: partial[cell] = { kinetic energy, momentum } of the cell's particles
__global__ void __launch_bounds__(kDiagBlock)
k_cell_energy(const Geometry g, const ParticleArrays p, const int64_t* __restrict__ cellStart, const int* __restrict__ count,
              double* __restrict__ partial) {
	__shared__ double sh[2 * kDiagBlock];
	const int cell = blockIdx.x;
	const int n = count[cell];
	const int64_t base = cellStart[cell];
	double ke = 0.0, mom = 0.0;
	for (int i = threadIdx.x; i < n; i += kDiagBlock) {
		const double* q = p.d + slot_index(base + i, g.nf);
		const double vx = q[3 * kFieldStride], vy = q[4 * kFieldStride], vz = q[5 * kFieldStride];
		const double charge = g.nf == kFields ? q[6 * kFieldStride] : g.uq;
		const double qom = g.nf == kFields ? q[7 * kFieldStride] : g.uqom;
		double v2 = 0.0; v2 += vx * vx; v2 += vy * vy; v2 += vz * vz;        // sumOfSquares: index order
		const double mass = charge / qom;
		ke += 0.5 * mass * v2;                                                  // cell.h:1317
		mom += mass * sqrt(v2);                                                 // cell.h:1332
	}
	block_sum2(ke, mom, sh);
	if (threadIdx.x == 0) { partial[2 * cell] = ke; partial[2 * cell + 1] = mom; }
}

// CTA b: partial[b] = { E energy, B energy } of its share of the interior nodes (node -> CTA assignment is static)
__global__ void __launch_bounds__(kDiagBlock)
k_field_energy(const Geometry g, const double* __restrict__ EB, const double* __restrict__ Bext, double* __restrict__ partial, int chunks) {
	__shared__ double sh[2 * kDiagBlock];
	const int64_t ex = g.nx + 1, ey = g.ny + 1, ez = g.nz + 1;               // interior nodes per axis: indices 1 .. n+1
	const int64_t nodes = ex * ey * ez;
	const int64_t per = (nodes + chunks - 1) / chunks;
	const int64_t lo = blockIdx.x * per, hi = min(nodes, lo + per);
	const double vol = 0.5 * g.w[0] * g.w[1] * g.w[2];                         // field.h:448
	const double fourPI = 16.0 * atan(1.0);
	double eE = 0.0, eB = 0.0;
	for (int64_t i = lo + threadIdx.x; i < hi; i += kDiagBlock) {
		const int64_t z = i % ez, y = (i / ez) % ey, x = i / (ez * ey);
		const int64_t node = ((x + 1) * (g.ny + 3) + (y + 1)) * (g.nz + 3) + (z + 1);
		const double* f = EB + node * 6;
		const double* bx = Bext + node * 3;
		double e2 = 0.0; e2 += f[0] * f[0]; e2 += f[1] * f[1]; e2 += f[2] * f[2];
		const double b0 = f[3] + bx[0], b1 = f[4] + bx[1], b2 = f[5] + bx[2];
		double b2s = 0.0; b2s += b0 * b0; b2s += b1 * b1; b2s += b2 * b2;
		eE += vol * e2 / fourPI;                                               // field.h:452
		eB += vol * b2s / fourPI;
	}
	block_sum2(eE, eB, sh);
	if (threadIdx.x == 0) { partial[2 * blockIdx.x] = eE; partial[2 * blockIdx.x + 1] = eB; }
}

// one CTA: out[0..1] = sum over n pairs, fixed order
__global__ void __launch_bounds__(kDiagBlock)
k_sum_pairs(const double* __restrict__ partial, int64_t n, double* __restrict__ out) {
	__shared__ double sh[2 * kDiagBlock];
	double a = 0.0, b = 0.0;
	for (int64_t i = threadIdx.x; i < n; i += kDiagBlock) { a += partial[2 * i]; b += partial[2 * i + 1]; }
	block_sum2(a, b, sh);
	if (threadIdx.x == 0) { out[0] = a; out[1] = b; }
}

void launch_particle_energy(const Geometry& g, const ParticleArrays& p, const int64_t* cellStart, const int* count,
                            double* partial, double* out2, cudaStream_t st) {
	if (g.ncellsLocal > 0) k_cell_energy<<<g.ncellsLocal, kDiagBlock, 0, st>>>(g, p, cellStart, count, partial);
	k_sum_pairs<<<1, kDiagBlock, 0, st>>>(partial, g.ncellsLocal, out2);
}

void launch_field_energy(const Geometry& g, const double* EB, const double* Bext, double* partial, int chunks, double* out2, cudaStream_t st) {
	k_field_energy<<<chunks, kDiagBlock, 0, st>>>(g, EB, Bext, partial, chunks);
	k_sum_pairs<<<1, kDiagBlock, 0, st>>>(partial, chunks, out2);
}

} // namespace ipic
