// Explicit field update of the reference as sm_100a stencil kernels (FP64, -fmad=false).
//
//   solveFieldForward   reference field.h:294-337  (computeCurlB :162-193, computeCurlE :198-229, interpC2N :234-247)
//   solveFieldLeapfrog  reference field.h:342-373  (the `if(1)` branch)
//   updateFieldsOnBoundaries reference field.h:407-440
//
// The reference updates E, Bc and B *in place* while looping over the nodes; run sequentially (as its
// tests and any single-threaded build do) node p therefore sees already-updated Bc at p-d and
// not-yet-updated E at p+d (d in {0,1}^3, d != 0).  Those dependencies only ever point across the
// hyperplanes x+y+z = const, so sweeping the hyperplanes in increasing order with all nodes of one
// plane in parallel reproduces the sequential result bit for bit.
#include "field_kernels.h"

#include <cooperative_groups.h>
#include <cstdlib>
#include <cstring>

namespace ipic {

struct FieldDims { int fx, fy, fz, bx, by, bz, jx, jy, jz; };

__device__ __forceinline__ FieldDims dims(const Geometry& g) {
	return { g.nx + 3, g.ny + 3, g.nz + 3, g.nx + 2, g.ny + 2, g.nz + 2, g.nx + 1, g.ny + 1, g.nz + 1 };
}

#define EBN(x, y, z) (EB + ((((int64_t)(x)) * d.fy + (y)) * d.fz + (z)) * 6)
#define BCN(x, y, z) (Bc + ((((int64_t)(x)) * d.by + (y)) * d.bz + (z)) * 3)
#define JN(x, y, z)  (J  + ((((int64_t)(x)) * d.jy + (y)) * d.jz + (z)) * 3)

// lexicographic `pos < bcfield.size() - 1` (reference field.h:315, :366)
__device__ __forceinline__ bool ghost_guard(const Geometry& g, int x, int y, int z) {
	if (x < g.nx + 1) return true; if (x > g.nx + 1) return false;
	if (y < g.ny + 1) return true; if (y > g.ny + 1) return false;
	return z < g.nz + 1;
}

__device__ void solve_forward_node(const Geometry& g, double* EB, double* Bc, const double* J, int x, int y, int z) {
	const FieldDims d = dims(g);
	double V[2][2][2][3];
	#pragma unroll
	for (int i = 0; i < 2; i++)
		#pragma unroll
		for (int j = 0; j < 2; j++)
			#pragma unroll
			for (int k = 0; k < 2; k++) { const double* b = BCN(x - i, y - j, z - k); V[i][j][k][0] = b[0]; V[i][j][k][1] = b[1]; V[i][j][k][2] = b[2]; }
	double zdy, ydz, xdz, zdx, ydx, xdy;
	// computeCurlB, term order as written in the reference (field.h:179-192)
	zdy = .25 * (V[0][0][0][2] - V[0][1][0][2] + V[0][0][1][2] - V[0][1][1][2] + V[1][0][0][2] - V[1][1][0][2] + V[1][0][1][2] - V[1][1][1][2]) / g.w[1];
	ydz = .25 * (V[0][0][0][1] - V[0][0][1][1] + V[1][0][0][1] - V[1][0][1][1] + V[0][1][0][1] - V[0][1][1][1] + V[1][1][0][1] - V[1][1][1][1]) / g.w[2];
	xdz = .25 * (V[0][0][0][0] - V[0][0][1][0] + V[1][0][0][0] - V[1][0][1][0] + V[0][1][0][0] - V[0][1][1][0] + V[1][1][0][0] - V[1][1][1][0]) / g.w[2];
	zdx = .25 * (V[0][0][0][2] - V[1][0][0][2] + V[0][0][1][2] - V[1][0][1][2] + V[0][1][0][2] - V[1][1][0][2] + V[0][1][1][2] - V[1][1][1][2]) / g.w[0];
	ydx = .25 * (V[0][0][0][1] - V[1][0][0][1] + V[0][0][1][1] - V[1][0][1][1] + V[0][1][0][1] - V[1][1][0][1] + V[0][1][1][1] - V[1][1][1][1]) / g.w[0];
	xdy = .25 * (V[0][0][0][0] - V[0][1][0][0] + V[0][0][1][0] - V[0][1][1][0] + V[1][0][0][0] - V[1][1][0][0] + V[1][0][1][0] - V[1][1][1][0]) / g.w[1];
	const double curlB[3] = { zdy - ydz, xdz - zdx, ydx - xdy };
	double* E = EBN(x, y, z);
	const double* Jn = JN(x - 1, y - 1, z - 1);
	#pragma unroll
	for (int a = 0; a < 3; a++) E[a] += (g.c * curlB[a] - Jn[a]) * g.dt;
	if (ghost_guard(g, x, y, z)) {
		#pragma unroll
		for (int i = 0; i < 2; i++)
			#pragma unroll
			for (int j = 0; j < 2; j++)
				#pragma unroll
				for (int k = 0; k < 2; k++) { const double* e = EBN(x + i, y + j, z + k); V[i][j][k][0] = e[0]; V[i][j][k][1] = e[1]; V[i][j][k][2] = e[2]; }
		// computeCurlE (field.h:215-228)
		zdy = .25 * (V[0][1][0][2] - V[0][0][0][2] + V[0][1][1][2] - V[0][0][1][2] + V[1][1][0][2] - V[1][0][0][2] + V[1][1][1][2] - V[1][0][1][2]) / g.w[1];
		ydz = .25 * (V[0][0][1][1] - V[0][0][0][1] + V[1][0][1][1] - V[1][0][0][1] + V[0][1][1][1] - V[0][1][0][1] + V[1][1][1][1] - V[1][1][0][1]) / g.w[2];
		xdz = .25 * (V[0][0][1][0] - V[0][0][0][0] + V[1][0][1][0] - V[1][0][0][0] + V[0][1][1][0] - V[0][1][0][0] + V[1][1][1][0] - V[1][1][0][0]) / g.w[2];
		zdx = .25 * (V[1][0][0][2] - V[0][0][0][2] + V[1][0][1][2] - V[0][0][1][2] + V[1][1][0][2] - V[0][1][0][2] + V[1][1][1][2] - V[0][1][1][2]) / g.w[0];
		ydx = .25 * (V[1][0][0][1] - V[0][0][0][1] + V[1][0][1][1] - V[0][0][1][1] + V[1][1][0][1] - V[0][1][0][1] + V[1][1][1][1] - V[0][1][1][1]) / g.w[0];
		xdy = .25 * (V[0][1][0][0] - V[0][0][0][0] + V[0][1][1][0] - V[0][0][1][0] + V[1][1][0][0] - V[1][0][0][0] + V[1][1][1][0] - V[1][0][1][0]) / g.w[1];
		const double curlE[3] = { zdy - ydz, xdz - zdx, ydx - xdy };
		double* b = BCN(x, y, z);
		#pragma unroll
		for (int a = 0; a < 3; a++) b[a] -= g.c * curlE[a] * g.dt;
	}
	// interpC2N (field.h:234-247)
	double s[3] = { 0.0, 0.0, 0.0 };
	#pragma unroll
	for (int i = 0; i < 2; i++)
		#pragma unroll
		for (int j = 0; j < 2; j++)
			#pragma unroll
			for (int k = 0; k < 2; k++) { const double* b = BCN(x - i, y - j, z - k); s[0] += b[0]; s[1] += b[1]; s[2] += b[2]; }
	E[3] = .125 * s[0]; E[4] = .125 * s[1]; E[5] = .125 * s[2];
}

__device__ void solve_leapfrog_node(const Geometry& g, double* EB, double* Bc, const double* J, int x, int y, int z) {
	const FieldDims d = dims(g);
	const double dt = g.dt, c = g.c;
	double* E = EBN(x, y, z);
	const double* Jn = JN(x - 1, y - 1, z - 1);
	double* B = BCN(x, y, z);
	const double* Bmx = BCN(x - 1, y, z);
	const double* Bmy = BCN(x, y - 1, z);
	const double* Bmz = BCN(x, y, z - 1);
	// field.h:353-363
	E[0] = E[0] + dt * (c * c * ((B[2] - Bmy[2]) / g.w[1] - (B[1] - Bmz[1]) / g.w[2]) - Jn[0]);
	E[1] = E[1] + dt * (c * c * ((B[0] - Bmz[0]) / g.w[2] - (B[2] - Bmx[2]) / g.w[0]) - Jn[1]);
	E[2] = E[2] + dt * (c * c * ((B[1] - Bmx[1]) / g.w[0] - (B[0] - Bmy[0]) / g.w[1]) - Jn[2]);
	if (ghost_guard(g, x, y, z)) {
		const double* Epx = EBN(x + 1, y, z);
		const double* Epy = EBN(x, y + 1, z);
		const double* Epz = EBN(x, y, z + 1);
		// field.h:367-371
		B[0] = B[0] - dt * ((Epy[2] - E[2]) / g.w[1] - (Epz[1] - E[1]) / g.w[2]);
		B[1] = B[1] - dt * ((Epz[0] - E[0]) / g.w[2] - (Epx[2] - E[2]) / g.w[0]);
		B[2] = B[2] - dt * ((Epx[1] - E[1]) / g.w[0] - (Epy[0] - E[0]) / g.w[1]);
	}
}

// one hyperplane x+y+z == plane of the node range [1, n+2)^3
template<bool FORWARD>
__global__ void k_solver_plane(const Geometry g, double* EB, double* Bc, const double* J, const int* stopFlag, int plane) {
	if (*stopFlag) return;
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	const int ex = g.nx + 1, ey = g.ny + 1;
	if (t >= ex * ey) return;
	const int x = 1 + t / ey, y = 1 + t % ey, z = plane - x - y;
	if (z < 1 || z > g.nz + 1) return;
	if (FORWARD) solve_forward_node(g, EB, Bc, J, x, y, z); else solve_leapfrog_node(g, EB, Bc, J, x, y, z);
}

template<bool FORWARD>
__global__ void k_solver_node(const Geometry g, double* EB, double* Bc, const double* J, int x, int y, int z) {
	if (FORWARD) solve_forward_node(g, EB, Bc, J, x, y, z); else solve_leapfrog_node(g, EB, Bc, J, x, y, z);
}

// ---- the sweep over the hyperplanes ----------------------------------------------------------------------------------
// What the sweep costs is its dependency chain: 99 planes at 32^3, each a few hundred nodes of ~80 scattered 8-byte loads.
// Measured on B200 (scripts/probe_solver.py, forward solver, per step): one launch per plane 0.72 ms (a kernel boundary per
// plane); ONE CTA walking all planes with __syncthreads 1.11 ms (a single SM's load/store unit serialises the plane: ncu
// lg_throttle 4.6, long scoreboard 8.4 per issue); a cooperative launch of as many CTAs as the per-plane grid had, with a
// grid-wide barrier per plane 0.535 ms; one 8-CTA cluster with the cluster barrier per plane 0.497 ms (leapfrog 0.565 -> 0.484;
// profiles/r2y_solver_*.txt): the default.  IPIC_SOLVER_MODE = coop | planes | one | cluster selects; a cooperative launch
// that the device refuses falls back to one launch per plane, a cluster launch that does not fit to the cooperative one.

template<bool FORWARD>
__global__ void __launch_bounds__(512)
k_solver_wavefront(const Geometry g, double* EB, double* Bc, const double* J, const int* stopFlag) {
	if (*stopFlag) return;
	const int ex = g.nx + 1, ey = g.ny + 1;
	const int lo = 3, hi = ex + ey + (g.nz + 1);
	for (int plane = lo; plane <= hi; plane++) {
		for (int t = threadIdx.x; t < ex * ey; t += blockDim.x) {
			const int x = 1 + t / ey, y = 1 + t % ey, z = plane - x - y;
			if (z < 1 || z > g.nz + 1) continue;
			if (FORWARD) solve_forward_node(g, EB, Bc, J, x, y, z); else solve_leapfrog_node(g, EB, Bc, J, x, y, z);
		}
		__syncthreads();
	}
}

// all CTAs of a cooperative launch walk the planes together: grid.sync() orders the plane's stores before the next plane's loads
template<bool FORWARD>
__global__ void __launch_bounds__(128)
k_solver_coop(const Geometry g, double* EB, double* Bc, const double* J, const int* stopFlag) {
	cooperative_groups::grid_group grid = cooperative_groups::this_grid();
	if (*stopFlag) return;                               // the same word for every thread of the grid: nobody reaches a barrier
	const int ex = g.nx + 1, ey = g.ny + 1;
	const int lo = 3, hi = ex + ey + (g.nz + 1);
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	const bool mine = t < ex * ey;
	const int x = 1 + t / ey, y = 1 + t % ey;
	for (int plane = lo; plane <= hi; plane++) {
		const int z = plane - x - y;
		if (mine && z >= 1 && z <= g.nz + 1) {
			if (FORWARD) solve_forward_node(g, EB, Bc, J, x, y, z); else solve_leapfrog_node(g, EB, Bc, J, x, y, z);
		}
		grid.sync();
	}
}

// the same walk by ONE thread-block cluster (8 CTAs, or 16 where the grid needs them): the barrier per plane is the cluster's
// hardware barrier (barrier.cluster, release / acquire at cluster scope: a few hundred cycles) instead of the grid-wide one
template<bool FORWARD>
__global__ void __launch_bounds__(512)
k_solver_cluster(const Geometry g, double* EB, double* Bc, const double* J, const int* stopFlag) {
	cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
	if (*stopFlag) return;                               // the same word for every thread of the cluster: nobody reaches a barrier
	const int ex = g.nx + 1, ey = g.ny + 1;
	const int lo = 3, hi = ex + ey + (g.nz + 1);
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	const bool mine = t < ex * ey;
	const int x = 1 + t / ey, y = 1 + t % ey;
	for (int plane = lo; plane <= hi; plane++) {
		const int z = plane - x - y;
		if (mine && z >= 1 && z <= g.nz + 1) {
			if (FORWARD) solve_forward_node(g, EB, Bc, J, x, y, z); else solve_leapfrog_node(g, EB, Bc, J, x, y, z);
		}
		cluster.sync();
	}
}

// 0 planes, 1 one CTA, 2 cooperative, 3 one cluster, 4 (default) one cluster of the portable size 8 where it covers the grid
// (up to 62^3 cells), else cooperative
static int solver_mode() {
	static int v = -1;
	if (v < 0) {
		const char* e = std::getenv("IPIC_SOLVER_MODE");
		v = 4;
		if (e && std::strcmp(e, "planes") == 0) v = 0;
		if (e && std::strcmp(e, "one") == 0) v = 1;
		if (e && std::strcmp(e, "coop") == 0) v = 2;
		if (e && std::strcmp(e, "cluster") == 0) v = 3;
	}
	return v;
}

// one cluster that covers all (x, y) columns with at most 512 threads per CTA, if there is one; false: not launched
template<bool FORWARD>
static bool launch_solver_cluster(const Geometry& g, double* EB, double* Bc, const double* J, const int* stopFlag, cudaStream_t st, bool portableOnly) {
	const int threads = (g.nx + 1) * (g.ny + 1);
	int ctas = 8;
	if ((threads + ctas - 1) / ctas > 512) { if (portableOnly) return false; ctas = 16; }
	const int perCta = (((threads + ctas - 1) / ctas + 31) / 32) * 32;
	if (perCta > 512) return false;
	auto kernel = k_solver_cluster<FORWARD>;
	if (ctas > 8 && cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); return false; }
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(perCta); cfg.dynamicSmemBytes = 0; cfg.stream = st;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = ctas; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr; cfg.numAttrs = 1;
	if (cudaLaunchKernelEx(&cfg, kernel, g, EB, Bc, J, stopFlag) == cudaSuccess) return true;
	cudaGetLastError();
	return false;
}

int launch_field_solver(const Geometry& g, bool forward, double* EB, double* Bc, const double* J, const int* stopFlag, cudaStream_t st) {
	const int threads = (g.nx + 1) * (g.ny + 1);
	const int lo = 3, hi = (g.nx + 1) + (g.ny + 1) + (g.nz + 1);
	const int mode = solver_mode();
	if (mode == 1) {
		if (forward) k_solver_wavefront<true><<<1, 512, 0, st>>>(g, EB, Bc, J, stopFlag);
		else k_solver_wavefront<false><<<1, 512, 0, st>>>(g, EB, Bc, J, stopFlag);
		return 1;
	}
	if (mode >= 3) {
		if (forward ? launch_solver_cluster<true>(g, EB, Bc, J, stopFlag, st, mode == 4) : launch_solver_cluster<false>(g, EB, Bc, J, stopFlag, st, mode == 4)) return 1;
	}
	if (mode >= 2) {
		Geometry gg = g;
		void* args[] = { &gg, &EB, &Bc, &J, &stopFlag };
		const void* kernel = forward ? (const void*)k_solver_coop<true> : (const void*)k_solver_coop<false>;
		if (cudaLaunchCooperativeKernel(kernel, dim3((threads + 127) / 128), dim3(128), args, 0, st) == cudaSuccess) return 1;
		cudaGetLastError();                              // refused (grid too large to be co-resident, or no support): per-plane launches
	}
	int launches = 0;
	for (int p = lo; p <= hi; p++) {
		if (forward) k_solver_plane<true><<<(threads + 127) / 128, 128, 0, st>>>(g, EB, Bc, J, stopFlag, p);
		else k_solver_plane<false><<<(threads + 127) / 128, 128, 0, st>>>(g, EB, Bc, J, stopFlag, p);
		launches++;
	}
	return launches;
}

void launch_field_solver_node(const Geometry& g, bool forward, double* EB, double* Bc, const double* J, int x, int y, int z, cudaStream_t st) {
	if (forward) k_solver_node<true><<<1, 1, 0, st>>>(g, EB, Bc, J, x, y, z);
	else k_solver_node<false><<<1, 1, 0, st>>>(g, EB, Bc, J, x, y, z);
}

// ---- periodic ghost update, faces only (reference field.h:407-440) -------------------------------

// grid of n^3 elements of W doubles; copies plane n-2 -> 0 and 1 -> n-1 along each axis for the
// in-face indices [1, n-1)
__global__ void k_boundaries(double* gdata, int n, int W, const int* stopFlag) {
	if (*stopFlag) return;
	const int end = n - 1;
	const int m = end - 1;                       // number of in-face indices per direction
	const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const int64_t total = (int64_t)m * m * 6 * W;
	if (t >= total) return;
	const int w = (int)(t % W);
	const int face = (int)((t / W) % 6);
	const int b = 1 + (int)((t / (W * 6)) % m), a = 1 + (int)(t / ((int64_t)W * 6 * m));
	int dx, dy, dz, sx, sy, sz;
	switch (face) {
		case 0: dx = 0;   dy = a; dz = b; sx = end - 1; sy = a; sz = b; break;
		case 1: dx = end; dy = a; dz = b; sx = 1;       sy = a; sz = b; break;
		case 2: dx = a; dy = 0;   dz = b; sx = a; sy = end - 1; sz = b; break;
		case 3: dx = a; dy = end; dz = b; sx = a; sy = 1;       sz = b; break;
		case 4: dx = a; dy = b; dz = 0;   sx = a; sy = b; sz = end - 1; break;
		default: dx = a; dy = b; dz = end; sx = a; sy = b; sz = 1; break;
	}
	gdata[(((int64_t)dx * n + dy) * n + dz) * W + w] = gdata[(((int64_t)sx * n + sy) * n + sz) * W + w];
}

void launch_update_boundaries(const Geometry& g, double* EB, double* Bext, double* Bc, const int* stopFlag, cudaStream_t st) {
	auto run = [&](double* data, int n, int W) {
		const int64_t total = (int64_t)(n - 2) * (n - 2) * 6 * W;
		if (total > 0) k_boundaries<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(data, n, W, stopFlag);
	};
	run(EB, g.nx + 3, 6);
	run(Bext, g.nx + 3, 3);
	run(Bc, g.nx + 2, 3);
}

// ---- layout transforms ---------------------------------------------------------------------------

__global__ void k_split_field(const double* __restrict__ aos9, double* __restrict__ EB, double* __restrict__ Bext, int64_t nodes) {
	const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= nodes * 9) return;
	const int64_t n = t / 9; const int k = (int)(t % 9);
	if (k < 6) EB[n * 6 + k] = aos9[t]; else Bext[n * 3 + (k - 6)] = aos9[t];
}

__global__ void k_merge_field(double* __restrict__ aos9, const double* __restrict__ EB, const double* __restrict__ Bext, int64_t nodes) {
	const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= nodes * 9) return;
	const int64_t n = t / 9; const int k = (int)(t % 9);
	aos9[t] = k < 6 ? EB[n * 6 + k] : Bext[n * 3 + (k - 6)];
}

void launch_split_field(const double* aos9, double* EB, double* Bext, int64_t nodes, cudaStream_t st) {
	k_split_field<<<(unsigned)((nodes * 9 + 255) / 256), 256, 0, st>>>(aos9, EB, Bext, nodes);
}
void launch_merge_field(double* aos9, const double* EB, const double* Bext, int64_t nodes, cudaStream_t st) {
	k_merge_field<<<(unsigned)((nodes * 9 + 255) / 256), 256, 0, st>>>(aos9, EB, Bext, nodes);
}

} // namespace ipic
