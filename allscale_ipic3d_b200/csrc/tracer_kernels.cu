// pic_tracer on the device (reference code/app/src/pic_tracer.cxx:77-148): independent test particles in the
// analytic Earth dipole, no cells and no migration; every `frame_interval` steps the particle's cell (taken
// BEFORE the push of that step) is counted into the density frame t / frame_interval.
//
// One thread per particle for all T steps: the state lives in registers, HBM sees 64 B per particle once and
// one 32-bit atomic per (particle, frame).  Integer atomics commute, so the frames are deterministic.  The
// arithmetic is the mover's (ipic_device.cuh: dipole_push -- same expressions as
// pic_tracer.cxx:102-124), compiled with -fmad=false.
#include "tracer_kernels.h"

namespace ipic {

__global__ void __launch_bounds__(256)
k_trace(const Geometry g, const double* __restrict__ aos, int64_t n, int T, int frameInterval, unsigned* __restrict__ counts) {
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const double2* rec = reinterpret_cast<const double2*>(aos + i * 8);
	const double2 a = rec[0], b = rec[1], c = rec[2], d = rec[3];
	double px = a.x, py = a.y, pz = b.x, vx = b.y, vy = c.x, vz = c.y;
	const double qom = d.y;

	// universe box (pic_tracer.cxx:83-85): size = cellWidth (*) size, high = origin + size
	const double sx = g.w[0] * (double)g.nx, sy = g.w[1] * (double)g.ny, sz = g.w[2] * (double)g.nz;
	const double lx = g.origin[0], ly = g.origin[1], lz = g.origin[2];
	const double hx = lx + sx, hy = ly + sy, hz = lz + sz;

	px += g.objectCenter[0]; py += g.objectCenter[1]; pz += g.objectCenter[2];          // :87
	{
		const double dx = px - g.objectCenter[0], dy = py - g.objectCenter[1], dz = pz - g.objectCenter[2];
		double r2 = 0.0; r2 += dx * dx; r2 += dy * dy; r2 += dz * dz;
		if (r2 < g.r2planet) return;                                                    // :89-93 (strict)
	}
	if (px > hx || px < lx || py > hy || py < ly || pz > hz || pz < lz) return;         // :95-98

	const bool plainDipole = fabs(g.magTemp) > 1e-100 && fabs(g.magTemp) < 1e100 && g.pic > 0.0 && g.dt > 0.0 && g.dt < 1e100;
	for (int t = 0; t < T; t++) {
		// getCellCoordinates (hdr/cell.h:63-66): truncation of (p - origin) / width
		const long long cx = (long long)((px - lx) / g.w[0]), cy = (long long)((py - ly) / g.w[1]), cz = (long long)((pz - lz) / g.w[2]);
		dipole_push(g, plainDipole, px, py, pz, vx, vy, vz, qom);                       // :106-127 (dipole field, sub-cycling, Boris)
		// periodic wrap (:129-133)
		if (px > hx) px -= sx;
		if (px < lx) px += sx;
		if (py > hy) py -= sy;
		if (py < ly) py += sy;
		if (pz > hz) pz -= sz;
		if (pz < lz) pz += sz;
		const double dx = px - g.objectCenter[0], dy = py - g.objectCenter[1], dz = pz - g.objectCenter[2];
		double r2 = 0.0; r2 += dx * dx; r2 += dy * dy; r2 += dz * dz;
		if (r2 <= g.r2planet) return;                                                   // :137-141
		if (t % frameInterval == 0) {
			// a particle sitting exactly on the upper face has coordinate == size: the reference writes out of bounds there
			if (cx >= 0 && cx < g.nx && cy >= 0 && cy < g.ny && cz >= 0 && cz < g.nz)
				atomicAdd(counts + ((((int64_t)(t / frameInterval) * g.nx + cx) * g.ny + cy) * g.nz + cz), 1u);
		}
	}
}

void launch_trace(const Geometry& g, const double* aos, int64_t n, int T, int frameInterval, unsigned* counts, cudaStream_t st) {
	if (n <= 0) return;
	k_trace<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g, aos, n, T, frameInterval, counts);
}

} // namespace ipic
