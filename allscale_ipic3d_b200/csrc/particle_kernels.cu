// Hand-written sm_100a kernels of the particle path: fused mover + classify (+ deposit) and import.
//
// One CTA owns one cell.  A cell's particles are streamed once from generation `in` (coalesced FP64
// SoA loads), pushed, classified, and written once to generation `out`:
//   * stayers  -> head of the cell's segment, order preserved (ballot + prefix-scan compaction),
//   * leavers  -> tail of the same segment, written backwards (the cell's outbox) with a direction tag,
//   * absorbed -> dropped.
// The import kernel then appends, for every target cell, the matching outbox entries of its 26
// neighbours in the reference's canonical direction order.  No global atomics are used for data
// placement, so the result (including per-cell particle order) is deterministic and equals the
// reference's (exportParticles/importParticles, cell.h:975-1310).
//
// Compile with -fmad=false (bit-faithful FP64, see ipic_device.cuh).
#include "ipic_device.cuh"
#include "particle_kernels.h"

namespace ipic {

constexpr int kBlock = 256;
constexpr int kWarps = kBlock / 32;

// ---- TMA (bulk async copy) helpers: global -> shared, completion on an mbarrier ------------------

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"WAIT_LOOP:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@p bra WAIT_DONE;\n"
		"bra WAIT_LOOP;\n"
		"WAIT_DONE:\n"
		"}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// streaming (read-once / write-once) accesses: keep L1 clean, let L2 evict first
__device__ __forceinline__ double ld_stream(const double* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(double* p, double v) { __stcs(p, v); }

// ---- fused mover + classify (+ deposit) ----------------------------------------------------------

template<int ITEMS, bool INTERP, bool DEPOSIT, bool MIGRATE, bool PUSH>
__global__ void __launch_bounds__(kBlock)
k_move(const Geometry g, const ParticleArrays in, const ParticleArrays out, const int64_t* __restrict__ cellStart,
       const int* __restrict__ countIn, int* __restrict__ countOut, int* __restrict__ nLeave,
       unsigned char* __restrict__ tag, const double* __restrict__ fieldEB, double* __restrict__ cellJ,
       const int* __restrict__ stopFlag, int zBegin, int zCount) {

	if (*stopFlag) return;

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int xy = blockIdx.x / zCount;
	const int zl = zBegin + (blockIdx.x - xy * zCount);          // owned z-plane within the slab
	const int cx = xy / g.ny, cy = xy - cx * g.ny, cz = g.z0 + zl;
	const int cell = xy * g.lzs + zl + g.ghost;

	__shared__ int sStay[ITEMS * kWarps];
	__shared__ int sLeave[ITEMS * kWarps];
	__shared__ int sTotal[2];
	__shared__ alignas(16) double sCorner[48];
	__shared__ uint64_t sBar;
	__shared__ double sRed[DEPOSIT ? kWarps * 24 : 1];

	const int n = countIn[cell];
	const int64_t base = cellStart[cell];
	const int cap = (int)(cellStart[cell + 1] - base);

	// geometry of this cell (reference universe_properties.h:125-142)
	double org[3], ctr[3];
	org[0] = g.origin[0] + (double)cx * g.w[0];
	org[1] = g.origin[1] + (double)cy * g.w[1];
	org[2] = g.origin[2] + (double)cz * g.w[2];
	ctr[0] = org[0] + g.hw[0]; ctr[1] = org[1] + g.hw[1]; ctr[2] = org[2] + g.hw[2];

	if (INTERP && PUSH) {
		// stage the 8 corner (E,B) values of this cell: 4 bulk copies of one z-pair of nodes (96 B) each
		if (tid == 0) mbar_init(&sBar, 1);
		__syncthreads();
		if (tid == 0 && n > 0) {
			mbar_expect_tx(&sBar, 4 * 96);
			const int fy = g.ny + 3, fz = g.nz + 3;
			#pragma unroll
			for (int i = 0; i < 2; i++)
				#pragma unroll
				for (int j = 0; j < 2; j++) {
					const int64_t node = ((int64_t)(cx + i + 1) * fy + (cy + j + 1)) * fz + (cz + 1);
					tma_bulk_g2s(&sCorner[(i * 2 + j) * 12], fieldEB + node * 6, 96, &sBar);
				}
		}
		if (n > 0) mbar_wait(&sBar, 0);
	}

	double acc[DEPOSIT ? 24 : 1];
	if (DEPOSIT) {
		#pragma unroll
		for (int a = 0; a < 24; a++) acc[a] = 0.0;
	}

	int kept = 0, left = 0;
	for (int c0 = 0; c0 < n; c0 += kBlock * ITEMS) {
		double px[ITEMS], py[ITEMS], pz[ITEMS], vx[ITEMS], vy[ITEMS], vz[ITEMS], pq[ITEMS], pm[ITEMS];
		unsigned code[ITEMS];
		#pragma unroll
		for (int k = 0; k < ITEMS; k++) {
			const int idx = c0 + k * kBlock + tid;
			if (idx < n) {
				const int64_t s = base + idx;
				px[k] = ld_stream(in.f[0] + s); py[k] = ld_stream(in.f[1] + s); pz[k] = ld_stream(in.f[2] + s);
				vx[k] = ld_stream(in.f[3] + s); vy[k] = ld_stream(in.f[4] + s); vz[k] = ld_stream(in.f[5] + s);
				pq[k] = ld_stream(in.f[6] + s); pm[k] = ld_stream(in.f[7] + s);
			}
		}
		#pragma unroll
		for (int k = 0; k < ITEMS; k++) {
			const int idx = c0 + k * kBlock + tid;
			code[k] = kTagNone;
			if (idx < n) {
				if (DEPOSIT || (INTERP && PUSH)) {
					// fractional position in the cell (reference cell.h:810, :933)
					const double rx = (px[k] - org[0]) / g.w[0], ry = (py[k] - org[1]) / g.w[1], rz = (pz[k] - org[2]) / g.w[2];
					if (DEPOSIT) {
						// projectToDensityField weights (reference cell.h:815-816); corner (a,b,c) of this cell
						// is the node the reference reaches with (i,j,k) = (-a,-b,-c)
						const double jx = pq[k] * vx[k], jy = pq[k] * vy[k], jz = pq[k] * vz[k];
						#pragma unroll
						for (int a = 0; a < 2; a++)
							#pragma unroll
							for (int b = 0; b < 2; b++)
								#pragma unroll
								for (int c = 0; c < 2; c++) {
									const double fac = (a == 0 ? (1 - rx) : rx) * (b == 0 ? (1 - ry) : ry) * (c == 0 ? (1 - rz) : rz);
									const int o = ((a * 2 + b) * 2 + c) * 3;
									acc[o] += jx * fac; acc[o + 1] += jy * fac; acc[o + 2] += jz * fac;
								}
					}
					if (INTERP && PUSH) {
						const Vec3 E = trilinear(sCorner, 6, rx, ry, rz, g.vol);
						const Vec3 B = trilinear(sCorner + 3, 6, rx, ry, rz, g.vol);
						push_with_fields(g, px[k], py[k], pz[k], vx[k], vy[k], vz[k], pm[k], E, B);
					}
				}
				if (!INTERP && PUSH) {
					Vec3 E, B;
					dipole_field(g, px[k], py[k], pz[k], E, B);
					push_with_fields(g, px[k], py[k], pz[k], vx[k], vy[k], vz[k], pm[k], E, B);
				}
				if (PUSH) code[k] = MIGRATE ? classify(g, cx, cy, cz, ctr, px[k], py[k], pz[k], vx[k], vy[k], vz[k]) : 13u;
			}
		}

		if (!PUSH) continue;     // deposit-only pass: nothing is written back

		// stable compaction: order index of an element is k*kBlock + tid
		unsigned preStay[ITEMS], preLeave[ITEMS];
		#pragma unroll
		for (int k = 0; k < ITEMS; k++) {
			const bool st = code[k] == 13u, lv = code[k] < 27u && code[k] != 13u;
			const unsigned bs = __ballot_sync(0xffffffffu, st), bl = __ballot_sync(0xffffffffu, lv);
			const unsigned lt = (1u << lane) - 1u;
			preStay[k] = __popc(bs & lt); preLeave[k] = __popc(bl & lt);
			if (lane == 0) { sStay[k * kWarps + warp] = __popc(bs); sLeave[k * kWarps + warp] = __popc(bl); }
		}
		__syncthreads();
		if (warp == 0) {
			int a = (lane < ITEMS * kWarps) ? sStay[lane] : 0, b = (lane < ITEMS * kWarps) ? sLeave[lane] : 0;
			int ia = a, ib = b;
			#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				const int ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o);
				if (lane >= o) { ia += ta; ib += tb; }
			}
			if (lane < ITEMS * kWarps) { sStay[lane] = ia - a; sLeave[lane] = ib - b; }
			if (lane == 31) { sTotal[0] = ia; sTotal[1] = ib; }
		}
		__syncthreads();
		#pragma unroll
		for (int k = 0; k < ITEMS; k++) {
			const bool st = code[k] == 13u, lv = code[k] < 27u && code[k] != 13u;
			if (st || lv) {
				int64_t d;
				if (st) d = base + kept + sStay[k * kWarps + warp] + (int)preStay[k];
				else {
					d = base + cap - 1 - (left + sLeave[k * kWarps + warp] + (int)preLeave[k]);
					tag[d] = (unsigned char)code[k];
				}
				st_stream(out.f[0] + d, px[k]); st_stream(out.f[1] + d, py[k]); st_stream(out.f[2] + d, pz[k]);
				st_stream(out.f[3] + d, vx[k]); st_stream(out.f[4] + d, vy[k]); st_stream(out.f[5] + d, vz[k]);
				st_stream(out.f[6] + d, pq[k]); st_stream(out.f[7] + d, pm[k]);
			}
		}
		kept += sTotal[0]; left += sTotal[1];
		if (c0 + kBlock * ITEMS < n) __syncthreads();
	}
	if (PUSH && tid == 0) { countOut[cell] = kept; nLeave[cell] = left; }

	if (DEPOSIT) {
		// deterministic block reduction of the 8x3 corner sums: lanes (xor tree), then warps in order
		#pragma unroll
		for (int a = 0; a < 24; a++) {
			double v = acc[a];
			#pragma unroll
			for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
			if (lane == 0) sRed[warp * 24 + a] = v;
		}
		__syncthreads();
		if (tid < 24) {
			double v = sRed[tid];
			#pragma unroll
			for (int w = 1; w < kWarps; w++) v += sRed[w * 24 + tid];
			cellJ[(((int64_t)cx * g.ny + cy) * g.nz + cz) * 24 + tid] = v;      // indexed by GLOBAL cell id
		}
	}
}

// ---- import: append the 26 incoming outboxes in canonical order (reference cell.h:1195-1310) ------

__global__ void __launch_bounds__(kBlock)
k_import(const Geometry g, const ParticleArrays src, const int64_t* __restrict__ srcStart, const int* __restrict__ srcLeave,
         const unsigned char* __restrict__ tag, const ParticleArrays dst, const int64_t* __restrict__ dstStart,
         const int* __restrict__ stay, int* __restrict__ dstCount, int reserveOwnOutbox, int* __restrict__ stopFlag, int stepIndex, int zBegin, int zCount) {

	if (*(volatile int*)stopFlag) return;
	const int lane = threadIdx.x & 31;
	const int task = blockIdx.x * kWarps + (threadIdx.x >> 5);
	if (task >= g.nx * g.ny * zCount) return;
	const int xy = task / zCount;
	const int zl = zBegin + (task - xy * zCount);
	const int cx = xy / g.ny, cy = xy - cx * g.ny;
	const int cell = xy * g.lzs + zl + g.ghost;

	const int64_t dbase = dstStart[cell];
	const int dcap = (int)(dstStart[cell + 1] - dbase);
	const int room = dcap - (reserveOwnOutbox ? srcLeave[cell] : 0);
	int w = stay[cell];

	for (int d = 0; d < 27; d++) {
		if (d == 13) continue;
		const int a = d / 9, b = (d / 3) % 3, c = d % 3;
		// source cell = target + (a-1,b-1,c-1) with periodic address (transfer_buffer.h, cell.h:1012-1048)
		const int sx = (cx + a - 1 + g.nx) % g.nx, sy = (cy + b - 1 + g.ny) % g.ny;
		int szl = zl + c - 1;                         // local z; ghost planes hold the neighbours' rows
		if (!g.ghost) szl = (szl + g.lz) % g.lz;
		const int scell = (sx * g.ny + sy) * g.lzs + szl + g.ghost;
		const int nl = srcLeave[scell];
		if (nl == 0) continue;
		const unsigned want = (unsigned)((2 - a) * 9 + (2 - b) * 3 + (2 - c));
		const int64_t send = srcStart[scell + 1] - 1;     // outbox entry j lives at send - j
		for (int j0 = 0; j0 < nl; j0 += 32) {
			const int j = j0 + lane;
			const bool hit = j < nl && tag[send - j] == want;
			const unsigned m = __ballot_sync(0xffffffffu, hit);
			if (hit) {
				const int o = w + __popc(m & ((1u << lane) - 1u));
				if (o < room) {
					const int64_t s = send - j, t = dbase + o;
					#pragma unroll
					for (int f = 0; f < kFields; f++) dst.f[f][t] = src.f[f][s];
				}
			}
			w += __popc(m);
		}
	}
	if (lane == 0) {
		if (w > room) atomicMin(stopFlag + 1, stepIndex), atomicExch(stopFlag + 2, 1);   // recorded; flag raised by k_commit
		dstCount[cell] = w;
	}
}

// raises the stop flag after an import pass that overflowed (kept separate so that every cell of the
// failing step still gets its import attempt recorded in dstCount == required size)
__global__ void k_commit_flag(int* stopFlag) {
	if (stopFlag[2]) stopFlag[0] = 1;
}

// ---- node-centric gather of the per-cell corner sums (reference cell.h:787-824) -------------------

__global__ void k_gather_current(const Geometry g, const double* __restrict__ cellJ, double* __restrict__ J, const int* __restrict__ stopFlag) {
	if (*stopFlag) return;
	const int64_t nn = (int64_t)(g.nx + 1) * (g.ny + 1) * (g.nz + 1);
	const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= nn * 3) return;
	const int comp = (int)(t % 3);
	const int64_t node = t / 3;
	const int z = (int)(node % (g.nz + 1)), y = (int)((node / (g.nz + 1)) % (g.ny + 1)), x = (int)(node / ((int64_t)(g.nz + 1) * (g.ny + 1)));
	double Js = 0.0;
	for (int i = -1; i < 1; i++) for (int j = -1; j < 1; j++) for (int k = -1; k < 1; k++) {
		int ux = x + i, uy = y + j, uz = z + k;
		ux = ux < 0 ? g.nx - 1 : (ux >= g.nx ? 0 : ux);
		uy = uy < 0 ? g.ny - 1 : (uy >= g.ny ? 0 : uy);
		uz = uz < 0 ? g.nz - 1 : (uz >= g.nz ? 0 : uz);
		const int64_t cell = ((int64_t)ux * g.ny + uy) * g.nz + uz;       // cellJ is indexed by GLOBAL cell id
		Js += cellJ[cell * 24 + (((-i) * 2 + (-j)) * 2 + (-k)) * 3 + comp];
	}
	J[t] = (Js / g.vol) / 8.0;
}

// ---- invariant check (reference cell.h:48-59, :1168-1184) ----------------------------------------

__global__ void k_count_misplaced(const Geometry g, const ParticleArrays p, const int64_t* __restrict__ cellStart,
                                  const int* __restrict__ count, unsigned long long* __restrict__ bad) {
	const int xyz = blockIdx.x;
	const int zl = xyz % g.lz, xy = xyz / g.lz;
	const int cx = xy / g.ny, cy = xy - cx * g.ny, cz = g.z0 + zl;
	const int cell = xy * g.lzs + zl + g.ghost;
	const double c0 = g.origin[0] + (double)cx * g.w[0] + g.w[0] / 2.0;
	const double c1 = g.origin[1] + (double)cy * g.w[1] + g.w[1] / 2.0;
	const double c2 = g.origin[2] + (double)cz * g.w[2] + g.w[2] / 2.0;
	const int64_t base = cellStart[cell];
	unsigned long long local = 0;
	for (int i = threadIdx.x; i < count[cell]; i += blockDim.x) {
		const bool ok = fabs(p.f[0][base + i] - c0) <= g.hw[0] && fabs(p.f[1][base + i] - c1) <= g.hw[1] && fabs(p.f[2][base + i] - c2) <= g.hw[2];
		if (!ok) local++;
	}
	if (local) atomicAdd(bad, local);
}

// ---- AoS <-> SoA transforms for upload / download ------------------------------------------------

// aos holds the owned cells' particles cell-major; aosStart[c] = prefix of counts over owned cells (c in
// owned-cell order (x*ny + y)*lz + zl)
__global__ void k_scatter_aos(const Geometry g, const double* __restrict__ aos, const int64_t* __restrict__ aosStart,
                              const ParticleArrays p, const int64_t* __restrict__ cellStart) {
	const int oc = blockIdx.x;
	const int zl = oc % g.lz, xy = oc / g.lz;
	const int cell = xy * g.lzs + zl + g.ghost;
	const int64_t a0 = aosStart[oc], n = aosStart[oc + 1] - a0, base = cellStart[cell];
	for (int64_t e = threadIdx.x; e < n * 8; e += blockDim.x) p.f[e & 7][base + (e >> 3)] = aos[a0 * 8 + e];
}

__global__ void k_gather_aos(const Geometry g, double* __restrict__ aos, const int64_t* __restrict__ aosStart,
                             const ParticleArrays p, const int64_t* __restrict__ cellStart) {
	const int oc = blockIdx.x;
	const int zl = oc % g.lz, xy = oc / g.lz;
	const int cell = xy * g.lzs + zl + g.ghost;
	const int64_t a0 = aosStart[oc], n = aosStart[oc + 1] - a0, base = cellStart[cell];
	for (int64_t e = threadIdx.x; e < n * 8; e += blockDim.x) aos[a0 * 8 + e] = p.f[e & 7][base + (e >> 3)];
}

// copy the live head of every owned cell from one layout to another (relayout)
__global__ void k_copy_heads(const Geometry g, const ParticleArrays src, const int64_t* __restrict__ srcStart,
                             const ParticleArrays dst, const int64_t* __restrict__ dstStart, const int* __restrict__ count) {
	const int oc = blockIdx.x;
	const int zl = oc % g.lz, xy = oc / g.lz;
	const int cell = xy * g.lzs + zl + g.ghost;
	const int64_t s0 = srcStart[cell], d0 = dstStart[cell];
	const int n = count[cell];
	for (int f = 0; f < kFields; f++)
		for (int i = threadIdx.x; i < n; i += blockDim.x) dst.f[f][d0 + i] = src.f[f][s0 + i];
}

// ---- launchers -----------------------------------------------------------------------------------

template<int ITEMS, bool INTERP, bool DEPOSIT>
static void launch_move_t(const MoveArgs& a, cudaStream_t st) {
	const int grid = a.g.nx * a.g.ny * a.zCount;
	if (grid == 0) return;
#define IPIC_MOVE_ARGS a.g, a.in, a.out, a.cellStart, a.countIn, a.countOut, a.nLeave, a.tag, a.fieldEB, a.cellJ, a.stopFlag, a.zBegin, a.zCount
	if (!a.push) { if (DEPOSIT) k_move<ITEMS, false, DEPOSIT, false, false><<<grid, kBlock, 0, st>>>(IPIC_MOVE_ARGS); }
	else if (a.migrate) k_move<ITEMS, INTERP, DEPOSIT, true, true><<<grid, kBlock, 0, st>>>(IPIC_MOVE_ARGS);
	else k_move<ITEMS, INTERP, DEPOSIT, false, true><<<grid, kBlock, 0, st>>>(IPIC_MOVE_ARGS);
#undef IPIC_MOVE_ARGS
}

template<int ITEMS>
static void launch_move_i(const MoveArgs& a, cudaStream_t st) {
	if (a.interpolate) { if (a.deposit) launch_move_t<ITEMS, true, true>(a, st); else launch_move_t<ITEMS, true, false>(a, st); }
	else { if (a.deposit) launch_move_t<ITEMS, false, true>(a, st); else launch_move_t<ITEMS, false, false>(a, st); }
}

void launch_move(const MoveArgs& a, cudaStream_t st) {
	if (a.itemsPerThread >= 4) launch_move_i<4>(a, st);
	else if (a.itemsPerThread >= 2) launch_move_i<2>(a, st);
	else launch_move_i<1>(a, st);
}

void launch_import(const ImportArgs& a, cudaStream_t st) {
	const int tasks = a.g.nx * a.g.ny * a.zCount;
	if (tasks == 0) return;
	k_import<<<(tasks + kWarps - 1) / kWarps, kBlock, 0, st>>>(a.g, a.src, a.srcStart, a.srcLeave, a.tag, a.dst, a.dstStart, a.stay, a.dstCount,
	                                                         a.reserveOwnOutbox, a.stopFlag, a.stepIndex, a.zBegin, a.zCount);
	k_commit_flag<<<1, 1, 0, st>>>(a.stopFlag);
}

void launch_gather_current(const Geometry& g, const double* cellJ, double* J, const int* stopFlag, cudaStream_t st) {
	const int64_t n = (int64_t)(g.nx + 1) * (g.ny + 1) * (g.nz + 1) * 3;
	k_gather_current<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g, cellJ, J, stopFlag);
}

void launch_count_misplaced(const Geometry& g, const ParticleArrays& p, const int64_t* cellStart, const int* count, unsigned long long* bad, cudaStream_t st) {
	k_count_misplaced<<<g.nx * g.ny * g.lz, 128, 0, st>>>(g, p, cellStart, count, bad);
}

void launch_scatter_aos(const Geometry& g, const double* aos, const int64_t* aosStart, const ParticleArrays& p, const int64_t* cellStart, cudaStream_t st) {
	k_scatter_aos<<<g.nx * g.ny * g.lz, 256, 0, st>>>(g, aos, aosStart, p, cellStart);
}

void launch_gather_aos(const Geometry& g, double* aos, const int64_t* aosStart, const ParticleArrays& p, const int64_t* cellStart, cudaStream_t st) {
	k_gather_aos<<<g.nx * g.ny * g.lz, 256, 0, st>>>(g, aos, aosStart, p, cellStart);
}

void launch_copy_heads(const Geometry& g, const ParticleArrays& src, const int64_t* srcStart, const ParticleArrays& dst, const int64_t* dstStart, const int* count, cudaStream_t st) {
	k_copy_heads<<<g.nx * g.ny * g.lz, 256, 0, st>>>(g, src, srcStart, dst, dstStart, count);
}

} // namespace ipic
