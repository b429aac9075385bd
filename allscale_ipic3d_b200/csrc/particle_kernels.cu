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

#include <algorithm>
#include <cstdlib>

namespace ipic {

constexpr int kBlock = 128;                 // threads per CTA of the mover (one CTA per cell)
constexpr int kWarps = kBlock / 32;
#ifndef IPIC_CHUNK_ITERS
#define IPIC_CHUNK_ITERS 4           // chunk = 4 x 128 = 512 particles (24 KiB of shared memory in the one-species layout)
#endif
#ifndef IPIC_CHUNK_MINB
#define IPIC_CHUNK_MINB 5
#endif
constexpr int kRedRow = 16 * 9 + 8;          // padded row of the deposit reduction buffer (doubles): 16 segments of 8 (+1)

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
// try_wait suspends the thread in hardware until the phase completes or the time hint (ns) expires, so
// the surrounding loop executes a handful of instructions per wait instead of spinning on the issue slots
// the other warps need (ncu: the hint-less spin was 25 % of all executed instructions)
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
	uint32_t ok;
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
		"selp.u32 %0, 1, 0, p;\n"
		"}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
	return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// non-blocking probe of a phase
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
	uint32_t ok;
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
		"selp.u32 %0, 1, 0, p;\n"
		"}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
	return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
	while (!mbar_try_wait(bar, parity)) { }
}

// streaming (write-once) stores: let L2 evict first
__device__ __forceinline__ void st_stream(double* p, double v) { __stcs(p, v); }

// ---- direction-sorted index of a cell's outbox ----------------------------------------------------
//
// The import wants, for a target cell and a direction, "the entries of that neighbour's outbox with tag t".  Instead of
// letting all 26 neighbours scan all tags of an outbox, a stable counting sort of every outbox is left behind:
//   dirOff[t], t = 0..27 : number of entries with a tag below t (dirOff[27] = L, the number of leavers)
//   order[i], i < L      : the outbox entry (0 = first leaver of the cell) that is i-th in (tag, outbox order)
// Up to 32 leavers (every cell of the uniform plasma) the mover does it itself, in the registers of one warp, from the tags
// it still has in shared memory (sort_outbox_small).  Fuller outboxes are only MARKED by the mover (dirOff[0] = -2) and
// sorted by k_sort_outbox right after it, one warp per marked cell: measured on the dipole deck (~90 leavers per cell), a
// mover CTA that sits on its registers and shared memory while its warps sort costs 30-120 us per step, the separate
// kernel about 10.  Outboxes beyond kSortCap entries, and those of the chunked mover, carry dirOff[*] = -1: "not
// indexed, scan the tags".
constexpr int kSortCap = 2048;
constexpr int kSortPending = -2;
constexpr int kBigMarker = -4;             // dirOff[*] = kBigMarker - id: outbox number `id` of the big-outbox index (k_big_scan), see below

// called by ONE warp once the (at most 32) tags of the cell are in sTags
__device__ __forceinline__ void sort_outbox_small(const unsigned char* sTags, int L, int lane,
                                                  unsigned short* __restrict__ order, int* __restrict__ dirOff, int* __restrict__ sortFlag) {
	if (L > 32) {
		// (the flag is ONE word for the whole grid: tens of thousands of CTAs storing to it serialise in its L2 slice at ~9 ns
		// apiece -- measured +0.28 ms on the dipole deck -- so it is written only by the few CTAs that still read it as 0)
		// sortFlag[0]: this step's k_sort_outbox has work (reset by k_commit_flag); sortFlag[1]: sticky, for the host's decision
		// whether to launch k_sort_outbox at all in the following steps (ipic_ctx::sortLikely)
		// (all 28 words are marked: the import looks at the word of ITS direction, and must not take a stale prefix for an index
		// when the sort pass is skipped)
		if (lane < 28) dirOff[lane] = kSortPending;
		if (lane == 0) {
			if (*(volatile int*)sortFlag == 0) *sortFlag = 1;
			if (*(volatile int*)(sortFlag + 1) == 0) sortFlag[1] = 1;
		}
		return;
	}
	// rank of entry `lane` = entries with a smaller tag + earlier entries with the same tag
	const unsigned t = lane < L ? sTags[lane] : 255u;
	int below = 0, offd = 0;
	for (int k = 0; k < L; k++) {
		const unsigned tk = __shfl_sync(0xffffffffu, t, k);
		below += tk < t;
		offd += tk < (unsigned)lane;
	}
	const unsigned peers = __match_any_sync(0xffffffffu, t);
	if (lane < L) order[below + __popc(peers & ((1u << lane) - 1u))] = (unsigned short)lane;
	if (lane < 28) dirOff[lane] = offd;
}

// one warp per cell of the planes zBegin, zBegin + zStride, ...; only cells the mover marked are touched
__global__ void __launch_bounds__(256)
k_sort_outbox(const Geometry g, const int64_t* __restrict__ cellStart, const int* __restrict__ nLeave, const unsigned char* __restrict__ tag,
              unsigned short* __restrict__ order, int* __restrict__ dirOff, const int* __restrict__ stopFlag, const int* __restrict__ sortFlag,
              int zBegin, int zCount, int zStride) {
	__shared__ int sHistAll[8][28];
	if (*stopFlag || !*sortFlag) return;
	const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
	int* sHist = sHistAll[wp];
	const unsigned lt = (1u << lane) - 1u;
	const int tasks = g.nx * g.ny * zCount;
	for (int task = blockIdx.x * 8 + wp; task < tasks; task += gridDim.x * 8) {
	const int xy = task / zCount;
	const int cell = xy * g.lz + zBegin + (task - xy * zCount) * zStride;
	int* off = dirOff + (int64_t)cell * 28;
	if (off[0] != kSortPending) continue;
	const int L = nLeave[cell];
	if (L > kSortCap) {
		if (lane < 28) off[lane] = -1;
		continue;
	}
	__syncwarp();
	const int64_t start = cellStart[cell], send = cellStart[cell + 1] - 1;      // outbox entry j lives at slot send - j
	if (lane < 28) sHist[lane] = 0;
	__syncwarp();
	for (int j0 = 0; j0 < L; j0 += 32) {
		const int j = j0 + lane;
		const unsigned t = j < L ? tag[send - j] : 255u;
		const unsigned peers = __match_any_sync(0xffffffffu, t);
		if (j < L && (peers & lt) == 0u) sHist[t] += __popc(peers);          // one lane per tag value: no conflicts
		__syncwarp();
	}
	const int v = lane < 27 ? sHist[lane] : 0;
	int incl = v;
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
	__syncwarp();
	if (lane < 28) { sHist[lane] = incl - v; off[lane] = incl - v; }          // lane 27: v = 0, incl = L
	__syncwarp();
	for (int j0 = 0; j0 < L; j0 += 32) {
		const int j = j0 + lane;
		const unsigned t = j < L ? tag[send - j] : 255u;
		const unsigned peers = __match_any_sync(0xffffffffu, t);
		if (j < L) order[start + sHist[t] + __popc(peers & lt)] = (unsigned short)j;
		__syncwarp();
		if (j < L && (peers & lt) == 0u) sHist[t] += __popc(peers);
		__syncwarp();
	}
	}
}

// ---- fused mover + classify (+ deposit) ----------------------------------------------------------
//
// One CTA of 128 threads per cell; a tile is ITEMS x 128 consecutive particles of the cell = ITEMS x 4
// AoSoA blocks = one contiguous chunk of HBM.  At CTA start one thread queues TMA bulk copies
// (cp.async.bulk, completion on an mbarrier per stage) for the first `stages` tiles, i.e. normally for
// the WHOLE cell, so all of a cell's bytes are in flight before the first particle is pushed; the
// warps then consume tile after tile out of shared memory: push, classify, one ballot-based stable
// compaction (one __syncthreads per tile, after which the stage is re-armed for tile t+stages), and
// coalesced streaming stores straight from registers.  3-4 CTAs are resident per SM.

constexpr int kMaxStages = 8;
#ifndef IPIC_MINB_DEP
#define IPIC_MINB_DEP 5          // resident CTAs per SM the deposit variants are compiled for (register cap 65536/128/N)
#endif
#ifndef IPIC_MINB_NODEP
#define IPIC_MINB_NODEP 7
#endif

template<int NF, bool INTERP, bool DEPOSIT, bool MIGRATE, bool PUSH>
__global__ void __launch_bounds__(kBlock, DEPOSIT ? IPIC_MINB_DEP : IPIC_MINB_NODEP)
k_move(const Geometry g, const ParticleArrays in, const ParticleArrays out, const int64_t* __restrict__ cellStart,
       const int* __restrict__ countIn, int* __restrict__ countOut, int* __restrict__ nLeave,
       unsigned char* __restrict__ tag, unsigned short* __restrict__ order, int* __restrict__ dirOff,
       const double* __restrict__ fieldEB, double* __restrict__ cellJ,
       const int* __restrict__ stopFlag, int* __restrict__ sortFlag, int zBegin, int zCount, int zStride, int stages) {

	constexpr int ITEMS = 1;                                     // particles per thread per tile (one wins: more resident warps)
	constexpr int TILE = ITEMS * kBlock;
	constexpr int TILE_DOUBLES = TILE * NF;
	constexpr int BLOCK_DOUBLES = 32 * NF;                       // one AoSoA block

	if (*stopFlag) return;

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int xy = blockIdx.x / zCount;
	const int zl = zBegin + (blockIdx.x - xy * zCount) * zStride;   // owned z-plane within the slab (planes zBegin, zBegin + zStride, ...)
	const int cx = xy / g.ny, cy = xy - cx * g.ny, cz = g.z0 + zl;
	const int cell = xy * g.lz + zl;

	extern __shared__ __align__(128) double sTile[];             // stages x TILE_DOUBLES
	__shared__ uint64_t sFull[kMaxStages];
	__shared__ int sCnt[2][ITEMS * kWarps];                      // per (item, warp): stayers | leavers << 16
	__shared__ alignas(16) double sCorner[48];
	__shared__ uint64_t sBar;
	__shared__ double sRed[DEPOSIT ? 8 * kRedRow : 1];           // deposit reduction: 8 sums x 128 partials per pass
	__shared__ double sJ[24];                                    // the cell's 8x3 corner sums
	__shared__ double sGeo[6];                                   // cell origin, cell centre
	__shared__ unsigned char sTags[32];                          // tags of the first 32 outbox entries, for sort_outbox_small

	const int n = countIn[cell];
	const int64_t base = cellStart[cell];
	const int cap = (int)(cellStart[cell + 1] - base);
	const double* __restrict__ src = in.d + base * NF;           // base is a multiple of 32 slots
	double* __restrict__ dst = out.d + base * NF;
	const int nTiles = (n + TILE - 1) / TILE;

	if (DEPOSIT && tid < 24) sJ[tid] = 0.0;
	// geometry of this cell (reference universe_properties.h:125-142), computed once and parked in shared memory:
	// kept in registers the compiler re-derives it for every tile (int->double conversions included) to stay
	// under the launch bound, which costs more issue slots than six broadcast LDS
	if (tid >= 32 && tid < 35) {
		const int a = tid - 32;
		const int cc = a == 0 ? cx : (a == 1 ? cy : cz);
		const double o = g.origin[a] + (double)cc * g.w[a];
		sGeo[a] = o;
		sGeo[3 + a] = o + g.hw[a];
	}
	if (tid == 0) {
		for (int s = 0; s < stages; s++) mbar_init(&sFull[s], 1);
		if (INTERP && PUSH) mbar_init(&sBar, 1);
	}
	__syncthreads();

	// queue tile t (whole AoSoA blocks; the unused lanes of the last block are never looked at)
	auto issue = [&](int t, int s) {
		const int first = t * TILE;
		const int cnt = min(TILE, n - first);
		const uint32_t bytes = (uint32_t)((cnt + 31) >> 5) * (BLOCK_DOUBLES * 8);
		mbar_expect_tx(&sFull[s], bytes);
		tma_bulk_g2s(sTile + s * TILE_DOUBLES, src + (int64_t)first * NF, bytes, &sFull[s]);
	};
	if (tid == 0) {
		const int pre = min(stages, nTiles);
		for (int t = 0; t < pre; t++) issue(t, t);
		if (INTERP && PUSH && n > 0) {
			// the 8 corner (E,B) values of this cell: 4 bulk copies of one z-pair of nodes (96 B) each
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
	}

	const double* org = sGeo;
	const double* ctr = sGeo + 3;

	if (INTERP && PUSH && n > 0) mbar_wait(&sBar, 0);

	// CTA-uniform switches: a cell that touches no wall never reflects (reference cell.h:1102-1121 tests the
	// cell coordinate first); without an external dipole moment the analytic field is +-0 everywhere
	const bool wallCell = cx == 0 || cy == 0 || cz == 0 || cx == g.nx - 1 || cy == g.ny - 1 || cz == g.nz - 1;
	const bool fieldFree = !INTERP && g.magTemp == 0.0 && g.pic > 0.0 && g.dt > 0.0 && g.dt < 1e100;
	const bool plainDipole = !INTERP && fabs(g.magTemp) > 1e-100 && fabs(g.magTemp) < 1e100 && g.pic > 0.0 && g.dt > 0.0 && g.dt < 1e100;

	int kept = 0, left = 0, par = 0;
	uint32_t phase = 0;
	const unsigned lt = (1u << lane) - 1u;

	// The cell is worked through in batches of `stages` tiles (normally one batch = the whole cell).
	for (int tb = 0; tb < nTiles; tb += stages) {
		const int nb = min(stages, nTiles - tb);

		if (DEPOSIT) {
			// ---- phase A: current deposit of the batch (pre-push state), then a block reduction --------
			// Kept apart from the push so that the 24 accumulators are dead while the mover runs: one
			// kernel, one pass over HBM, but ~100 instead of ~160 registers (5 instead of 3 CTAs per SM).
			double acc[24];
			#pragma unroll
			for (int a = 0; a < 24; a++) acc[a] = 0.0;
			for (int b = 0; b < nb; b++) {
				mbar_wait(&sFull[b], phase);
				const double* tile = sTile + b * TILE_DOUBLES;
				#pragma unroll
				for (int k = 0; k < ITEMS; k++) {
					if ((tb + b) * TILE + k * kBlock + tid < n) {
						const double* p = tile + (k * kWarps + warp) * BLOCK_DOUBLES + lane;
						// projectToDensityField weights (reference cell.h:810-816); corner (a,b,c) of this cell is the
						// node the reference reaches with (i,j,k) = (-a,-b,-c).  The per-cell sums are accumulated in
						// a different order than the reference's sequential loop anyway, so this part uses a
						// reciprocal multiply and FMAs (J is compared at 1e-12, not bit for bit).
						const double rx = (p[0] - org[0]) * g.iw[0], ry = (p[32] - org[1]) * g.iw[1], rz = (p[64] - org[2]) * g.iw[2];
						const double ox = 1.0 - rx, oy = 1.0 - ry, oz = 1.0 - rz;
						const double q = NF == 8 ? p[192] : g.uq;
						const double jx = q * p[96], jy = q * p[128], jz = q * p[160];
						const double wxy[4] = { ox * oy, ox * ry, rx * oy, rx * ry };
						#pragma unroll
						for (int ab = 0; ab < 4; ab++) {
							const double f0 = wxy[ab] * oz, f1 = wxy[ab] * rz;
							const int o = ab * 6;
							acc[o] = __fma_rn(jx, f0, acc[o]); acc[o + 1] = __fma_rn(jy, f0, acc[o + 1]); acc[o + 2] = __fma_rn(jz, f0, acc[o + 2]);
							acc[o + 3] = __fma_rn(jx, f1, acc[o + 3]); acc[o + 4] = __fma_rn(jy, f1, acc[o + 4]); acc[o + 5] = __fma_rn(jz, f1, acc[o + 5]);
						}
					}
				}
			}
			// deterministic block reduction, 8 of the 24 sums per pass: thread (seg, i) parks its partial at
			// column seg*9+i (conflict-free), then thread (row, seg) adds the 8 partials of its segment in
			// fixed order and the 16 segments of a row are combined with 4 shuffles
			#pragma unroll
			for (int pass = 0; pass < 3; pass++) {
				#pragma unroll
				for (int r = 0; r < 8; r++) sRed[r * kRedRow + (tid >> 3) * 9 + (tid & 7)] = acc[pass * 8 + r];
				__syncthreads();
				{
					const int row = tid >> 4, seg = tid & 15;
					const double* r = sRed + row * kRedRow + seg * 9;
					double v = r[0];
					#pragma unroll
					for (int i = 1; i < 8; i++) v += r[i];
					v += __shfl_xor_sync(0xffffffffu, v, 1);
					v += __shfl_xor_sync(0xffffffffu, v, 2);
					v += __shfl_xor_sync(0xffffffffu, v, 4);
					v += __shfl_xor_sync(0xffffffffu, v, 8);
					if (seg == 0) sJ[pass * 8 + row] += v;
				}
				__syncthreads();
			}
		}

		// ---- phase B: push, classify, compact, store ------------------------------------------------
		for (int b = 0; b < nb; b++) {
			const int t = tb + b;
			const int c0 = t * TILE;
			if (!PUSH) {
				// deposit-only pass: the stage is free once phase A has read it (the reduction's barriers are behind us)
				if (tid == 0 && t + stages < nTiles) issue(t + stages, b);
				continue;
			}
			double px[ITEMS], py[ITEMS], pz[ITEMS], vx[ITEMS], vy[ITEMS], vz[ITEMS], pq[ITEMS], pm[ITEMS];
			unsigned code[ITEMS];
			if (!DEPOSIT) mbar_wait(&sFull[b], phase);
			const double* tile = sTile + b * TILE_DOUBLES;
			#pragma unroll
			for (int k = 0; k < ITEMS; k++) {
				if (c0 + k * kBlock + tid < n) {
					const double* p = tile + (k * kWarps + warp) * BLOCK_DOUBLES + lane;
					px[k] = p[0]; py[k] = p[32]; pz[k] = p[64];
					vx[k] = p[96]; vy[k] = p[128]; vz[k] = p[160];
					if (NF == 8) { pq[k] = p[192]; pm[k] = p[224]; } else { pq[k] = g.uq; pm[k] = g.uqom; }
				}
			}
			#pragma unroll
			for (int k = 0; k < ITEMS; k++) {
				const int idx = c0 + k * kBlock + tid;
				code[k] = kTagNone;
				if (idx < n) {
					if (INTERP) {
						// fractional position in the cell (reference cell.h:933), kept as a true division: bit-exact
						const double rx = (px[k] - org[0]) / g.w[0], ry = (py[k] - org[1]) / g.w[1], rz = (pz[k] - org[2]) / g.w[2];
						const Vec3 E = trilinear(sCorner, 6, rx, ry, rz, g.vol);
						const Vec3 B = trilinear(sCorner + 3, 6, rx, ry, rz, g.vol);
						push_with_fields(g, px[k], py[k], pz[k], vx[k], vy[k], vz[k], pm[k], E, B);
					} else if (!(fieldFree && field_free_push(g, px[k], py[k], pz[k], vx[k], vy[k], vz[k], pm[k]))) {
						dipole_push(g, plainDipole, px[k], py[k], pz[k], vx[k], vy[k], vz[k], pm[k]);
					}
					code[k] = MIGRATE ? classify(g, wallCell, cx, cy, cz, ctr, px[k], py[k], pz[k], vx[k], vy[k], vz[k]) : 13u;
				}
			}

			// stable compaction: order index of an element is k*kBlock + tid
			unsigned bs[ITEMS], bl[ITEMS];
			#pragma unroll
			for (int k = 0; k < ITEMS; k++) {
				bs[k] = __ballot_sync(0xffffffffu, code[k] == 13u);
				bl[k] = __ballot_sync(0xffffffffu, code[k] < 27u && code[k] != 13u);
				if (lane == 0) sCnt[par][k * kWarps + warp] = __popc(bs[k]) | (__popc(bl[k]) << 16);
			}
			__syncthreads();                                     // every thread has its tile in registers: stage is free
			if (tid == 0 && t + stages < nTiles) issue(t + stages, b);

			int run = 0, mine[ITEMS];
			#pragma unroll
			for (int k = 0; k < ITEMS; k++)
				#pragma unroll
				for (int w = 0; w < kWarps; w++) {
					if (w == warp) mine[k] = run;
					run += sCnt[par][k * kWarps + w];
				}
			#pragma unroll
			for (int k = 0; k < ITEMS; k++) {
				const bool st = code[k] == 13u, lv = code[k] < 27u && code[k] != 13u;
				if (st || lv) {
					int d;
					if (st) d = kept + (mine[k] & 0xffff) + __popc(bs[k] & lt);
					else {
						const int j = left + (mine[k] >> 16) + __popc(bl[k] & lt);       // outbox entry
						d = cap - 1 - j;
						tag[base + d] = (unsigned char)code[k];
						if (j < 32) sTags[j] = (unsigned char)code[k];
					}
					double* q = dst + (int64_t)(d >> 5) * BLOCK_DOUBLES + (d & 31);
					st_stream(q, px[k]); st_stream(q + 32, py[k]); st_stream(q + 64, pz[k]);
					st_stream(q + 96, vx[k]); st_stream(q + 128, vy[k]); st_stream(q + 160, vz[k]);
					if (NF == 8) { st_stream(q + 192, pq[k]); st_stream(q + 224, pm[k]); }
				}
			}
			kept += run & 0xffff; left += run >> 16;
			par ^= 1;
		}
		phase ^= 1;
	}
	if (PUSH && tid == 0) { countOut[cell] = kept; nLeave[cell] = left; }
	if (DEPOSIT && tid < 24) cellJ[cellj_index(g, cx, cy, cz) + tid] = sJ[tid];      // GLOBAL cell id, z-major: a slab is contiguous
	if (MIGRATE && PUSH) {
		__syncthreads();                                         // the tags of the cell's first leavers are in sTags
		if (warp == 0) sort_outbox_small(sTags, left, lane, order + base, dirOff + (int64_t)cell * 28, sortFlag);
	}
}

// ---- fused mover + deposit, warp-specialised ---------------------------------------------------------
//
// The analytic mover with the current deposit fused in (the headline configuration).  k_move<.., DEPOSIT> runs the
// deposit and the push as two phases per batch of tiles in the same four warps: the 24 corner accumulators must be
// dead while the mover's registers are live, the tiles are walked twice, and every batch pays a block reduction.
// Here a FIFTH warp does the deposit: it walks every tile of the cell out of the same shared-memory stages the push
// warps consume (one pass over HBM), keeps its 24 accumulators for the whole cell and reduces them once, with
// shuffles, at the end.  It is also the TMA producer: a stage is re-armed once the push warps (named barrier + one
// arrival) and the deposit warp (one arrival) are both done with it.  The FP64-heavy deposit stream and the
// integer-heavy classify/compaction stream of the push warps interleave on the SM's schedulers.
// The push warps' code is k_move's push phase unchanged, so particles, counts and tags are identical; the per-cell
// corner sums are accumulated in a different (fixed) order: J stays within 1e-12 of the reference's.

#ifndef IPIC_MINB_WS
#define IPIC_MINB_WS 4           // B200 sweep (scripts/sweep_ws.sh): 5 CTAs/SM caps the kernel at 72 registers and the accumulators spill
#endif

template<int NF, bool MIGRATE>
__global__ void __launch_bounds__(kBlock + 32, IPIC_MINB_WS)
k_move_ws(const Geometry g, const ParticleArrays in, const ParticleArrays out, const int64_t* __restrict__ cellStart,
          const int* __restrict__ countIn, int* __restrict__ countOut, int* __restrict__ nLeave,
          unsigned char* __restrict__ tag, unsigned short* __restrict__ order, int* __restrict__ dirOff,
          double* __restrict__ cellJ, const int* __restrict__ stopFlag, int* __restrict__ sortFlag, int zBegin, int zCount, int zStride, int stages) {

	constexpr int TILE = kBlock;
	constexpr int TILE_DOUBLES = TILE * NF;
	constexpr int BLOCK_DOUBLES = 32 * NF;

	if (*stopFlag) return;

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int xy = blockIdx.x / zCount;
	const int zl = zBegin + (blockIdx.x - xy * zCount) * zStride;
	const int cx = xy / g.ny, cy = xy - cx * g.ny, cz = g.z0 + zl;
	const int cell = xy * g.lz + zl;

	extern __shared__ __align__(128) double sTile[];             // stages x TILE_DOUBLES
	__shared__ uint64_t sFull[kMaxStages], sEmpty[kMaxStages];
	__shared__ int sCnt[2][kWarps];                              // per push warp: stayers | leavers << 16
	__shared__ double sGeo[6];                                   // cell origin, cell centre
	__shared__ unsigned char sTags[32];                          // tags of the first 32 outbox entries, for sort_outbox_small

	const int n = countIn[cell];
	const int64_t base = cellStart[cell];
	const int cap = (int)(cellStart[cell + 1] - base);
	const double* __restrict__ src = in.d + base * NF;
	double* __restrict__ dst = out.d + base * NF;
	const int nTiles = (n + TILE - 1) / TILE;

	if (tid >= 32 && tid < 35) {
		const int a = tid - 32;
		const int cc = a == 0 ? cx : (a == 1 ? cy : cz);
		const double o = g.origin[a] + (double)cc * g.w[a];
		sGeo[a] = o;
		sGeo[3 + a] = o + g.hw[a];
	}
	if (tid == 0) {
		for (int s = 0; s < stages; s++) { mbar_init(&sFull[s], 1); mbar_init(&sEmpty[s], 2); }
	}
	__syncthreads();
	const double* org = sGeo;
	const double* ctr = sGeo + 3;

	if (warp == kWarps) {
		// ================= deposit warp (and TMA producer) =================
		auto issue = [&](int t, int s) {
			const int first = t * TILE;
			const int cnt = min(TILE, n - first);
			const uint32_t bytes = (uint32_t)((cnt + 31) >> 5) * (BLOCK_DOUBLES * 8);
			mbar_expect_tx(&sFull[s], bytes);
			tma_bulk_g2s(sTile + s * TILE_DOUBLES, src + (int64_t)first * NF, bytes, &sFull[s]);
		};
		int nextIssue = min(stages, nTiles);                     // tiles [0, nextIssue) are on their way
		if (lane == 0) for (int t = 0; t < nextIssue; t++) issue(t, t);
		double acc[24];
		#pragma unroll
		for (int a = 0; a < 24; a++) acc[a] = 0.0;
		int s = 0;
		uint32_t phase = 0;
		for (int t = 0; t < nTiles; t++) {
			// re-arm the stages both sides are done with (tile T goes where tile T - stages was); block only for the tile needed now
			if (lane == 0) {
				while (nextIssue < nTiles && nextIssue - stages < t) {
					const int ss = nextIssue % stages;
					const uint32_t ph = (uint32_t)(nextIssue / stages - 1) & 1u;
					if (nextIssue == t) mbar_wait(&sEmpty[ss], ph);
					else if (!mbar_test(&sEmpty[ss], ph)) break;
					issue(nextIssue, ss);
					nextIssue++;
				}
			}
			__syncwarp();
			mbar_wait(&sFull[s], phase);
			const double* tile = sTile + s * TILE_DOUBLES;
			const int left = n - t * TILE;
			#pragma unroll
			for (int k = 0; k < kWarps; k++) {
				if (k * 32 + lane < left) {
					const double* p = tile + k * BLOCK_DOUBLES + lane;
					// projectToDensityField (reference cell.h:810-816) weighs q*v with the trilinear basis (1-r | r) per axis.  The
					// 8 corner weights span the monomials {1, x, y, z, xy, xz, yz, xyz} of the fractional position: accumulate
					// the 8 x 3 monomial moments (37 FP64 operations per particle instead of 48) and expand once per cell.
					const double rx = (p[0] - org[0]) * g.iw[0], ry = (p[32] - org[1]) * g.iw[1], rz = (p[64] - org[2]) * g.iw[2];
					const double q = NF == 8 ? p[192] : g.uq;
					const double jv[3] = { q * p[96], q * p[128], q * p[160] };
					const double xy = rx * ry, xz = rx * rz, yz = ry * rz;
					const double xyz = xy * rz;
					#pragma unroll
					for (int c = 0; c < 3; c++) {
						acc[c] += jv[c];
						acc[3 + c] = __fma_rn(jv[c], rx, acc[3 + c]);
						acc[6 + c] = __fma_rn(jv[c], ry, acc[6 + c]);
						acc[9 + c] = __fma_rn(jv[c], rz, acc[9 + c]);
						acc[12 + c] = __fma_rn(jv[c], xy, acc[12 + c]);
						acc[15 + c] = __fma_rn(jv[c], xz, acc[15 + c]);
						acc[18 + c] = __fma_rn(jv[c], yz, acc[18 + c]);
						acc[21 + c] = __fma_rn(jv[c], xyz, acc[21 + c]);
					}
				}
			}
			__syncwarp();
			if (lane == 0) mbar_arrive(&sEmpty[s]);
			if (++s == stages) { s = 0; phase ^= 1; }
		}
		// the cell's 24 monomial moments: fixed-shape shuffle tree (every lane ends up with all of them) ...
		#pragma unroll
		for (int a = 0; a < 24; a++) {
			double v = acc[a];
			v += __shfl_xor_sync(0xffffffffu, v, 16);
			v += __shfl_xor_sync(0xffffffffu, v, 8);
			v += __shfl_xor_sync(0xffffffffu, v, 4);
			v += __shfl_xor_sync(0xffffffffu, v, 2);
			v += __shfl_xor_sync(0xffffffffu, v, 1);
			acc[a] = v;
		}
		// ... expanded into the 8 corner sums, corner (a,b,c) = weight (a ? x : 1-x)(b ? y : 1-y)(c ? z : 1-z); lane = corner*3 + comp
		double mine = 0.0;
		#pragma unroll
		for (int c = 0; c < 3; c++) {
			const double S1 = acc[c], Sx = acc[3 + c], Sy = acc[6 + c], Sz = acc[9 + c], Sxy = acc[12 + c], Sxz = acc[15 + c], Syz = acc[18 + c], Sxyz = acc[21 + c];
			const double corner[8] = {
				S1 - Sx - Sy - Sz + Sxy + Sxz + Syz - Sxyz,      // (0,0,0)
				Sz - Sxz - Syz + Sxyz,                           // (0,0,1)
				Sy - Sxy - Syz + Sxyz,                           // (0,1,0)
				Syz - Sxyz,                                      // (0,1,1)
				Sx - Sxy - Sxz + Sxyz,                           // (1,0,0)
				Sxz - Sxyz,                                      // (1,0,1)
				Sxy - Sxyz,                                      // (1,1,0)
				Sxyz };                                          // (1,1,1)
			#pragma unroll
			for (int k = 0; k < 8; k++) if (lane == k * 3 + c) mine = corner[k];
		}
		if (lane < 24) cellJ[cellj_index(g, cx, cy, cz) + lane] = mine;      // GLOBAL cell id, z-major: a slab is contiguous
		return;
	}

	// ================= push warps =================
	const bool wallCell = cx == 0 || cy == 0 || cz == 0 || cx == g.nx - 1 || cy == g.ny - 1 || cz == g.nz - 1;
	const bool fieldFree = g.magTemp == 0.0 && g.pic > 0.0 && g.dt > 0.0 && g.dt < 1e100;
	const bool plainDipole = fabs(g.magTemp) > 1e-100 && fabs(g.magTemp) < 1e100 && g.pic > 0.0 && g.dt > 0.0 && g.dt < 1e100;
	const unsigned lt = (1u << lane) - 1u;
	int kept = 0, left = 0, par = 0, s = 0;
	uint32_t phase = 0;
	for (int t = 0; t < nTiles; t++) {
		const int idx = t * TILE + tid;
		double px, py, pz, vx, vy, vz, pq, pm;
		unsigned code = kTagNone;
		mbar_wait(&sFull[s], phase);
		if (idx < n) {
			const double* p = sTile + s * TILE_DOUBLES + warp * BLOCK_DOUBLES + lane;
			px = p[0]; py = p[32]; pz = p[64];
			vx = p[96]; vy = p[128]; vz = p[160];
			if (NF == 8) { pq = p[192]; pm = p[224]; } else { pq = g.uq; pm = g.uqom; }
			if (!(fieldFree && field_free_push(g, px, py, pz, vx, vy, vz, pm))) dipole_push(g, plainDipole, px, py, pz, vx, vy, vz, pm);
			code = MIGRATE ? classify(g, wallCell, cx, cy, cz, ctr, px, py, pz, vx, vy, vz) : 13u;
		}
		const unsigned bs = __ballot_sync(0xffffffffu, code == 13u);
		const unsigned bl = __ballot_sync(0xffffffffu, code < 27u && code != 13u);
		if (lane == 0) sCnt[par][warp] = __popc(bs) | (__popc(bl) << 16);
		asm volatile("bar.sync 1, %0;" ::"n"(kBlock) : "memory");        // every push thread has its particle in registers
		if (tid == 0) mbar_arrive(&sEmpty[s]);
		int run = 0, mine = 0;
		#pragma unroll
		for (int w = 0; w < kWarps; w++) {
			if (w == warp) mine = run;
			run += sCnt[par][w];
		}
		const bool st = code == 13u, lv = code < 27u && code != 13u;
		if (st || lv) {
			int d;
			if (st) d = kept + (mine & 0xffff) + __popc(bs & lt);
			else {
				const int j = left + (mine >> 16) + __popc(bl & lt);                 // outbox entry
				d = cap - 1 - j;
				tag[base + d] = (unsigned char)code;
				if (j < 32) sTags[j] = (unsigned char)code;
			}
			double* q = dst + (int64_t)(d >> 5) * BLOCK_DOUBLES + (d & 31);
			st_stream(q, px); st_stream(q + 32, py); st_stream(q + 64, pz);
			st_stream(q + 96, vx); st_stream(q + 128, vy); st_stream(q + 160, vz);
			if (NF == 8) { st_stream(q + 192, pq); st_stream(q + 224, pm); }
		}
		kept += run & 0xffff; left += run >> 16;
		par ^= 1;
		if (++s == stages) { s = 0; phase ^= 1; }
	}
	if (tid == 0) { countOut[cell] = kept; nLeave[cell] = left; }
	if (MIGRATE) {
		asm volatile("bar.sync 1, %0;" ::"n"(kBlock) : "memory");        // the push warps: the tags of the cell's first leavers are in sTags
		if (warp == 0) sort_outbox_small(sTags, left, lane, order + base, dirOff + (int64_t)cell * 28, sortFlag);
	}
}

// ---- chunked mover: several CTAs per cell ----------------------------------------------------------
//
// k_move gives a cell to ONE CTA, which is the right shape while cells hold a few hundred to a few thousand
// particles.  When cells are large (weak scaling on a fixed grid: 7812 particles per cell at 8 GPUs, so only 4096
// work items per slab) or very uneven (the reference's :E: / :B: loads pile most of the plasma into a few cells)
// that serialises the work.  k_move_chunk splits every cell segment into chunks of kChunk slots, one CTA each:
//
//   1. the chunk (<= kChunk particles, one contiguous piece of HBM) is brought into shared memory by ONE TMA bulk copy;
//   2. deposit (pre-push state, 24 corner sums -> block reduction), then push + classify IN PLACE in shared memory;
//      per 32-particle block the numbers of stayers / leavers (ballots) are kept, and scanned over the chunk;
//   3. a cell with several chunks needs, for a stable compaction, the totals of the chunks before this one: a
//      single-pass chained scan with decoupled look-back over one 64-bit status word per chunk
//      (flag | stayers | leavers).  Chunks are handed out through a ticket counter, so every predecessor of a
//      running chunk is itself running or done -- the look-back cannot wait for a CTA that was never scheduled;
//   4. the chunk is written out of shared memory: stayers to the head of the cell's segment at (prefix + offset),
//      leavers backwards into the outbox at the tail, exactly where the one-CTA kernel puts them.
//
// Results (particle order, counts, tags) are identical to k_move's; the per-cell corner sums of the deposit are
// the chunk partials added in chunk order by the cell's last chunk.

constexpr int kChunkIters = IPIC_CHUNK_ITERS;      // 32-particle blocks per warp and chunk
constexpr int kChunk = kChunkIters * kBlock;       // particles per chunk
constexpr int kChunkBlocks = kChunk / 32;

#define kStatusAggregate (1ull << 62)
#define kStatusPrefix (2ull << 62)
#define kStatusMask (3ull << 62)
__device__ __forceinline__ unsigned long long pack_status(unsigned long long flag, unsigned stay, unsigned leave) {
	return flag | ((unsigned long long)stay << 31) | (unsigned long long)leave;
}
__device__ __forceinline__ unsigned status_stay(unsigned long long w) { return (unsigned)((w >> 31) & 0x7fffffffu); }
__device__ __forceinline__ unsigned status_leave(unsigned long long w) { return (unsigned)(w & 0x7fffffffu); }

template<int NF, bool INTERP, bool DEPOSIT, bool MIGRATE>
__global__ void __launch_bounds__(kBlock, IPIC_CHUNK_MINB)
k_move_chunk(const Geometry g, const ParticleArrays in, const ParticleArrays out, const int64_t* __restrict__ cellStart,
             const int* __restrict__ countIn, int* __restrict__ countOut, int* __restrict__ nLeave,
             unsigned char* __restrict__ tag, const double* __restrict__ fieldEB, double* __restrict__ cellJ, double* chunkJ,
             const int* __restrict__ stopFlag, const int2* __restrict__ chunkTable, int rangeStart, unsigned* ticket,
             unsigned long long* status, int* __restrict__ dirOff) {

	constexpr int BLOCK_DOUBLES = 32 * NF;
	if (*stopFlag) return;

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	extern __shared__ __align__(128) double sTile[];             // kChunk x NF doubles, AoSoA blocks as in HBM
	__shared__ uint64_t sFull, sBar;
	__shared__ int sPos;
	__shared__ unsigned char sCode[kChunk];
	__shared__ int sBlk[kChunkBlocks];                           // per 32-particle block: stayers | leavers << 16
	__shared__ int sOff[2][kChunkBlocks];                        // exclusive scans over the blocks: [0] stayers, [1] leavers
	__shared__ int sPrefix[2];                                   // stayers / leavers of the cell's earlier chunks
	__shared__ alignas(16) double sCorner[48];
	__shared__ double sRed[DEPOSIT ? 8 * kRedRow : 1];
	__shared__ double sJ[24];
	__shared__ double sGeo[6];

	if (tid == 0) sPos = rangeStart + (int)atomicAdd(ticket, 1u);
	__syncthreads();
	const int pos = sPos;
	const int2 ref = chunkTable[pos];
	const int cell = ref.x, j = ref.y;
	const int xy = cell / g.lz, zl = cell - xy * g.lz;
	const int cx = xy / g.ny, cy = xy - cx * g.ny, cz = g.z0 + zl;

	const int n = countIn[cell];
	const int lo = j * kChunk;
	const int cnt = min(kChunk, n - lo);
	if (cnt <= 0 && !(j == 0 && n == 0)) return;                 // chunk beyond the population: nobody waits for it
	const bool last = lo + kChunk >= n;                          // the cell's last populated chunk
	if (MIGRATE && last && tid < 28) dirOff[(int64_t)cell * 28 + tid] = -1;       // outboxes of this kernel carry no direction index: the import scans their tags
	const int64_t base = cellStart[cell];
	const int cap = (int)(cellStart[cell + 1] - base);
	double* __restrict__ dst = out.d + base * NF;

	if (tid >= 32 && tid < 35) {
		const int a = tid - 32;
		const int cc = a == 0 ? cx : (a == 1 ? cy : cz);
		const double o = g.origin[a] + (double)cc * g.w[a];
		sGeo[a] = o;
		sGeo[3 + a] = o + g.hw[a];
	}
	if (tid == 0) {
		mbar_init(&sFull, 1);
		if (INTERP) mbar_init(&sBar, 1);
	}
	__syncthreads();
	if (tid == 0 && cnt > 0) {
		const uint32_t bytes = (uint32_t)((cnt + 31) >> 5) * (BLOCK_DOUBLES * 8);
		mbar_expect_tx(&sFull, bytes);
		tma_bulk_g2s(sTile, in.d + (base + lo) * NF, bytes, &sFull);
		if (INTERP) {
			mbar_expect_tx(&sBar, 4 * 96);
			const int fy = g.ny + 3, fz = g.nz + 3;
			#pragma unroll
			for (int i = 0; i < 2; i++)
				#pragma unroll
				for (int jj = 0; jj < 2; jj++) {
					const int64_t node = ((int64_t)(cx + i + 1) * fy + (cy + jj + 1)) * fz + (cz + 1);
					tma_bulk_g2s(&sCorner[(i * 2 + jj) * 12], fieldEB + node * 6, 96, &sBar);
				}
		}
	}
	const double* org = sGeo;
	const double* ctr = sGeo + 3;
	const bool wallCell = cx == 0 || cy == 0 || cz == 0 || cx == g.nx - 1 || cy == g.ny - 1 || cz == g.nz - 1;
	const bool fieldFree = !INTERP && g.magTemp == 0.0 && g.pic > 0.0 && g.dt > 0.0 && g.dt < 1e100;
	const bool plainDipole = !INTERP && fabs(g.magTemp) > 1e-100 && fabs(g.magTemp) < 1e100 && g.pic > 0.0 && g.dt > 0.0 && g.dt < 1e100;
	const unsigned lt = (1u << lane) - 1u;

	if (cnt > 0) {
		if (INTERP) mbar_wait(&sBar, 0);
		mbar_wait(&sFull, 0);
	}

	if (DEPOSIT) {
		double acc[24];
		#pragma unroll
		for (int a = 0; a < 24; a++) acc[a] = 0.0;
		for (int it = 0; it < kChunkIters; it++) {
			const int idx = it * kBlock + tid;
			if (idx < cnt) {
				const double* p = sTile + (it * kWarps + warp) * BLOCK_DOUBLES + lane;
				const double rx = (p[0] - org[0]) * g.iw[0], ry = (p[32] - org[1]) * g.iw[1], rz = (p[64] - org[2]) * g.iw[2];
				const double ox = 1.0 - rx, oy = 1.0 - ry, oz = 1.0 - rz;
				const double q = NF == 8 ? p[192] : g.uq;
				const double jx = q * p[96], jy = q * p[128], jz = q * p[160];
				const double wxy[4] = { ox * oy, ox * ry, rx * oy, rx * ry };
				#pragma unroll
				for (int ab = 0; ab < 4; ab++) {
					const double f0 = wxy[ab] * oz, f1 = wxy[ab] * rz;
					const int o = ab * 6;
					acc[o] = __fma_rn(jx, f0, acc[o]); acc[o + 1] = __fma_rn(jy, f0, acc[o + 1]); acc[o + 2] = __fma_rn(jz, f0, acc[o + 2]);
					acc[o + 3] = __fma_rn(jx, f1, acc[o + 3]); acc[o + 4] = __fma_rn(jy, f1, acc[o + 4]); acc[o + 5] = __fma_rn(jz, f1, acc[o + 5]);
				}
			}
		}
		#pragma unroll
		for (int pass = 0; pass < 3; pass++) {
			#pragma unroll
			for (int r = 0; r < 8; r++) sRed[r * kRedRow + (tid >> 3) * 9 + (tid & 7)] = acc[pass * 8 + r];
			__syncthreads();
			{
				const int row = tid >> 4, seg = tid & 15;
				const double* r = sRed + row * kRedRow + seg * 9;
				double v = r[0];
				#pragma unroll
				for (int i = 1; i < 8; i++) v += r[i];
				v += __shfl_xor_sync(0xffffffffu, v, 1);
				v += __shfl_xor_sync(0xffffffffu, v, 2);
				v += __shfl_xor_sync(0xffffffffu, v, 4);
				v += __shfl_xor_sync(0xffffffffu, v, 8);
				if (seg == 0) sJ[pass * 8 + row] = v;
			}
			__syncthreads();
		}
		// a cell of several chunks: park the partial where the cell's last chunk will find it (visible before the status word)
		if (n > kChunk && tid < 24) { chunkJ[(int64_t)pos * 24 + tid] = sJ[tid]; __threadfence(); }
	}

	// ---- push + classify in place; per-block counts ---------------------------------------------
	for (int it = 0; it < kChunkIters; it++) {
		const int idx = it * kBlock + tid;
		unsigned code = kTagNone;
		if (idx < cnt) {
			double* p = sTile + (it * kWarps + warp) * BLOCK_DOUBLES + lane;
			double px = p[0], py = p[32], pz = p[64], vx = p[96], vy = p[128], vz = p[160];
			const double pm = NF == 8 ? p[224] : g.uqom;
			if (INTERP) {
				const double rx = (px - org[0]) / g.w[0], ry = (py - org[1]) / g.w[1], rz = (pz - org[2]) / g.w[2];
				const Vec3 E = trilinear(sCorner, 6, rx, ry, rz, g.vol);
				const Vec3 B = trilinear(sCorner + 3, 6, rx, ry, rz, g.vol);
				push_with_fields(g, px, py, pz, vx, vy, vz, pm, E, B);
			} else if (!(fieldFree && field_free_push(g, px, py, pz, vx, vy, vz, pm))) {
				dipole_push(g, plainDipole, px, py, pz, vx, vy, vz, pm);
			}
			code = MIGRATE ? classify(g, wallCell, cx, cy, cz, ctr, px, py, pz, vx, vy, vz) : 13u;
			p[0] = px; p[32] = py; p[64] = pz; p[96] = vx; p[128] = vy; p[160] = vz;
		}
		if (idx < kChunk) sCode[idx] = (unsigned char)code;
		const unsigned bs = __ballot_sync(0xffffffffu, code == 13u);
		const unsigned bl = __ballot_sync(0xffffffffu, code < 27u && code != 13u);
		if (lane == 0) sBlk[it * kWarps + warp] = __popc(bs) | (__popc(bl) << 16);
	}
	__syncthreads();

	// ---- scan over the chunk's blocks, then the cell-level prefix (look-back) ----------------------
	if (warp == 0) {
		unsigned totS = 0, totL = 0;
		#pragma unroll
		for (int b0 = 0; b0 < kChunkBlocks; b0 += 32) {
			const int v = (b0 + lane < kChunkBlocks) ? sBlk[b0 + lane] : 0;
			int s = v & 0xffff, l = v >> 16;
			int is = s, il = l;
			#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const int ts = __shfl_up_sync(0xffffffffu, is, d), tl = __shfl_up_sync(0xffffffffu, il, d);
				if (lane >= d) { is += ts; il += tl; }
			}
			if (b0 + lane < kChunkBlocks) { sOff[0][b0 + lane] = (int)totS + is - s; sOff[1][b0 + lane] = (int)totL + il - l; }
			totS += (unsigned)__shfl_sync(0xffffffffu, is, 31);
			totL += (unsigned)__shfl_sync(0xffffffffu, il, 31);
		}
		unsigned preS = 0, preL = 0;
		if (n > kChunk) {
			volatile unsigned long long* st = status;
			if (!last && lane == 0) st[pos] = pack_status(j == 0 ? kStatusPrefix : kStatusAggregate, totS, totL);
			int remaining = j, look = pos - 1;
			while (remaining > 0) {
				const int k = min(32, remaining);
				unsigned long long w;
				do { w = lane < k ? st[look - lane] : kStatusPrefix; } while (__any_sync(0xffffffffu, (w & kStatusMask) == 0));
				const unsigned pm = __ballot_sync(0xffffffffu, lane < k && (w & kStatusMask) == kStatusPrefix);
				const int stopLane = pm ? __ffs(pm) - 1 : k - 1;          // nearest predecessor that already knows its inclusive prefix
				const unsigned s = lane <= stopLane ? status_stay(w) : 0u, l = lane <= stopLane ? status_leave(w) : 0u;
				preS += __reduce_add_sync(0xffffffffu, s);
				preL += __reduce_add_sync(0xffffffffu, l);
				if (pm) break;
				look -= k; remaining -= k;
			}
			if (!last && j > 0 && lane == 0) st[pos] = pack_status(kStatusPrefix, preS + totS, preL + totL);
			__threadfence();                                             // the predecessors' deposit partials are read after their status words
		}
		if (lane == 0) {
			sPrefix[0] = (int)preS; sPrefix[1] = (int)preL;
			if (last) { countOut[cell] = (int)(preS + totS); nLeave[cell] = (int)(preL + totL); }
		}
	}
	__syncthreads();

	// ---- write-out from shared memory --------------------------------------------------------------
	const int preS = sPrefix[0], preL = sPrefix[1];
	for (int it = 0; it < kChunkIters; it++) {
		const int idx = it * kBlock + tid;
		const unsigned code = idx < kChunk ? sCode[idx] : kTagNone;
		const bool stay = code == 13u, leave = code < 27u && code != 13u;
		const unsigned bs = __ballot_sync(0xffffffffu, stay), bl = __ballot_sync(0xffffffffu, leave);
		if (stay || leave) {
			const int blk = it * kWarps + warp;
			int d;
			if (stay) d = preS + sOff[0][blk] + __popc(bs & lt);
			else {
				d = cap - 1 - (preL + sOff[1][blk] + __popc(bl & lt));
				tag[base + d] = (unsigned char)code;
			}
			const double* p = sTile + blk * BLOCK_DOUBLES + lane;
			double* q = dst + (int64_t)(d >> 5) * BLOCK_DOUBLES + (d & 31);
			st_stream(q, p[0]); st_stream(q + 32, p[32]); st_stream(q + 64, p[64]);
			st_stream(q + 96, p[96]); st_stream(q + 128, p[128]); st_stream(q + 160, p[160]);
			if (NF == 8) { st_stream(q + 192, p[192]); st_stream(q + 224, p[224]); }
		}
	}

	if (DEPOSIT) {
		if (n <= kChunk) {
			if (tid < 24) cellJ[cellj_index(g, cx, cy, cz) + tid] = sJ[tid];
		} else if (last) {
			// every earlier chunk has published a status word by now (look-back), hence its partial: add them in chunk
			// order; 5 strided partial sums per component, combined in fixed order
			__shared__ double sSum[5][24];
			const int grp = tid / 24, comp = tid - grp * 24;
			if (grp < 5) {
				double v = 0.0;
				for (int k = grp; k < j; k += 5) v += __ldcg(chunkJ + (int64_t)(pos - j + k) * 24 + comp);
				sSum[grp][comp] = v;
			}
			__syncthreads();
			if (tid < 24) {
				double v = sSum[0][tid];
				#pragma unroll
				for (int k = 1; k < 5; k++) v += sSum[k][tid];
				cellJ[cellj_index(g, cx, cy, cz) + tid] = v + sJ[tid];
			}
		}
	}
}

// ---- import: append the 26 incoming outboxes in canonical order (reference cell.h:1195-1310) ------
//
// One warp per target cell.  Lane d looks at the neighbour of direction d (lexicographic (0,0,0)..(2,2,2),
// 13 = self skipped): source cell = target + (a-1,b-1,c-1) with periodic address (transfer_buffer.h,
// cell.h:1012-1048); the entries of that neighbour's outbox tagged (2-a,2-b,2-c) = 26 - d are ours.
//
// The movers leave a DIRECTION-SORTED INDEX of every outbox behind (sort_outbox): dirOff[cell][t] = number of
// entries with a tag below t, order[cellStart + i] = the outbox entry that is i-th in (tag, outbox order).  Lane d
// therefore reads two integers to learn how many particles arrive from its neighbour and where they are -- no tag
// is scanned, and only particles that really arrive are touched.  An exclusive scan over the lanes puts the
// directions in canonical order, every lane lists its arrivals in a shared-memory window, and the warp copies 32
// arrivals per trip.  Neighbours whose outbox carries no index (more than kSortCap leavers, or written by the
// chunked mover) are scanned tag by tag, direction after direction.
//
// SLAB: the universe is split into z-slabs; a neighbour below the slab's first plane / above its last
// one lives on another GPU, and its outbox is the matching column of the face message received from
// that GPU (k_face_copy), where the entries of one direction are contiguous.

#ifndef IPIC_IMPORT_BLOCK
#define IPIC_IMPORT_BLOCK 256
#endif
#ifndef IPIC_IMPORT_MINB
#define IPIC_IMPORT_MINB 4       // resident CTAs per SM the import is compiled for
#endif
constexpr int kImportBlock = IPIC_IMPORT_BLOCK;
constexpr int kImportWarps = kImportBlock / 32;
constexpr int kListWindow = 256;           // arrivals listed per pass of a warp

// one arrival: fields of record `rec` of a face message (AoS, always 8 fields) or of slot `slot` of generation src
__device__ __forceinline__ void copy_arrival(bool remote, const double* __restrict__ msg, int64_t rec, const double* __restrict__ srcd, int64_t slot,
                                             int nf, double* __restrict__ t) {
	double v[kFields];
	if (remote) {
		const double2* s = reinterpret_cast<const double2*>(msg + rec * kFields);
		#pragma unroll
		for (int f = 0; f < kFields; f += 2) { const double2 x = s[f >> 1]; v[f] = x.x; v[f + 1] = x.y; }
	} else {
		const double* s = srcd + slot_index(slot, nf);
		#pragma unroll
		for (int f = 0; f < kFields; f++) if (f < nf) v[f] = s[f * kFieldStride];
	}
	#pragma unroll
	for (int f = 0; f < kFields; f++) if (f < nf) t[f * kFieldStride] = v[f];
}

template<bool SLAB>
__global__ void __launch_bounds__(kImportBlock, IPIC_IMPORT_MINB)
k_import(const Geometry g, const ParticleArrays src, const int64_t* __restrict__ srcStart, const int* __restrict__ srcLeave,
         const unsigned char* __restrict__ tag, const unsigned short* __restrict__ order, const int* __restrict__ dirOff,
         const ParticleArrays dst, const int64_t* __restrict__ dstStart,
         const int* __restrict__ stay, int* __restrict__ dstCount, int reserveOwnOutbox, int* __restrict__ stopFlag, int stepIndex, int zBegin, int zCount, int zStride,
         const FaceMsg fromDown, const FaceMsg fromUp, const BigOutboxes big) {

	__shared__ int sListAll[kImportWarps][kListWindow];       // arrivals of the current window: direction << 24 | index within the direction
	__shared__ int64_t sSrcA[kImportWarps][27];               // per direction: first order slot (local) / first record (remote)
	__shared__ int64_t sSrcB[kImportWarps][27];               // per direction: last slot of the neighbour's segment (local)
	if (*(volatile int*)stopFlag) return;
	const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
	const int task = blockIdx.x * kImportWarps + wp;
	if (task >= g.nx * g.ny * zCount) return;
	const int xy = task / zCount;
	const int zl = zBegin + (task - xy * zCount) * zStride;
	const int cx = xy / g.ny, cy = xy - cx * g.ny;
	const int cell = xy * g.lz + zl;

	const int64_t dbase = dstStart[cell];
	const int dcap = (int)(dstStart[cell + 1] - dbase);
	const int room = dcap - (reserveOwnOutbox ? srcLeave[cell] : 0);
	int w = stay[cell];

	auto isRemote = [&](int d) { const int c = d % 3; return SLAB && ((c == 0 && zl == 0) || (c == 2 && zl == g.lz - 1)); };

	// lane d: what arrives from its neighbour
	int cnt = 0, scell = 0, bigId = -1;
	int64_t A = 0, B = 0;
	bool unsorted = false, corrupt = false;
	if (lane < 27 && lane != 13) {
		const int a = lane / 9, b = (lane / 3) % 3, c = lane % 3;
		const int sx = (cx + a - 1 + g.nx) % g.nx, sy = (cy + b - 1 + g.ny) % g.ny;
		if (isRemote(lane)) {
			const FaceMsg& m = c == 0 ? fromDown : fromUp;
			const int* o = m.offsets + (sx * g.ny + sy) * 9 + (2 - a) * 3 + (2 - b);
			const int first = o[0];
			cnt = o[1] - first;
			A = first;
			if (first < 0 || cnt < 0 || first + cnt > m.cap) { corrupt = true; cnt = 0; }
		} else {
			const int szl = (zl + c - 1 + g.lz) % g.lz;
			scell = (sx * g.ny + sy) * g.lz + szl;
			const int* o = dirOff + (int64_t)scell * 28 + (26 - lane);
			const int o0 = o[0];
			B = srcStart[scell + 1] - 1;                  // outbox entry j lives at slot B - j
			if (o0 <= kBigMarker) {
				// a big outbox (k_big_index): only its size per direction is looked up here, k_big_copy moves the particles
				bigId = kBigMarker - o0;
				const int* bo = big.off + bigId * 28 + (26 - lane);
				cnt = bo[1] - bo[0];
			} else if (o0 < 0) unsorted = true;
			else { cnt = o[1] - o0; A = srcStart[scell] + o0; }
		}
	}
	if (SLAB && __any_sync(0xffffffffu, corrupt)) {
		// the sender ran out of message space (it reports that itself) or the message is damaged
		if (lane == 0) { atomicMin(stopFlag + 1, stepIndex); atomicExch(stopFlag + 3, 2); atomicExch(stopFlag, 1); }
		return;
	}

	if (!__any_sync(0xffffffffu, unsorted || bigId >= 0 || cnt >= (1 << 24))) {
		int incl = cnt;
		#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
		const int dirBase = incl - cnt;                      // lane d: place of its first arrival in the canonical order
		const int total = __shfl_sync(0xffffffffu, incl, 31);
		if (total == 0) {
			if (lane == 0) dstCount[cell] = w;
			return;
		}
		int* sList = sListAll[wp];
		if (lane < 27) { sSrcA[wp][lane] = A; sSrcB[wp][lane] = B; }
		for (int win0 = 0; win0 < total; win0 += kListWindow) {
			const int winEnd = min(total, win0 + kListWindow);
			__syncwarp();
			for (int p = max(dirBase, win0); p < min(dirBase + cnt, winEnd); p++) sList[p - win0] = (lane << 24) | (p - dirBase);
			__syncwarp();
			for (int i0 = win0; i0 < winEnd; i0 += 32) {
				const int i = i0 + lane;
				if (i < winEnd && w + i < room) {
					const int e = sList[i - win0];
					const int d = e >> 24, k = e & 0xffffff;
					const int64_t a0 = sSrcA[wp][d];
					double* t = dst.d + slot_index(dbase + w + i, g.nf);
					if (isRemote(d)) copy_arrival(true, (d % 3 == 0 ? fromDown : fromUp).particles, a0 + k, nullptr, 0, g.nf, t);
					else copy_arrival(false, nullptr, 0, src.d, sSrcB[wp][d] - (int64_t)order[a0 + k], g.nf, t);
				}
			}
		}
		w += total;
	} else {
		// some neighbour's outbox has no index: direction after direction; indexed / remote ranges are copied, the others scanned
		for (int d = 0; d < 27; d++) {
			if (d == 13) continue;
			const bool uns = __shfl_sync(0xffffffffu, (int)unsorted, d) != 0;
			const int64_t a0 = __shfl_sync(0xffffffffu, A, d), b0 = __shfl_sync(0xffffffffu, B, d);
			const int bid = __shfl_sync(0xffffffffu, bigId, d);
			if (bid >= 0) {
				// arrivals out of a big outbox: reserve their place in the canonical order and tell k_big_copy where it is
				if (lane == 0) { big.dst[bid * 27 + (26 - d)] = dbase + w; big.room[bid * 27 + (26 - d)] = dbase + room; }
				w += __shfl_sync(0xffffffffu, cnt, d);
				continue;
			}
			if (!uns) {
				const int n = __shfl_sync(0xffffffffu, cnt, d);
				const bool remote = isRemote(d);
				for (int k0 = 0; k0 < n; k0 += 32) {
					const int k = k0 + lane;
					if (k < n && w + k < room) {
						double* t = dst.d + slot_index(dbase + w + k, g.nf);
						if (remote) copy_arrival(true, (d % 3 == 0 ? fromDown : fromUp).particles, a0 + k, nullptr, 0, g.nf, t);
						else copy_arrival(false, nullptr, 0, src.d, b0 - (int64_t)order[a0 + k], g.nf, t);
					}
				}
				w += n;
				continue;
			}
			const int sc = __shfl_sync(0xffffffffu, scell, d);
			const int nl = srcLeave[sc];
			const unsigned want = (unsigned)(26 - d);
			for (int j0 = 0; j0 < nl; j0 += 32) {
				const int j = j0 + lane;
				const bool hit = j < nl && tag[b0 - j] == want;
				const unsigned m = __ballot_sync(0xffffffffu, hit);
				if (hit) {
					const int o = w + __popc(m & ((1u << lane) - 1u));
					if (o < room) copy_arrival(false, nullptr, 0, src.d, b0 - j, g.nf, dst.d + slot_index(dbase + o, g.nf));
				}
				w += __popc(m);
			}
		}
	}
	if (lane == 0) {
		if (w > room) atomicMin(stopFlag + 1, stepIndex), atomicExch(stopFlag + 2, 1);   // recorded; flag raised by k_commit
		dstCount[cell] = w;
	}
}

// ---- big outboxes: index and copy spread over the whole GPU ---------------------------------------------
//
// The reference's :E: / :B: loads pile most of the plasma into a few cells, and a cell that holds millions of particles can
// send millions of them to ONE neighbour in one step.  Such an outbox carries no direction index (it comes from the chunked
// mover, or holds more than kSortCap entries), so every target warp above would scan all of its tags and copy its share 32
// entries per trip -- one warp for millions of particles.  For the cells the host lists as big (segment capacity of tens of
// thousands of slots, capi.cu bigCellSlots(); universes on one GPU) the work is cut into segments of kBigSeg outbox
// entries, one warp each:
//   k_big_hist   per segment: how many entries carry each tag
//   k_big_scan   per big cell: exclusive scan of those counts over the segments, tag by tag; totals -> big.off (what the
//                import needs to place the arrivals of every direction) and dirOff[cell][*] = kBigMarker - id
//   k_import     sees the marker: counts the arrivals, reserves their slots in canonical order, copies nothing, and
//                leaves the first slot and the end of the usable room per (big cell, tag) in big.dst / big.room
//   k_big_copy   per segment: entry j with tag t goes to big.dst[t] + (entries with tag t in earlier segments) + (earlier
//                entries with tag t in this segment) -- the position the one-warp scan would have given it.

constexpr int kBigSeg = 1024;              // outbox entries per warp task
constexpr int kBigWarps = 8;

__global__ void __launch_bounds__(kBigWarps * 32)
k_big_hist(const BigOutboxes big, const int64_t* __restrict__ cellStart, const int* __restrict__ nLeave, const unsigned char* __restrict__ tag,
           const int* __restrict__ stopFlag) {
	__shared__ int sCntAll[kBigWarps][28];
	if (*stopFlag) return;
	const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
	const int task = blockIdx.x * kBigWarps + wp;
	if (task >= big.totalSegs) return;
	const int2 ref = big.segTable[task];
	const int id = ref.x, seg = ref.y;
	const int cell = big.cell[id];
	const int L = nLeave[cell];
	const int j0 = seg * kBigSeg;
	if (j0 >= L) return;                                  // beyond the outbox: the scan does not read this segment's counts
	const int jEnd = min(L, j0 + kBigSeg);
	const int64_t send = cellStart[cell + 1] - 1;         // outbox entry j lives at slot send - j
	int* sCnt = sCntAll[wp];
	const unsigned lt = (1u << lane) - 1u;
	if (lane < 27) sCnt[lane] = 0;                        // tags of outbox entries are 0..26 (never 13)
	__syncwarp();
	for (int jj = j0; jj < jEnd; jj += 32) {
		const int j = jj + lane;
		const unsigned t = j < jEnd ? tag[send - j] : 255u;
		const unsigned peers = __match_any_sync(0xffffffffu, t);
		if (t < 27u && (peers & lt) == 0u) sCnt[t] += __popc(peers);      // one lane per tag value: no conflicts
		__syncwarp();
	}
	const int nseg = big.segFirst[id + 1] - big.segFirst[id];
	if (lane < 27) big.rank[(int64_t)big.segFirst[id] * 27 + (int64_t)lane * nseg + seg] = sCnt[lane];
}

// one CTA of 27 warps per big cell: warp t scans the counts of tag t over the cell's segments
__global__ void __launch_bounds__(27 * 32)
k_big_scan(const BigOutboxes big, const int* __restrict__ nLeave, int* __restrict__ dirOff, const int* __restrict__ stopFlag) {
	__shared__ int sTot[32];
	if (*stopFlag) return;
	const int lane = threadIdx.x & 31, t = threadIdx.x >> 5;
	const int id = blockIdx.x;
	const int cell = big.cell[id];
	const int L = nLeave[cell];
	const int used = (L + kBigSeg - 1) / kBigSeg;
	const int nseg = big.segFirst[id + 1] - big.segFirst[id];
	int* r = big.rank + (int64_t)big.segFirst[id] * 27 + (int64_t)t * nseg;
	int running = 0;
	for (int s0 = 0; s0 < used; s0 += 32) {
		const int sidx = s0 + lane;
		const int v = sidx < used ? r[sidx] : 0;
		int incl = v;
		#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
		if (sidx < used) r[sidx] = running + incl - v;
		running += __shfl_sync(0xffffffffu, incl, 31);
	}
	if (lane == 0) sTot[t] = running;
	if (threadIdx.x >= 27 && threadIdx.x < 32) sTot[threadIdx.x] = 0;
	__syncthreads();
	if (t == 0) {
		const int v = lane < 27 ? sTot[lane] : 0;
		int incl = v;
		#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
		if (lane < 28) {
			big.off[id * 28 + lane] = incl - v;               // lane 27: v = 0, incl = L
			dirOff[(int64_t)cell * 28 + lane] = kBigMarker - id;
		}
	}
}

__global__ void __launch_bounds__(kBigWarps * 32)
k_big_copy(const BigOutboxes big, const Geometry g, const ParticleArrays src, const int64_t* __restrict__ srcStart, const int* __restrict__ nLeave,
           const unsigned char* __restrict__ tag, const ParticleArrays dst, const int* __restrict__ stopFlag) {
	__shared__ int sCntAll[kBigWarps][28];
	__shared__ long long sDstAll[kBigWarps][28], sRoomAll[kBigWarps][28];
	if (*stopFlag) return;
	const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
	const int task = blockIdx.x * kBigWarps + wp;
	if (task >= big.totalSegs) return;
	const int2 ref = big.segTable[task];
	const int id = ref.x, seg = ref.y;
	const int cell = big.cell[id];
	const int L = nLeave[cell];
	const int j0 = seg * kBigSeg;
	if (j0 >= L) return;
	const int jEnd = min(L, j0 + kBigSeg);
	const int64_t send = srcStart[cell + 1] - 1;
	const int nseg = big.segFirst[id + 1] - big.segFirst[id];
	int* sCnt = sCntAll[wp];
	long long* sDst = sDstAll[wp];
	long long* sRoom = sRoomAll[wp];
	const unsigned lt = (1u << lane) - 1u;
	if (lane < 27) {
		sCnt[lane] = big.rank[(int64_t)big.segFirst[id] * 27 + (int64_t)lane * nseg + seg];     // entries with this tag in earlier segments
		sDst[lane] = big.dst[id * 27 + lane];
		sRoom[lane] = big.room[id * 27 + lane];
	}
	__syncwarp();
	for (int jj = j0; jj < jEnd; jj += 32) {
		const int j = jj + lane;
		const unsigned t = j < jEnd ? tag[send - j] : 255u;
		const unsigned peers = __match_any_sync(0xffffffffu, t);
		if (t < 27u) {
			const long long slot = sDst[t] + sCnt[t] + __popc(peers & lt);
			if (slot < sRoom[t]) copy_arrival(false, nullptr, 0, src.d, send - j, g.nf, dst.d + slot_index(slot, g.nf));
		}
		__syncwarp();
		if (t < 27u && (peers & lt) == 0u) sCnt[t] += __popc(peers);
		__syncwarp();
	}
}

// raises the stop flag after an import pass that overflowed (kept separate so that every cell of the
// failing step still gets its import attempt recorded in dstCount == required size)
__global__ void k_commit_flag(int* stopFlag) {
	if (stopFlag[2]) stopFlag[0] = 1;
	stopFlag[4] = 0;                       // "some outbox needs k_sort_outbox": the movers of the next step raise it again
}


// ---- z-slab faces: pack the leavers that cross the slab boundary ------------------------------------
//
// Face 0 = the slab's lowest owned z-plane, leavers with direction k = 0 (to the neighbour below);
// face 1 = the highest owned plane, k = 2 (to the neighbour above).  A face message is
//   offsets[nx*ny*9 + 1]  exclusive prefix of the entries per ((x,y) column, (i,j) of the direction code); last = total
//   particles[cap][8]     AoS, columns in (x*ny + y) order, within a column by direction code, then outbox order
// i.e. for the receiver's import kernel a column is the outbox it would have found in the neighbour's cell (the
// reference's transfer buffers of that cell, transfer_buffer.h:15-95), restricted to the entries that can concern
// it and already sorted by direction: arrival order stays canonical across slab boundaries.

constexpr int kFaceBlock = 256;
constexpr int kFaceWarps = kFaceBlock / 32;

__global__ void __launch_bounds__(kFaceBlock)
k_face_count(const Geometry g, const int64_t* __restrict__ cellStart, const int* __restrict__ nLeave,
             const unsigned char* __restrict__ tag, const int* __restrict__ dirOff, int* __restrict__ faceCnt, const int* __restrict__ stopFlag) {
	const int lane = threadIdx.x & 31;
	const int nxy = g.nx * g.ny;
	const int task = blockIdx.x * kFaceWarps + (threadIdx.x >> 5);
	if (task >= 2 * nxy) return;
	int mine = 0;                                        // lane ab < 9: entries of direction code ab*3 + (face ? 2 : 0)
	if (!*stopFlag) {
		const int f = task / nxy, xy = task - f * nxy;
		const int cell = xy * g.lz + (f ? g.lz - 1 : 0);
		const int* off = dirOff + (int64_t)cell * 28;
		const unsigned want = f ? 2u : 0u;
		if (off[0] >= 0) {
			if (lane < 9) mine = off[lane * 3 + want + 1] - off[lane * 3 + want];
		} else {
			const int nl = nLeave[cell];
			const int64_t send = cellStart[cell + 1] - 1;
			for (int j0 = 0; j0 < nl; j0 += 32) {
				const int j = j0 + lane;
				const unsigned t = j < nl ? tag[send - j] : 255u;
				#pragma unroll
				for (int ab = 0; ab < 9; ab++) {
					const int c = __popc(__ballot_sync(0xffffffffu, t == ab * 3 + want));
					if (lane == ab) mine += c;
				}
			}
		}
	}
	if (lane < 9) faceCnt[task * 9 + lane] = mine;
}

// one CTA of 1024 threads per face: exclusive scan of the (column, direction) counts into the message header
__global__ void __launch_bounds__(1024)
k_face_scan(const Geometry g, const int* __restrict__ faceCnt, FaceMsg m0, FaceMsg m1, int* __restrict__ totals) {
	__shared__ int sWarp[32];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int n = g.nx * g.ny * 9;
	const FaceMsg m = blockIdx.x ? m1 : m0;
	const int* cnt = faceCnt + blockIdx.x * n;
	int carry = 0;
	for (int base = 0; base < n; base += 1024) {
		const int i = base + tid;
		const int v = i < n ? cnt[i] : 0;
		int incl = v;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
		if (lane == 31) sWarp[warp] = incl;
		__syncthreads();
		if (warp == 0) {
			int w = sWarp[lane];
			#pragma unroll
			for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, d); if (lane >= d) w += t; }
			sWarp[lane] = w;
		}
		__syncthreads();
		if (i < n) m.offsets[i] = carry + (warp ? sWarp[warp - 1] : 0) + incl - v;
		carry += sWarp[31];
		__syncthreads();
	}
	if (tid == 0) {
		m.offsets[n] = carry;
		totals[blockIdx.x] = carry;          // the host compares it with the buffer capacity and packs again into a bigger one if needed
	}
}

__global__ void __launch_bounds__(kFaceBlock)
k_face_copy(const Geometry g, const ParticleArrays src, const int64_t* __restrict__ cellStart, const int* __restrict__ nLeave,
            const unsigned char* __restrict__ tag, const unsigned short* __restrict__ order, const int* __restrict__ dirOff, FaceMsg m0, FaceMsg m1) {
	const int lane = threadIdx.x & 31;
	const int nxy = g.nx * g.ny;
	const int task = blockIdx.x * kFaceWarps + (threadIdx.x >> 5);
	if (task >= 2 * nxy) return;
	const int f = task / nxy, xy = task - f * nxy;
	const FaceMsg m = f ? m1 : m0;
	const int* mo = m.offsets + xy * 9;
	if (mo[9] - mo[0] <= 0) return;
	const int cell = xy * g.lz + (f ? g.lz - 1 : 0);
	const int64_t send = cellStart[cell + 1] - 1, start = cellStart[cell];
	const int* off = dirOff + (int64_t)cell * 28;
	const bool sorted = off[0] >= 0;
	const unsigned want = f ? 2u : 0u;
	auto emit = [&](int64_t slot, int o) {
		if (o >= m.cap) return;
		const double* s = src.d + slot_index(slot, g.nf);
		double v[kFields];
		v[6] = g.uq; v[7] = g.uqom;
		#pragma unroll
		for (int k = 0; k < kFields; k++) if (k < g.nf) v[k] = s[k * kFieldStride];
		double2* d = reinterpret_cast<double2*>(m.particles + (int64_t)o * kFields);
		#pragma unroll
		for (int k = 0; k < kFields; k += 2) d[k >> 1] = make_double2(v[k], v[k + 1]);
	};
	for (int ab = 0; ab < 9; ab++) {
		const int first = mo[ab], total = mo[ab + 1] - first;
		if (total <= 0) continue;
		const unsigned code = ab * 3 + want;
		if (sorted) {
			const int o0 = off[code];
			for (int k = lane; k < total; k += 32) emit(send - (int64_t)order[start + o0 + k], first + k);
		} else {
			const int nl = nLeave[cell];
			int w = 0;
			for (int j0 = 0; j0 < nl; j0 += 32) {
				const int j = j0 + lane;
				const bool hit = j < nl && tag[send - j] == code;
				const unsigned mask = __ballot_sync(0xffffffffu, hit);
				if (hit) emit(send - j, first + w + __popc(mask & ((1u << lane) - 1u)));
				w += __popc(mask);
			}
		}
	}
}

// ---- node-centric gather of the per-cell corner sums (reference cell.h:787-824) -------------------

// node planes [zNode0, zNode0 + zNodes) of the global (nz + 1)-plane node grid
__global__ void k_gather_current(const Geometry g, const double* __restrict__ cellJ, double* __restrict__ J, const int* __restrict__ stopFlag,
                                 int zNode0, int zNodes) {
	if (*stopFlag) return;
	const int64_t nn = (int64_t)(g.nx + 1) * (g.ny + 1) * zNodes;
	const int64_t tt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (tt >= nn * 3) return;
	const int comp = (int)(tt % 3);
	const int64_t sub = tt / 3;
	const int z = zNode0 + (int)(sub % zNodes), y = (int)((sub / zNodes) % (g.ny + 1)), x = (int)(sub / ((int64_t)zNodes * (g.ny + 1)));
	const int64_t t = ((((int64_t)x * (g.ny + 1) + y) * (g.nz + 1)) + z) * 3 + comp;
	double Js = 0.0;
	for (int i = -1; i < 1; i++) for (int j = -1; j < 1; j++) for (int k = -1; k < 1; k++) {
		int ux = x + i, uy = y + j, uz = z + k;
		ux = ux < 0 ? g.nx - 1 : (ux >= g.nx ? 0 : ux);
		uy = uy < 0 ? g.ny - 1 : (uy >= g.ny ? 0 : uy);
		uz = uz < 0 ? g.nz - 1 : (uz >= g.nz ? 0 : uz);
		Js += cellJ[cellj_index(g, ux, uy, uz) + (((-i) * 2 + (-j)) * 2 + (-k)) * 3 + comp];   // cellJ is indexed by GLOBAL cell id
	}
	J[t] = (Js / g.vol) / 8.0;
}

// ---- invariant check (reference cell.h:48-59, :1168-1184) ----------------------------------------

__global__ void k_count_misplaced(const Geometry g, const ParticleArrays p, const int64_t* __restrict__ cellStart,
                                  const int* __restrict__ count, unsigned long long* __restrict__ bad) {
	const int xyz = blockIdx.x;
	const int zl = xyz % g.lz, xy = xyz / g.lz;
	const int cx = xy / g.ny, cy = xy - cx * g.ny, cz = g.z0 + zl;
	const int cell = xy * g.lz + zl;
	const double c0 = g.origin[0] + (double)cx * g.w[0] + g.w[0] / 2.0;
	const double c1 = g.origin[1] + (double)cy * g.w[1] + g.w[1] / 2.0;
	const double c2 = g.origin[2] + (double)cz * g.w[2] + g.w[2] / 2.0;
	const int64_t base = cellStart[cell];
	unsigned long long local = 0;
	for (int i = threadIdx.x; i < count[cell]; i += blockDim.x) {
		const double* q = p.d + slot_index(base + i, g.nf);
		const bool ok = fabs(q[0] - c0) <= g.hw[0] && fabs(q[kFieldStride] - c1) <= g.hw[1] && fabs(q[2 * kFieldStride] - c2) <= g.hw[2];
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
	const int cell = xy * g.lz + zl;
	const int64_t a0 = aosStart[oc], n = aosStart[oc + 1] - a0, base = cellStart[cell];
	for (int64_t e = threadIdx.x; e < n * 8; e += blockDim.x)
		if ((int)(e & 7) < g.nf) p.d[slot_index(base + (e >> 3), g.nf) + (e & 7) * kFieldStride] = aos[a0 * 8 + e];
}

__global__ void k_gather_aos(const Geometry g, double* __restrict__ aos, const int64_t* __restrict__ aosStart,
                             const ParticleArrays p, const int64_t* __restrict__ cellStart) {
	const int oc = blockIdx.x;
	const int zl = oc % g.lz, xy = oc / g.lz;
	const int cell = xy * g.lz + zl;
	const int64_t a0 = aosStart[oc], n = aosStart[oc + 1] - a0, base = cellStart[cell];
	for (int64_t e = threadIdx.x; e < n * 8; e += blockDim.x) {
		const int f = (int)(e & 7);
		aos[a0 * 8 + e] = f < g.nf ? p.d[slot_index(base + (e >> 3), g.nf) + f * kFieldStride] : (f == 6 ? g.uq : g.uqom);
	}
}

// copy the live head of every owned cell from one layout to another (relayout)
__global__ void k_copy_heads(const Geometry g, const ParticleArrays src, const int64_t* __restrict__ srcStart,
                             const ParticleArrays dst, const int64_t* __restrict__ dstStart, const int* __restrict__ count,
                             int srcNf, double srcQ, double srcQom) {
	const int oc = blockIdx.x;
	const int zl = oc % g.lz, xy = oc / g.lz;
	const int cell = xy * g.lz + zl;
	const int64_t s0 = srcStart[cell], d0 = dstStart[cell];
	const int n = count[cell];
	// both segment starts are multiples of 32 slots: whole AoSoA blocks line up, copy them double by double
	// (srcNf == g.nf), or field by field when the copy also widens the layout from 6 to 8 fields
	if (srcNf == g.nf) {
		const int64_t doubles = ((int64_t)(n + 31) >> 5) * (32 * g.nf);
		const double* s = src.d + s0 * g.nf;
		double* d = dst.d + d0 * g.nf;
		for (int64_t i = threadIdx.x; i < doubles; i += blockDim.x) d[i] = s[i];
	} else {
		for (int64_t e = threadIdx.x; e < (int64_t)n * 8; e += blockDim.x) {
			const int f = (int)(e & 7);
			const int64_t slot = e >> 3;
			dst.d[slot_index(d0 + slot, g.nf) + f * kFieldStride] = f < srcNf ? src.d[slot_index(s0 + slot, srcNf) + f * kFieldStride] : (f == 6 ? srcQ : srcQom);
		}
	}
}

// ---- device self-test of the arithmetic helpers ---------------------------------------------------

// counts inputs for which div3_ge1 differs (bitwise) from IEEE division; inputs are pseudo-random with
// exponents spread over the range the Boris rotation can produce, and exact zeros mixed in
__global__ void k_selftest_division(uint64_t n, uint64_t seed, unsigned long long* __restrict__ bad) {
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint64_t st = seed + i * 0x9E3779B97F4A7C15ull;
	auto next = [&]() { st ^= st >> 30; st *= 0xBF58476D1CE4E5B9ull; st ^= st >> 27; st *= 0x94D049BB133111EBull; st ^= st >> 31; return st; };
	auto value = [&]() {
		const uint64_t r = next();
		if ((r & 15) == 0) return (r & 16) ? -0.0 : 0.0;
		const double m = 1.0 + (double)(next() >> 11) * (1.0 / 9007199254740992.0);
		const int e = (int)((r >> 8) % 121) - 60;
		return ((r & 32) ? -m : m) * exp2((double)e);
	};
	const double tx = value(), ty = value(), tz = value();
	double t2 = 0.0 + tx * tx; t2 += ty * ty; t2 += tz * tz;
	const double den = 1.0 + t2;
	double x = 2.0 * tx, y = 2.0 * ty, z = 2.0 * tz;
	const double wx = x / den, wy = y / den, wz = z / den;
	div3_ge1(x, y, z, den);
	if (__double_as_longlong(x) != __double_as_longlong(wx) || __double_as_longlong(y) != __double_as_longlong(wy) ||
	    __double_as_longlong(z) != __double_as_longlong(wz)) atomicAdd(bad, 1ull);
	// the quotient magTemp / |p|^5 of dipole_push: any numerator and denominator with magnitudes in (1e-100, 1e100)
	auto wide = [&]() {
		const uint64_t r = next();
		const double m = 1.0 + (double)(next() >> 11) * (1.0 / 9007199254740992.0);
		const int e = (int)((r >> 8) % 661) - 330;
		return ((r & 32) ? -m : m) * exp2((double)e);
	};
	const double num = wide(), dd = fabs(wide());
	const double want = num / dd, got = div_by_rcp(num, dd, rcp_newton(dd));
	if (__double_as_longlong(want) != __double_as_longlong(got)) atomicAdd(bad, 1ull);
}

void launch_selftest_division(uint64_t n, uint64_t seed, unsigned long long* bad, cudaStream_t st) {
	k_selftest_division<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, seed, bad);
}

// ---- launchers -----------------------------------------------------------------------------------

// dynamic shared memory of one CTA: the tile stages, and at least the deposit reduction buffer
static size_t move_smem_bytes(int nf, int stages) {
	return (size_t)stages * kBlock * nf * sizeof(double);
}

// Opt-in to more than 48 KiB of dynamic shared memory.  The attribute belongs to the (kernel, device) pair, so it is
// set once per device the kernel is launched on (in-process groups drive several devices from one process); `mask`
// is the per-instantiation set of devices already configured.  A failure stays in cudaGetLastError for the caller's
// check after the launch.
template<typename Kernel>
static void optin_dynamic_smem(Kernel kernel, size_t bytes, unsigned long long& mask) {
	int dev = 0;
	if (cudaGetDevice(&dev) != cudaSuccess) return;
	const unsigned long long bit = 1ull << (dev & 63);
	if (mask & bit) return;
	if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) == cudaSuccess) mask |= bit;
}

template<int NF, bool INTERP, bool DEPOSIT, bool MIGRATE, bool PUSH>
static void launch_move_k(const MoveArgs& a, int grid, cudaStream_t st) {
	static unsigned long long configured = 0;        // devices on which this instantiation has its shared-memory opt-in
	auto kernel = k_move<NF, INTERP, DEPOSIT, MIGRATE, PUSH>;
	optin_dynamic_smem(kernel, move_smem_bytes(NF, kMaxStages), configured);
	const int stages = std::max(1, std::min(a.stages, kMaxStages));
	kernel<<<grid, kBlock, move_smem_bytes(NF, stages), st>>>(a.g, a.in, a.out, a.cellStart, a.countIn, a.countOut, a.nLeave, a.tag, a.order, a.dirOff,
	                                                          a.fieldEB, a.cellJ, a.stopFlag, const_cast<int*>(a.stopFlag) + 4, a.zBegin, a.zCount, a.zStride > 0 ? a.zStride : 1, stages);
}

template<int NF, bool MIGRATE>
static void launch_move_ws(const MoveArgs& a, int grid, cudaStream_t st) {
	static unsigned long long configured = 0;
	auto kernel = k_move_ws<NF, MIGRATE>;
	optin_dynamic_smem(kernel, move_smem_bytes(NF, kMaxStages), configured);
	const int stages = std::max(1, std::min(a.stages, kMaxStages));
	kernel<<<grid, kBlock + 32, move_smem_bytes(NF, stages), st>>>(a.g, a.in, a.out, a.cellStart, a.countIn, a.countOut, a.nLeave, a.tag, a.order, a.dirOff,
	                                                               a.cellJ, a.stopFlag, const_cast<int*>(a.stopFlag) + 4, a.zBegin, a.zCount, a.zStride > 0 ? a.zStride : 1, stages);
}

static bool use_warp_specialised() {
	static int v = -1;
	if (v < 0) { const char* e = std::getenv("IPIC_WS"); v = e ? (std::atoi(e) != 0) : 1; }
	return v != 0;
}
// tuning knob: the warp-specialised kernel for the analytic dipole as well (IPIC_WS_DIPOLE=1)
static bool ws_for_dipole() {
	static int v = -1;
	if (v < 0) { const char* e = std::getenv("IPIC_WS_DIPOLE"); v = e ? (std::atoi(e) != 0) : 0; }
	return v != 0;
}

// the deposit warp pays off when the push is light (field-free plasma: the push warps are integer / issue bound); with the
// analytic dipole every push carries a pow and ~150 FP64 operations, and the four-warp kernel's 5 CTAs/SM of pushers win
bool move_uses_ws(const Geometry& g, bool deposit, bool interpolate, bool push) {
	return deposit && !interpolate && push && (g.magTemp == 0.0 || ws_for_dipole()) && use_warp_specialised();
}

template<int NF, bool INTERP, bool DEPOSIT>
static void launch_move_t(const MoveArgs& a, cudaStream_t st) {
	const int grid = a.g.nx * a.g.ny * a.zCount;
	if (grid == 0) return;
	if (move_uses_ws(a.g, DEPOSIT, INTERP, a.push)) {
		if (a.migrate) launch_move_ws<NF, true>(a, grid, st); else launch_move_ws<NF, false>(a, grid, st);
		return;
	}
	if (!a.push) { if (DEPOSIT) launch_move_k<NF, false, DEPOSIT, false, false>(a, grid, st); }
	else if (a.migrate) launch_move_k<NF, INTERP, DEPOSIT, true, true>(a, grid, st);
	else launch_move_k<NF, INTERP, DEPOSIT, false, true>(a, grid, st);
}

template<int NF>
static void launch_move_i(const MoveArgs& a, cudaStream_t st) {
	if (a.interpolate) { if (a.deposit) launch_move_t<NF, true, true>(a, st); else launch_move_t<NF, true, false>(a, st); }
	else { if (a.deposit) launch_move_t<NF, false, true>(a, st); else launch_move_t<NF, false, false>(a, st); }
}

template<int NF, bool INTERP, bool DEPOSIT, bool MIGRATE>
static void launch_chunk_k(const MoveArgs& a, cudaStream_t st) {
	static unsigned long long configured = 0;
	auto kernel = k_move_chunk<NF, INTERP, DEPOSIT, MIGRATE>;
	const size_t smem = (size_t)kChunk * NF * sizeof(double);
	optin_dynamic_smem(kernel, smem, configured);
	cudaMemsetAsync(a.chunkStatus + a.chunkFirst, 0, (size_t)a.chunkCount * sizeof(unsigned long long), st);
	cudaMemsetAsync(a.chunkTicket, 0, sizeof(unsigned), st);
	kernel<<<a.chunkCount, kBlock, smem, st>>>(a.g, a.in, a.out, a.cellStart, a.countIn, a.countOut, a.nLeave, a.tag, a.fieldEB, a.cellJ, a.chunkJ,
	                                          a.stopFlag, a.chunkTable, a.chunkFirst, a.chunkTicket, a.chunkStatus, a.dirOff);
}

template<int NF>
static void launch_chunk_i(const MoveArgs& a, cudaStream_t st) {
	if (a.interpolate) {
		if (a.deposit) { if (a.migrate) launch_chunk_k<NF, true, true, true>(a, st); else launch_chunk_k<NF, true, true, false>(a, st); }
		else { if (a.migrate) launch_chunk_k<NF, true, false, true>(a, st); else launch_chunk_k<NF, true, false, false>(a, st); }
	} else {
		if (a.deposit) { if (a.migrate) launch_chunk_k<NF, false, true, true>(a, st); else launch_chunk_k<NF, false, true, false>(a, st); }
		else { if (a.migrate) launch_chunk_k<NF, false, false, true>(a, st); else launch_chunk_k<NF, false, false, false>(a, st); }
	}
}

int chunk_particles() { return kChunk; }

int launch_move(const MoveArgs& a, cudaStream_t st) {
	if (a.chunked && a.push) {
		if (a.chunkCount <= 0) return 0;
		if (a.g.nf == 6) launch_chunk_i<6>(a, st);
		else launch_chunk_i<8>(a, st);
		return 1;
	}
	if (a.g.nf == 6) launch_move_i<6>(a, st);
	else launch_move_i<8>(a, st);
	if (!(a.push && a.migrate && a.sortPass)) return 1;
	// outboxes of more than 32 entries were only marked: sort them now (returns at once when the movers marked none).  Without
	// this pass the import still works: it scans the tags of marked outboxes like those of the chunked mover.
	const int tasks = a.g.nx * a.g.ny * a.zCount;
	if (tasks == 0) return 1;
	k_sort_outbox<<<std::min((tasks + 7) / 8, 148 * 8), 256, 0, st>>>(a.g, a.cellStart, a.nLeave, a.tag, a.order, a.dirOff, a.stopFlag, a.stopFlag + 4,
	                                              a.zBegin, a.zCount, a.zStride > 0 ? a.zStride : 1);
	return 2;
}

int big_outbox_segment() { return kBigSeg; }

// index of the big outboxes the movers of this step left behind (after the movers, before the import)
int launch_big_index(const BigOutboxes& big, const int64_t* cellStart, const int* nLeave, const unsigned char* tag, int* dirOff, const int* stopFlag, cudaStream_t st) {
	if (big.nBig <= 0) return 0;
	cudaMemsetAsync(big.dst, 0, (size_t)big.nBig * 27 * sizeof(long long), st);      // a (cell, tag) nobody imports copies nothing: room 0
	cudaMemsetAsync(big.room, 0, (size_t)big.nBig * 27 * sizeof(long long), st);
	k_big_hist<<<(big.totalSegs + kBigWarps - 1) / kBigWarps, kBigWarps * 32, 0, st>>>(big, cellStart, nLeave, tag, stopFlag);
	k_big_scan<<<big.nBig, 27 * 32, 0, st>>>(big, nLeave, dirOff, stopFlag);
	return 2;
}

// the import proper, then the particles of the big outboxes, then (commit) the stop flag of an overflowed pass
int launch_import(const ImportArgs& a, bool commit, cudaStream_t st) {
	const int tasks = a.g.nx * a.g.ny * a.zCount;
	if (tasks == 0) return 0;
	const int grid = (tasks + kImportWarps - 1) / kImportWarps;
	int launched = 1;
	if (a.slab) k_import<true><<<grid, kImportBlock, 0, st>>>(a.g, a.src, a.srcStart, a.srcLeave, a.tag, a.order, a.dirOff, a.dst, a.dstStart, a.stay, a.dstCount,
	                                                        a.reserveOwnOutbox, a.stopFlag, a.stepIndex, a.zBegin, a.zCount, a.zStride > 0 ? a.zStride : 1, a.fromDown, a.fromUp, a.big);
	else k_import<false><<<grid, kImportBlock, 0, st>>>(a.g, a.src, a.srcStart, a.srcLeave, a.tag, a.order, a.dirOff, a.dst, a.dstStart, a.stay, a.dstCount,
	                                                    a.reserveOwnOutbox, a.stopFlag, a.stepIndex, a.zBegin, a.zCount, a.zStride > 0 ? a.zStride : 1, FaceMsg{}, FaceMsg{}, a.big);
	if (a.big.nBig > 0) {
		k_big_copy<<<(a.big.totalSegs + kBigWarps - 1) / kBigWarps, kBigWarps * 32, 0, st>>>(a.big, a.g, a.src, a.srcStart, a.srcLeave, a.tag, a.dst, a.stopFlag);
		launched++;
	}
	if (commit) { k_commit_flag<<<1, 1, 0, st>>>(a.stopFlag); launched++; }
	return launched;
}

void launch_face_pack(const Geometry& g, const ParticleArrays& src, const int64_t* cellStart, const int* nLeave, const unsigned char* tag,
                      const unsigned short* order, const int* dirOff,
                      int* faceCnt, const FaceMsg& down, const FaceMsg& up, int* totals, int* stopFlag, int stepIndex, cudaStream_t st) {
	(void)stepIndex;
	const int tasks = 2 * g.nx * g.ny;
	const int grid = (tasks + kFaceWarps - 1) / kFaceWarps;
	k_face_count<<<grid, kFaceBlock, 0, st>>>(g, cellStart, nLeave, tag, dirOff, faceCnt, stopFlag);
	k_face_scan<<<2, 1024, 0, st>>>(g, faceCnt, down, up, totals);
	k_face_copy<<<grid, kFaceBlock, 0, st>>>(g, src, cellStart, nLeave, tag, order, dirOff, down, up);
}

void launch_gather_current(const Geometry& g, const double* cellJ, double* J, const int* stopFlag, cudaStream_t st, int zNode0, int zNodes) {
	if (zNodes < 0) { zNode0 = 0; zNodes = g.nz + 1; }
	const int64_t n = (int64_t)(g.nx + 1) * (g.ny + 1) * zNodes * 3;
	k_gather_current<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g, cellJ, J, stopFlag, zNode0, zNodes);
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

void launch_copy_heads(const Geometry& g, const ParticleArrays& src, const int64_t* srcStart, const ParticleArrays& dst, const int64_t* dstStart, const int* count,
                       int srcNf, double srcQ, double srcQom, cudaStream_t st) {
	k_copy_heads<<<g.nx * g.ny * g.lz, 256, 0, st>>>(g, src, srcStart, dst, dstStart, count, srcNf, srcQ, srcQom);
}

// ---- uniform-species test on the staged AoS upload: flag != 0 if any particle's (q, qom) bits differ from the first one's
__global__ void k_check_species(const double* __restrict__ aos, int64_t n, int* __restrict__ flag) {
	const long long q0 = __double_as_longlong(aos[6]), m0 = __double_as_longlong(aos[7]);
	bool differs = false;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
		const double2 v = *reinterpret_cast<const double2*>(aos + i * 8 + 6);
		differs = differs || __double_as_longlong(v.x) != q0 || __double_as_longlong(v.y) != m0;
	}
	if (differs) atomicOr(flag, 1);
}

void launch_check_species(const double* aos, int64_t n, int* flag, cudaStream_t st) {
	if (n <= 0) return;
	const int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
	k_check_species<<<grid, 256, 0, st>>>(aos, n, flag);
}

} // namespace ipic
