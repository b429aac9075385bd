// C ABI (include/ipic3d_b200.h) over the sm_100a kernels: context, HBM layout, step orchestration.
//
// What simulateSteps keeps in locals (TransferBuffers, scratch; reference simulator.h:153-156) lives
// in ipic_ctx.  The loop body of simulateSteps (simulator.h:181-230) is ipic_step.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <limits>
#include <vector>

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>          // header-only; ranges cost nothing unless a profiler is attached

#include "../../include/ipic3d_b200.h"
#include "comm.h"
#include "diag_kernels.h"
#include "field_kernels.h"
#include "init_kernels.h"
#include "ipic_device.cuh"
#include "particle_kernels.h"
#include "tracer_kernels.h"

using namespace ipic;

namespace {

thread_local std::string g_lastError;

int fail(int code, const char* fmt, ...) {
	char buf[512];
	va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
	g_lastError = buf;
	return code;
}

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(IPIC_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

inline int64_t alignUp(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// NVTX range over a scope: the C ABI entry points and the phases of a slab step show up by name in Nsight timelines
struct NvtxRange {
	explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
	~NvtxRange() { nvtxRangePop(); }
};

} // namespace

struct ipic_group;

struct ipic_ctx {
	int device = 0, rank = 0, nranks = 1;
	int zBegin = 0, zCount = 0;                    // the slab's z-planes [zBegin, zBegin + zCount) of the global grid
	std::vector<int> slabZ0, slabLz;               // every rank's planes (known after the negotiation before the first step)
	ipic_props props{};
	Geometry g{};
	cudaStream_t stream = nullptr;

	// particle storage: two generations with a common layout
	int64_t slots = 0;
	int slotsNf = 0;                               // fields per particle the buffers were allocated for
	bool allowUniformSpecies = true;               // ipic_set_uniform_species
	ParticleArrays buf[2]{};
	int cur = 0;                                   // generation holding the live particles
	std::vector<int64_t> hCellStart;               // ncellsLocal + 1
	int64_t* dCellStart = nullptr;
	int* dCount = nullptr;                         // live particles per local cell (head of segment, generation `cur`)
	int* dStay = nullptr;                          // stayers per cell after the mover (head of segment, other generation)
	int* dLeave = nullptr;                         // outbox entries per local cell
	unsigned char* dTag = nullptr;                 // direction tag per slot (outbox entries only)
	unsigned short* dOrder = nullptr;              // per slot: the movers' direction-sorted index of each cell's outbox (sort_outbox)
	ZigguratTables* dZiggurat = nullptr;           // tables of the deck generator (ipic_generate_normal_speed), made on first use
	int* dDirOff = nullptr;                        // per local cell: 28 prefix counts over the direction codes of its outbox (-1: not indexed)
	// chunked mover (k_move_chunk): chunk table of the current layout, ordered plane 0 | interior planes | last plane
	bool chunked = false;                          // chosen at upload / re-layout from the populations (IPIC_CHUNKED=0/1 forces)
	int2* dChunkTable = nullptr;
	unsigned long long* dChunkStatus = nullptr;
	unsigned* dChunkTicket = nullptr;              // [0] plane 0, [1] interior, [2] last plane, [3] whole slab
	double* dChunkJ = nullptr;
	int chunkRange[4] = { 0, 0, 0, 0 };            // table offsets: plane 0 = [r0, r1), interior = [r1, r2), last plane = [r2, r3)
	int64_t chunkCapacity = 0;                     // chunks the device arrays were allocated for
	// big-outbox index (k_big_hist / k_big_scan / k_big_copy): cells of a single-GPU universe whose segment holds at least
	// bigCellSlots() slots, rebuilt with the chunk table; nBig = 0: none
	BigOutboxes big{};
	int64_t bigCellCapacity = 0, bigSegCapacity = 0;   // what the device arrays behind `big` were allocated for
	int* dStop = nullptr;                          // 8 control words: [0] stop, [1] first failing step, [2] cell overflow seen in this import pass, [3] exchange failure (1 face message full on this rank, 2 unusable message received), [4] an outbox awaits k_sort_outbox
	bool haveCells = false;
	int64_t meanPopulation = 0;                    // particles per owned cell at upload (drives the mover's tile shape)
	double slack = 0.25;                           // head-room per cell as a fraction of its population
	int64_t minPad = 64;                           // ... and in absolute slots
	double neighbourShare = 0.0;                   // adaptive head-room per cell: this share of the fullest of its 26 neighbours (0 = off)
	int64_t padFloor = 0;                          // adaptive absolute head-room: a share of the MEAN population, for universes whose sparse
	                                               // cells fill up (dipole runs start with 12 particles in the outer cells); doubles per re-layout
	bool userHeadroom = false;                     // ipic_set_headroom was called: no automatic floor at upload

	// fields (always the global grids; replicated per slab)
	double* dEB = nullptr;                         // (n+3)^3 x 6
	double* dBext = nullptr;                       // (n+3)^3 x 3
	double* dBc = nullptr;                         // (n+2)^3 x 3
	double* dJ = nullptr;                          // (n+1)^3 x 3
	double* dCellJ = nullptr;                      // global cells x 24
	int64_t nField = 0, nBc = 0, nJ = 0;

	// staging for AoS transfers
	double* dStage = nullptr; int64_t stageDoubles = 0;
	int64_t* dAosStart = nullptr;

	// measurement
	bool profiling = false;
	std::vector<cudaEvent_t> events;               // pairs (start, end) around the mover / import launches
	std::vector<int> eventKind;                    // per pair: 0 mover, 1 import
	size_t eventsUsed = 0;
	double moverMs = 0, importMs = 0;
	double phaseMs[4] = { 0, 0, 0, 0 };             // slab steps: [0] movers (span), [1] imports, [2] waiting for the exchange, [3] waiting for J + field phase
	uint64_t launches = 0;
	uint64_t relayouts = 0;                        // cell-segment overflows repaired so far
	// k_sort_outbox follows the movers only while outboxes of more than 32 entries occur: unknown after an upload (assume yes),
	// then whatever the movers reported (dStop[5], sticky) when the host last looked -- a wrong guess costs speed, never results
	bool sortLikely = true;
	// batch enqueued by ipic_step_async and not yet settled: generation at its start, steps, flags
	uint64_t asyncSteps = 0;
	int asyncGenAtStart = 0;
	uint32_t asyncFlags = 0;

	// z-slab exchange (nranks > 1): NCCL communicator (one process per GPU) or an in-process group
	ipic_comm* comm = nullptr;
	std::shared_ptr<ipic_group> group;
	cudaStream_t commStream = nullptr;
	cudaStream_t bStream = nullptr;                // high priority: movers of the boundary planes + face packing
	cudaEvent_t evPacked = nullptr, evRecv = nullptr, evMoved = nullptr, evJ = nullptr, evMainReady = nullptr;
	int64_t userMsgCap = 0;                        // ipic_set_exchange_capacity (0 = automatic)
	int64_t facePopulation = 0;                    // particles in the fuller of the two boundary planes at upload
	bool exchangeDirty = true;                     // message capacity must be (re)negotiated before the next step
	int sendCap = 0, recvCap = 0;                  // capacity of the face message buffers, in particles (grown on demand)
	int* dTotals = nullptr;                        // [0],[1] entries packed into my down / up face; [2],[3] entries coming from below / above
	int* hTotals = nullptr;                        // pinned mirror
	cudaEvent_t moverSpanEnd = nullptr;            // profiling: pending end of the movers' span
	char* sendMsg[2] = { nullptr, nullptr };       // [0] face 0 -> slab below, [1] face 1 -> slab above
	char* recvMsg[2] = { nullptr, nullptr };       // [0] from the slab below, [1] from the slab above
	int* dFaceCnt = nullptr;
	int64_t* dNegotiate = nullptr;                 // nranks x SlabOffer scratch
	bool emptyUpload = true;                       // the last upload held no particles
	int* hStop = nullptr;                          // pinned mirror of dStop, refreshed after every slab step
	cudaEvent_t evStepDone = nullptr;
};

// contexts of one process that form a complete set of slabs and exchange by peer copies
struct ipic_group { std::vector<ipic_ctx*> members; };

namespace {

int ownedCells(const ipic_ctx* c) { return c->g.nx * c->g.ny * c->g.lz; }

void freeParticleBuffers(ipic_ctx* c) {
	for (int b = 0; b < 2; b++) { if (c->buf[b].d) cudaFree(c->buf[b].d); c->buf[b].d = nullptr; }
	if (c->dTag) cudaFree(c->dTag);
	c->dTag = nullptr;
	if (c->dOrder) cudaFree(c->dOrder);
	c->dOrder = nullptr;
	c->slots = 0;
	c->slotsNf = 0;
}

int allocParticleBuffers(ParticleArrays out[2], unsigned char** tag, unsigned short** order, int64_t slots, int nf) {
	// slots is a multiple of 32: whole AoSoA-32 blocks of 32*nf doubles
	for (int b = 0; b < 2; b++) CK(cudaMalloc(&out[b].d, (size_t)std::max<int64_t>(slots, 32) * nf * sizeof(double)));
	CK(cudaMalloc(tag, (size_t)std::max<int64_t>(slots, 1)));
	CK(cudaMalloc(order, (size_t)std::max<int64_t>(slots, 1) * sizeof(unsigned short)));
	return IPIC_OK;
}

// capacity of a cell segment that currently needs room for n particles (localFloor: what its surroundings suggest, below)
int64_t capacityFor(const ipic_ctx* c, int64_t n, int64_t localFloor = 0) {
	const int64_t pad = std::max<int64_t>(std::max<int64_t>((int64_t)std::ceil((double)n * c->slack), c->minPad), std::max(c->padFloor, localFloor));
	return alignUp(n + pad, kSlotAlign);
}

// segment layout of the slab's cells: capacity for the current population plus head-room.  With neighbourShare > 0 (set for
// universes that hold sparse cells next to full ones) a cell also gets room for that share of its fullest neighbour's
// population: particles diffuse from full cells into the empty ones beside them (the shell around the Earth decks' inner
// box; the cell ahead of a beam), and how much can arrive is a matter of the neighbours, not of the slab's mean.
void buildLayout(const ipic_ctx* c, const std::vector<int64_t>& need, std::vector<int64_t>& start) {
	const Geometry& g = c->g;
	start.assign((size_t)g.ncellsLocal + 1, 0);
	int64_t pos = 0;
	for (int cell = 0; cell < g.ncellsLocal; cell++) {
		int64_t localFloor = 0;
		if (c->neighbourShare > 0.0) {
			const int xy = cell / g.lz, zl = cell - xy * g.lz, cx = xy / g.ny, cy = xy - cx * g.ny;
			int64_t fullest = 0;
			for (int a = -1; a <= 1; a++) for (int b = -1; b <= 1; b++) for (int d = -1; d <= 1; d++) {
				const int x = (cx + a + g.nx) % g.nx, y = (cy + b + g.ny) % g.ny;
				const int z = std::min(std::max(zl + d, 0), g.lz - 1);              // beyond the slab: the boundary plane stands in
				fullest = std::max(fullest, need[(size_t)((x * g.ny + y) * g.lz + z)]);
			}
			localFloor = (int64_t)(c->neighbourShare * (double)fullest);
		}
		start[(size_t)cell] = pos;
		pos += capacityFor(c, need[(size_t)cell], localFloor);
	}
	start[(size_t)g.ncellsLocal] = pos;
}

// slots a cell segment must hold for its outbox to be indexed and copied by many warps (IPIC_BIG_CAP overrides; IPIC_BIG_IMPORT=0
// switches the big-outbox path off: the import then scans such outboxes with the target cell's one warp, as it does for small ones)
int64_t bigCellSlots() {
	int64_t v = 32768;
	if (const char* e = std::getenv("IPIC_BIG_CAP")) v = std::max<int64_t>(32, std::atoll(e));
	if (const char* e = std::getenv("IPIC_BIG_IMPORT")) if (std::atoi(e) == 0) v = std::numeric_limits<int64_t>::max();
	return v;
}

void freeBigOutboxTables(ipic_ctx* c) {
	cudaFree(const_cast<int*>(c->big.cell)); cudaFree(const_cast<int*>(c->big.segFirst)); cudaFree(const_cast<int2*>(c->big.segTable));
	cudaFree(c->big.rank); cudaFree(c->big.off); cudaFree(c->big.dst); cudaFree(c->big.room);
	c->big = BigOutboxes{};
	c->bigCellCapacity = c->bigSegCapacity = 0;
}

// The big cells of the layout in c->hCellStart (universes on one GPU: in a slab context a big outbox may have to cross to
// another GPU, which the face messages do).  Called with the chunk table, i.e. at upload and after a re-layout.
int buildBigOutboxTables(ipic_ctx* c) {
	c->big.nBig = 0; c->big.totalSegs = 0;
	if (c->nranks > 1) return IPIC_OK;
	const int64_t least = bigCellSlots();
	const int seg = big_outbox_segment();
	std::vector<int> cells, segFirst(1, 0);
	std::vector<int2> segTable;
	for (int cell = 0; cell < c->g.ncellsLocal; cell++) {
		const int64_t cap = c->hCellStart[(size_t)cell + 1] - c->hCellStart[(size_t)cell];
		if (cap < least) continue;
		const int id = (int)cells.size();
		const int nseg = (int)((cap + seg - 1) / seg);
		cells.push_back(cell);
		for (int k = 0; k < nseg; k++) segTable.push_back(make_int2(id, k));
		segFirst.push_back((int)segTable.size());
	}
	if (cells.empty()) return IPIC_OK;
	if ((int64_t)cells.size() > c->bigCellCapacity || (int64_t)segTable.size() > c->bigSegCapacity) {
		freeBigOutboxTables(c);
		const size_t nc = cells.size() + cells.size() / 4 + 1, ns = segTable.size() + segTable.size() / 4 + 1;
		int* dCell = nullptr; int* dFirst = nullptr; int2* dTable = nullptr;
		CK(cudaMalloc(&dCell, nc * sizeof(int)));
		c->big.cell = dCell;
		CK(cudaMalloc(&dFirst, (nc + 1) * sizeof(int)));
		c->big.segFirst = dFirst;
		CK(cudaMalloc(&dTable, ns * sizeof(int2)));
		c->big.segTable = dTable;
		CK(cudaMalloc(&c->big.rank, ns * 27 * sizeof(int)));
		CK(cudaMalloc(&c->big.off, nc * 28 * sizeof(int)));
		CK(cudaMalloc(&c->big.dst, nc * 27 * sizeof(long long)));
		CK(cudaMalloc(&c->big.room, nc * 27 * sizeof(long long)));
		c->bigCellCapacity = (int64_t)nc; c->bigSegCapacity = (int64_t)ns;
	}
	CK(cudaMemcpyAsync(const_cast<int*>(c->big.cell), cells.data(), cells.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
	CK(cudaMemcpyAsync(const_cast<int*>(c->big.segFirst), segFirst.data(), segFirst.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
	CK(cudaMemcpyAsync(const_cast<int2*>(c->big.segTable), segTable.data(), segTable.size() * sizeof(int2), cudaMemcpyHostToDevice, c->stream));
	CK(cudaStreamSynchronize(c->stream));            // the host vectors go out of scope
	c->big.nBig = (int)cells.size();
	c->big.totalSegs = (int)segTable.size();
	return IPIC_OK;
}

// Chunk table of the layout in c->hCellStart, and the choice between one CTA per cell (k_move) and several
// (k_move_chunk): the latter when cells are big enough that whole cells are too few / too uneven work items.
int buildChunkTable(ipic_ctx* c, int64_t maxPopulation) {
	const Geometry& g = c->g;
	const int per = chunk_particles();
	// measured on B200 (scripts/probe_loads.py): one CTA per cell wins up to ~50 k particles in the fullest cell (k_move
	// 0.36 ms vs 0.43 ms on the :E: load); beyond that a single CTA becomes the critical path (:B: load, 6.5 M particles in
	// one cell: 28.8 ms vs 0.59 ms chunked)
	c->chunked = maxPopulation > 128 * (int64_t)per;
	if (const char* e = std::getenv("IPIC_CHUNKED")) c->chunked = std::atoi(e) != 0;
	if (!c->chunked) return buildBigOutboxTables(c);
	std::vector<int2> table;
	auto addPlanes = [&](int zFirst, int zEnd) {
		for (int xy = 0; xy < g.nx * g.ny; xy++) for (int zl = zFirst; zl < zEnd; zl++) {
			const int cell = xy * g.lz + zl;
			const int64_t cap = c->hCellStart[(size_t)cell + 1] - c->hCellStart[(size_t)cell];
			const int nch = (int)std::max<int64_t>(1, (cap + per - 1) / per);
			for (int j = 0; j < nch; j++) table.push_back(make_int2(cell, j));
		}
	};
	c->chunkRange[0] = 0;
	addPlanes(0, std::min(1, g.lz));
	c->chunkRange[1] = (int)table.size();
	addPlanes(1, std::max(1, g.lz - 1));
	c->chunkRange[2] = (int)table.size();
	if (g.lz > 1) addPlanes(g.lz - 1, g.lz);
	c->chunkRange[3] = (int)table.size();
	if ((int64_t)table.size() > c->chunkCapacity) {
		cudaFree(c->dChunkTable); cudaFree(c->dChunkStatus); cudaFree(c->dChunkJ);
		c->dChunkTable = nullptr; c->dChunkStatus = nullptr; c->dChunkJ = nullptr; c->chunkCapacity = 0;
		const size_t cap = table.size() + table.size() / 4;
		CK(cudaMalloc(&c->dChunkTable, cap * sizeof(int2)));
		CK(cudaMalloc(&c->dChunkStatus, cap * sizeof(unsigned long long)));
		CK(cudaMalloc(&c->dChunkJ, cap * 24 * sizeof(double)));
		c->chunkCapacity = (int64_t)cap;
	}
	if (!c->dChunkTicket) CK(cudaMalloc(&c->dChunkTicket, 4 * sizeof(unsigned)));
	CK(cudaMemcpyAsync(c->dChunkTable, table.data(), table.size() * sizeof(int2), cudaMemcpyHostToDevice, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	return buildBigOutboxTables(c);
}

int ensureStage(ipic_ctx* c, int64_t doubles) {
	if (c->stageDoubles >= doubles) return IPIC_OK;
	if (c->dStage) cudaFree(c->dStage);
	c->dStage = nullptr; c->stageDoubles = 0;
	CK(cudaMalloc(&c->dStage, (size_t)std::max<int64_t>(doubles, 1) * sizeof(double)));
	c->stageDoubles = doubles;
	return IPIC_OK;
}

int fetchOwnedCounts(ipic_ctx* c, const int* dArray, std::vector<int>& out) {
	out.resize((size_t)ownedCells(c));
	CK(cudaMemcpyAsync(out.data(), dArray, out.size() * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	return IPIC_OK;
}

void fillGeometry(ipic_ctx* c) {
	const ipic_props& p = c->props;
	Geometry& g = c->g;
	g.nx = (int)p.size[0]; g.ny = (int)p.size[1]; g.nz = (int)p.size[2];
	g.lz = c->zCount > 0 ? c->zCount : g.nz / c->nranks;
	g.z0 = c->zCount > 0 ? c->zBegin : c->rank * g.lz;
	g.ncellsLocal = g.nx * g.ny * g.lz;
	for (int a = 0; a < 3; a++) {
		g.w[a] = p.cellWidth[a];
		g.hw[a] = p.cellWidth[a] / 2.0;
		g.iw[a] = 1.0 / p.cellWidth[a];
		g.origin[a] = p.origin[a];
		g.objectCenter[a] = p.objectCenter[a];
	}
	g.dt = p.dt;
	g.c = p.speedOfLight;
	g.planetRadius = p.planetRadius;
	g.r2planet = p.planetRadius * p.planetRadius;
	g.magTemp = -p.externalMagneticField[2] * std::pow(p.planetRadius, 3);   // host libm, as the reference (cell.h:919)
	g.pic = M_PI * p.speedOfLight;                                           // cell.h:949
	g.picSafe = g.pic * (1.0 - 1e-6);
	g.vol = p.cellWidth[0] * p.cellWidth[1] * p.cellWidth[2];
	g.nf = kFields; g.uq = 0.0; g.uqom = 0.0;
}

// Tile shape of the mover (measured on B200, profiles/README.md): the kernel is bound by how many warps
// are resident, so one particle per thread (72 / 96 registers, 7 / 5 CTAs per SM) wins over two, and few
// stages win over many (shared memory caps the CTAs per SM): 2 stages = 16 KiB for the plain mover; with
// the fused deposit a batch of 4 stages amortises the per-batch block reduction of the 24 corner sums.
void pickTileShape(const ipic_ctx* c, bool deposit, bool pushes, int& stages) {
	const int64_t avg = c->meanPopulation;
	const int64_t tile = 128;
	// fused deposit, warp-specialised (k_move_ws): a deep ring decouples the deposit warp from the push warps (sweep: 8 > 6 > 4)
	const bool ws = move_uses_ws(c->g, deposit, !pushes, true);    // launch_move's choice of k_move_ws (callers pass pushes = !interpolate)
	// (8 stages of the 8-field layout are 64 KiB: six keep four CTAs per SM)
	stages = (int)std::max<int64_t>(1, std::min<int64_t>((avg + avg / 8 + tile - 1) / tile, ws ? (c->g.nf == kFields ? 6 : 8) : (deposit ? 4 : 2)));
	if (const char* e = std::getenv("IPIC_STAGES")) { const int v = std::atoi(e); if (v >= 1 && v <= 8) stages = v; }   // tuning knob
}

cudaEvent_t nextEvent(ipic_ctx* c) {
	if (c->eventsUsed == c->events.size()) { cudaEvent_t e; cudaEventCreate(&e); c->events.push_back(e); }
	return c->events[c->eventsUsed++];
}

int drainProfile(ipic_ctx* c) {
	for (size_t i = 0; i + 1 < c->eventsUsed; i += 2) {
		float a = 0;
		if (cudaEventElapsedTime(&a, c->events[i], c->events[i + 1]) != cudaSuccess) { cudaGetLastError(); continue; }   // a span cut short by a repair
		const int kind = c->eventKind[i / 2];
		if (kind == 0) c->moverMs += a;
		if (kind == 1) c->importMs += a;
		c->phaseMs[kind & 3] += a;
	}
	c->eventsUsed = 0;
	c->eventKind.clear();
	return IPIC_OK;
}

// CUDA-event bracket around a group of launches on the context's stream (only while profiling)
struct ProfSpan {
	ipic_ctx* c; cudaStream_t st; cudaEvent_t end = nullptr;
	ProfSpan(ipic_ctx* ctx, int kind, cudaStream_t stream = nullptr) : c(ctx), st(stream ? stream : ctx->stream) {
		if (!c->profiling) return;
		cudaEvent_t start = nextEvent(c); end = nextEvent(c);
		c->eventKind.push_back(kind);
		cudaEventRecord(start, st);
	}
	~ProfSpan() { if (end) cudaEventRecord(end, st); }
};

FaceMsg faceView(const ipic_ctx* c, char* base, int cap) {
	const int nxy = c->g.nx * c->g.ny;
	FaceMsg m;
	m.offsets = reinterpret_cast<int*>(base);
	const size_t offP = (size_t)alignUp((int64_t)(nxy * 9 + 1) * 4, 128);
	m.particles = reinterpret_cast<double*>(base + offP);
	m.cap = cap;
	return m;
}

// where the import kernel finds the outboxes of the cells beyond the slab's first / last plane
void setImportSources(const ipic_ctx* c, ImportArgs& ia) {
	ia.slab = c->nranks > 1;
	ia.big = c->big;
	if (ia.slab) { ia.fromDown = faceView(c, c->recvMsg[0], c->recvCap); ia.fromUp = faceView(c, c->recvMsg[1], c->recvCap); }
}

// Re-establish a valid layout after the import of `failedStep` ran out of room: the generation that
// was being written still holds every stayer (head) and every outbox (tail) intact, dCount holds the
// size every cell needs.  Build a roomier layout, move the heads over and redo that import.
int relayoutAfterOverflow(ipic_ctx* c, int outGen, int failedStep) {
	std::vector<int> need;
	int rc = fetchOwnedCounts(c, c->dCount, need);
	if (rc) return rc;
	std::vector<int64_t> need64(need.begin(), need.end());
	std::vector<int64_t> newStart;
	c->slack = std::min(c->slack * 1.5 + 0.1, 4.0);
	{
		int64_t total = 0;
		for (int64_t v : need64) total += v;
		const int64_t mean = total / std::max<int64_t>(1, (int64_t)need64.size());
		c->padFloor = std::min<int64_t>(std::max<int64_t>(2 * c->padFloor, mean / 4), 2 * mean + 64);
	}
	buildLayout(c, need64, newStart);
	ParticleArrays nb[2]{}; unsigned char* ntag = nullptr; unsigned short* norder = nullptr;
	int64_t* dNewStart = nullptr;
	// until the new storage is swapped in, a failure frees it again
	struct Fresh { ParticleArrays* nb; unsigned char** tag; unsigned short** order; int64_t** start; bool armed;
	               ~Fresh() { if (armed) { cudaFree(nb[0].d); cudaFree(nb[1].d); cudaFree(*tag); cudaFree(*order); cudaFree(*start); } } } fresh{ nb, &ntag, &norder, &dNewStart, true };
	const int64_t nslots = newStart.back();
	{
		// the new storage lives beside the old one until the heads and the failed import have been redone
		size_t freeB = 0, totalB = 0;
		const double needB = (double)nslots * (2.0 * c->g.nf * sizeof(double) + 1 + sizeof(unsigned short));
		if (cudaMemGetInfo(&freeB, &totalB) == cudaSuccess && needB > (double)freeB)
			return fail(IPIC_ERR_CAPACITY, "rank %d: a cell segment overflowed in step %d and the re-layout needs %.1f GB beside the live storage, %.1f GB are free: "
			                               "use more GPUs (thinner slabs) or give the cells more head-room up front (ipic_set_headroom)",
			            c->rank, failedStep, needB / 1e9, (double)freeB / 1e9);
	}
	rc = allocParticleBuffers(nb, &ntag, &norder, nslots, c->g.nf);
	if (rc) return rc;
	CK(cudaMalloc(&dNewStart, newStart.size() * sizeof(int64_t)));
	CK(cudaMemcpyAsync(dNewStart, newStart.data(), newStart.size() * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
	CK(cudaMemsetAsync(c->dStop, 0, 4 * sizeof(int), c->stream));      // [4] (sort flag) stays: the failed step's outbox index is in use
	int big = 0x7fffffff;
	CK(cudaMemcpyAsync(c->dStop + 1, &big, sizeof(int), cudaMemcpyHostToDevice, c->stream));
	launch_copy_heads(c->g, c->buf[outGen], c->dCellStart, nb[outGen], dNewStart, c->dStay, c->g.nf, c->g.uq, c->g.uqom, c->stream);
	ImportArgs ia{};
	ia.g = c->g; ia.src = c->buf[outGen]; ia.srcStart = c->dCellStart; ia.srcLeave = c->dLeave; ia.tag = c->dTag;
	ia.order = c->dOrder; ia.dirOff = c->dDirOff;          // the failed step's outbox index, slots of the OLD layout
	ia.dst = nb[outGen]; ia.dstStart = dNewStart; ia.stay = c->dStay; ia.dstCount = c->dCount;
	ia.reserveOwnOutbox = 0; ia.stopFlag = c->dStop; ia.stepIndex = failedStep; ia.zBegin = 0; ia.zCount = c->g.lz; ia.zStride = 1;
	setImportSources(c, ia);              // (with the failed step's big-outbox index, which describes the OLD layout's outboxes)
	c->launches += 1 + launch_import(ia, true, c->stream);
	CK(cudaStreamSynchronize(c->stream));
	int stop[4];
	CK(cudaMemcpy(stop, c->dStop, sizeof stop, cudaMemcpyDeviceToHost));
	// swap in the new storage
	fresh.armed = false;
	freeParticleBuffers(c);
	cudaFree(c->dCellStart);
	c->buf[0] = nb[0]; c->buf[1] = nb[1]; c->dTag = ntag; c->dOrder = norder; c->dCellStart = dNewStart; c->slots = nslots; c->slotsNf = c->g.nf; c->hCellStart = newStart;
	if (stop[0]) return fail(IPIC_ERR_INTERNAL, "relayout did not resolve the overflow of step %d", failedStep);
	return buildChunkTable(c, *std::max_element(need64.begin(), need64.end()));
}

// Leave the uniform-species layout: same segments, 8 fields per particle, q and qom materialised.  Needed when
// slabs turn out to carry different species (the live generation is converted; the other one is scratch).
int convertToFullLayout(ipic_ctx* c) {
	if (c->g.nf == kFields) return IPIC_OK;
	CK(cudaSetDevice(c->device));
	ParticleArrays nb[2]{}; unsigned char* ntag = nullptr; unsigned short* norder = nullptr;
	int rc = allocParticleBuffers(nb, &ntag, &norder, c->slots, kFields);
	if (rc) return rc;
	Geometry wide = c->g;
	wide.nf = kFields;
	launch_copy_heads(wide, c->buf[c->cur], c->dCellStart, nb[c->cur], c->dCellStart, c->dCount, c->g.nf, c->g.uq, c->g.uqom, c->stream);
	c->launches++;
	CK(cudaStreamSynchronize(c->stream));
	const int64_t slots = c->slots;
	freeParticleBuffers(c);
	c->buf[0] = nb[0]; c->buf[1] = nb[1]; c->dTag = ntag; c->dOrder = norder; c->slots = slots; c->slotsNf = kFields;
	c->g = wide;
	return IPIC_OK;
}

} // namespace

extern "C" {

// settles a batch left by ipic_step_async (waits, repairs a cell-segment overflow, redoes the skipped steps); defined with ipic_sync
static int settleAsync(ipic_ctx* c);

const char* ipic_last_error(void) { return g_lastError.c_str(); }
const char* ipic_version(void) { return "ipic3d_b200 0.1 (sm_100a)"; }

int ipic_device_count(int* count) {
	if (!count) return fail(IPIC_ERR_ARG, "count is null");
	CK(cudaGetDeviceCount(count));
	return IPIC_OK;
}

int ipic_alloc_pinned(size_t bytes, void** out) {
	if (!out) return fail(IPIC_ERR_ARG, "null argument");
	CK(cudaMallocHost(out, bytes ? bytes : 1));
	return IPIC_OK;
}

int ipic_free_pinned(void* p) {
	if (p) CK(cudaFreeHost(p));
	return IPIC_OK;
}

int ipic_create_slab(const ipic_props* props, int device, int rank, int nranks, ipic_ctx** out) {
	if (!props || !out) return fail(IPIC_ERR_ARG, "null argument");
	if (nranks < 1 || rank < 0 || rank >= nranks) return fail(IPIC_ERR_ARG, "bad rank %d of %d", rank, nranks);
	if (props->size[2] % nranks != 0) return fail(IPIC_ERR_ARG, "nz=%lld is not divisible by %d slabs", (long long)props->size[2], nranks);
	const int lz = (int)(props->size[2] / nranks);
	return ipic_create_slab_range(props, device, rank, nranks, rank * lz, lz, out);
}

int ipic_create_slab_range(const ipic_props* props, int device, int rank, int nranks, int zBegin, int zCount, ipic_ctx** out) {
	if (!props || !out) return fail(IPIC_ERR_ARG, "null argument");
	if (nranks < 1 || rank < 0 || rank >= nranks) return fail(IPIC_ERR_ARG, "bad rank %d of %d", rank, nranks);
	if (zBegin < 0 || zCount < 1 || zBegin + zCount > props->size[2]) return fail(IPIC_ERR_ARG, "bad z-range [%d, %d) of %lld planes", zBegin, zBegin + zCount, (long long)props->size[2]);
	if (nranks == 1 && (zBegin != 0 || zCount != props->size[2])) return fail(IPIC_ERR_ARG, "a single context holds all z-planes");
	for (int a = 0; a < 3; a++) {
		if (props->size[a] <= 0 || props->size[a] > (1 << 20)) return fail(IPIC_ERR_ARG, "bad grid size");
		if (!(props->cellWidth[a] > 0)) return fail(IPIC_ERR_ARG, "cell width must be positive");
	}
	if (!(props->dt > 0) || !(props->speedOfLight > 0) || !(props->planetRadius >= 0)) return fail(IPIC_ERR_ARG, "dt, speedOfLight must be positive and planetRadius non-negative");
	int ndev = 0;
	CK(cudaGetDeviceCount(&ndev));
	if (device < 0 || device >= ndev) return fail(IPIC_ERR_CUDA, "CUDA device %d not available (%d devices)", device, ndev);
	CK(cudaSetDevice(device));
	// a failure below (CK returns) must not leak the half-built context
	struct Guard { ipic_ctx* c; ~Guard() { if (c) ipic_destroy(c); } } guard{ new ipic_ctx() };
	ipic_ctx* c = guard.c;
	c->device = device; c->rank = rank; c->nranks = nranks; c->props = *props;
	c->zBegin = zBegin; c->zCount = zCount;
	fillGeometry(c);
	CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
	const Geometry& g = c->g;
	CK(cudaMalloc(&c->dCellStart, ((size_t)g.ncellsLocal + 1) * sizeof(int64_t)));
	CK(cudaMalloc(&c->dCount, (size_t)g.ncellsLocal * sizeof(int)));
	CK(cudaMalloc(&c->dStay, (size_t)g.ncellsLocal * sizeof(int)));
	CK(cudaMalloc(&c->dLeave, (size_t)g.ncellsLocal * sizeof(int)));
	CK(cudaMemset(c->dCount, 0, (size_t)g.ncellsLocal * sizeof(int)));
	CK(cudaMemset(c->dStay, 0, (size_t)g.ncellsLocal * sizeof(int)));
	CK(cudaMemset(c->dLeave, 0, (size_t)g.ncellsLocal * sizeof(int)));
	CK(cudaMalloc(&c->dDirOff, (size_t)std::max(1, g.ncellsLocal) * 28 * sizeof(int)));
	CK(cudaMemset(c->dDirOff, 0, (size_t)std::max(1, g.ncellsLocal) * 28 * sizeof(int)));
	CK(cudaMalloc(&c->dStop, 8 * sizeof(int)));
	int stop[8] = { 0, 0x7fffffff, 0, 0, 0, 0, 0, 0 };
	CK(cudaMemcpy(c->dStop, stop, sizeof stop, cudaMemcpyHostToDevice));
	c->nField = (int64_t)(g.nx + 3) * (g.ny + 3) * (g.nz + 3);
	c->nBc = (int64_t)(g.nx + 2) * (g.ny + 2) * (g.nz + 2);
	c->nJ = (int64_t)(g.nx + 1) * (g.ny + 1) * (g.nz + 1);
	CK(cudaMalloc(&c->dEB, (size_t)c->nField * 6 * sizeof(double)));
	CK(cudaMalloc(&c->dBext, (size_t)c->nField * 3 * sizeof(double)));
	CK(cudaMalloc(&c->dBc, (size_t)c->nBc * 3 * sizeof(double)));
	CK(cudaMalloc(&c->dJ, (size_t)c->nJ * 3 * sizeof(double)));
	CK(cudaMalloc(&c->dCellJ, (size_t)g.nx * g.ny * g.nz * 24 * sizeof(double)));
	CK(cudaMemset(c->dEB, 0, (size_t)c->nField * 6 * sizeof(double)));
	CK(cudaMemset(c->dBext, 0, (size_t)c->nField * 3 * sizeof(double)));
	CK(cudaMemset(c->dBc, 0, (size_t)c->nBc * 3 * sizeof(double)));
	CK(cudaMemset(c->dJ, 0, (size_t)c->nJ * 3 * sizeof(double)));
	CK(cudaMemset(c->dCellJ, 0, (size_t)g.nx * g.ny * g.nz * 24 * sizeof(double)));
	CK(cudaMalloc(&c->dAosStart, ((size_t)ownedCells(c) + 1) * sizeof(int64_t)));
	if (nranks > 1) {
		int prLow = 0, prHigh = 0;
		CK(cudaDeviceGetStreamPriorityRange(&prLow, &prHigh));
		// both side streams run at high priority: the block scheduler takes their (few) CTAs -- the boundary planes' movers, the
		// face packing, NCCL's send / receive kernels -- ahead of the tens of thousands of interior-mover CTAs queued on the main stream
		CK(cudaStreamCreateWithPriority(&c->commStream, cudaStreamNonBlocking, prHigh));
		CK(cudaStreamCreateWithPriority(&c->bStream, cudaStreamNonBlocking, prHigh));
		for (cudaEvent_t* e : { &c->evPacked, &c->evRecv, &c->evMoved, &c->evJ, &c->evMainReady }) CK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
		CK(cudaMalloc(&c->dTotals, 4 * sizeof(int)));
		CK(cudaMemset(c->dTotals, 0, 4 * sizeof(int)));
		CK(cudaMallocHost(&c->hTotals, 4 * sizeof(int)));
		std::memset(c->hTotals, 0, 4 * sizeof(int));
		CK(cudaMalloc(&c->dFaceCnt, (size_t)2 * g.nx * g.ny * 9 * sizeof(int)));
		CK(cudaMalloc(&c->dNegotiate, (size_t)nranks * 4 * sizeof(int64_t)));
		CK(cudaMallocHost(&c->hStop, 8 * sizeof(int)));
		std::memset(c->hStop, 0, 8 * sizeof(int));
		CK(cudaEventCreateWithFlags(&c->evStepDone, cudaEventDisableTiming));
	}
	guard.c = nullptr;
	*out = c;
	return IPIC_OK;
}

int ipic_create(const ipic_props* props, int device, ipic_ctx** out) { return ipic_create_slab(props, device, 0, 1, out); }

int ipic_destroy(ipic_ctx* c) {
	if (!c) return IPIC_OK;
	cudaSetDevice(c->device);
	if (c->stream) cudaStreamSynchronize(c->stream);
	if (c->commStream) cudaStreamSynchronize(c->commStream);
	if (c->bStream) cudaStreamSynchronize(c->bStream);
	if (c->comm) comm_destroy(c->comm);
	if (c->group) {
		// the remaining members can no longer step; drop the back references
		for (ipic_ctx*& m : c->group->members) if (m == c) m = nullptr;
		c->group.reset();
	}
	for (int i = 0; i < 2; i++) { cudaFree(c->sendMsg[i]); cudaFree(c->recvMsg[i]); }
	cudaFree(c->dFaceCnt); cudaFree(c->dNegotiate);
	if (c->hStop) cudaFreeHost(c->hStop);
	if (c->evStepDone) cudaEventDestroy(c->evStepDone);
	for (cudaEvent_t e : { c->evPacked, c->evRecv, c->evMoved, c->evJ, c->evMainReady }) if (e) cudaEventDestroy(e);
	cudaFree(c->dTotals);
	if (c->hTotals) cudaFreeHost(c->hTotals);
	if (c->commStream) cudaStreamDestroy(c->commStream);
	if (c->bStream) cudaStreamDestroy(c->bStream);
	freeParticleBuffers(c);
	cudaFree(c->dCellStart); cudaFree(c->dCount); cudaFree(c->dStay); cudaFree(c->dLeave); cudaFree(c->dStop); cudaFree(c->dDirOff);
	cudaFree(c->dChunkTable); cudaFree(c->dChunkStatus); cudaFree(c->dChunkTicket); cudaFree(c->dChunkJ);
	freeBigOutboxTables(c);
	cudaFree(c->dEB); cudaFree(c->dBext); cudaFree(c->dBc); cudaFree(c->dJ); cudaFree(c->dCellJ);
	cudaFree(c->dStage); cudaFree(c->dAosStart); cudaFree(c->dZiggurat);
	for (auto e : c->events) cudaEventDestroy(e);
	if (c->stream) cudaStreamDestroy(c->stream);
	delete c;
	return IPIC_OK;
}

int ipic_set_headroom(ipic_ctx* c, double slack, int64_t minPad) {
	if (!c || !(slack >= 0) || minPad < 0) return fail(IPIC_ERR_ARG, "bad head-room");
	c->slack = slack; c->minPad = minPad;
	c->userHeadroom = true; c->padFloor = 0; c->neighbourShare = 0.0;
	return IPIC_OK;
}

int ipic_set_exchange_capacity(ipic_ctx* c, int64_t particles) {
	if (!c || particles < 0 || particles > ((int64_t)1 << 26)) return fail(IPIC_ERR_ARG, "bad exchange capacity");
	c->userMsgCap = particles;
	c->exchangeDirty = true;
	return IPIC_OK;
}

int ipic_set_uniform_species(ipic_ctx* c, int enable) {
	if (!c) return fail(IPIC_ERR_ARG, "null context");
	c->allowUniformSpecies = enable != 0;
	return IPIC_OK;
}

int ipic_particle_fields(ipic_ctx* c, int* fields) {
	if (!c || !fields) return fail(IPIC_ERR_ARG, "null argument");
	*fields = c->g.nf;
	return IPIC_OK;
}

namespace {

// Segment layout, storage and per-cell counters for the populations `need` (owned cells, slab order) in a layout of
// nf fields per particle; the particles themselves are written afterwards (scatter of an upload, device generator).
int establishLayout(ipic_ctx* c, const std::vector<int64_t>& need, int nf, double uq, double uqom) {
	const int nOwned = ownedCells(c);
	int64_t total = 0;
	std::vector<int> localCount((size_t)c->g.ncellsLocal, 0);
	for (int i = 0; i < nOwned; i++) { total += need[(size_t)i]; localCount[(size_t)i] = (int)need[(size_t)i]; }
	c->g.nf = nf; c->g.uq = uq; c->g.uqom = uqom;
	if (!c->userHeadroom) {
		// sparse cells next to full ones fill up towards the mean (the dipole runs start with 12 particles per cell outside the
		// inner box and ~1000 inside; diffusion levels that out): give every cell room for the mean population and a quarter,
		// instead of paying a re-layout every few steps.  Evenly filled universes (:U:) keep the tight layout.
		const int64_t mean = total / std::max(1, nOwned);
		const int64_t least = need.empty() ? 0 : *std::min_element(need.begin(), need.end());
		c->padFloor = least < mean / 4 ? mean + mean / 4 : 0;
		// ... and for what its fullest neighbour can send it: the slab's mean says little when most of the slab is empty (medium.inp
		// on 8 GPUs: the outer slabs hold 8 empty planes beside 6 full ones, a floor of 1.25 x their mean overflowed within ten steps)
		c->neighbourShare = least < mean / 4 ? 0.6 : 0.0;
	}
	std::vector<int64_t> start;
	buildLayout(c, need, start);
	if (!c->userHeadroom && c->padFloor > 0) {
		// the adaptive floor must fit the GPU: two generations of nf doubles, a tag byte and an order entry per slot; what is
		// allocated now would be freed first.  (large.inp on 8 GPUs: 885 M particles per slab, a floor of 1.25 x the mean
		// population on 512 k cells would ask for 190 GB.)  The floor shrinks until 85 % of the memory suffices.
		size_t freeB = 0, totalB = 0;
		if (cudaMemGetInfo(&freeB, &totalB) == cudaSuccess) {
			const double perSlot = 2.0 * nf * sizeof(double) + 1 + sizeof(unsigned short);
			const double have = 0.85 * ((double)freeB + (double)c->slots * (2.0 * c->slotsNf * sizeof(double) + 3));
			while ((c->padFloor > 64 || c->neighbourShare > 0.01) && (double)start.back() * perSlot > have) {
				c->padFloor /= 2;
				c->neighbourShare /= 2;
				buildLayout(c, need, start);
			}
		}
	}
	if (start.back() > c->slots || start.back() < c->slots / 2 || c->slotsNf != nf) {
		freeParticleBuffers(c);
		int rc = allocParticleBuffers(c->buf, &c->dTag, &c->dOrder, start.back(), nf);
		if (rc) return rc;
		c->slots = start.back();
		c->slotsNf = nf;
	}
	c->hCellStart = start;
	int rc = buildChunkTable(c, need.empty() ? 0 : *std::max_element(need.begin(), need.end()));
	if (rc) return rc;
	CK(cudaMemcpyAsync(c->dCellStart, start.data(), start.size() * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
	CK(cudaMemcpyAsync(c->dCount, localCount.data(), localCount.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
	CK(cudaMemsetAsync(c->dLeave, 0, (size_t)c->g.ncellsLocal * sizeof(int), c->stream));
	CK(cudaStreamSynchronize(c->stream));     // the host vectors above go out of scope
	c->haveCells = true;
	c->sortLikely = true;
	c->emptyUpload = total == 0;
	c->meanPopulation = total / std::max(1, nOwned);
	if (c->nranks > 1) {
		int64_t lo = 0, hi = 0;
		for (int xy = 0; xy < c->g.nx * c->g.ny; xy++) { lo += need[(size_t)xy * c->g.lz]; hi += need[(size_t)xy * c->g.lz + c->g.lz - 1]; }
		c->facePopulation = std::max(lo, hi);
		c->exchangeDirty = true;
	}
	return IPIC_OK;
}

} // namespace

int ipic_upload_cells(ipic_ctx* c, const uint64_t* counts, const double* aos) {
	NvtxRange nvtx("ipic_upload_cells");
	if (!c || !counts) return fail(IPIC_ERR_ARG, "null argument");
	CK(cudaSetDevice(c->device));
	{ int rcA = settleAsync(c); if (rcA) return rcA; }
	const int nOwned = ownedCells(c);
	std::vector<int64_t> need((size_t)nOwned), aosStart((size_t)nOwned + 1, 0);
	for (int i = 0; i < nOwned; i++) {
		if (counts[i] > 0x3fffffffu) return fail(IPIC_ERR_ARG, "cell %d holds %llu particles (limit 2^30 per cell)", i, (unsigned long long)counts[i]);
		need[i] = (int64_t)counts[i];
		aosStart[i + 1] = aosStart[i] + need[i];
	}
	const int64_t total = aosStart[nOwned];
	if (total > 0 && !aos) return fail(IPIC_ERR_ARG, "particles_aos is null");
	// stage the AoS records on the device and look at the species: one (q, qom) for all particles -> 6-field layout
	int rc = ensureStage(c, total * 8 + 8);
	if (rc) return rc;
	int nf = kFields;
	double uq = 0.0, uqom = 0.0;
	if (total) {
		CK(cudaMemcpyAsync(c->dStage, aos, (size_t)total * 64, cudaMemcpyHostToDevice, c->stream));
		if (c->allowUniformSpecies) {
			int* dFlag = reinterpret_cast<int*>(c->dStage + total * 8);
			CK(cudaMemsetAsync(dFlag, 0, sizeof(int), c->stream));
			launch_check_species(c->dStage, total, dFlag, c->stream);
			c->launches++;
			int mixed = 1;
			CK(cudaMemcpyAsync(&mixed, dFlag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
			CK(cudaStreamSynchronize(c->stream));
			if (!mixed) { nf = 6; uq = aos[6]; uqom = aos[7]; }
		}
	}
	rc = establishLayout(c, need, nf, uq, uqom);
	if (rc) return rc;
	CK(cudaMemcpyAsync(c->dAosStart, aosStart.data(), aosStart.size() * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
	if (total) {
		launch_scatter_aos(c->g, c->dStage, c->dAosStart, c->buf[c->cur], c->dCellStart, c->stream);
		c->launches++;
	}
	CK(cudaStreamSynchronize(c->stream));     // aosStart goes out of scope
	return IPIC_OK;
}

int ipic_generate_uniform(ipic_ctx* c, uint64_t numParticles) {
	if (!c) return fail(IPIC_ERR_ARG, "null context");
	CK(cudaSetDevice(c->device));
	{ int rcA = settleAsync(c); if (rcA) return rcA; }
	const Geometry& g = c->g;
	const uint64_t numCells = (uint64_t)g.nx * g.ny * g.nz;
	if (numParticles / numCells + 1 > 0x3fffffffu) return fail(IPIC_ERR_ARG, "%llu particles over %llu cells exceed 2^30 per cell", (unsigned long long)numParticles, (unsigned long long)numCells);
	std::vector<int64_t> need((size_t)ownedCells(c));
	for (int xy = 0; xy < g.nx * g.ny; xy++) for (int zl = 0; zl < g.lz; zl++) {
		const uint64_t lin = (uint64_t)xy * g.nz + (uint64_t)(g.z0 + zl);      // the reference's flatten, z fastest (cell.h:521-523)
		need[(size_t)xy * g.lz + zl] = (int64_t)(numParticles / numCells + (lin < numParticles % numCells ? 1 : 0));
	}
	const bool oneSpecies = c->allowUniformSpecies;
	int rc = establishLayout(c, need, oneSpecies ? 6 : kFields, oneSpecies ? -1.0 : 0.0, oneSpecies ? -25.0 : 0.0);
	if (rc) return rc;
	launch_init_uniform(c->g, c->buf[c->cur], c->dCellStart, c->dCount, c->stream);
	c->launches++;
	CK(cudaStreamSynchronize(c->stream));
	CK(cudaGetLastError());
	return IPIC_OK;
}

// the reference's zigset (hdr/ziggurat_normal_distribution.h:107-125) with the host's libm, like the reference itself
static void zigguratTables(ZigguratTables& t) {
	const double m1 = 2147483648.0;
	double dn = 3.442619855899, tn = dn;
	const double vn = 9.91256303526217e-3;
	const double q = vn / std::exp(-.5 * dn * dn);
	t.kn[0] = (unsigned long long)((dn / q) * m1);
	t.kn[1] = 0;
	t.wn[0] = (float)(q / m1);
	t.wn[127] = (float)(dn / m1);
	t.fn[0] = 1.;
	t.fn[127] = (float)std::exp(-.5 * dn * dn);
	for (int i = 126; i >= 1; --i) {
		dn = std::sqrt(-2. * std::log(vn / dn + std::exp(-.5 * dn * dn)));
		t.kn[i + 1] = (unsigned long long)((dn / tn) * m1);
		tn = dn;
		t.fn[i] = (float)std::exp(-.5 * dn * dn);
		t.wn[i] = (float)(dn / m1);
	}
}

int ipic_generate_normal_speed(ipic_ctx* c, const uint64_t* counts, const double* mean3, const double* stddev3, double q, double qom) {
	if (!c || !counts || !mean3 || !stddev3) return fail(IPIC_ERR_ARG, "null argument");
	CK(cudaSetDevice(c->device));
	{ int rcA = settleAsync(c); if (rcA) return rcA; }
	const int nOwned = ownedCells(c);
	std::vector<int64_t> need((size_t)nOwned);
	for (int i = 0; i < nOwned; i++) {
		if (counts[i] > 0x3fffffffu) return fail(IPIC_ERR_ARG, "cell %d holds %llu particles (limit 2^30 per cell)", i, (unsigned long long)counts[i]);
		need[(size_t)i] = (int64_t)counts[i];
	}
	const bool oneSpecies = c->allowUniformSpecies;
	int rc = establishLayout(c, need, oneSpecies ? 6 : kFields, q, qom);
	if (rc) return rc;
	if (!c->dZiggurat) {
		ZigguratTables t;
		zigguratTables(t);
		CK(cudaMalloc(&c->dZiggurat, sizeof t));
		CK(cudaMemcpy(c->dZiggurat, &t, sizeof t, cudaMemcpyHostToDevice));
	}
	launch_init_deck(c->g, c->buf[c->cur], c->dCellStart, c->dCount, c->dZiggurat, mean3, stddev3, c->stream);
	c->launches += 2;
	CK(cudaStreamSynchronize(c->stream));
	CK(cudaGetLastError());
	return IPIC_OK;
}

int ipic_cell_counts(ipic_ctx* c, uint64_t* counts) {
	if (!c || !counts) return fail(IPIC_ERR_ARG, "null argument");
	CK(cudaSetDevice(c->device));
	{ int rcA = settleAsync(c); if (rcA) return rcA; }
	std::vector<int> owned;
	int rc = fetchOwnedCounts(c, c->dCount, owned);
	if (rc) return rc;
	for (size_t i = 0; i < owned.size(); i++) counts[i] = (uint64_t)owned[i];
	return IPIC_OK;
}

int ipic_total_particles(ipic_ctx* c, uint64_t* total) {
	if (!c || !total) return fail(IPIC_ERR_ARG, "null argument");
	std::vector<uint64_t> counts((size_t)ownedCells(c));
	int rc = ipic_cell_counts(c, counts.data());
	if (rc) return rc;
	uint64_t t = 0; for (uint64_t v : counts) t += v;
	*total = t;
	return IPIC_OK;
}

int ipic_download_cells(ipic_ctx* c, uint64_t* counts, double* aos, uint64_t capacity) {
	NvtxRange nvtx("ipic_download_cells");
	if (!c || !counts) return fail(IPIC_ERR_ARG, "null argument");
	CK(cudaSetDevice(c->device));
	{ int rcA = settleAsync(c); if (rcA) return rcA; }
	std::vector<int> owned;
	int rc = fetchOwnedCounts(c, c->dCount, owned);
	if (rc) return rc;
	std::vector<int64_t> aosStart(owned.size() + 1, 0);
	for (size_t i = 0; i < owned.size(); i++) { counts[i] = (uint64_t)owned[i]; aosStart[i + 1] = aosStart[i] + owned[i]; }
	const int64_t total = aosStart.back();
	if ((uint64_t)total > capacity) return fail(IPIC_ERR_CAPACITY, "download needs room for %lld particles, caller offers %llu", (long long)total, (unsigned long long)capacity);
	if (total == 0) return IPIC_OK;
	if (!aos) return fail(IPIC_ERR_ARG, "particles_aos is null");
	rc = ensureStage(c, total * 8);
	if (rc) return rc;
	CK(cudaMemcpyAsync(c->dAosStart, aosStart.data(), aosStart.size() * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
	launch_gather_aos(c->g, c->dStage, c->dAosStart, c->buf[c->cur], c->dCellStart, c->stream);
	c->launches++;
	CK(cudaMemcpyAsync(aos, c->dStage, (size_t)total * 64, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	return IPIC_OK;
}

int ipic_upload_fields(ipic_ctx* c, const double* field, const double* bc, const double* J) {
	if (!c) return fail(IPIC_ERR_ARG, "null context");
	CK(cudaSetDevice(c->device));
	{ int rcA = settleAsync(c); if (rcA) return rcA; }
	if (field) {
		int rc = ensureStage(c, c->nField * 9);
		if (rc) return rc;
		CK(cudaMemcpyAsync(c->dStage, field, (size_t)c->nField * 72, cudaMemcpyHostToDevice, c->stream));
		launch_split_field(c->dStage, c->dEB, c->dBext, c->nField, c->stream);
		c->launches++;
		CK(cudaStreamSynchronize(c->stream));
	}
	if (bc) CK(cudaMemcpy(c->dBc, bc, (size_t)c->nBc * 24, cudaMemcpyHostToDevice));
	if (J) CK(cudaMemcpy(c->dJ, J, (size_t)c->nJ * 24, cudaMemcpyHostToDevice));
	return IPIC_OK;
}

int ipic_download_fields(ipic_ctx* c, double* field, double* bc, double* J) {
	if (!c) return fail(IPIC_ERR_ARG, "null context");
	CK(cudaSetDevice(c->device));
	{ int rcA = settleAsync(c); if (rcA) return rcA; }
	CK(cudaStreamSynchronize(c->stream));
	if (field) {
		int rc = ensureStage(c, c->nField * 9);
		if (rc) return rc;
		launch_merge_field(c->dStage, c->dEB, c->dBext, c->nField, c->stream);
		c->launches++;
		CK(cudaMemcpyAsync(field, c->dStage, (size_t)c->nField * 72, cudaMemcpyDeviceToHost, c->stream));
		CK(cudaStreamSynchronize(c->stream));
	}
	if (bc) CK(cudaMemcpy(bc, c->dBc, (size_t)c->nBc * 24, cudaMemcpyDeviceToHost));
	if (J) CK(cudaMemcpy(J, c->dJ, (size_t)c->nJ * 24, cudaMemcpyDeviceToHost));
	return IPIC_OK;
}

// ---- one iteration of the time loop --------------------------------------------------------------

namespace {

struct StepFlags {
	bool interp, project, solve, migrate, fuseDeposit;
	uint32_t raw;
	explicit StepFlags(uint32_t f) : raw(f) {
		interp = f & IPIC_MOVER_INTERPOLATED;
		project = f & IPIC_PROJECT_CURRENT;
		solve = f & (IPIC_SOLVER_FORWARD | IPIC_SOLVER_LEAPFROG);
		migrate = !(f & IPIC_SKIP_MIGRATION);
		// The reference order is project -> boundaries/solve -> move -> import (simulator.h:189-223).  The
		// analytic mover never reads the grid, so there the projection is fused into the mover pass and the
		// field phase runs after it with identical results; the interpolating mover needs the new fields
		// first, so its projection is a separate read-only pass.
		fuseDeposit = project && !interp;
		// Who needs the current where?  A field solver runs replicated on every slab (its in-place node order is global) and
		// needs J everywhere: the slabs all-gather their per-cell corner sums.  Without a solver every slab only builds J on
		// its own node planes, for which it needs ONE plane of corner sums from each z-neighbour (a halo exchange).
		jHalo = project && !solve;
	}
	bool jHalo;
};

MoveArgs moveArgs(ipic_ctx* c, const StepFlags& f, int zBegin, int zCount, int zStride = 1) {
	const int in = c->cur, out = c->cur ^ 1;
	MoveArgs ma{};
	ma.g = c->g; ma.in = c->buf[in]; ma.out = c->buf[out]; ma.cellStart = c->dCellStart;
	ma.countIn = c->dCount; ma.countOut = c->dStay; ma.nLeave = c->dLeave; ma.tag = c->dTag; ma.order = c->dOrder; ma.dirOff = c->dDirOff;
	ma.fieldEB = c->dEB; ma.cellJ = c->dCellJ; ma.stopFlag = c->dStop;
	ma.zBegin = zBegin; ma.zCount = zCount; ma.zStride = zStride;
	ma.interpolate = f.interp; ma.migrate = f.migrate; ma.push = true; ma.deposit = f.fuseDeposit; ma.sortPass = c->sortLikely;
	pickTileShape(c, ma.deposit, !f.interp, ma.stages);
	ma.chunked = false;
	if (c->chunked && zStride == 1) {
		const int lz = c->g.lz;
		int grp = -1;
		if (zBegin == 0 && zCount == lz) grp = 3;
		else if (zBegin == 0 && zCount == 1) grp = 0;
		else if (zBegin == lz - 1 && zCount == 1) grp = 2;
		else if (zBegin == 1 && zCount == lz - 2) grp = 1;
		if (grp >= 0) {
			ma.chunked = true;
			ma.chunkTable = c->dChunkTable; ma.chunkStatus = c->dChunkStatus; ma.chunkJ = c->dChunkJ;
			ma.chunkTicket = c->dChunkTicket + grp;
			ma.chunkFirst = grp == 3 ? 0 : c->chunkRange[grp];
			ma.chunkCount = (grp == 3 ? c->chunkRange[3] : c->chunkRange[grp + 1]) - ma.chunkFirst;
		}
	}
	return ma;
}

void fieldPhase(ipic_ctx* c, const StepFlags& f) {
	if (f.raw & IPIC_UPDATE_BOUNDARIES) { launch_update_boundaries(c->g, c->dEB, c->dBext, c->dBc, c->dStop, c->stream); c->launches += 2; }
	if (f.solve) c->launches += launch_field_solver(c->g, (f.raw & IPIC_SOLVER_FORWARD) != 0, c->dEB, c->dBc, c->dJ, c->dStop, c->stream);
}

// deposit-only pass over the given planes (the projection of the interpolating mode)
void depositPass(ipic_ctx* c, const StepFlags& f, int zBegin, int zCount) {
	MoveArgs da = moveArgs(c, f, zBegin, zCount);
	da.push = false; da.deposit = true; da.interpolate = false;
	pickTileShape(c, true, false, da.stages);
	c->launches += launch_move(da, c->stream);
}

void importPlanes(ipic_ctx* c, int stepIndex, int zBegin, int zCount, bool commit, int zStride = 1) {
	if (zCount <= 0) return;
	const int out = c->cur ^ 1;
	ImportArgs ia{};
	ia.g = c->g; ia.src = c->buf[out]; ia.srcStart = c->dCellStart; ia.srcLeave = c->dLeave; ia.tag = c->dTag;
	ia.order = c->dOrder; ia.dirOff = c->dDirOff;
	ia.dst = c->buf[out]; ia.dstStart = c->dCellStart; ia.stay = c->dStay; ia.dstCount = c->dCount;
	ia.reserveOwnOutbox = 1; ia.stopFlag = c->dStop; ia.stepIndex = stepIndex; ia.zBegin = zBegin; ia.zCount = zCount; ia.zStride = zStride;
	setImportSources(c, ia);
	c->launches += launch_import(ia, commit, c->stream);
}

// event pair whose two halves are recorded in different phases (slab steps: the span of all movers)
void spanBegin(ipic_ctx* c, int kind, cudaEvent_t* pendingEnd) {
	*pendingEnd = nullptr;
	if (!c->profiling) return;
	cudaEvent_t start = nextEvent(c);
	*pendingEnd = nextEvent(c);
	c->eventKind.push_back(kind);
	cudaEventRecord(start, c->stream);
}
void spanEnd(ipic_ctx* c, cudaEvent_t* pendingEnd) {
	if (*pendingEnd) cudaEventRecord(*pendingEnd, c->stream);
	*pendingEnd = nullptr;
}

// the whole universe on one GPU: no exchange
int enqueueStepSingle(ipic_ctx* c, const StepFlags& f, int stepIndex) {
	NvtxRange nvtx("step:enqueue");
	if (f.project && f.interp) {
		depositPass(c, f, 0, c->g.lz);
		launch_gather_current(c->g, c->dCellJ, c->dJ, c->dStop, c->stream);
		c->launches++;
	}
	if (!f.fuseDeposit) fieldPhase(c, f);
	{
		ProfSpan span(c, 0);
		c->launches += launch_move(moveArgs(c, f, 0, c->g.lz), c->stream);
	}
	if (f.migrate && c->big.nBig > 0) {
		// outboxes of the universe's big cells: counted and ranked by many warps, so that no single warp has to import them
		ProfSpan span(c, 1);
		c->launches += launch_big_index(c->big, c->dCellStart, c->dLeave, c->dTag, c->dDirOff, c->dStop, c->stream);
	}
	if (f.fuseDeposit) {
		launch_gather_current(c->g, c->dCellJ, c->dJ, c->dStop, c->stream);
		c->launches++;
		fieldPhase(c, f);
	}
	if (f.migrate) {
		ProfSpan span(c, 1);
		importPlanes(c, stepIndex, 0, c->g.lz, true);
	} else {
		CK(cudaMemcpyAsync(c->dCount, c->dStay, (size_t)c->g.ncellsLocal * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
	}
	c->cur ^= 1;
	CK(cudaGetLastError());
	return IPIC_OK;
}

// ---- z-slab contexts ------------------------------------------------------------------------------
//
// A step of a slab is enqueued in phases.  With NCCL (one process per GPU) a context runs its phases
// back to back; an in-process group runs every phase over all members before the next one so that
// each cross-context event wait refers to an event recorded earlier in host order.
//
//   BOUNDARY  movers of the two boundary planes, pack both faces, totals to the host   [boundary stream, high priority]
//   INTERIOR  mover of the interior planes, concurrently (fills the SMs the boundary
//             launches leave idle)                                                     [main stream]
//   SETTLE    host: confirm the previous step (repair a cell overflow), read the packed
//             totals, swap them with the neighbours, grow message buffers if needed    [host, comm stream]
//   EXCHANGE  the face messages, exactly as large as their content                     [comm stream]
//   GATHER_J  all-gather of the slabs' per-cell corner sums -- overlaps the imports    [comm stream]
//   IMPORT    import of all planes; the out-of-slab neighbours of the boundary planes
//             are columns of the received messages                                     [main stream]
//   FIELDS    wait for the all-gather; J on every rank; ghost update + field solver    [main stream]
//
// (PRE_DEPOSIT / PRE_GATHER_J: the projection pass of the interpolating mode, which needs J and the
// field update before the mover.)  The host waits (SETTLE) happen while the GPU works through the movers
// of the same step, which were queued before.

enum SlabPhase { PH_PRE_DEPOSIT, PH_PRE_GATHER_J, PH_BOUNDARY, PH_INTERIOR, PH_SETTLE, PH_EXCHANGE, PH_GATHER_J, PH_IMPORT, PH_FIELDS, PH_COUNT };

size_t faceBytes(const ipic_ctx* c, int64_t cap) {
	const int nxy = c->g.nx * c->g.ny;
	return (size_t)(alignUp((int64_t)(nxy * 9 + 1) * 4, 128) + cap * 64);
}

int64_t wantedMsgCap(const ipic_ctx* c) {
	if (c->userMsgCap > 0) return c->userMsgCap;
	return std::min<int64_t>(std::max<int64_t>(c->facePopulation / 64, 16384), (int64_t)1 << 26);
}

int allocFaceBuffers(ipic_ctx* c, char* bufs[2], int64_t cap) {
	for (int i = 0; i < 2; i++) {
		cudaFree(bufs[i]);
		bufs[i] = nullptr;
		CK(cudaMalloc(&bufs[i], faceBytes(c, cap)));
		CK(cudaMemset(bufs[i], 0, faceBytes(c, cap)));
	}
	return IPIC_OK;
}

int allocExchange(ipic_ctx* c, int64_t cap) {
	CK(cudaSetDevice(c->device));
	CK(cudaStreamSynchronize(c->stream));
	CK(cudaStreamSynchronize(c->bStream));
	CK(cudaStreamSynchronize(c->commStream));
	int rc = allocFaceBuffers(c, c->sendMsg, cap);
	if (rc) return rc;
	rc = allocFaceBuffers(c, c->recvMsg, cap);
	if (rc) return rc;
	c->sendCap = c->recvCap = (int)cap;
	c->exchangeDirty = false;
	return IPIC_OK;
}

// What a slab contributes to the negotiation before the first step after an upload: its species state
// (0 = no particles, 6 = one species with the (q, qom) bits that follow, 8 = mixed).
struct SlabOffer { int64_t species, qBits, qomBits, zRange; };       // zRange = z0 << 32 | lz

SlabOffer offerOf(const ipic_ctx* c) {
	SlabOffer o{};
	o.species = c->emptyUpload ? 0 : c->g.nf;
	std::memcpy(&o.qBits, &c->g.uq, 8);
	std::memcpy(&o.qomBits, &c->g.uqom, 8);
	o.zRange = ((int64_t)c->g.z0 << 32) | (int64_t)c->g.lz;
	return o;
}

// All slabs use the same particle layout: 6 fields only if every slab that holds particles holds the same single
// species (arrivals from a neighbour must fit the layout); empty slabs adopt it; otherwise everybody widens to 8.
int agreeOnLayout(ipic_ctx* c, const std::vector<SlabOffer>& all) {
	// the slabs must tile the z-axis in rank order
	c->slabZ0.clear(); c->slabLz.clear();
	int next = 0;
	for (const SlabOffer& o : all) {
		const int z0 = (int)(o.zRange >> 32), lz = (int)(o.zRange & 0xffffffff);
		if (z0 != next || lz < 1) return fail(IPIC_ERR_ARG, "the slabs do not tile the z-axis in rank order (slab %d starts at plane %d, expected %d)", (int)c->slabZ0.size(), z0, next);
		c->slabZ0.push_back(z0); c->slabLz.push_back(lz);
		next = z0 + lz;
	}
	if (next != c->g.nz) return fail(IPIC_ERR_ARG, "the slabs cover %d of %d z-planes", next, c->g.nz);
	bool uniform = true, any = false;
	int64_t q = 0, m = 0;
	for (const SlabOffer& o : all) {
		if (o.species == 0) continue;
		if (o.species != 6) { uniform = false; break; }
		if (!any) { q = o.qBits; m = o.qomBits; any = true; }
		else if (o.qBits != q || o.qomBits != m) { uniform = false; break; }
	}
	if (!uniform || !any) return convertToFullLayout(c);
	if (c->g.nf == kFields) {          // an empty slab: its (8-field sized) buffers are roomy enough for the 6-field layout
		c->g.nf = 6;
		std::memcpy(&c->g.uq, &q, 8);
		std::memcpy(&c->g.uqom, &m, 8);
	}
	return IPIC_OK;
}

// before the first step after an upload: agree on the layout, set up the message buffers.  NCCL contexts must
// all enter this in the same ipic_step call.
int negotiateExchangeNccl(ipic_ctx* c) {
	if (!c->exchangeDirty) return IPIC_OK;
	CK(cudaSetDevice(c->device));
	std::vector<SlabOffer> all((size_t)c->nranks);
	const SlabOffer mine = offerOf(c);
	SlabOffer* d = reinterpret_cast<SlabOffer*>(c->dNegotiate);
	CK(cudaMemcpyAsync(d + c->rank, &mine, sizeof mine, cudaMemcpyHostToDevice, c->commStream));
	int rc = comm_allgather(c->comm, d + c->rank, d, sizeof(SlabOffer), c->commStream, g_lastError);
	if (rc) return rc;
	CK(cudaMemcpyAsync(all.data(), d, all.size() * sizeof(SlabOffer), cudaMemcpyDeviceToHost, c->commStream));
	CK(cudaStreamSynchronize(c->commStream));
	rc = agreeOnLayout(c, all);
	if (rc) return rc;
	return allocExchange(c, wantedMsgCap(c));
}

int negotiateExchangeGroup(ipic_group* grp) {
	bool dirty = false;
	std::vector<SlabOffer> all;
	for (ipic_ctx* m : grp->members) { dirty = dirty || m->exchangeDirty; all.push_back(offerOf(m)); }
	if (!dirty) return IPIC_OK;
	for (ipic_ctx* m : grp->members) {
		int rc = agreeOnLayout(m, all);
		if (rc) return rc;
		rc = allocExchange(m, wantedMsgCap(m));
		if (rc) return rc;
	}
	return IPIC_OK;
}

ipic_ctx* neighbour(const ipic_ctx* c, int dir) {      // in-process group only
	const int n = c->nranks;
	return c->group->members[(size_t)((c->rank + dir + n) % n)];
}

// every rank ends up with the corner sums of all cells (slab r is the r-th contiguous piece of dCellJ)
int allGatherCellJ(ipic_ctx* c) {
	const size_t planeBytes = (size_t)c->g.nx * c->g.ny * 24 * sizeof(double);      // cellJ is z-major: a slab is one contiguous piece
	CK(cudaStreamWaitEvent(c->commStream, c->evMoved, 0));
	CK(cudaStreamWaitEvent(c->commStream, c->evPacked, 0));
	if (c->comm) {
		std::vector<size_t> off((size_t)c->nranks), len((size_t)c->nranks);
		for (int r = 0; r < c->nranks; r++) { off[(size_t)r] = (size_t)c->slabZ0[(size_t)r] * planeBytes; len[(size_t)r] = (size_t)c->slabLz[(size_t)r] * planeBytes; }
		int rc = comm_allgatherv(c->comm, c->dCellJ, off.data(), len.data(), c->commStream, g_lastError);
		if (rc) return rc;
	} else {
		for (ipic_ctx* m : c->group->members) {
			if (m == c) continue;
			CK(cudaStreamWaitEvent(c->commStream, m->evMoved, 0));
			CK(cudaStreamWaitEvent(c->commStream, m->evPacked, 0));
			const size_t offD = (size_t)m->g.z0 * planeBytes / sizeof(double);
			CK(cudaMemcpyPeerAsync(c->dCellJ + offD, c->device, m->dCellJ + offD, m->device, (size_t)m->g.lz * planeBytes, c->commStream));
		}
	}
	CK(cudaEventRecord(c->evJ, c->commStream));
	return IPIC_OK;
}

// main stream: wait for the all-gather, then the node-centric gather of J
int finishCurrent(ipic_ctx* c, bool ownPlanesOnly = false) {
	CK(cudaStreamWaitEvent(c->stream, c->evJ, 0));
	if (ownPlanesOnly) launch_gather_current(c->g, c->dCellJ, c->dJ, c->dStop, c->stream, c->g.z0, c->g.lz + 1);
	else launch_gather_current(c->g, c->dCellJ, c->dJ, c->dStop, c->stream);
	c->launches++;
	return IPIC_OK;
}

// Halo of the current projection: the corner sums of my lowest / highest cell plane go to the slab below / above, theirs
// arrive at the places of the planes just outside my slab in the (global-sized) cellJ array -- everything the node planes
// z0 .. z0 + lz need (projectToDensityField reaches one cell down, periodically, hdr/cell.h:797-806).  The boundary planes'
// sums are complete once the boundary movers are done, so this runs under the interior mover.
int exchangeCellJHalo(ipic_ctx* c) {
	const Geometry& g = c->g;
	const size_t planeDoubles = (size_t)g.nx * g.ny * 24;
	double* myLow = c->dCellJ + (size_t)g.z0 * planeDoubles;
	double* myHigh = c->dCellJ + (size_t)(g.z0 + g.lz - 1) * planeDoubles;
	double* fromBelow = c->dCellJ + (size_t)((g.z0 - 1 + g.nz) % g.nz) * planeDoubles;
	double* fromAbove = c->dCellJ + (size_t)((g.z0 + g.lz) % g.nz) * planeDoubles;
	CK(cudaStreamWaitEvent(c->commStream, c->evPacked, 0));
	if (c->comm) {
		CommPiece sends[2] = { { myHigh, planeDoubles * sizeof(double), +1 }, { myLow, planeDoubles * sizeof(double), -1 } };
		CommPiece recvs[2] = { { fromBelow, planeDoubles * sizeof(double), -1 }, { fromAbove, planeDoubles * sizeof(double), +1 } };
		int rc = comm_sendrecv(c->comm, sends, 2, recvs, 2, c->commStream, g_lastError);
		if (rc) return rc;
	} else {
		ipic_ctx* down = neighbour(c, -1);
		ipic_ctx* up = neighbour(c, +1);
		CK(cudaStreamWaitEvent(c->commStream, down->evPacked, 0));
		CK(cudaStreamWaitEvent(c->commStream, up->evPacked, 0));
		CK(cudaMemcpyPeerAsync(fromBelow, c->device, down->dCellJ + (size_t)(down->g.z0 + down->g.lz - 1) * planeDoubles, down->device, planeDoubles * sizeof(double), c->commStream));
		CK(cudaMemcpyPeerAsync(fromAbove, c->device, up->dCellJ + (size_t)up->g.z0 * planeDoubles, up->device, planeDoubles * sizeof(double), c->commStream));
	}
	CK(cudaEventRecord(c->evJ, c->commStream));
	return IPIC_OK;
}

// in-process group: nobody may still be copying my slab of dCellJ (previous step) when I overwrite it
int waitPeersDoneWithCellJ(ipic_ctx* c) {
	if (!c->group) return IPIC_OK;
	for (ipic_ctx* m : c->group->members) if (m != c) CK(cudaStreamWaitEvent(c->stream, m->evJ, 0));
	return IPIC_OK;
}

// pack both faces of generation `out` into the send buffers and report the totals to the host (boundary stream)
int packFaces(ipic_ctx* c, int stepIndex) {
	const int out = c->cur ^ 1;
	launch_face_pack(c->g, c->buf[out], c->dCellStart, c->dLeave, c->dTag, c->dOrder, c->dDirOff, c->dFaceCnt, faceView(c, c->sendMsg[0], c->sendCap),
	                 faceView(c, c->sendMsg[1], c->sendCap), c->dTotals, c->dStop, stepIndex, c->bStream);
	c->launches += 3;
	CK(cudaMemcpyAsync(c->hTotals, c->dTotals, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->bStream));
	return IPIC_OK;
}

// the pieces of one face message that actually hold data: (column, direction) offsets and `total` particle records
struct FacePieces { void* p[2]; size_t bytes[2]; };
FacePieces piecesOf(const ipic_ctx* c, char* base, int cap, int total) {
	const FaceMsg m = faceView(c, base, cap);
	FacePieces f;
	f.p[0] = m.offsets;   f.bytes[0] = (size_t)(c->g.nx * c->g.ny * 9 + 1) * sizeof(int);
	f.p[1] = m.particles; f.bytes[1] = (size_t)total * 64;
	return f;
}

int slabPhase(ipic_ctx* c, int phase, const StepFlags& f, int stepIndex) {
	static const char* const kPhaseName[] = { "slab:pre_deposit", "slab:pre_gather_J", "slab:boundary_movers+pack", "slab:interior_mover", "slab:settle",
	                                          "slab:exchange", "slab:gather_J", "slab:import", "slab:fields" };
	NvtxRange nvtx(kPhaseName[phase]);
	CK(cudaSetDevice(c->device));
	const int lz = c->g.lz;
	switch (phase) {
	case PH_PRE_DEPOSIT:
		if (!(f.project && f.interp)) break;
		{ int rc = waitPeersDoneWithCellJ(c); if (rc) return rc; }
		depositPass(c, f, 0, lz);
		CK(cudaEventRecord(c->evMoved, c->stream));
		CK(cudaEventRecord(c->evPacked, c->stream));          // the all-gather waits for both events; nothing runs on the boundary stream here
		break;
	case PH_PRE_GATHER_J:
		if (!(f.project && f.interp)) break;
		if (f.jHalo) { int rc = exchangeCellJHalo(c); if (rc) return rc; }
		else { int rc = allGatherCellJ(c); if (rc) return rc; }
		{ int rc = finishCurrent(c, f.jHalo); if (rc) return rc; }
		break;
	case PH_BOUNDARY: {
		if (!f.fuseDeposit) fieldPhase(c, f);
		if (f.fuseDeposit) { int rc = waitPeersDoneWithCellJ(c); if (rc) return rc; }
		// everything queued on the main stream so far (previous imports, field phase) comes first
		spanBegin(c, 0, &c->moverSpanEnd);
		CK(cudaEventRecord(c->evMainReady, c->stream));
		CK(cudaStreamWaitEvent(c->bStream, c->evMainReady, 0));
		if (c->group && f.migrate) {
			// the neighbours pulled last step's messages out of my send buffers on their comm streams
			CK(cudaStreamWaitEvent(c->bStream, neighbour(c, -1)->evRecv, 0));
			CK(cudaStreamWaitEvent(c->bStream, neighbour(c, +1)->evRecv, 0));
		}
		if (lz > 1 && !c->chunked) {
			c->launches += launch_move(moveArgs(c, f, 0, 2, lz - 1), c->bStream);      // both boundary planes in one launch
		} else {
			c->launches += launch_move(moveArgs(c, f, 0, 1), c->bStream);
			if (lz > 1) c->launches += launch_move(moveArgs(c, f, lz - 1, 1), c->bStream);
		}
		if (f.migrate) { int rc = packFaces(c, stepIndex); if (rc) return rc; }
		CK(cudaEventRecord(c->evPacked, c->bStream));
		break;
	}
	case PH_INTERIOR:
		if (lz > 2) c->launches += launch_move(moveArgs(c, f, 1, lz - 2), c->stream);
		if (f.fuseDeposit) CK(cudaEventRecord(c->evMoved, c->stream));
		break;
	case PH_EXCHANGE: {
		if (!f.migrate) break;
		const int* t = c->hTotals;                            // settled: [0] my down face, [1] my up face, [2] from below, [3] from above
		const FacePieces sUp = piecesOf(c, c->sendMsg[1], c->sendCap, t[1]), sDown = piecesOf(c, c->sendMsg[0], c->sendCap, t[0]);
		const FacePieces rDown = piecesOf(c, c->recvMsg[0], c->recvCap, t[2]), rUp = piecesOf(c, c->recvMsg[1], c->recvCap, t[3]);
		CK(cudaStreamWaitEvent(c->commStream, c->evPacked, 0));
		if (c->comm) {
			CommPiece sends[4], recvs[4];
			for (int k = 0; k < 2; k++) {
				sends[k] = { sUp.p[k], sUp.bytes[k], +1 };   sends[2 + k] = { sDown.p[k], sDown.bytes[k], -1 };
				recvs[k] = { rDown.p[k], rDown.bytes[k], -1 }; recvs[2 + k] = { rUp.p[k], rUp.bytes[k], +1 };
			}
			int rc = comm_sendrecv(c->comm, sends, 4, recvs, 4, c->commStream, g_lastError);
			if (rc) return rc;
		} else {
			ipic_ctx* down = neighbour(c, -1);
			ipic_ctx* up = neighbour(c, +1);
			CK(cudaStreamWaitEvent(c->commStream, down->evPacked, 0));
			CK(cudaStreamWaitEvent(c->commStream, up->evPacked, 0));
			const FacePieces fromDown = piecesOf(down, down->sendMsg[1], down->sendCap, t[2]), fromUp = piecesOf(up, up->sendMsg[0], up->sendCap, t[3]);
			for (int k = 0; k < 2; k++) {
				if (rDown.bytes[k]) CK(cudaMemcpyPeerAsync(rDown.p[k], c->device, fromDown.p[k], down->device, rDown.bytes[k], c->commStream));
				if (rUp.bytes[k]) CK(cudaMemcpyPeerAsync(rUp.p[k], c->device, fromUp.p[k], up->device, rUp.bytes[k], c->commStream));
			}
		}
		CK(cudaEventRecord(c->evRecv, c->commStream));
		break;
	}
	case PH_GATHER_J:
		if (!f.fuseDeposit) break;
		if (f.jHalo) { int rc = exchangeCellJHalo(c); if (rc) return rc; }      // two planes, on the comm stream under the interior mover
		else { int rc = allGatherCellJ(c); if (rc) return rc; }                 // everything, under the imports
		break;
	case PH_IMPORT:
		CK(cudaStreamWaitEvent(c->stream, c->evPacked, 0));   // the boundary planes' movers
		spanEnd(c, &c->moverSpanEnd);
		if (f.migrate) {
			{
				ProfSpan span(c, 1);
				importPlanes(c, stepIndex, 1, lz - 2, false);
			}
			{
				ProfSpan wait(c, 2);
				CK(cudaStreamWaitEvent(c->stream, c->evRecv, 0));
			}
			ProfSpan span(c, 1);
			if (lz > 1) importPlanes(c, stepIndex, 0, 2, true, lz - 1);      // both boundary planes in one launch
			else importPlanes(c, stepIndex, 0, 1, true);
		} else {
			CK(cudaMemcpyAsync(c->dCount, c->dStay, (size_t)c->g.ncellsLocal * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
		}
		c->cur ^= 1;
		CK(cudaGetLastError());
		break;
	case PH_FIELDS:
		if (!f.fuseDeposit) break;
		{
			ProfSpan wait(c, 3);
			int rc = finishCurrent(c, f.jHalo);
			if (rc) return rc;
			fieldPhase(c, f);
		}
		break;
	default: break;
	}
	return IPIC_OK;
}

int enqueueStep(ipic_ctx* c, uint32_t flags, int stepIndex) {
	return enqueueStepSingle(c, StepFlags(flags), stepIndex);      // slab contexts are driven by runSlabSteps
}

int enqueueSteps(ipic_ctx* c, uint64_t first, uint64_t count, uint32_t flags) {
	for (uint64_t s = first; s < first + count; s++) {
		int rc = enqueueStep(c, flags, (int)std::min<uint64_t>(s, 0x7ffffff0u));
		if (rc) return rc;
	}
	return IPIC_OK;
}

int stopError(ipic_ctx* c, const int* stop) {
	if (stop[3] == 2) return fail(IPIC_ERR_INTERNAL, "rank %d: unusable face message received in step %d", c->rank, stop[1]);
	return fail(IPIC_ERR_CAPACITY, "rank %d: cell segment overflow in a multi-GPU run (step %d) could not be repaired", c->rank, stop[1]);
}

// Did the slab step whose PH_IMPORT was enqueued last complete without overflowing a cell segment?  If it
// did overflow: the generation being written still holds every stayer and every outbox, and the received
// face messages are untouched, so this rank re-layouts on its own and redoes that import; nothing of the
// NEXT step has left this GPU yet (see runSlabSteps), so the other ranks never see the hiccup.
int confirmSlabStep(ipic_ctx* c, bool* recovered) {
	*recovered = false;
	CK(cudaSetDevice(c->device));
	CK(cudaEventSynchronize(c->evStepDone));
	// (the sticky word is never cleared on slabs: a slab that once needed the sort pass keeps it)
	c->sortLikely = c->hStop[5] != 0;
	if (!c->hStop[0]) return IPIC_OK;
	if (c->hStop[3]) return stopError(c, c->hStop);
	const int failed = c->hStop[1];
	CK(cudaStreamSynchronize(c->stream));                    // drain the (skipped) launches queued behind the failure
	CK(cudaStreamSynchronize(c->bStream));
	int rc = relayoutAfterOverflow(c, c->cur, failed);
	if (rc) return rc;
	c->relayouts++;
	*recovered = true;
	return IPIC_OK;
}

// SETTLE, part 1: my packed totals are on the host; a face that does not fit its buffer gets a bigger one and is packed again
int settleSendSizes(ipic_ctx* c, int stepIndex) {
	CK(cudaSetDevice(c->device));
	CK(cudaEventSynchronize(c->evPacked));
	const int need = std::max(c->hTotals[0], c->hTotals[1]);
	if (need <= c->sendCap) return IPIC_OK;
	const int64_t cap = alignUp((int64_t)need + need / 2 + 1024, 1024);
	if (c->userMsgCap > 0) return fail(IPIC_ERR_CAPACITY, "rank %d: %d particles cross a slab face in step %d, exchange capacity fixed at %lld by ipic_set_exchange_capacity",
	                                   c->rank, need, stepIndex, (long long)c->userMsgCap);
	CK(cudaStreamSynchronize(c->commStream));               // nobody reads the old buffers any more (previous step confirmed)
	int rc = allocFaceBuffers(c, c->sendMsg, cap);
	if (rc) return rc;
	c->sendCap = (int)cap;
	rc = packFaces(c, stepIndex);
	if (rc) return rc;
	CK(cudaEventRecord(c->evPacked, c->bStream));
	CK(cudaEventSynchronize(c->evPacked));
	return IPIC_OK;
}

// SETTLE, part 2 (NCCL): swap totals with the z-neighbours; part 3: make room for what is coming
int settleRecvSizesNccl(ipic_ctx* c) {
	CommPiece sends[2] = { { c->dTotals + 1, sizeof(int), +1 }, { c->dTotals + 0, sizeof(int), -1 } };
	CommPiece recvs[2] = { { c->dTotals + 2, sizeof(int), -1 }, { c->dTotals + 3, sizeof(int), +1 } };
	int rc = comm_sendrecv(c->comm, sends, 2, recvs, 2, c->commStream, g_lastError);
	if (rc) return rc;
	CK(cudaMemcpyAsync(c->hTotals + 2, c->dTotals + 2, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->commStream));
	CK(cudaStreamSynchronize(c->commStream));
	return IPIC_OK;
}

int growRecvBuffers(ipic_ctx* c) {
	const int need = std::max(c->hTotals[2], c->hTotals[3]);
	if (need <= c->recvCap) return IPIC_OK;
	CK(cudaSetDevice(c->device));
	const int64_t cap = alignUp((int64_t)need + need / 2 + 1024, 1024);
	int rc = allocFaceBuffers(c, c->recvMsg, cap);         // the previous step (its imports read the old buffers) is confirmed
	if (rc) return rc;
	c->recvCap = (int)cap;
	return IPIC_OK;
}

bool phaseCommunicates(int ph, const StepFlags& f) {
	switch (ph) {
	case PH_PRE_GATHER_J: return f.project && f.interp;
	case PH_SETTLE: return f.migrate;
	case PH_GATHER_J: return f.fuseDeposit;
	default: return false;
	}
}

// Steps of slab contexts (one NCCL rank, or all members of an in-process group).  The host runs at most
// one step ahead of the device: before the first communication of step s is enqueued, step s-1 is
// confirmed (and repaired if a cell segment overflowed).  The launches of step s queued before that
// point (all movers + pack, or the projection pass) keep the GPU busy meanwhile.
int runSlabSteps(ipic_ctx** ctxs, int n, uint64_t numSteps, uint32_t flags, double* tFirst, double* tRest) {
	const StepFlags f(flags);
	std::vector<cudaEvent_t> ev((size_t)n * 3);
	for (int r = 0; r < n; r++) {
		CK(cudaSetDevice(ctxs[r]->device));
		for (int k = 0; k < 3; k++) CK(cudaEventCreate(&ev[(size_t)r * 3 + k]));
		CK(cudaEventRecord(ev[(size_t)r * 3], ctxs[r]->stream));
	}
	auto confirmAll = [&](int stepIndex, int uptoPhase) -> int {
		for (int r = 0; r < n; r++) {
			bool recovered = false;
			int rc = confirmSlabStep(ctxs[r], &recovered);
			if (rc) return rc;
			if (recovered && f.fuseDeposit) {
				// PH_FIELDS of the failed step was queued behind its import and did nothing (every kernel returns on the stop
				// flag): the all-gathered corner sums of that step are still in dCellJ (the next step's movers, which would
				// overwrite them, were skipped as well), so build J and advance the fields now, before those movers run again
				rc = finishCurrent(ctxs[r], f.jHalo); if (rc) return rc;
				fieldPhase(ctxs[r], f);
			}
			// the launches of the current step that were queued behind the failure did nothing: queue them again
			if (recovered) for (int q = 0; q < uptoPhase; q++) { rc = slabPhase(ctxs[r], q, f, stepIndex); if (rc) return rc; }
		}
		return IPIC_OK;
	};
	auto settle = [&](int stepIndex) -> int {
		for (int r = 0; r < n; r++) { int rc = settleSendSizes(ctxs[r], stepIndex); if (rc) return rc; }
		for (int r = 0; r < n; r++) {
			ipic_ctx* c = ctxs[r];
			if (c->comm) { int rc = settleRecvSizesNccl(c); if (rc) return rc; }
			else { c->hTotals[2] = neighbour(c, -1)->hTotals[1]; c->hTotals[3] = neighbour(c, +1)->hTotals[0]; }
			int rc = growRecvBuffers(c);
			if (rc) return rc;
		}
		return IPIC_OK;
	};
	for (uint64_t s = 0; s < numSteps; s++) {
		const int stepIndex = (int)std::min<uint64_t>(s, 0x7ffffff0u);
		bool confirmed = s == 0;
		for (int ph = 0; ph < PH_COUNT; ph++) {
			if (!confirmed && (phaseCommunicates(ph, f) || ph == PH_IMPORT)) {
				int rc = confirmAll(stepIndex, ph);
				if (rc) return rc;
				confirmed = true;
			}
			if (ph == PH_SETTLE) {
				if (f.migrate) { int rc = settle(stepIndex); if (rc) return rc; }
				continue;
			}
			for (int r = 0; r < n; r++) { int rc = slabPhase(ctxs[r], ph, f, stepIndex); if (rc) return rc; }
		}
		for (int r = 0; r < n; r++) {
			ipic_ctx* c = ctxs[r];
			CK(cudaSetDevice(c->device));
			if (s == 0) CK(cudaEventRecord(ev[(size_t)r * 3 + 1], c->stream));
			CK(cudaMemcpyAsync(c->hStop, c->dStop, 8 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
			CK(cudaEventRecord(c->evStepDone, c->stream));
		}
	}
	{ int rc = confirmAll(0, 0); if (rc) return rc; }
	float msFirst = 0, msRest = 0;
	for (int r = 0; r < n; r++) {
		ipic_ctx* c = ctxs[r];
		CK(cudaSetDevice(c->device));
		CK(cudaEventRecord(ev[(size_t)r * 3 + 2], c->stream));
		CK(cudaStreamSynchronize(c->stream));
		CK(cudaStreamSynchronize(c->bStream));
		CK(cudaStreamSynchronize(c->commStream));
		float a = 0, b = 0;
		CK(cudaEventElapsedTime(&a, ev[(size_t)r * 3], ev[(size_t)r * 3 + 1]));
		CK(cudaEventElapsedTime(&b, ev[(size_t)r * 3 + 1], ev[(size_t)r * 3 + 2]));
		msFirst = std::max(msFirst, a); msRest = std::max(msRest, b);
		if (c->profiling) { int rc = drainProfile(c); if (rc) return rc; }
		for (int k = 0; k < 3; k++) cudaEventDestroy(ev[(size_t)r * 3 + k]);
	}
	if (tFirst) *tFirst = msFirst / 1000.0;
	if (tRest) *tRest = msRest / 1000.0;
	return IPIC_OK;
}

} // namespace

static int checkStepArgs(ipic_ctx* c, uint32_t flags) {
	if (!c) return fail(IPIC_ERR_ARG, "null context");
	if (!c->haveCells) return fail(IPIC_ERR_ARG, "ipic_upload_cells has not been called");
	if ((flags & IPIC_SOLVER_FORWARD) && (flags & IPIC_SOLVER_LEAPFROG)) return fail(IPIC_ERR_ARG, "choose one field solver");
	if ((flags & IPIC_UPDATE_BOUNDARIES) && !(c->g.nx == c->g.ny && c->g.ny == c->g.nz)) return fail(IPIC_ERR_ARG, "updateFieldsOnBoundaries assumes a cubic universe (reference field.h:422)");
	if (c->nranks > 1 && !c->comm && !c->group) return fail(IPIC_ERR_COMM, "slab context %d/%d has no communicator (ipic_comm_init / ipic_group_create)", c->rank, c->nranks);
	return IPIC_OK;
}

static int checkSoloStep(ipic_ctx* c, uint32_t flags) {
	int rc = checkStepArgs(c, flags);
	if (rc) return rc;
	if (c->group) return fail(IPIC_ERR_ARG, "context %d/%d belongs to an in-process group: step it with ipic_group_step", c->rank, c->nranks);
	if (c->nranks > 1) return negotiateExchangeNccl(c);
	return IPIC_OK;
}

// Waits for a batch of single-GPU steps whose first step started from generation `genAtStart`, and repairs it if an
// import ran out of room: re-layout, redo that import, enqueue the steps that were skipped behind it, and so on until
// all `numSteps` are done.  msFirst / msRest (may be null) accumulate the device time between the three events.
static int finishSingleBatch(ipic_ctx* c, int genAtStart, uint64_t numSteps, uint32_t flags, cudaEvent_t* ev, float* msFirst, float* msRest) {
	bool firstBatch = true;
	for (;;) {
		CK(cudaStreamSynchronize(c->stream));
		if (ev) {
			float a = 0, b = 0;
			CK(cudaEventElapsedTime(&a, ev[0], ev[1]));
			CK(cudaEventElapsedTime(&b, ev[1], ev[2]));
			if (firstBatch) { *msFirst += a; *msRest += b; } else { *msRest += a + b; }      // only the very first step counts as "first"
		}
		firstBatch = false;
		if (c->profiling) { int rc = drainProfile(c); if (rc) return rc; }
		int stop[8];
		CK(cudaMemcpy(stop, c->dStop, sizeof stop, cudaMemcpyDeviceToHost));
		c->sortLikely = stop[5] != 0;
		if (stop[5]) CK(cudaMemsetAsync(c->dStop + 5, 0, sizeof(int), c->stream));      // ordered before the next batch's movers
		if (!stop[0]) return IPIC_OK;
		// import of step `stop[1]` of this batch overflowed; later steps were skipped on the device
		const int failed = stop[1];
		const int outGen = genAtStart ^ ((failed + 1) & 1);
		int rc = relayoutAfterOverflow(c, outGen, failed);
		if (rc) return rc;
		c->relayouts++;
		c->cur = outGen;
		if ((uint64_t)failed + 1 >= numSteps) return IPIC_OK;
		numSteps -= (uint64_t)failed + 1;
		genAtStart = c->cur;
		if (ev) CK(cudaEventRecord(ev[0], c->stream));
		rc = enqueueSteps(c, 0, 1, flags);
		if (rc) return rc;
		if (ev) CK(cudaEventRecord(ev[1], c->stream));
		rc = enqueueSteps(c, 1, numSteps - 1, flags);
		if (rc) return rc;
		if (ev) CK(cudaEventRecord(ev[2], c->stream));
	}
}

// A batch enqueued by ipic_step_async is settled before anything else looks at / changes the context's state.
static int settleAsync(ipic_ctx* c) {
	if (!c || c->asyncSteps == 0) return IPIC_OK;
	CK(cudaSetDevice(c->device));
	const uint64_t n = c->asyncSteps;
	c->asyncSteps = 0;
	return finishSingleBatch(c, c->asyncGenAtStart, n, c->asyncFlags, nullptr, nullptr, nullptr);
}

int ipic_step_async(ipic_ctx* c, uint64_t numSteps, uint32_t flags) {
	int rc = checkSoloStep(c, flags);
	if (rc) return rc;
	if (c->nranks > 1) return fail(IPIC_ERR_ARG, "ipic_step_async is for single-GPU contexts: slab steps confirm every step on the host (ipic_step)");
	CK(cudaSetDevice(c->device));
	if (c->asyncSteps && c->asyncFlags != flags) { rc = settleAsync(c); if (rc) return rc; }
	if (c->asyncSteps == 0) { c->asyncGenAtStart = c->cur; c->asyncFlags = flags; }
	// steps carry their index within the pending batch so that ipic_sync can locate an overflow
	rc = enqueueSteps(c, c->asyncSteps, numSteps, flags);
	if (rc) return rc;
	c->asyncSteps += numSteps;
	return IPIC_OK;
}

int ipic_sync(ipic_ctx* c) {
	if (!c) return fail(IPIC_ERR_ARG, "null context");
	CK(cudaSetDevice(c->device));
	int rc = settleAsync(c);
	if (rc) return rc;
	CK(cudaStreamSynchronize(c->stream));
	return IPIC_OK;
}

int ipic_step(ipic_ctx* c, uint64_t numSteps, uint32_t flags, double* tFirst, double* tRest) {
	NvtxRange nvtx("ipic_step");
	if (tFirst) *tFirst = 0;
	if (tRest) *tRest = 0;
	{ int rc = checkSoloStep(c, flags); if (rc) return rc; }
	if (numSteps == 0) return IPIC_OK;
	CK(cudaSetDevice(c->device));
	{ int rc = settleAsync(c); if (rc) return rc; }
	if (c->nranks > 1) return runSlabSteps(&c, 1, numSteps, flags, tFirst, tRest);
	cudaEvent_t ev[3];
	for (auto& e : ev) CK(cudaEventCreate(&e));
	float msFirst = 0, msRest = 0;
	{
		const int genAtStart = c->cur;
		CK(cudaEventRecord(ev[0], c->stream));
		int rc = enqueueSteps(c, 0, 1, flags);
		if (rc) return rc;
		CK(cudaEventRecord(ev[1], c->stream));
		// the remaining steps carry their index within this batch so that an overflow can be located
		rc = enqueueSteps(c, 1, numSteps - 1, flags);
		if (rc) return rc;
		CK(cudaEventRecord(ev[2], c->stream));
		rc = finishSingleBatch(c, genAtStart, numSteps, flags, ev, &msFirst, &msRest);
		if (rc) return rc;
	}
	for (auto& e : ev) cudaEventDestroy(e);
	if (tFirst) *tFirst = msFirst / 1000.0;
	if (tRest) *tRest = msRest / 1000.0;
	return IPIC_OK;
}

int ipic_group_create(const ipic_props* props, const int* devices, int ndev, ipic_ctx** out) {
	if (!props || ndev < 1) return fail(IPIC_ERR_ARG, "bad group arguments");
	if (props->size[2] % ndev != 0) return fail(IPIC_ERR_ARG, "nz=%lld is not divisible by %d slabs", (long long)props->size[2], ndev);
	const int lz = (int)(props->size[2] / ndev);
	std::vector<int> z0((size_t)ndev), n((size_t)ndev, lz);
	for (int r = 0; r < ndev; r++) z0[(size_t)r] = r * lz;
	return ipic_group_create_ranges(props, devices, ndev, z0.data(), n.data(), out);
}

int ipic_group_create_ranges(const ipic_props* props, const int* devices, int ndev, const int* zBegin, const int* zCount, ipic_ctx** out) {
	if (!props || !devices || !out || ndev < 1 || !zBegin || !zCount) return fail(IPIC_ERR_ARG, "bad group arguments");
	auto grp = std::make_shared<ipic_group>();
	for (int r = 0; r < ndev; r++) {
		ipic_ctx* c = nullptr;
		int rc = ipic_create_slab_range(props, devices[r], r, ndev, zBegin[r], zCount[r], &c);
		if (rc) { for (ipic_ctx* m : grp->members) { m->group.reset(); ipic_destroy(m); } return rc; }
		grp->members.push_back(c);
	}
	for (int r = 0; r < ndev; r++) {
		if (ndev > 1) grp->members[(size_t)r]->group = grp;
		out[r] = grp->members[(size_t)r];
	}
	// direct peer copies between distinct devices where the hardware allows it (cudaMemcpyPeerAsync works either way)
	for (int a = 0; a < ndev; a++) for (int b = 0; b < ndev; b++) {
		if (devices[a] == devices[b]) continue;
		int can = 0;
		if (cudaDeviceCanAccessPeer(&can, devices[a], devices[b]) == cudaSuccess && can) { cudaSetDevice(devices[a]); cudaDeviceEnablePeerAccess(devices[b], 0); }
	}
	cudaGetLastError();      // "peer access already enabled" is not an error
	return IPIC_OK;
}

int ipic_group_step(ipic_ctx** ctxs, int n, uint64_t numSteps, uint32_t flags, double* tFirst, double* tRest) {
	NvtxRange nvtx("ipic_group_step");
	if (tFirst) *tFirst = 0;
	if (tRest) *tRest = 0;
	if (!ctxs || n < 1 || !ctxs[0]) return fail(IPIC_ERR_ARG, "bad group arguments");
	if (n == 1) return ipic_step(ctxs[0], numSteps, flags, tFirst, tRest);
	ipic_group* grp = ctxs[0]->group.get();
	if (!grp || (int)grp->members.size() != n) return fail(IPIC_ERR_ARG, "the contexts are not the %d members of one group", n);
	for (int r = 0; r < n; r++) {
		if (ctxs[r] != grp->members[(size_t)r] || !ctxs[r]) return fail(IPIC_ERR_ARG, "the contexts are not the members of one group in rank order");
		int rc = checkStepArgs(ctxs[r], flags);
		if (rc) return rc;
	}
	if (numSteps == 0) return IPIC_OK;
	{ int rc = negotiateExchangeGroup(grp); if (rc) return rc; }
	return runSlabSteps(ctxs, n, numSteps, flags, tFirst, tRest);
}

int ipic_count_misplaced(ipic_ctx* c, uint64_t* misplaced) {
	if (!c || !misplaced) return fail(IPIC_ERR_ARG, "null argument");
	CK(cudaSetDevice(c->device));
	{ int rcA = settleAsync(c); if (rcA) return rcA; }
	unsigned long long* d = nullptr;
	CK(cudaMalloc(&d, sizeof(unsigned long long)));
	CK(cudaMemsetAsync(d, 0, sizeof(unsigned long long), c->stream));
	launch_count_misplaced(c->g, c->buf[c->cur], c->dCellStart, c->dCount, d, c->stream);
	c->launches++;
	unsigned long long h = 0;
	CK(cudaMemcpyAsync(&h, d, sizeof h, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	cudaFree(d);
	*misplaced = h;
	return IPIC_OK;
}

int ipic_diagnostics(ipic_ctx* c, double* out4) {
	if (!c || !out4) return fail(IPIC_ERR_ARG, "null argument");
	CK(cudaSetDevice(c->device));
	{ int rcA = settleAsync(c); if (rcA) return rcA; }
	const int64_t pairs = (int64_t)std::max(1, c->g.ncellsLocal) + kFieldEnergyChunks + 2;
	int rc = ensureStage(c, 2 * pairs);
	if (rc) return rc;
	double* cellPart = c->dStage;
	double* fieldPart = cellPart + 2 * (int64_t)std::max(1, c->g.ncellsLocal);
	double* res = fieldPart + 2 * kFieldEnergyChunks;                  // { KE, momentum, E energy, B energy }
	if (c->haveCells) launch_particle_energy(c->g, c->buf[c->cur], c->dCellStart, c->dCount, cellPart, res, c->stream);
	else CK(cudaMemsetAsync(res, 0, 2 * sizeof(double), c->stream));
	launch_field_energy(c->g, c->dEB, c->dBext, fieldPart, kFieldEnergyChunks, res + 2, c->stream);
	c->launches += c->haveCells ? 4 : 2;
	double h[4];
	CK(cudaMemcpyAsync(h, res, sizeof h, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	out4[0] = h[1]; out4[1] = h[2]; out4[2] = h[3]; out4[3] = h[0];
	return IPIC_OK;
}

int ipic_group_diagnostics(ipic_ctx** ctxs, int n, double* out4) {
	if (!ctxs || n < 1 || !out4) return fail(IPIC_ERR_ARG, "bad group arguments");
	double mom = 0.0, ke = 0.0;
	for (int r = 0; r < n; r++) {
		double d[4];
		int rc = ipic_diagnostics(ctxs[r], d);
		if (rc) return rc;
		mom += d[0]; ke += d[3];
		if (r == 0) { out4[1] = d[1]; out4[2] = d[2]; }
	}
	out4[0] = mom; out4[3] = ke;
	return IPIC_OK;
}

int ipic_trace(const ipic_props* props, int device, const double* aos, uint64_t n, int timesteps, int frameInterval, int numFrames,
               uint32_t* counts, double* seconds) {
	if (seconds) *seconds = 0;
	if (!props || !counts || (n && !aos)) return fail(IPIC_ERR_ARG, "null argument");
	if (timesteps < 0 || frameInterval < 1 || numFrames < 1) return fail(IPIC_ERR_ARG, "bad tracer arguments (timesteps %d, frame interval %d, frames %d)", timesteps, frameInterval, numFrames);
	if (timesteps > 0 && (timesteps - 1) / frameInterval >= numFrames) return fail(IPIC_ERR_ARG, "%d steps at one frame per %d steps need %d frames, %d offered", timesteps, frameInterval, (timesteps - 1) / frameInterval + 1, numFrames);
	for (int a = 0; a < 3; a++) if (props->size[a] <= 0 || props->size[a] > 4096 || !(props->cellWidth[a] > 0)) return fail(IPIC_ERR_ARG, "bad grid");
	if (!(props->dt > 0) || !(props->speedOfLight > 0) || !(props->planetRadius >= 0)) return fail(IPIC_ERR_ARG, "dt, speedOfLight must be positive and planetRadius non-negative");
	int ndev = 0;
	CK(cudaGetDeviceCount(&ndev));
	if (device < 0 || device >= ndev) return fail(IPIC_ERR_CUDA, "CUDA device %d not available (%d devices)", device, ndev);
	CK(cudaSetDevice(device));
	ipic_ctx geo;                                      // only to derive the Geometry the kernels take
	geo.props = *props;
	fillGeometry(&geo);
	const size_t cells = (size_t)numFrames * (size_t)geo.g.nx * geo.g.ny * geo.g.nz;
	const uint64_t chunk = 1u << 22;                   // particles per upload: the next chunk crosses PCIe while this one is traced
	unsigned* dCounts = nullptr;
	double* dStage[2] = { nullptr, nullptr };
	cudaStream_t st[2] = { nullptr, nullptr };
	cudaEvent_t ev[2] = { nullptr, nullptr };
	auto cleanup = [&]() { cudaFree(dCounts); for (int i = 0; i < 2; i++) { cudaFree(dStage[i]); if (st[i]) cudaStreamDestroy(st[i]); if (ev[i]) cudaEventDestroy(ev[i]); } };
	#define TCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return fail(IPIC_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } } while (0)
	TCK(cudaMalloc(&dCounts, std::max<size_t>(cells, 1) * sizeof(unsigned)));
	TCK(cudaMemset(dCounts, 0, std::max<size_t>(cells, 1) * sizeof(unsigned)));
	for (int i = 0; i < 2; i++) {
		TCK(cudaMalloc(&dStage[i], (size_t)std::min<uint64_t>(std::max<uint64_t>(n, 1), chunk) * 64));
		TCK(cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking));
		TCK(cudaEventCreate(&ev[i]));
	}
	TCK(cudaEventRecord(ev[0], st[0]));
	int k = 0;
	for (uint64_t first = 0; first < n; first += chunk, k ^= 1) {
		const uint64_t cnt = std::min(chunk, n - first);
		TCK(cudaMemcpyAsync(dStage[k], aos + first * 8, (size_t)cnt * 64, cudaMemcpyHostToDevice, st[k]));
		launch_trace(geo.g, dStage[k], (int64_t)cnt, timesteps, frameInterval, dCounts, st[k]);
	}
	TCK(cudaStreamSynchronize(st[0]));
	TCK(cudaStreamSynchronize(st[1]));
	TCK(cudaEventRecord(ev[1], st[0]));
	TCK(cudaEventSynchronize(ev[1]));
	float ms = 0;
	TCK(cudaEventElapsedTime(&ms, ev[0], ev[1]));
	TCK(cudaGetLastError());
	TCK(cudaMemcpy(counts, dCounts, cells * sizeof(unsigned), cudaMemcpyDeviceToHost));
	#undef TCK
	cleanup();
	if (seconds) *seconds = ms / 1000.0;
	return IPIC_OK;
}

int ipic_selftest(int device, uint64_t n, uint64_t seed, uint64_t* mismatches) {
	if (!mismatches) return fail(IPIC_ERR_ARG, "null argument");
	int ndev = 0;
	CK(cudaGetDeviceCount(&ndev));
	if (device < 0 || device >= ndev) return fail(IPIC_ERR_CUDA, "CUDA device %d not available (%d devices)", device, ndev);
	CK(cudaSetDevice(device));
	unsigned long long* d = nullptr;
	CK(cudaMalloc(&d, sizeof(unsigned long long)));
	CK(cudaMemset(d, 0, sizeof(unsigned long long)));
	launch_selftest_division(n, seed, d, nullptr);
	unsigned long long h = 0;
	CK(cudaMemcpy(&h, d, sizeof h, cudaMemcpyDeviceToHost));
	cudaFree(d);
	*mismatches = h;
	return IPIC_OK;
}

int ipic_profile_reset(ipic_ctx* c) {
	if (!c) return fail(IPIC_ERR_ARG, "null context");
	c->profiling = true; c->moverMs = 0; c->importMs = 0; c->launches = 0; c->eventsUsed = 0; c->eventKind.clear();
	for (double& v : c->phaseMs) v = 0;
	return IPIC_OK;
}

int ipic_profile_get(ipic_ctx* c, double* moverMs, double* importMs, uint64_t* launches) {
	if (!c) return fail(IPIC_ERR_ARG, "null context");
	CK(cudaSetDevice(c->device));
	CK(cudaStreamSynchronize(c->stream));
	int rc = drainProfile(c);
	if (rc) return rc;
	if (moverMs) *moverMs = c->moverMs;
	if (importMs) *importMs = c->importMs;
	if (launches) *launches = c->launches;
	return IPIC_OK;
}

int ipic_relayout_count(ipic_ctx* c, uint64_t* count) {
	if (!c || !count) return fail(IPIC_ERR_ARG, "null argument");
	*count = c->relayouts;
	return IPIC_OK;
}

int ipic_profile_phases(ipic_ctx* c, double* ms4) {
	if (!c || !ms4) return fail(IPIC_ERR_ARG, "null argument");
	CK(cudaSetDevice(c->device));
	CK(cudaStreamSynchronize(c->stream));
	int rc = drainProfile(c);
	if (rc) return rc;
	for (int i = 0; i < 4; i++) ms4[i] = c->phaseMs[i];
	return IPIC_OK;
}

int ipic_comm_unique_id(void* id128) { return comm_unique_id(id128, g_lastError); }

int ipic_comm_init(ipic_ctx* c, const void* id128) {
	if (!c || !id128) return fail(IPIC_ERR_ARG, "null argument");
	CK(cudaSetDevice(c->device));
	return comm_create(&c->comm, id128, c->rank, c->nranks, g_lastError);
}

} // extern "C"
