// Host-side launch interface of particle_kernels.cu.
#pragma once
#include "ipic_device.cuh"

namespace ipic {

struct MoveArgs {
	Geometry g;
	ParticleArrays in, out;
	const int64_t* cellStart;
	const int* countIn;
	int* countOut;             // number of stayers per cell (head of the segment in `out`)
	int* nLeave;
	unsigned char* tag;
	unsigned short* order;     // per slot: direction-sorted index of the cell's outbox (sort_outbox), written by the mover
	int* dirOff;               // per local cell: 28 exclusive prefix counts over the direction codes (-1: no index)
	const double* fieldEB;     // (n+3)^3 x 6 : E.xyz B.xyz per node (mode "interpolated")
	double* cellJ;             // cells x 24 corner sums (deposit)
	const int* stopFlag;
	int zBegin, zCount;        // owned z-planes zBegin, zBegin + zStride, ... (zCount of them) of the slab
	int zStride;               // 0 / 1: consecutive planes; lz - 1 with zCount 2: the slab's two boundary planes in one launch
	int stages;                // shared-memory tile stages per CTA (TMA ring)
	bool interpolate, deposit, migrate, push;   // push=false: deposit-only pass (no write-back)
	bool sortPass;             // follow the mover with k_sort_outbox (worth it when outboxes of more than 32 entries are expected)
	// chunked mover (several CTAs per cell, k_move_chunk): the chunks [chunkFirst, chunkFirst + chunkCount) of the table
	bool chunked;
	const int2* chunkTable;    // per chunk: (local cell id, index of the chunk within the cell)
	int chunkFirst, chunkCount;
	unsigned* chunkTicket;     // hands the chunks out in table order
	unsigned long long* chunkStatus;   // one look-back word per chunk
	double* chunkJ;            // 24 deposit partials per chunk (cells of several chunks)
};

// particles per chunk of the chunked mover
int chunk_particles();

// Index of the BIG outboxes of a step (cells of a single-GPU universe whose segment holds tens of thousands of slots): their
// entries are counted, ranked and copied by many warps (k_big_hist / k_big_scan / k_big_copy) instead of by the one warp
// that imports the target cell.  nBig = 0: no such cells, nothing is launched.
struct BigOutboxes {
	int nBig = 0, totalSegs = 0;
	const int* cell = nullptr;        // [nBig] local cell id
	const int* segFirst = nullptr;    // [nBig + 1] first entry of each big cell in segTable
	const int2* segTable = nullptr;   // [totalSegs] (big id, segment of big_outbox_segment() outbox entries within the cell)
	int* rank = nullptr;              // [totalSegs * 27] per cell, tag-major: entries with tag t in segment s; after the scan: in the segments before s
	int* off = nullptr;               // [nBig * 28] exclusive prefix of the outbox's entries over the 27 tags
	long long* dst = nullptr;         // [nBig * 27] first slot of the target cell's arrivals with (big cell, tag); written by the import
	long long* room = nullptr;        // [nBig * 27] end of the slots that target may use
};

struct ImportArgs {
	Geometry g;
	ParticleArrays src; const int64_t* srcStart; const int* srcLeave; const unsigned char* tag;
	const unsigned short* order; const int* dirOff;   // the movers' direction-sorted outbox index (slots of srcStart's layout)
	ParticleArrays dst; const int64_t* dstStart; const int* stay; int* dstCount;
	int reserveOwnOutbox;      // 1 while other cells may still read this cell's outbox out of `dst`
	int* stopFlag;             // [0] stop, [1] first failing step, [2] overflow seen in this pass
	int stepIndex;
	int zBegin, zCount, zStride;   // planes zBegin, zBegin + zStride, ... (zStride 0 / 1: consecutive)
	bool slab;                 // z-slab context: neighbours beyond the slab's first / last plane are columns of the face messages
	FaceMsg fromDown, fromUp;
	BigOutboxes big;           // outboxes indexed by launch_big_index in this step (their dirOff carries the marker)
};

// returns the number of kernels launched (the mover; after a migrating push also k_sort_outbox, which indexes the outboxes
// the mover only marked).  stopFlag points at the context's 8 control words: [0] stop ... [4] "an outbox awaits sorting"
int launch_move(const MoveArgs& a, cudaStream_t st);
// does launch_move run the warp-specialised kernel (k_move_ws: a fifth warp deposits) for these switches?
bool move_uses_ws(const Geometry& g, bool deposit, bool interpolate, bool push);
// `commit`: also raise the stop flag if any import pass of this step overflowed a cell segment; returns the number of kernels launched
int launch_import(const ImportArgs& a, bool commit, cudaStream_t st);
// outbox entries per segment of the big-outbox index
int big_outbox_segment();
// after the movers of a step, before its import: count and rank the entries of the big outboxes; returns the number of kernels launched
int launch_big_index(const BigOutboxes& big, const int64_t* cellStart, const int* nLeave, const unsigned char* tag, int* dirOff, const int* stopFlag, cudaStream_t st);
// z-slab exchange: pack both faces of the slab into their messages (3 launches)
void launch_face_pack(const Geometry& g, const ParticleArrays& src, const int64_t* cellStart, const int* nLeave, const unsigned char* tag,
                      const unsigned short* order, const int* dirOff, int* faceCnt, const FaceMsg& down, const FaceMsg& up, int* totals, int* stopFlag, int stepIndex, cudaStream_t st);
// J on the node planes [zNode0, zNode0 + zNodes) (default: all nz + 1 planes) from the per-cell corner sums
void launch_gather_current(const Geometry& g, const double* cellJ, double* J, const int* stopFlag, cudaStream_t st, int zNode0 = 0, int zNodes = -1);
void launch_count_misplaced(const Geometry& g, const ParticleArrays& p, const int64_t* cellStart, const int* count, unsigned long long* bad, cudaStream_t st);
void launch_scatter_aos(const Geometry& g, const double* aos, const int64_t* aosStart, const ParticleArrays& p, const int64_t* cellStart, cudaStream_t st);
void launch_gather_aos(const Geometry& g, double* aos, const int64_t* aosStart, const ParticleArrays& p, const int64_t* cellStart, cudaStream_t st);
// copy the live head of every cell into another layout; `g` describes the DESTINATION (its nf), srcNf / srcQ / srcQom the source
void launch_copy_heads(const Geometry& g, const ParticleArrays& src, const int64_t* srcStart, const ParticleArrays& dst, const int64_t* dstStart, const int* count,
                       int srcNf, double srcQ, double srcQom, cudaStream_t st);
void launch_check_species(const double* aos, int64_t n, int* flag, cudaStream_t st);

void launch_selftest_division(uint64_t n, uint64_t seed, unsigned long long* bad, cudaStream_t st);

} // namespace ipic
