/*
 * ipic3d_b200.h — C ABI of the B200-native particle hot path of AllScale iPIC3D.
 *
 * This is the drop-in boundary: plain C types, caller-allocated buffers, no callbacks, no torch
 * types.  It is what the reference's C++ entry point
 *     simulateSteps<Projector,FieldSolver,Mover>(numSteps, Universe&)
 *                                   (reference code/app/include/ipic3d/app/simulator.h:42-47,127-242)
 * calls underneath in this repository (see host/include/ipic3d/app/simulator.h and INTEGRATION.md).
 *
 * Data conventions (all doubles are IEEE binary64, all arrays are dense, caller-owned):
 *   cells     : N = nx*ny*nz, flattened (x*ny + y)*nz + z           (reference cell.h:521-523)
 *   particles : AoS of 8 doubles {x,y,z,vx,vy,vz,q,qom} = reference `struct Particle`
 *               (particle.h:9-19), stored cell-major in canonical per-cell order.
 *   field     : (nx+3)(ny+3)(nz+3) nodes x 9 doubles {E.xyz,B.xyz,Bext.xyz} = `FieldNode`   (field.h:14-18)
 *   bcfield   : (nx+2)(ny+2)(nz+2) cells x 3 doubles {Bc.xyz}          = `BcFieldCell` (field.h:20-23)
 *   density   : (nx+1)(ny+1)(nz+1) nodes x 3 doubles {J.xyz}           = `DensityNode` (field.h:25-27)
 *   all three flattened like the cells (last index fastest), sizes as in universe.h:36.
 *
 * Every function returns 0 on success and a negative ipic_status on failure; the message of the
 * last failure on the calling thread is available from ipic_last_error().  There is NO CPU
 * fallback: without a CUDA device every compute entry point fails with IPIC_ERR_CUDA.
 */
#ifndef IPIC3D_B200_H
#define IPIC3D_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#  define IPIC_API __attribute__((visibility("default")))
#else
#  define IPIC_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
	IPIC_OK            =  0,
	IPIC_ERR_ARG       = -1,   /* bad argument / state */
	IPIC_ERR_CUDA      = -2,   /* CUDA runtime failure (incl. no device) */
	IPIC_ERR_CAPACITY  = -3,   /* caller buffer too small */
	IPIC_ERR_COMM      = -4,   /* NCCL failure / not initialised */
	IPIC_ERR_INTERNAL  = -5
} ipic_status;

/* POD mirror of ipic3d::UniverseProperties (reference universe_properties.h:52-75). */
typedef struct {
	int64_t size[3];              /* grid size in cells                    (:57)  */
	double  cellWidth[3];         /*                                       (:59)  */
	double  dt;                   /*                                       (:61)  */
	double  speedOfLight;         /*                                       (:63)  */
	double  planetRadius;         /*                                       (:65)  */
	double  objectCenter[3];      /*                                       (:67)  */
	double  origin[3];            /*                                       (:70)  */
	double  externalMagneticField[3]; /*                                   (:72)  */
} ipic_props;

/* Step flags: which of the reference's policy functors are active (simulator.h:246-275). */
enum {
	IPIC_MOVER_DIPOLE_ANALYTIC = 0u,      /* detail::default_particle_mover as shipped: E=0, analytic dipole B (cell.h:896-964) */
	IPIC_MOVER_INTERPOLATED    = 1u << 0, /* the commented-out corner gather + trilinearInterpolationF2P (cell.h:905-937, :867-885) */
	IPIC_PROJECT_CURRENT       = 1u << 1, /* detail::default_particle_to_field_projector → projectToDensityField (cell.h:787-824);
	                                         the per-cell sums are formed in another (fixed, reproducible) order than the reference's
	                                         sequential loop: J agrees within 1e-12 of max|J|, not bit for bit */
	IPIC_SOLVER_FORWARD        = 1u << 2, /* detail::forward_field_solver  → solveFieldForward  (field.h:294-337) */
	IPIC_SOLVER_LEAPFROG       = 1u << 3, /* detail::leapfrog_field_solver → solveFieldLeapfrog (field.h:342-402) */
	IPIC_UPDATE_BOUNDARIES     = 1u << 4, /* updateFieldsOnBoundaries before the solve (field.h:407-440, simulator.h:204) */
	IPIC_SKIP_MIGRATION        = 1u << 5  /* diagnostics only: run mover without export/import */
};

typedef struct ipic_ctx ipic_ctx;

/* Library / device ------------------------------------------------------------------------- */
IPIC_API const char* ipic_last_error(void);
IPIC_API const char* ipic_version(void);
IPIC_API int  ipic_device_count(int* count);

/* Page-locked host memory for the AoS transfer buffers (so that uploads / downloads run at PCIe speed). */
IPIC_API int  ipic_alloc_pinned(size_t bytes, void** out);
IPIC_API int  ipic_free_pinned(void* p);

/* Context: one z-slab of the universe on one GPU.  For a single GPU the slab is the whole
 * universe (rank 0 of 1).  Replaces the locals of simulateSteps (TransferBuffers, scratch;
 * simulator.h:153-156). `device` is a CUDA ordinal. */
IPIC_API int  ipic_create(const ipic_props* props, int device, ipic_ctx** out);
IPIC_API int  ipic_create_slab(const ipic_props* props, int device, int rank, int nranks, ipic_ctx** out);
/* Slabs need not be equally thick: rank r owns the z-planes [z_begin, z_begin + z_count); the ranks' ranges must tile
 * [0, nz) in rank order (checked when the slabs first step together).  For universes whose particles are not spread
 * evenly along z -- the Earth decks fill only the inner 75 % box -- the caller cuts the slabs by particle count. */
IPIC_API int  ipic_create_slab_range(const ipic_props* props, int device, int rank, int nranks, int z_begin, int z_count, ipic_ctx** out);
IPIC_API int  ipic_destroy(ipic_ctx* ctx);

/* Head-room of a cell's segment in HBM: room for n*(1+slack) particles, at least n+min_pad.  When an
 * import still overflows a segment the context re-layouts transparently (see ipic_relayout_count). */
IPIC_API int  ipic_set_headroom(ipic_ctx* ctx, double slack, int64_t min_pad);
/* Number of times a cell segment overflowed and the context re-laid out its particle storage (the step
 * that overflowed is redone; results are unaffected).  Slab contexts repair themselves the same way:
 * ipic_step confirms every slab step on the host before the next step's exchange is enqueued. */
IPIC_API int  ipic_relayout_count(ipic_ctx* ctx, uint64_t* count);

/* Multi-GPU wiring (z-slab neighbours, NCCL send/recv).  unique_id is NCCL's 128-byte id,
 * produced on rank 0 by ipic_comm_unique_id and distributed by the caller. */
IPIC_API int  ipic_comm_unique_id(void* unique_id_128);
IPIC_API int  ipic_comm_init(ipic_ctx* ctx, const void* unique_id_128);

/* Particle layout in HBM.  When every uploaded particle carries the same (q, qom) bit patterns -- true for all
 * of the reference's workloads, which run one species -- the context keeps the two constants aside and streams
 * 6 instead of 8 doubles per particle (96 instead of 128 bytes of HBM traffic per push); downloads, face
 * messages and results are unaffected.  Slabs agree on the layout before their first step.  enable = 0 forces
 * the general 8-field layout (takes effect at the next upload).  ipic_particle_fields reports 6 or 8. */
IPIC_API int  ipic_set_uniform_species(ipic_ctx* ctx, int enable);
IPIC_API int  ipic_particle_fields(ipic_ctx* ctx, int* fields);

/* Face messages of the z-slab exchange are sent exactly as large as their content (the host reads the packed
 * totals while the movers of the same step run) and their buffers grow on demand.  particles > 0 pins the
 * buffer capacity instead (a step that needs more fails with IPIC_ERR_CAPACITY); 0 = automatic (default). */
IPIC_API int  ipic_set_exchange_capacity(ipic_ctx* ctx, int64_t particles);

/* Single-process multi-GPU: ndev slab contexts (rank r on devices[r]; devices may repeat) that exchange
 * by peer copies instead of NCCL.  This is what a single-process caller such as the reference's
 * simulateSteps uses to drive all GPUs of a box.  Upload / download each member with the functions
 * below (rank r owns the z-planes [r*nz/ndev, (r+1)*nz/ndev)); step all members together with
 * ipic_group_step; destroy each member with ipic_destroy. */
IPIC_API int  ipic_group_create(const ipic_props* props, const int* devices, int ndev, ipic_ctx** out_ctxs);
IPIC_API int  ipic_group_create_ranges(const ipic_props* props, const int* devices, int ndev, const int* z_begin, const int* z_count, ipic_ctx** out_ctxs);
IPIC_API int  ipic_group_step(ipic_ctx** ctxs, int n, uint64_t num_steps, uint32_t flags, double* t_first, double* t_rest);

/* State transfer.  Replaces direct access to universe.cells / .field / .bcfield / .currentDensity
 * (universe.h:12-27).  `counts` has one entry per cell of the GLOBAL grid when the context is the
 * whole universe, or of the slab (nx*ny*nz_local, same flatten) for slab contexts. */
IPIC_API int  ipic_upload_cells(ipic_ctx* ctx, const uint64_t* counts, const double* particles_aos);
/* The universe of the reference's `:U:N` benchmark created directly in HBM: initCells(properties, N, distribution::uniform<>)
 * (cell.h:638-697) for the context's cells -- population N / cells (+1 for the first N % cells cells of the z-fastest
 * flatten), per cell two std::minstd_rand engines seeded (x*1023 + y)*1023 + z + 1 / + 2, positions uniform in the cell's
 * box and velocities uniform in [-0.2, 0.2]^3 through std::uniform_real_distribution<double> (libstdc++ generate_canonical:
 * two engine calls per double), electrons (q -1, qom -25).  Bit-identical to the host generator: the LCG is jumped ahead
 * per particle, the floating-point steps are the library's, unfused.  Replaces ipic_upload_cells for this universe. */
IPIC_API int  ipic_generate_uniform(ipic_ctx* ctx, uint64_t num_particles);
/* The universe of `ipic3d <deck>` created directly in HBM: phase 2 of initCells<Distribution> (cell.h:596-632) for the context's
 * cells with the distribution main() builds (ipic3d.cxx:72-83, distribution::uniform_pos_normal_speed): per cell, seed
 * (x*1023 + y)*1023 + z, positions uniform in the cell's box from std::minstd_rand(seed) (cell.h:138-174), velocity components
 * mean + stddev * ziggurat sample (cell.h:222-257, ziggurat_normal_distribution.h: 64-bit SHR3 state 123456789 ^ s with
 * s = the seed after generic_particle_generator::seed's two (1023 s) & s steps, cell.h:281-287), charge q and charge-to-mass
 * ratio qom for every particle (cell.h:624-625).  counts (owned cells, slab order) are the populations of phase 1, which is
 * sequential by construction and stays with the caller (host mirror: EarthDeck::cellPopulations).  Bit-identical to the host
 * generator.  Replaces ipic_upload_cells for such a universe: nothing crosses PCIe but the counts. */
IPIC_API int  ipic_generate_normal_speed(ipic_ctx* ctx, const uint64_t* counts, const double* mean3, const double* stddev3, double q, double qom);
IPIC_API int  ipic_download_cells(ipic_ctx* ctx, uint64_t* counts, double* particles_aos, uint64_t capacity_particles);
IPIC_API int  ipic_cell_counts(ipic_ctx* ctx, uint64_t* counts);
IPIC_API int  ipic_total_particles(ipic_ctx* ctx, uint64_t* total);
IPIC_API int  ipic_upload_fields(ipic_ctx* ctx, const double* field, const double* bcfield, const double* density);
/* Slab contexts: E, B, Bext, Bc are replicated on every slab (the reference's solvers update in place in a global node
 * order, so every slab runs them on the whole -- tiny -- grid).  J: with a field solver among the step flags every slab
 * holds all of it (the slabs all-gather their per-cell corner sums); with the projection alone (IPIC_PROJECT_CURRENT, no
 * solver) a slab builds J only on its own node planes z_begin .. z_begin + z_count from a one-plane halo of its z-neighbours'
 * corner sums, and `density` is valid on those planes only. */
IPIC_API int  ipic_download_fields(ipic_ctx* ctx, double* field, double* bcfield, double* density);

/* The hot path: numSteps iterations of the loop body of simulateSteps (simulator.h:181-230).
 * Blocking; returns after the device is idle.  t_first / t_rest (seconds, may be NULL) mirror
 * DurationMeasurement{firstStep, remainingSteps} (simulator.h:37-40). */
IPIC_API int  ipic_step(ipic_ctx* ctx, uint64_t num_steps, uint32_t flags, double* t_first, double* t_rest);

/* Same, but only enqueues the work on the context's stream; pair with ipic_sync.  Single-GPU contexts only. */
IPIC_API int  ipic_step_async(ipic_ctx* ctx, uint64_t num_steps, uint32_t flags);
IPIC_API int  ipic_sync(ipic_ctx* ctx);

/* Invariant check mirroring verifyCorrectParticlesPositionInCell (cell.h:1168-1184): number of
 * particles outside the closed box of their cell. */
IPIC_API int  ipic_count_misplaced(ipic_ctx* ctx, uint64_t* misplaced);

/* Conserved quantities of the state on the device, the four columns writeOutputData prints into
 * <base>ConservedQuantities.out (simulator.h:71-91): out4 = { total particle momentum sum (q/qom)*|v|
 * (getTotalParticlesEnergy(cells, getParticlesMomentum), cell.h:1330-1355), E field energy, B + Bext field
 * energy (getFieldEnergy, field.h:443-459), total kinetic energy (getParticlesKineticEnergy, cell.h:1315-1325) }.
 * Deterministic tree reductions on the device: no particle leaves the GPU.  For a slab context the two
 * particle sums cover the slab's own cells (the field grids are replicated); ipic_group_diagnostics adds the
 * members' sums in rank order. */
IPIC_API int  ipic_diagnostics(ipic_ctx* ctx, double* out4);
IPIC_API int  ipic_group_diagnostics(ipic_ctx** ctxs, int n, double* out4);

/* pic_tracer (code/app/src/pic_tracer.cxx:77-148, traceParticle): traces n independent particles (AoS records as
 * above; the caller generates them, pic_tracer.cxx:236-262) for `timesteps` steps in the analytic dipole of `props`
 * with periodic wrap-around and absorption by the planet, and counts each surviving particle into the density frame
 * t / frame_interval at the cell it occupied before step t, every frame_interval steps.
 * counts[num_frames][nx][ny][nz] (uint32, ParticleCount::flatten, pic_tracer.cxx:66-68) is overwritten.
 * `seconds` (may be NULL): device time of upload + tracing.  No context needed; particles are streamed through
 * the GPU in chunks, so n is not limited by HBM. */
IPIC_API int  ipic_trace(const ipic_props* props, int device, const double* particles_aos, uint64_t n, int timesteps,
                         int frame_interval, int num_frames, uint32_t* counts, double* seconds);

/* Device self-test of the FP64 helpers that replace library code on the hot path (shared-reciprocal
 * division of the Boris rotation): number of random inputs (out of n) whose result is not bit-identical
 * to IEEE division.  Must be 0. */
IPIC_API int  ipic_selftest(int device, uint64_t n, uint64_t seed, uint64_t* mismatches);

/* Measurement support: device time (ms) of the mover / import kernels accumulated since the last
 * reset, and the number of kernel launches issued. */
IPIC_API int  ipic_profile_reset(ipic_ctx* ctx);
IPIC_API int  ipic_profile_get(ipic_ctx* ctx, double* mover_ms, double* import_ms, uint64_t* launches);
/* Slab steps, accumulated ms on the main stream: [0] all movers (span from the first launch to the last one's end),
 * [1] import kernels, [2] waiting for the face exchange, [3] waiting for the all-gathered current + J + field phase. */
IPIC_API int  ipic_profile_phases(ipic_ctx* ctx, double* ms4);

#ifdef __cplusplus
}
#endif
#endif /* IPIC3D_B200_H */
