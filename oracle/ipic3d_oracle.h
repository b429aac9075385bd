/*
 * TEST INFRASTRUCTURE ONLY — CPU oracle for the particle hot path of AllScale iPIC3D.
 * Plain-C restatement of the reference algorithm; see ipic3d_oracle.c for the file:line map.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 *
 * PARITY PIN: checked bit-for-bit (memcmp of every particle and field value) against the
 * reference's own headers compiled here (oracle/_ref/libipic3d_ref.so, see ref_capi.cc) by
 * tests/test_oracle_vs_reference.py, and against the reference's known-answer vectors
 * (tests/test_oracle_kat.py, tests/golden/).
 */
#ifndef IPIC3D_ORACLE_H
#define IPIC3D_ORACLE_H
#include <stdint.h>
#include "ipic3d_b200.h"   /* ipic_props, step flags */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_universe orc_universe;

orc_universe* orc_create(const ipic_props* props);
void orc_destroy(orc_universe* u);

void orc_set_cells(orc_universe* u, const uint64_t* counts, const double* aos);
void orc_get_counts(const orc_universe* u, uint64_t* counts);
void orc_get_particles(const orc_universe* u, double* aos);
uint64_t orc_total_particles(const orc_universe* u);
uint64_t orc_count_misplaced(const orc_universe* u);
/* { total momentum, E energy, B+Bext energy, total kinetic energy }: the columns of ConservedQuantities.out */
void orc_diagnostics(const orc_universe* u, double* out4);

void orc_set_field(orc_universe* u, const double* field);     /* (n+3)^3 x 9 */
void orc_get_field(const orc_universe* u, double* field);
void orc_set_bcfield(orc_universe* u, const double* bc);      /* (n+2)^3 x 3 */
void orc_get_bcfield(const orc_universe* u, double* bc);
void orc_set_density(orc_universe* u, const double* J);       /* (n+1)^3 x 3 */
void orc_get_density(const orc_universe* u, double* J);
void orc_init_fields(orc_universe* u, const double* drift, const double* B0);

/* n iterations of the time-loop body; flags = IPIC_* step flags (0 = the loop as shipped). */
void orc_step(orc_universe* u, uint64_t n, uint32_t flags);

/* unit level */
void orc_update_velocity(double* p8, const double* E, const double* B, double dt);
void orc_update_position(double* p8, double dt);
void orc_trilinear(const double* corners24, const double* pos, double vol, double* out3);
void orc_project_node(orc_universe* u, int64_t x, int64_t y, int64_t z);
void orc_solve_forward_node(orc_universe* u, int64_t x, int64_t y, int64_t z);
void orc_solve_leapfrog_node(orc_universe* u, int64_t x, int64_t y, int64_t z);
void orc_update_boundaries(orc_universe* u);
/* mover for one particle in dipole-analytic mode; returns the number of sub-cycles used */
int  orc_move_particle_dipole(const ipic_props* props, double* p8);

/* pic_tracer's traceParticle over n particles (AoS); counts[frames][nx][ny][nz] is accumulated into */
void orc_trace(const ipic_props* props, const double* aos, uint64_t n, int T, int frame_interval, uint32_t* counts);

#ifdef __cplusplus
}
#endif
#endif
