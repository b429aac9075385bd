/*
 * TEST INFRASTRUCTURE ONLY — CPU oracle for the particle hot path of AllScale iPIC3D.
 *
 * A plain-C restatement (scalars and flat arrays, no templates, no runtime) of what the
 * reference computes per time step.  It is the checker for the CUDA path; it is never the thing
 * shipped or measured (except as bench.py's "port" cpu_baseline).  "hdr/" below abbreviates
 * /root/reference/code/app/include/ipic3d/app/.
 *
 *   orc_update_velocity     <- Particle::updateVelocity         hdr/particle.h:31-49
 *   orc_update_position     <- Particle::updatePosition         hdr/particle.h:24-26
 *   move_cell_dipole        <- moveParticles (live code)        hdr/cell.h:896-964
 *   move_cell_interpolated  <- moveParticles (commented gather) hdr/cell.h:905-937 + :947-960
 *   orc_trilinear           <- trilinearInterpolationF2P        hdr/cell.h:867-885
 *   export_cell             <- exportParticles                  hdr/cell.h:975-1159
 *   import_cell             <- importParticles                  hdr/cell.h:1195-1310
 *   orc_project_node        <- projectToDensityField            hdr/cell.h:787-824
 *   curl_b / curl_e         <- computeCurlB / computeCurlE      hdr/field.h:162-229
 *   interp_c2n / interp_n2c <- interpC2N / interpN2C            hdr/field.h:234-267
 *   orc_solve_forward_node  <- solveFieldForward                hdr/field.h:294-337
 *   orc_solve_leapfrog_node <- solveFieldLeapfrog (branch if(1)) hdr/field.h:342-373
 *   orc_update_boundaries   <- updateFieldsOnBoundaries         hdr/field.h:407-440
 *   orc_init_fields         <- initFields/getDipoleFieldAt/initBcFields hdr/field.h:43-137
 *   orc_step                <- simulateSteps loop body          hdr/simulator.h:181-230
 *   orc_trace               <- traceParticle                    code/app/src/pic_tracer.cxx:77-148
 *   orc_diagnostics         <- writeOutputData columns          hdr/simulator.h:71-91 (getFieldEnergy hdr/field.h:443-459,
 *                              getParticlesKineticEnergy / getParticlesMomentum / getTotalParticlesEnergy hdr/cell.h:1315-1355)
 *
 * Floating point: every expression keeps the reference's association order; compile with
 * -ffp-contract=off (no FMA), as x86-64 GCC does for the reference itself.
 *
 * PARITY PIN: bit-exact against oracle/_ref (the reference's headers compiled with the AllScale
 * shim) in tests/test_oracle_vs_reference.py; KATs in tests/test_oracle_kat.py.
 */
#include "ipic3d_oracle.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double* p; uint64_t n, cap; } plist;   /* particles: 8 doubles each */

struct orc_universe {
	ipic_props props;
	int64_t nx, ny, nz, ncells;
	plist* cells;       /* ncells */
	plist* outbox;      /* ncells * 27, indexed [source cell][i*9+j*3+k], (i,j,k) as in export */
	double* field;      /* (n+3)^3 * 9 : E, B, Bext */
	double* bc;         /* (n+2)^3 * 3 */
	double* J;          /* (n+1)^3 * 3 */
};

static void plist_push(plist* l, const double* p8) {
	if (l->n == l->cap) { l->cap = l->cap ? l->cap * 2 : 16; l->p = (double*)realloc(l->p, l->cap * 64); }
	memcpy(l->p + l->n * 8, p8, 64);
	l->n++;
}

static int64_t cell_index(const orc_universe* u, int64_t x, int64_t y, int64_t z) { return (x * u->ny + y) * u->nz + z; }
static int64_t fidx(const orc_universe* u, int64_t x, int64_t y, int64_t z) { return ((x * (u->ny + 3) + y) * (u->nz + 3) + z) * 9; }
static int64_t bidx(const orc_universe* u, int64_t x, int64_t y, int64_t z) { return ((x * (u->ny + 2) + y) * (u->nz + 2) + z) * 3; }
static int64_t jidx(const orc_universe* u, int64_t x, int64_t y, int64_t z) { return ((x * (u->ny + 1) + y) * (u->nz + 1) + z) * 3; }

orc_universe* orc_create(const ipic_props* props) {
	orc_universe* u = (orc_universe*)calloc(1, sizeof(orc_universe));
	u->props = *props;
	u->nx = props->size[0]; u->ny = props->size[1]; u->nz = props->size[2];
	u->ncells = u->nx * u->ny * u->nz;
	u->cells = (plist*)calloc((size_t)u->ncells, sizeof(plist));
	u->outbox = (plist*)calloc((size_t)u->ncells * 27, sizeof(plist));
	u->field = (double*)calloc((size_t)((u->nx + 3) * (u->ny + 3) * (u->nz + 3) * 9), sizeof(double));
	u->bc = (double*)calloc((size_t)((u->nx + 2) * (u->ny + 2) * (u->nz + 2) * 3), sizeof(double));
	u->J = (double*)calloc((size_t)((u->nx + 1) * (u->ny + 1) * (u->nz + 1) * 3), sizeof(double));
	return u;
}

void orc_destroy(orc_universe* u) {
	if (!u) return;
	for (int64_t c = 0; c < u->ncells; c++) free(u->cells[c].p);
	for (int64_t c = 0; c < u->ncells * 27; c++) free(u->outbox[c].p);
	free(u->cells); free(u->outbox); free(u->field); free(u->bc); free(u->J); free(u);
}

void orc_set_cells(orc_universe* u, const uint64_t* counts, const double* aos) {
	for (int64_t c = 0; c < u->ncells; c++) {
		plist* l = &u->cells[c];
		if (l->cap < counts[c]) { l->cap = counts[c]; l->p = (double*)realloc(l->p, l->cap * 64); }
		l->n = counts[c];
		if (counts[c]) memcpy(l->p, aos, counts[c] * 64);
		aos += counts[c] * 8;
	}
}
void orc_get_counts(const orc_universe* u, uint64_t* counts) { for (int64_t c = 0; c < u->ncells; c++) counts[c] = u->cells[c].n; }
void orc_get_particles(const orc_universe* u, double* aos) {
	for (int64_t c = 0; c < u->ncells; c++) { if (u->cells[c].n) memcpy(aos, u->cells[c].p, u->cells[c].n * 64); aos += u->cells[c].n * 8; }
}
uint64_t orc_total_particles(const orc_universe* u) { uint64_t t = 0; for (int64_t c = 0; c < u->ncells; c++) t += u->cells[c].n; return t; }

#define ORC_COPY(NAME_SET, NAME_GET, MEMBER, EXTRA, W) \
void NAME_SET(orc_universe* u, const double* b) { memcpy(u->MEMBER, b, (size_t)((u->nx + EXTRA) * (u->ny + EXTRA) * (u->nz + EXTRA) * W) * sizeof(double)); } \
void NAME_GET(const orc_universe* u, double* b) { memcpy(b, u->MEMBER, (size_t)((u->nx + EXTRA) * (u->ny + EXTRA) * (u->nz + EXTRA) * W) * sizeof(double)); }
ORC_COPY(orc_set_field, orc_get_field, field, 3, 9)
ORC_COPY(orc_set_bcfield, orc_get_bcfield, bc, 2, 3)
ORC_COPY(orc_set_density, orc_get_density, J, 1, 3)

/* ---- geometry: universe_properties.h:125-142 ------------------------------------------------ */

static void origin_of_cell(const ipic_props* P, const int64_t pos[3], double out[3]) {
	for (int a = 0; a < 3; a++) out[a] = P->origin[a] + (double)pos[a] * P->cellWidth[a];
}
static void center_of_cell(const ipic_props* P, const int64_t pos[3], double out[3]) {
	origin_of_cell(P, pos, out);
	for (int a = 0; a < 3; a++) out[a] = out[a] + P->cellWidth[a] / 2.0;
}

/* isInside, cell.h:48-59 */
static int is_inside(const ipic_props* P, const int64_t pos[3], const double* p8) {
	double c[3]; center_of_cell(P, pos, c);
	for (int a = 0; a < 3; a++) if (!(fabs(p8[a] - c[a]) <= P->cellWidth[a] / 2.0)) return 0;
	return 1;
}

uint64_t orc_count_misplaced(const orc_universe* u) {
	uint64_t bad = 0;
	for (int64_t x = 0; x < u->nx; x++) for (int64_t y = 0; y < u->ny; y++) for (int64_t z = 0; z < u->nz; z++) {
		const plist* l = &u->cells[cell_index(u, x, y, z)];
		int64_t pos[3] = { x, y, z };
		for (uint64_t i = 0; i < l->n; i++) if (!is_inside(&u->props, pos, l->p + i * 8)) bad++;
	}
	return bad;
}

/* ---- conserved quantities: simulator.h:71-91 ---------------------------------------------------
 * out4 = { total momentum, E energy, B(+Bext) energy, total kinetic energy }; sequential folds in the
 * reference's order: cells lexicographic, particles in cell order, per-cell sum first (cell.h:1315-1355);
 * field nodes [1, size+2) lexicographic (field.h:443-459). */
void orc_diagnostics(const orc_universe* u, double* out4) {
	double mom = 0.0, ke = 0.0;
	for (int64_t c = 0; c < u->ncells; c++) {
		const plist* l = &u->cells[c];
		double cm = 0.0, ck = 0.0;
		for (uint64_t i = 0; i < l->n; i++) {
			const double* p = l->p + i * 8;
			double v2 = 0.0; v2 += p[3] * p[3]; v2 += p[4] * p[4]; v2 += p[5] * p[5];
			ck += 0.5 * (p[6] / p[7]) * v2;
			cm += (p[6] / p[7]) * sqrt(v2);
		}
		mom += cm; ke += ck;
	}
	const double vol = 0.5 * u->props.cellWidth[0] * u->props.cellWidth[1] * u->props.cellWidth[2];
	const double fourPI = 16.0 * atan(1.0);
	double eE = 0.0, eB = 0.0;
	for (int64_t x = 1; x < u->nx + 2; x++) for (int64_t y = 1; y < u->ny + 2; y++) for (int64_t z = 1; z < u->nz + 2; z++) {
		const double* f = u->field + fidx(u, x, y, z);
		double e2 = 0.0; e2 += f[0] * f[0]; e2 += f[1] * f[1]; e2 += f[2] * f[2];
		const double b0 = f[3] + f[6], b1 = f[4] + f[7], b2 = f[5] + f[8];
		double bb = 0.0; bb += b0 * b0; bb += b1 * b1; bb += b2 * b2;
		eE += vol * e2 / (fourPI);
		eB += vol * bb / (fourPI);
	}
	out4[0] = mom; out4[1] = eE; out4[2] = eB; out4[3] = ke;
}

/* ---- Boris pusher: particle.h:24-49 ---------------------------------------------------------- */

void orc_update_velocity(double* p, const double* E, const double* B, double dt) {
	double* v = p + 3;
	double qom = p[7];
	double k = qom * 0.5 * dt;
	double t[3] = { k * B[0], k * B[1], k * B[2] };
	double t_mag2 = 0.0; for (int a = 0; a < 3; a++) t_mag2 += t[a] * t[a];
	double den = 1.0 + t_mag2;
	double s[3] = { (2.0 * t[0]) / den, (2.0 * t[1]) / den, (2.0 * t[2]) / den };
	double vm[3] = { v[0] + k * E[0], v[1] + k * E[1], v[2] + k * E[2] };
	double vp[3] = { vm[0] + (vm[1] * t[2] - vm[2] * t[1]), vm[1] + (vm[2] * t[0] - vm[0] * t[2]), vm[2] + (vm[0] * t[1] - vm[1] * t[0]) };
	double vq[3] = { vm[0] + (vp[1] * s[2] - vp[2] * s[1]), vm[1] + (vp[2] * s[0] - vp[0] * s[2]), vm[2] + (vp[0] * s[1] - vp[1] * s[0]) };
	v[0] = vq[0] + k * E[0]; v[1] = vq[1] + k * E[1]; v[2] = vq[2] + k * E[2];
}

void orc_update_position(double* p, double dt) { for (int a = 0; a < 3; a++) p[a] += p[3 + a] * dt; }

/* int(double) exactly as x86-64 cvttsd2si does it for the reference build: out of range / NaN -> INT_MIN */
static int x86_double_to_int(double v) {
	if (!(v > -2147483649.0 && v < 2147483648.0)) return INT_MIN;
	return (int)v;
}

/* sub-cycling + Boris, cell.h:947-960 */
static int push_with_fields(const ipic_props* P, double* p, const double* E, const double* B) {
	double B_mag = 0.0; for (int a = 0; a < 3; a++) B_mag += B[a] * B[a];
	double dt_sub = M_PI * P->speedOfLight / (4.0 * fabs(p[7]) * B_mag);
	int sub_cycles = x86_double_to_int(P->dt / dt_sub) + 1;
	if (sub_cycles > 100) sub_cycles = 100;
	dt_sub = P->dt / (double)sub_cycles;
	for (int c = 0; c < sub_cycles; c++) { orc_update_velocity(p, E, B, dt_sub); orc_update_position(p, dt_sub); }
	return sub_cycles;
}

/* cell.h:919-945: analytic z-dipole at the particle's absolute position, E = 0 */
int orc_move_particle_dipole(const ipic_props* P, double* p) {
	double magneticFieldTemp = -P->externalMagneticField[2] * pow(P->planetRadius, 3);
	double r2 = 0.0; for (int a = 0; a < 3; a++) r2 += p[a] * p[a];
	double fac1 = magneticFieldTemp / pow(r2, 2.5);
	double E[3] = { 0.0, 0.0, 0.0 }, B[3];
	B[0] = 3.0 * p[0] * p[2] * fac1;
	B[1] = 3.0 * p[1] * p[2] * fac1;
	B[2] = (2.0 * (p[2] * p[2]) - p[0] * p[0] - p[1] * p[1]) * fac1;
	return push_with_fields(P, p, E, B);
}

/* pic_tracer.cxx:77-148: n independent particles, T steps, periodic wrap, absorption; counts[frame][x][y][z] += 1
 * every frame_interval steps at the cell the particle occupied BEFORE that step's push.  The field expressions of
 * pic_tracer.cxx:106-112 are those of cell.h:919-945 (pow(v, 2) == v * v exactly in glibc). */
void orc_trace(const ipic_props* P, const double* aos, uint64_t n, int T, int frame_interval, uint32_t* counts) {
	double usize[3], low[3], high[3];
	for (int a = 0; a < 3; a++) { usize[a] = P->cellWidth[a] * (double)P->size[a]; low[a] = P->origin[a]; high[a] = P->origin[a] + usize[a]; }
	const double R2 = P->planetRadius * P->planetRadius;
	for (uint64_t i = 0; i < n; i++) {
		double p[8];
		memcpy(p, aos + i * 8, 64);
		for (int a = 0; a < 3; a++) p[a] += P->objectCenter[a];
		double r2 = 0.0; for (int a = 0; a < 3; a++) { double d = p[a] - P->objectCenter[a]; r2 += d * d; }
		if (r2 < R2) continue;
		int outside = 0;
		for (int a = 0; a < 3; a++) if (p[a] > high[a] || p[a] < low[a]) outside = 1;
		if (outside) continue;
		for (int t = 0; t < T; t++) {
			int64_t c[3];
			for (int a = 0; a < 3; a++) c[a] = (int64_t)((p[a] - P->origin[a]) / P->cellWidth[a]);
			orc_move_particle_dipole(P, p);
			for (int a = 0; a < 3; a++) {
				if (p[a] > high[a]) p[a] -= usize[a];
				if (p[a] < low[a]) p[a] += usize[a];
			}
			r2 = 0.0; for (int a = 0; a < 3; a++) { double d = p[a] - P->objectCenter[a]; r2 += d * d; }
			if (r2 <= R2) break;
			if (t % frame_interval == 0 && c[0] >= 0 && c[0] < P->size[0] && c[1] >= 0 && c[1] < P->size[1] && c[2] >= 0 && c[2] < P->size[2])
				counts[(((int64_t)(t / frame_interval) * P->size[0] + c[0]) * P->size[1] + c[1]) * P->size[2] + c[2]] += 1;
		}
	}
}

static void move_cell_dipole(orc_universe* u, plist* cell) {
	for (uint64_t i = 0; i < cell->n; i++) orc_move_particle_dipole(&u->props, cell->p + i * 8);
}

/* cell.h:867-885 (including the division by the cell volume) */
void orc_trilinear(const double* c, const double* pos, double vol, double* out) {
	double res[3] = { 0.0, 0.0, 0.0 };
	for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) for (int k = 0; k < 2; k++) {
		double fac = (i == 0 ? (1 - pos[0]) : pos[0]) * (j == 0 ? (1 - pos[1]) : pos[1]) * (k == 0 ? (1 - pos[2]) : pos[2]);
		const double* cc = c + ((i * 2 + j) * 2 + k) * 3;
		for (int a = 0; a < 3; a++) res[a] += cc[a] * fac;
	}
	for (int a = 0; a < 3; a++) out[a] = res[a] / vol;
}

/* cell.h:905-937 (commented in the reference) followed by :947-960 */
static void move_cell_interpolated(orc_universe* u, plist* cell, const int64_t pos[3]) {
	if (cell->n == 0) return;
	const ipic_props* P = &u->props;
	double Es[24], Bs[24];
	for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) for (int k = 0; k < 2; k++) {
		const double* node = u->field + fidx(u, pos[0] + i + 1, pos[1] + j + 1, pos[2] + k + 1);
		memcpy(Es + ((i * 2 + j) * 2 + k) * 3, node, 24);
		memcpy(Bs + ((i * 2 + j) * 2 + k) * 3, node + 3, 24);
	}
	double org[3]; origin_of_cell(P, pos, org);
	double vol = P->cellWidth[0] * P->cellWidth[1] * P->cellWidth[2];
	for (uint64_t n = 0; n < cell->n; n++) {
		double* p = cell->p + n * 8;
		double rel[3], E[3], B[3];
		for (int a = 0; a < 3; a++) rel[a] = (p[a] - org[a]) / P->cellWidth[a];
		orc_trilinear(Es, rel, vol, E);
		orc_trilinear(Bs, rel, vol, B);
		push_with_fields(P, p, E, B);
	}
}

/* ---- migration: cell.h:975-1159, :1195-1310, transfer_buffer.h ------------------------------- */

static void export_cell(orc_universe* u, const int64_t pos[3]) {
	const ipic_props* P = &u->props;
	int64_t c = cell_index(u, pos[0], pos[1], pos[2]);
	plist* cell = &u->cells[c];
	plist* out = u->outbox + c * 27;
	for (int d = 0; d < 27; d++) out[d].n = 0;             /* the 27 buffer clears (:1050-1086) */
	double center[3], hw[3];
	center_of_cell(P, pos, center);
	for (int a = 0; a < 3; a++) hw[a] = P->cellWidth[a] / 2.0;
	uint64_t kept = 0;
	for (uint64_t n = 0; n < cell->n; n++) {
		double* p = cell->p + n * 8;
		double rel[3];
		for (int a = 0; a < 3; a++) rel[a] = p[a] - center[a];
		int invert = 0;                                        /* half-cell "reflection" (:1102-1121) */
		for (int a = 0; a < 3; a++) {
			if (pos[a] == 0 && rel[a] < -hw[a]) { invert = 1; p[a] = p[a] + hw[a]; }
			else if (pos[a] == P->size[a] - 1 && rel[a] > hw[a]) { invert = 1; p[a] = p[a] - hw[a]; }
		}
		if (invert) for (int a = 0; a < 3; a++) p[3 + a] *= (-1);
		double r2 = 0.0;                                       /* absorption by the planet (:1123-1128) */
		for (int a = 0; a < 3; a++) { double d = p[a] - P->objectCenter[a]; r2 += d * d; }
		if (r2 <= P->planetRadius * P->planetRadius) continue;
		for (int a = 0; a < 3; a++) rel[a] = p[a] - center[a];
		if (fabs(rel[0]) > hw[0] || fabs(rel[1]) > hw[1] || fabs(rel[2]) > hw[2]) {
			int i = (rel[0] < -hw[0]) ? 0 : ((rel[0] > hw[0]) ? 2 : 1);
			int j = (rel[1] < -hw[1]) ? 0 : ((rel[1] > hw[1]) ? 2 : 1);
			int k = (rel[2] < -hw[2]) ? 0 : ((rel[2] > hw[2]) ? 2 : 1);
			plist_push(&out[i * 9 + j * 3 + k], p);
		} else {
			if (kept != n) memcpy(cell->p + kept * 8, p, 64);   /* stable: stayers keep their order */
			kept++;
		}
	}
	cell->n = kept;
}

static void import_cell(orc_universe* u, const int64_t pos[3]) {
	int64_t c = cell_index(u, pos[0], pos[1], pos[2]);
	plist* cell = &u->cells[c];
	const int64_t size[3] = { u->nx, u->ny, u->nz };
	/* buffers in TransferDirection order (0,0,0)..(2,2,2) without (1,1,1) (:1231-1306).  The buffer
	 * (target, dir=(a,b,c)) is the one the source cell target+(a-1,b-1,c-1) (periodic address)
	 * filled through its neighbors[2-a][2-b][2-c] entry (:1012-1048). */
	for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) for (int d = 0; d < 3; d++) {
		if (a == 1 && b == 1 && d == 1) continue;
		int64_t s[3] = { (pos[0] + (a - 1) + size[0]) % size[0], (pos[1] + (b - 1) + size[1]) % size[1], (pos[2] + (d - 1) + size[2]) % size[2] };
		const plist* in = &u->outbox[cell_index(u, s[0], s[1], s[2]) * 27 + (2 - a) * 9 + (2 - b) * 3 + (2 - d)];
		for (uint64_t n = 0; n < in->n; n++) plist_push(cell, in->p + n * 8);
	}
}

/* ---- current deposition: cell.h:787-824 ------------------------------------------------------ */

void orc_project_node(orc_universe* u, int64_t x, int64_t y, int64_t z) {
	const ipic_props* P = &u->props;
	const int64_t node[3] = { x, y, z };
	double Js[3] = { 0.0, 0.0, 0.0 };
	for (int i = -1; i < 1; i++) for (int j = -1; j < 1; j++) for (int k = -1; k < 1; k++) {
		int64_t cur[3] = { node[0] + i, node[1] + j, node[2] + k };
		for (int a = 0; a < 3; a++) cur[a] = (cur[a] < 0) ? P->size[a] - 1 : ((cur[a] >= P->size[a]) ? 0 : cur[a]);
		double org[3]; origin_of_cell(P, cur, org);
		const plist* cell = &u->cells[cell_index(u, cur[0], cur[1], cur[2])];
		for (uint64_t n = 0; n < cell->n; n++) {
			const double* p = cell->p + n * 8;
			double rel[3];
			for (int a = 0; a < 3; a++) rel[a] = (p[a] - org[a]) / P->cellWidth[a];
			double fac = (i == 0 ? (1 - rel[0]) : rel[0]) * (j == 0 ? (1 - rel[1]) : rel[1]) * (k == 0 ? (1 - rel[2]) : rel[2]);
			for (int a = 0; a < 3; a++) Js[a] += p[6] * p[3 + a] * fac;
		}
	}
	double vol = P->cellWidth[0] * P->cellWidth[1] * P->cellWidth[2];
	double* J = u->J + jidx(u, x, y, z);
	for (int a = 0; a < 3; a++) J[a] = (Js[a] / vol) / 8.0;
}

/* ---- field stencils: field.h:162-267 --------------------------------------------------------- */

/* V(i,j,k,comp) for the 2x2x2 neighbourhood; the reference's long sums are kept term by term. */
static void curl_from_corners(const double V[2][2][2][3], const double w[3], int forward, double curl[3]) {
	double zdy, ydz, xdz, zdx, ydx, xdy;
	if (!forward) { /* computeCurlB: V[i][j][k] = Bc[pos-(i,j,k)]  (field.h:162-193) */
		zdy = .25 * (V[0][0][0][2] - V[0][1][0][2] + V[0][0][1][2] - V[0][1][1][2] + V[1][0][0][2] - V[1][1][0][2] + V[1][0][1][2] - V[1][1][1][2]) / w[1];
		ydz = .25 * (V[0][0][0][1] - V[0][0][1][1] + V[1][0][0][1] - V[1][0][1][1] + V[0][1][0][1] - V[0][1][1][1] + V[1][1][0][1] - V[1][1][1][1]) / w[2];
		xdz = .25 * (V[0][0][0][0] - V[0][0][1][0] + V[1][0][0][0] - V[1][0][1][0] + V[0][1][0][0] - V[0][1][1][0] + V[1][1][0][0] - V[1][1][1][0]) / w[2];
		zdx = .25 * (V[0][0][0][2] - V[1][0][0][2] + V[0][0][1][2] - V[1][0][1][2] + V[0][1][0][2] - V[1][1][0][2] + V[0][1][1][2] - V[1][1][1][2]) / w[0];
		ydx = .25 * (V[0][0][0][1] - V[1][0][0][1] + V[0][0][1][1] - V[1][0][1][1] + V[0][1][0][1] - V[1][1][0][1] + V[0][1][1][1] - V[1][1][1][1]) / w[0];
		xdy = .25 * (V[0][0][0][0] - V[0][1][0][0] + V[0][0][1][0] - V[0][1][1][0] + V[1][0][0][0] - V[1][1][0][0] + V[1][0][1][0] - V[1][1][1][0]) / w[1];
	} else {        /* computeCurlE: V[i][j][k] = E[pos+(i,j,k)]   (field.h:198-229) */
		zdy = .25 * (V[0][1][0][2] - V[0][0][0][2] + V[0][1][1][2] - V[0][0][1][2] + V[1][1][0][2] - V[1][0][0][2] + V[1][1][1][2] - V[1][0][1][2]) / w[1];
		ydz = .25 * (V[0][0][1][1] - V[0][0][0][1] + V[1][0][1][1] - V[1][0][0][1] + V[0][1][1][1] - V[0][1][0][1] + V[1][1][1][1] - V[1][1][0][1]) / w[2];
		xdz = .25 * (V[0][0][1][0] - V[0][0][0][0] + V[1][0][1][0] - V[1][0][0][0] + V[0][1][1][0] - V[0][1][0][0] + V[1][1][1][0] - V[1][1][0][0]) / w[2];
		zdx = .25 * (V[1][0][0][2] - V[0][0][0][2] + V[1][0][1][2] - V[0][0][1][2] + V[1][1][0][2] - V[0][1][0][2] + V[1][1][1][2] - V[0][1][1][2]) / w[0];
		ydx = .25 * (V[1][0][0][1] - V[0][0][0][1] + V[1][0][1][1] - V[0][0][1][1] + V[1][1][0][1] - V[0][1][0][1] + V[1][1][1][1] - V[0][1][1][1]) / w[0];
		xdy = .25 * (V[0][1][0][0] - V[0][0][0][0] + V[0][1][1][0] - V[0][0][1][0] + V[1][1][0][0] - V[1][0][0][0] + V[1][1][1][0] - V[1][0][1][0]) / w[1];
	}
	curl[0] = zdy - ydz; curl[1] = xdz - zdx; curl[2] = ydx - xdy;
}

static void curl_b(const orc_universe* u, const int64_t pos[3], double curl[3]) {
	double V[2][2][2][3];
	for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) for (int k = 0; k < 2; k++)
		memcpy(V[i][j][k], u->bc + bidx(u, pos[0] - i, pos[1] - j, pos[2] - k), 24);
	curl_from_corners(V, u->props.cellWidth, 0, curl);
}

static void curl_e(const orc_universe* u, const int64_t pos[3], double curl[3]) {
	double V[2][2][2][3];
	for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) for (int k = 0; k < 2; k++)
		memcpy(V[i][j][k], u->field + fidx(u, pos[0] + i, pos[1] + j, pos[2] + k), 24);
	curl_from_corners(V, u->props.cellWidth, 1, curl);
}

static void interp_c2n(orc_universe* u, const int64_t pos[3]) {   /* field.h:234-247 */
	double s[3] = { 0.0, 0.0, 0.0 };
	for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) for (int k = 0; k < 2; k++) {
		const double* b = u->bc + bidx(u, pos[0] - i, pos[1] - j, pos[2] - k);
		for (int a = 0; a < 3; a++) s[a] += b[a];
	}
	double* B = u->field + fidx(u, pos[0], pos[1], pos[2]) + 3;
	for (int a = 0; a < 3; a++) B[a] = .125 * s[a];
}

static void interp_n2c(orc_universe* u, const int64_t pos[3]) {   /* field.h:252-267 */
	double s[3] = { 0.0, 0.0, 0.0 };
	for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) for (int k = 0; k < 2; k++) {
		const double* b = u->field + fidx(u, pos[0] + i, pos[1] + j, pos[2] + k) + 3;
		for (int a = 0; a < 3; a++) s[a] += b[a];
	}
	double* Bc = u->bc + bidx(u, pos[0], pos[1], pos[2]);
	for (int a = 0; a < 3; a++) Bc[a] = .125 * s[a];
}

/* lexicographic `pos < bcfield.size() - 1` (field.h:315, :366; Vector operator<) */
static int ghost_guard(const orc_universe* u, const int64_t pos[3]) {
	const int64_t lim[3] = { u->nx + 1, u->ny + 1, u->nz + 1 };
	for (int a = 0; a < 3; a++) { if (pos[a] < lim[a]) return 1; if (pos[a] > lim[a]) return 0; }
	return 0;
}

void orc_solve_forward_node(orc_universe* u, int64_t x, int64_t y, int64_t z) {
	const ipic_props* P = &u->props;
	const int64_t pos[3] = { x, y, z };
	double curlB[3];
	curl_b(u, pos, curlB);
	double* E = u->field + fidx(u, x, y, z);
	const double* J = u->J + jidx(u, x - 1, y - 1, z - 1);
	for (int a = 0; a < 3; a++) E[a] += (P->speedOfLight * curlB[a] - J[a]) * P->dt;
	if (ghost_guard(u, pos)) {
		double curlE[3];
		curl_e(u, pos, curlE);
		double* Bc = u->bc + bidx(u, x, y, z);
		for (int a = 0; a < 3; a++) Bc[a] -= P->speedOfLight * curlE[a] * P->dt;
	}
	interp_c2n(u, pos);
}

void orc_solve_leapfrog_node(orc_universe* u, int64_t x, int64_t y, int64_t z) {
	const ipic_props* P = &u->props;
	const int64_t pos[3] = { x, y, z };
	const double dt = P->dt, c = P->speedOfLight;
	const double* w = P->cellWidth;
	double* E = u->field + fidx(u, x, y, z);
	const double* J = u->J + jidx(u, x - 1, y - 1, z - 1);
	double* B = u->bc + bidx(u, x, y, z);
	const double* Bmx = u->bc + bidx(u, x - 1, y, z);
	const double* Bmy = u->bc + bidx(u, x, y - 1, z);
	const double* Bmz = u->bc + bidx(u, x, y, z - 1);
	E[0] = E[0] + dt * (c * c * ((B[2] - Bmy[2]) / w[1] - (B[1] - Bmz[1]) / w[2]) - J[0]);
	E[1] = E[1] + dt * (c * c * ((B[0] - Bmz[0]) / w[2] - (B[2] - Bmx[2]) / w[0]) - J[1]);
	E[2] = E[2] + dt * (c * c * ((B[1] - Bmx[1]) / w[0] - (B[0] - Bmy[0]) / w[1]) - J[2]);
	if (ghost_guard(u, pos)) {
		const double* Epx = u->field + fidx(u, x + 1, y, z);
		const double* Epy = u->field + fidx(u, x, y + 1, z);
		const double* Epz = u->field + fidx(u, x, y, z + 1);
		B[0] = B[0] - dt * ((Epy[2] - E[2]) / w[1] - (Epz[1] - E[1]) / w[2]);
		B[1] = B[1] - dt * ((Epz[0] - E[0]) / w[2] - (Epx[2] - E[2]) / w[0]);
		B[2] = B[2] - dt * ((Epx[1] - E[1]) / w[0] - (Epy[0] - E[0]) / w[1]);
	}
}

/* field.h:407-440: periodic ghost copy, faces only, cubic grid assumed (uses the x extent) */
static void boundaries_one(double* g, int64_t n, int W) {
	int64_t end = n - 1;
#define G(X, Y, Z) (g + (((X) * n + (Y)) * n + (Z)) * W)
	for (int64_t a = 1; a < end; a++) for (int64_t b = 1; b < end; b++) {
		memcpy(G(0, a, b), G(end - 1, a, b), W * 8); memcpy(G(end, a, b), G(1, a, b), W * 8);
		memcpy(G(a, 0, b), G(a, end - 1, b), W * 8); memcpy(G(a, end, b), G(a, 1, b), W * 8);
		memcpy(G(a, b, 0), G(a, b, end - 1), W * 8); memcpy(G(a, b, end), G(a, b, 1), W * 8);
	}
#undef G
}
void orc_update_boundaries(orc_universe* u) {
	boundaries_one(u->field, u->nx + 3, 9);
	boundaries_one(u->bc, u->nx + 2, 3);
}

/* field.h:43-137 */
void orc_init_fields(orc_universe* u, const double* drift, const double* B0) {
	const ipic_props* P = &u->props;
	memset(u->field, 0, (size_t)((u->nx + 3) * (u->ny + 3) * (u->nz + 3) * 9) * 8);
	memset(u->bc, 0, (size_t)((u->nx + 2) * (u->ny + 2) * (u->nz + 2) * 3) * 8);
	memset(u->J, 0, (size_t)((u->nx + 1) * (u->ny + 1) * (u->nz + 1) * 3) * 8);
	double ebc[3] = { -1.0 * (drift[1] * B0[2] - drift[2] * B0[1]), -1.0 * (drift[2] * B0[0] - drift[0] * B0[2]), -1.0 * (drift[0] * B0[1] - drift[1] * B0[0]) };
	double a = P->planetRadius;
	for (int64_t x = 1; x < u->nx + 2; x++) for (int64_t y = 1; y < u->ny + 2; y++) for (int64_t z = 1; z < u->nz + 2; z++) {
		const int64_t np[3] = { x - 1, y - 1, z - 1 };
		double loc[3]; origin_of_cell(P, np, loc);           /* getLocationForFields == same formula */
		double d[3]; for (int q = 0; q < 3; q++) d[q] = loc[q] - P->objectCenter[q];
		double r2 = 0.0; for (int q = 0; q < 3; q++) r2 += d[q] * d[q];
		double* n = u->field + fidx(u, x, y, z);
		n[0] = ebc[0]; n[1] = ebc[1]; n[2] = ebc[2];
		if (r2 > a * a) {
			double fac1 = -P->externalMagneticField[2] * pow(a, 3) / pow(r2, 2.5);
			n[6] = 3.0 * d[0] * d[2] * fac1;
			n[7] = 3.0 * d[1] * d[2] * fac1;
			n[8] = (2.0 * d[2] * d[2] - d[0] * d[0] - d[1] * d[1]) * fac1;
		} else { n[6] = 0.0; n[7] = 0.0; n[8] = 0.0; }
		for (int q = 0; q < 3; q++) n[3 + q] = B0[q] + n[6 + q];
	}
	for (int64_t x = 1; x < u->nx + 1; x++) for (int64_t y = 1; y < u->ny + 1; y++) for (int64_t z = 1; z < u->nz + 1; z++) {
		const int64_t pos[3] = { x, y, z };
		interp_n2c(u, pos);
	}
}

/* ---- time loop: simulator.h:181-230 (STEP 1/2 as in the commented code, off unless flagged) ---- */

void orc_step(orc_universe* u, uint64_t nsteps, uint32_t flags) {
	for (uint64_t s = 0; s < nsteps; s++) {
		if (flags & IPIC_PROJECT_CURRENT)
			for (int64_t x = 0; x <= u->nx; x++) for (int64_t y = 0; y <= u->ny; y++) for (int64_t z = 0; z <= u->nz; z++) orc_project_node(u, x, y, z);
		if (flags & IPIC_UPDATE_BOUNDARIES) orc_update_boundaries(u);
		if (flags & (IPIC_SOLVER_FORWARD | IPIC_SOLVER_LEAPFROG))
			for (int64_t x = 1; x < u->nx + 2; x++) for (int64_t y = 1; y < u->ny + 2; y++) for (int64_t z = 1; z < u->nz + 2; z++) {
				if (flags & IPIC_SOLVER_FORWARD) orc_solve_forward_node(u, x, y, z); else orc_solve_leapfrog_node(u, x, y, z);
			}
		for (int64_t x = 0; x < u->nx; x++) for (int64_t y = 0; y < u->ny; y++) for (int64_t z = 0; z < u->nz; z++) {
			const int64_t pos[3] = { x, y, z };
			plist* cell = &u->cells[cell_index(u, x, y, z)];
			if (flags & IPIC_MOVER_INTERPOLATED) move_cell_interpolated(u, cell, pos); else move_cell_dipole(u, cell);
			if (!(flags & IPIC_SKIP_MIGRATION)) export_cell(u, pos);
		}
		if (!(flags & IPIC_SKIP_MIGRATION))
			for (int64_t x = 0; x < u->nx; x++) for (int64_t y = 0; y < u->ny; y++) for (int64_t z = 0; z < u->nz; z++) {
				const int64_t pos[3] = { x, y, z };
				import_cell(u, pos);
			}
	}
}
