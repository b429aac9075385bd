/*
 * TEST DOUBLE (tests/ only) of the few C-ABI entry points the reference-side binding calls
 * (host/ref_binding/ipic3d/app/simulator.h), implemented on the CPU oracle.  It exists so that the binding's host logic --
 * flattening the reference's Grid<Cell,3> into counts + AoS, passing the fields, scattering the results back -- can be
 * checked end to end on a machine without a GPU: the reference's unmodified driver, built with the binding and run against
 * this double, must write the same .out as the reference itself.  Built into a scratch directory by
 * tests/test_reference_binding.py and put in front of the real library with LD_LIBRARY_PATH; never shipped, never linked
 * by the product.
 */
#include <stdlib.h>
#include <string.h>
#include "ipic3d_b200.h"
#include "ipic3d_oracle.h"

struct ipic_ctx { orc_universe* u; ipic_props props; };

const char* ipic_last_error(void) { return "mock C ABI (CPU oracle)"; }

int ipic_create(const ipic_props* props, int device, ipic_ctx** out) {
	(void)device;
	if (!props || !out) return IPIC_ERR_ARG;
	ipic_ctx* c = (ipic_ctx*)calloc(1, sizeof *c);
	c->props = *props;
	c->u = orc_create(props);
	*out = c;
	return IPIC_OK;
}

int ipic_destroy(ipic_ctx* c) { if (c) { orc_destroy(c->u); free(c); } return IPIC_OK; }

int ipic_upload_cells(ipic_ctx* c, const uint64_t* counts, const double* aos) { orc_set_cells(c->u, counts, aos); return IPIC_OK; }

int ipic_upload_fields(ipic_ctx* c, const double* field, const double* bc, const double* J) {
	if (field) orc_set_field(c->u, field);
	if (bc) orc_set_bcfield(c->u, bc);
	if (J) orc_set_density(c->u, J);
	return IPIC_OK;
}

int ipic_step(ipic_ctx* c, uint64_t n, uint32_t flags, double* t_first, double* t_rest) {
	orc_step(c->u, n, flags);
	if (t_first) *t_first = 0.0;
	if (t_rest) *t_rest = 0.0;
	return IPIC_OK;
}

int ipic_total_particles(ipic_ctx* c, uint64_t* total) { *total = orc_total_particles(c->u); return IPIC_OK; }

int ipic_download_cells(ipic_ctx* c, uint64_t* counts, double* aos, uint64_t capacity) {
	if (orc_total_particles(c->u) > capacity) return IPIC_ERR_CAPACITY;
	orc_get_counts(c->u, counts);
	orc_get_particles(c->u, aos);
	return IPIC_OK;
}

int ipic_download_fields(ipic_ctx* c, double* field, double* bc, double* J) {
	if (field) orc_get_field(c->u, field);
	if (bc) orc_get_bcfield(c->u, bc);
	if (J) orc_get_density(c->u, J);
	return IPIC_OK;
}
