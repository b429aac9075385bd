// Host-side launch interface of diag_kernels.cu (conserved-quantity reductions).
#pragma once
#include "ipic_device.cuh"

namespace ipic {

constexpr int kFieldEnergyChunks = 148;      // CTAs of the field-energy pass (static node -> CTA assignment)

// out2[0] = total kinetic energy, out2[1] = total momentum of the owned cells; partial: 2 doubles per owned cell
void launch_particle_energy(const Geometry& g, const ParticleArrays& p, const int64_t* cellStart, const int* count,
                            double* partial, double* out2, cudaStream_t st);
// out2[0] = E energy, out2[1] = (B + Bext) energy over the interior nodes; partial: 2 doubles per chunk
void launch_field_energy(const Geometry& g, const double* EB, const double* Bext, double* partial, int chunks, double* out2, cudaStream_t st);

} // namespace ipic
