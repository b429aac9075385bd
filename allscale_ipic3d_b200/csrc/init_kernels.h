// Host-side launch interface of init_kernels.cu (device generator of the `:U:N` benchmark universe).
#pragma once
#include "ipic_device.cuh"

namespace ipic {

// fills the live head of every owned cell (count[cell] particles) with the reference's uniform-plasma particles
void launch_init_uniform(const Geometry& g, const ParticleArrays& p, const int64_t* cellStart, const int* count, cudaStream_t st);

// tables of the reference's ziggurat (hdr/ziggurat_normal_distribution.h:107-125, zigset), computed on the host
struct ZigguratTables { unsigned long long kn[128]; float wn[128]; float fn[128]; };

// the universe of an `.inp` deck: uniform positions per cell + normal velocities (two launches), count[cell] particles per cell
void launch_init_deck(const Geometry& g, const ParticleArrays& p, const int64_t* cellStart, const int* count, const ZigguratTables* tables,
                      const double mean[3], const double stddev[3], cudaStream_t st);

} // namespace ipic
