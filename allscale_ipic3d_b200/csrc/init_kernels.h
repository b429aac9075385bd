// Host-side launch interface of init_kernels.cu (device generator of the `:U:N` benchmark universe).
#pragma once
#include "ipic_device.cuh"

namespace ipic {

// fills the live head of every owned cell (count[cell] particles) with the reference's uniform-plasma particles
void launch_init_uniform(const Geometry& g, const ParticleArrays& p, const int64_t* cellStart, const int* count, cudaStream_t st);

} // namespace ipic
