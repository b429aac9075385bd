// Host-side launch interface of tracer_kernels.cu (pic_tracer: test particles in the analytic dipole).
#pragma once
#include "ipic_device.cuh"

namespace ipic {

// traces particles [0, n) of the AoS array for T steps; counts[frame][x][y][z] += 1 every frameInterval steps
void launch_trace(const Geometry& g, const double* aos, int64_t n, int T, int frameInterval, unsigned* counts, cudaStream_t st);

} // namespace ipic
