// Host-side launch interface of field_kernels.cu (explicit field stencils, ghost update, layout transforms).
#pragma once
#include "ipic_device.cuh"

namespace ipic {

// AoS FieldNode{E,B,Bext} (9 doubles) <-> device layout EB[node][6] + Bext[node][3]
void launch_split_field(const double* aos9, double* EB, double* Bext, int64_t nodes, cudaStream_t st);
void launch_merge_field(double* aos9, const double* EB, const double* Bext, int64_t nodes, cudaStream_t st);

// updateFieldsOnBoundaries (reference field.h:407-440); two launches
void launch_update_boundaries(const Geometry& g, double* EB, double* Bext, double* Bc, const int* stopFlag, cudaStream_t st);

// solveFieldForward / solveFieldLeapfrog over all nodes [1, n+2)^3 with the reference's in-place,
// lexicographic-order semantics, realised as a hyperplane wavefront; returns the number of launches
int launch_field_solver(const Geometry& g, bool forward, double* EB, double* Bc, const double* J, const int* stopFlag, cudaStream_t st);

// single-node variants (the granularity the reference's unit tests exercise)
void launch_field_solver_node(const Geometry& g, bool forward, double* EB, double* Bc, const double* J, int x, int y, int z, cudaStream_t st);

} // namespace ipic
