// z-slab neighbour exchange over NCCL (send/recv), resolved at run time with dlopen so that the
// single-GPU library has no link-time dependency on NCCL.
#pragma once
#include <cstdint>
#include <string>
#include <cuda_runtime.h>

struct ipic_comm;

int comm_unique_id(void* id128, std::string& err);
int comm_create(ipic_comm** out, const void* id128, int rank, int nranks, std::string& err);
void comm_destroy(ipic_comm* c);
// exchange `bytes` with both z-neighbours: sendUp -> rank+1, sendDown -> rank-1 (periodic), receive likewise
int comm_exchange(ipic_comm* c, const void* sendUp, void* recvFromDown, const void* sendDown, void* recvFromUp, size_t bytes, cudaStream_t st, std::string& err);
// all-gather equally sized pieces in rank order
int comm_allgather(ipic_comm* c, const void* send, void* recv, size_t bytesPerRank, cudaStream_t st, std::string& err);
