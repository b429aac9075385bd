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
// one piece of a z-neighbour exchange: `dir` +1 = the slab above (rank+1), -1 = the slab below (rank-1), periodic
struct CommPiece { void* p; size_t bytes; int dir; };
// all sends and receives in ONE NCCL group; pieces of zero bytes are skipped (both sides know the sizes).  With two
// ranks both neighbours are the same peer: NCCL matches the pieces per peer in the order given, so everybody lists
// sends as [to above..., to below...] and receives as [from below..., from above...].
int comm_sendrecv(ipic_comm* c, const CommPiece* sends, int nSends, const CommPiece* recvs, int nRecvs, cudaStream_t st, std::string& err);
// all-gather equally sized pieces in rank order
int comm_allgather(ipic_comm* c, const void* send, void* recv, size_t bytesPerRank, cudaStream_t st, std::string& err);
// all-gather pieces of different sizes in place: rank r's piece is bytes[r] at base + offsets[r] (one broadcast per rank, one NCCL group)
int comm_allgatherv(ipic_comm* c, void* base, const size_t* offsets, const size_t* bytes, cudaStream_t st, std::string& err);
