// NCCL point-to-point plumbing for the z-slab decomposition (see comm.h).
#include "comm.h"

#include <dlfcn.h>
#include <cstdio>
#include <cstring>

#include "../../include/ipic3d_b200.h"

namespace {

// minimal NCCL surface (nccl.h, 2.x ABI)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0 };

struct NcclApi {
	void* handle = nullptr;
	int (*GetUniqueId)(ncclUniqueId*) = nullptr;
	int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	int (*CommDestroy)(ncclComm_t) = nullptr;
	int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
	int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	int (*GroupStart)() = nullptr;
	int (*GroupEnd)() = nullptr;
	const char* (*GetErrorString)(int) = nullptr;
};

NcclApi* api(std::string& err) {
	static NcclApi a;
	static bool tried = false;
	if (a.handle) return &a;
	if (tried) { err = "NCCL library not available"; return nullptr; }
	tried = true;
	const char* names[] = { "libnccl.so.2", "libnccl.so" };
	for (const char* n : names) { a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (a.handle) break; }
	if (!a.handle) { err = std::string("cannot load NCCL: ") + dlerror(); return nullptr; }
#define SYM(field, name) a.field = (decltype(a.field))dlsym(a.handle, name); if (!a.field) { err = "NCCL symbol missing: " name; a.handle = nullptr; return nullptr; }
	SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
	SYM(Send, "ncclSend") SYM(Recv, "ncclRecv") SYM(AllGather, "ncclAllGather") SYM(Broadcast, "ncclBroadcast") SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd")
	SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
	return &a;
}

} // namespace

struct ipic_comm { ncclComm_t comm = nullptr; int rank = 0, nranks = 1; };

#define NC(call) do { int r_ = (call); if (r_ != ncclSuccess) { err = std::string(#call " failed: ") + a->GetErrorString(r_); return IPIC_ERR_COMM; } } while (0)

int comm_unique_id(void* id128, std::string& err) {
	NcclApi* a = api(err);
	if (!a) return IPIC_ERR_COMM;
	if (!id128) { err = "null id"; return IPIC_ERR_ARG; }
	ncclUniqueId id;
	NC(a->GetUniqueId(&id));
	std::memcpy(id128, &id, sizeof id);
	return IPIC_OK;
}

int comm_create(ipic_comm** out, const void* id128, int rank, int nranks, std::string& err) {
	NcclApi* a = api(err);
	if (!a) return IPIC_ERR_COMM;
	ncclUniqueId id;
	std::memcpy(&id, id128, sizeof id);
	ipic_comm* c = new ipic_comm();
	c->rank = rank; c->nranks = nranks;
	NC(a->CommInitRank(&c->comm, nranks, id, rank));
	*out = c;
	return IPIC_OK;
}

void comm_destroy(ipic_comm* c) {
	if (!c) return;
	std::string err;
	NcclApi* a = api(err);
	if (a && c->comm) a->CommDestroy(c->comm);
	delete c;
}

int comm_sendrecv(ipic_comm* c, const CommPiece* sends, int nSends, const CommPiece* recvs, int nRecvs, cudaStream_t st, std::string& err) {
	NcclApi* a = api(err);
	if (!a || !c) { if (a) err = "no communicator"; return IPIC_ERR_COMM; }
	const int up = (c->rank + 1) % c->nranks, down = (c->rank + c->nranks - 1) % c->nranks;
	NC(a->GroupStart());
	for (int i = 0; i < nSends; i++) if (sends[i].bytes) NC(a->Send(sends[i].p, sends[i].bytes, ncclInt8, sends[i].dir > 0 ? up : down, c->comm, st));
	for (int i = 0; i < nRecvs; i++) if (recvs[i].bytes) NC(a->Recv(recvs[i].p, recvs[i].bytes, ncclInt8, recvs[i].dir > 0 ? up : down, c->comm, st));
	NC(a->GroupEnd());
	return IPIC_OK;
}

int comm_allgather(ipic_comm* c, const void* send, void* recv, size_t bytesPerRank, cudaStream_t st, std::string& err) {
	NcclApi* a = api(err);
	if (!a || !c) { if (a) err = "no communicator"; return IPIC_ERR_COMM; }
	NC(a->AllGather(send, recv, bytesPerRank, ncclInt8, c->comm, st));
	return IPIC_OK;
}

int comm_allgatherv(ipic_comm* c, void* base, const size_t* offsets, const size_t* bytes, cudaStream_t st, std::string& err) {
	NcclApi* a = api(err);
	if (!a || !c) { if (a) err = "no communicator"; return IPIC_ERR_COMM; }
	NC(a->GroupStart());
	for (int r = 0; r < c->nranks; r++) {
		if (!bytes[r]) continue;
		char* piece = static_cast<char*>(base) + offsets[r];
		NC(a->Broadcast(piece, piece, bytes[r], ncclInt8, r, c->comm, st));
	}
	NC(a->GroupEnd());
	return IPIC_OK;
}
