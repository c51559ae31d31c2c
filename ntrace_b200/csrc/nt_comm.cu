// ntrace_b200 — multi-GPU replication of the resident BVH through NCCL, inside the C ABI.
//
// NEW functionality: the reference is single-context (src/framework/gpu/CudaModule.hpp:92-97) and has no collectives.  north_star /
// one process per GPU, the BVH built on one rank and replicated to the others with an NCCL broadcast over NVLink, rays
// sharded with no collective on the ray path.  NCCL is bound at run time (dlopen of libnccl.so.2: the library still links only cudart,
// and a single-GPU host never needs NCCL to be installed); the handful of prototypes used are declared here, as nccl.h declares them.
// The communicator bootstrap is the host's: rank 0 asks for a 128-byte unique id and ships it to the other ranks by any means (a file,
// a socket, MPI, torch.distributed), then every rank calls nt_comm_init.
#include "../../include/ntrace_b200.h"
#include "nt_common.cuh"

#include <dlfcn.h>
#include <cstdlib>
#include <cstring>

namespace nt {
namespace {

typedef struct { char internal[128]; } NcclUniqueId;      // nccl.h:37-38
typedef void* NcclComm;
enum { kNcclUint8 = 1, kNcclFloat64 = 8, kNcclSum = 0, kNcclMax = 2 };   // nccl.h: ncclDataType_t / ncclRedOp_t

struct Nccl {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    NcclComm comm = nullptr;
    int numRanks = 0, rank = -1;
    DevBuf scratch;
} n;

int load_nccl()
{
    if (n.lib) return 0;
    const char* names[] = {getenv("NTRACE_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* name : names) {
        if (!name || !*name) continue;
        n.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
        if (n.lib) break;
    }
    if (!n.lib) { set_error(std::string("ntrace_b200: cannot load NCCL (libnccl.so.2; set NTRACE_NCCL_LIB): ") + (dlerror() ? dlerror() : "")); return 1; }
    bool ok = true;
    auto sym = [&](const char* s) { void* p = dlsym(n.lib, s); if (!p) ok = false; return p; };
    n.GetUniqueId = (int (*)(NcclUniqueId*))sym("ncclGetUniqueId");
    n.CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int))sym("ncclCommInitRank");
    n.CommDestroy = (int (*)(NcclComm))sym("ncclCommDestroy");
    n.Broadcast = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))sym("ncclBroadcast");
    n.AllReduce = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))sym("ncclAllReduce");
    n.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    if (!ok) { set_error("ntrace_b200: the NCCL library lacks a required symbol"); dlclose(n.lib); n.lib = nullptr; return 1; }
    return 0;
}

} // namespace

bool check_nccl(int rc, const char* what)
{
    if (rc == 0) return true;
    set_error(std::string("NCCL error in ") + what + ": " + (n.GetErrorString ? n.GetErrorString(rc) : "?"));
    return false;
}
#define NT_NCCL(call) do { if (!::nt::check_nccl((call), #call)) return 1; } while (0)

int comm_unique_id(void* out128)
{
    if (!out128) { set_error("ntrace_b200: null output"); return 1; }
    if (load_nccl()) return 1;
    NcclUniqueId id;
    NT_NCCL(n.GetUniqueId(&id));
    memcpy(out128, &id, 128);
    return 0;
}

int comm_init(int numRanks, int rank, const void* uniqueId128)
{
    if (numRanks < 1 || rank < 0 || rank >= numRanks || !uniqueId128) { set_error("ntrace_b200: bad communicator arguments"); return 1; }
    if (n.comm) { set_error("ntrace_b200: communicator already initialised; call nt_comm_destroy first"); return 1; }
    if (load_nccl()) return 1;
    NcclUniqueId id;
    memcpy(&id, uniqueId128, 128);
    NT_NCCL(n.CommInitRank(&n.comm, numRanks, id, rank));
    n.numRanks = numRanks; n.rank = rank;
    return 0;
}

int comm_destroy()
{
    if (n.comm) { n.CommDestroy(n.comm); n.comm = nullptr; }
    n.scratch.release();
    n.numRanks = 0; n.rank = -1;
    return 0;
}

bool comm_ready() { return n.comm != nullptr; }
int comm_rank() { return n.rank; }
int comm_size() { return n.numRanks; }

int comm_broadcast_bytes(void* devPtr, size_t bytes, int root, cudaStream_t stream)
{
    if (!bytes) return 0;
    NT_NCCL(n.Broadcast(devPtr, devPtr, bytes, kNcclUint8, root, n.comm, stream));
    return 0;
}

// host-side convenience for the drivers above the ABI: element-wise sum (op 0) or max (op 1) of `count` doubles over the ranks
int comm_allreduce_f64(double* hostValues, int count, int op, cudaStream_t stream)
{
    if (!n.comm) { set_error("ntrace_b200: no communicator (nt_comm_init)"); return 1; }
    if (count <= 0 || !hostValues || (op != 0 && op != 1)) { set_error("ntrace_b200: bad all-reduce arguments"); return 1; }
    NT_CUDA(n.scratch.reserve((size_t)count * 8));
    NT_CUDA(cudaMemcpyAsync(n.scratch.p, hostValues, (size_t)count * 8, cudaMemcpyHostToDevice, stream));
    NT_NCCL(n.AllReduce(n.scratch.p, n.scratch.p, (size_t)count, kNcclFloat64, op == 0 ? kNcclSum : kNcclMax, n.comm, stream));
    NT_CUDA(cudaMemcpyAsync(hostValues, n.scratch.p, (size_t)count * 8, cudaMemcpyDeviceToHost, stream));
    NT_CUDA(cudaStreamSynchronize(stream));
    return 0;
}

} // namespace nt
