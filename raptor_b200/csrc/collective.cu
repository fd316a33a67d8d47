// raptor_b200/csrc/collective.cu -- the one multi-GPU call of the path: the OPTIONAL all-gather of trajectory slabs for a central learner.
// Environments shard by global id with no collective on the rollout path; when a learner wants every rank's dataset, the slabs the collection kernels wrote
// are gathered over NVLink / NVSwitch with NCCL, on the handle's stream (ordered behind the kernel that produced the slab, no host synchronisation).
// NCCL is bound at run time (dlopen of the libnccl.so.2 the process already carries, e.g. torch's): the engine library has no link-time dependency on it, and a
// single-GPU deployment never needs it.  The communicator is the caller's (ncclCommInitRank, one rank per GPU / process).
#include <dlfcn.h>
#include "handle.h"

using b200l2f::fail;

namespace {
using nccl_all_gather_t = int (*)(const void* sendbuff, void* recvbuff, size_t sendcount, int datatype, void* comm, cudaStream_t stream);
using nccl_comm_count_t = int (*)(void* comm, int* count);
using nccl_error_string_t = const char* (*)(int);
struct Nccl { void* lib = nullptr; nccl_all_gather_t all_gather = nullptr; nccl_comm_count_t comm_count = nullptr; nccl_error_string_t error_string = nullptr; std::string err; };
Nccl& nccl(){
    static Nccl n = [](){
        Nccl r;
        const char* env = std::getenv("B200L2F_NCCL_LIB");
        const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
        for(const char* name : names){
            if(!name) continue;
            r.lib = dlopen(name, RTLD_NOW | RTLD_NOLOAD);             // the copy already in the process (torch's bundled NCCL), so both sides share one library
            if(!r.lib) r.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if(r.lib) break;
        }
        if(!r.lib){ r.err = "libnccl.so.2 not found (set B200L2F_NCCL_LIB to its path)"; return r; }
        r.all_gather = (nccl_all_gather_t)dlsym(r.lib, "ncclAllGather");
        r.comm_count = (nccl_comm_count_t)dlsym(r.lib, "ncclCommCount");
        r.error_string = (nccl_error_string_t)dlsym(r.lib, "ncclGetErrorString");
        if(!r.all_gather || !r.comm_count) r.err = "libnccl does not export ncclAllGather / ncclCommCount";
        return r;
    }();
    return n;
}
constexpr int NCCL_FLOAT32 = 7;   // ncclDataType_t ncclFloat32 (nccl.h)
}  // namespace

extern "C" {

int b200l2f_allgather_trajectories(b200l2f_handle* h, void* nccl_comm, const float* send, float* recv, size_t count_per_rank, int32_t* n_ranks_out){
    if(!h) return fail(h, B200L2F_ERR_ARGUMENT, "allgather_trajectories: null handle");
    CU(cudaSetDevice(h->cfg.device));
    if(!nccl_comm || !send || !recv) return fail(h, B200L2F_ERR_ARGUMENT, "allgather_trajectories: null argument");
    Nccl& n = nccl();
    if(!n.err.empty()) return fail(h, B200L2F_ERR_UNSUPPORTED, "allgather_trajectories: " + n.err);
    int ranks = 0;
    int rc = n.comm_count(nccl_comm, &ranks);
    if(rc != 0) return fail(h, B200L2F_ERR_CUDA, std::string("allgather_trajectories: ncclCommCount: ") + (n.error_string ? n.error_string(rc) : "error"));
    if(n_ranks_out) *n_ranks_out = ranks;
    if(count_per_rank == 0) return B200L2F_OK;
    rc = n.all_gather(send, recv, count_per_rank, NCCL_FLOAT32, nccl_comm, h->stream);
    if(rc != 0) return fail(h, B200L2F_ERR_CUDA, std::string("allgather_trajectories: ncclAllGather: ") + (n.error_string ? n.error_string(rc) : "error"));
    return B200L2F_OK;
}

}  // extern "C"
