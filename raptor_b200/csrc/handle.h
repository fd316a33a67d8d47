// raptor_b200/csrc/handle.h -- the opaque handle behind include/b200_l2f.h and the host-side helpers every translation unit of the
// engine shares (status / error plumbing, pinned staging, AoS <-> SoA transposes).  The engine is split into several .cu files so that
// the heavy kernel instantiations compile in parallel; none of them calls device code of another (no relocatable device code needed).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>
#include <new>
#include <type_traits>

#include "../../include/b200_l2f.h"
#include "kernels.cuh"

namespace b200l2f {
enum SpecKind { KIND_DEFAULT = 0, KIND_RAPTOR = 1, KIND_TEACHER = 2 };
std::string& create_error();   // thread-local message of the last failed b200l2f_create (engine.cu)
}

struct b200l2f_handle {
    b200l2f_config cfg{};
    int kind = 0; bool dr = false; int H = 0, obs_dim = 0, sdim = 0;
    int n = 0;
    cudaStream_t stream = nullptr; bool own_stream = false;
    float* d_params = nullptr;       // [145][n]
    float* d_env_row = nullptr;      // [145]
    float h_env_row[B200L2F_PARAMS_DIM];
    std::vector<float*> d_state;     // slots x [sdim][n]
    uint64_t* d_rng = nullptr;
    int* d_flags = nullptr;          // [0] error flag, [1] parameter features
    bool features_dirty = true; int features = 0;
    // the feature pass of the LAST parameter upload, enqueued right behind it on the main stream: results land in page-locked host words, refresh_features
    // then waits for the event only (no launch + synchronise in front of the rollout).  Any later writer of d_params bumps params_version, which invalidates it.
    cudaEvent_t features_ready = nullptr; int* h_features = nullptr; float* h_row0_pinned = nullptr; int64_t params_version = 0, features_version = -1;
    bool params_follow_env_row = false;   // every column was filled from h_env_row (initial / sampled parameters, collect's resets) and not edited since
    float row0[B200L2F_PARAMS_DIM];   // parameter row of environment 0 (uniform MDP constants of the fused kernels)
    // actor
    bool policy_loaded = false; b200l2f_policy_desc pol{};
    float* d_blob = nullptr; size_t blob_floats = 0;
    float* d_hidden = nullptr; int* d_gru_step = nullptr;
    int* d_sched = nullptr; size_t sched_ints = 0;     // work counter + per-tile progress of the time-chunked scheduler
    float* d_acc_ret = nullptr; int* d_acc_len = nullptr;
    float* d_tc_image = nullptr;     // tensor-core weight image (hi/lo planes of the three B operands, TMA source)
    float* d_ts_image = nullptr;     // the same with scaled GRU gate rows (k_rollout_raptor_ts)
    float* d_mlp_tc_image = nullptr; // same for an MLP actor (MlpTcImage<IN, OUT>), null when the actor has no tensor-core instantiation
    // critic of the learner feed (learner.cu): MLP blob with the single output row padded to 4, and its tensor-core operand image
    bool critic_loaded = false; int critic_std = 0, critic_gemm = 0;
    float* d_critic_blob = nullptr; float* d_critic_tc_image = nullptr;
    double* d_colstats = nullptr; size_t colstats_doubles = 0;
    // DAgger data path (dagger.cu): teacher weights (tensor-core images + blobs), steady-state offsets, recorded student rollout
    float* d_teacher_images = nullptr; float* d_teacher_blobs = nullptr; float* d_teacher_offsets = nullptr;
    int n_teachers = 0, episodes_per_teacher = 0, teacher_gemm = 0;
    float* d_dg_states = nullptr; uint8_t* d_dg_term = nullptr; size_t dg_state_floats = 0;
    int* d_dg_eplen = nullptr; int* d_dg_offsets = nullptr; float* d_dg_returns = nullptr;
    std::vector<float> h_image;      // k-major actor image passed by value to the fused kernels (constant-bank weights)
    bool weights_in_constant_bank = false; bool rolled = false;
    // PPO runner bookkeeping (rl/components/on_policy_runner/on_policy_runner.h: episode_step, episode_return, truncated)
    int* d_episode_step = nullptr; float* d_episode_return = nullptr; uint8_t* d_truncated = nullptr;
    // staging
    void* h_pinned = nullptr; size_t pinned_bytes = 0;
    cudaEvent_t pinned_read = nullptr; bool pinned_read_pending = false;   // an async H2D copy out of h_pinned may still be in flight
    void* d_stage = nullptr; size_t stage_bytes = 0;
    // asynchronous host <-> device pipeline (b200l2f_*_async): two copy streams beside the main stream, one upload staging buffer per kind and one download
    // staging buffer, each guarded by a (ready, free) event pair so that transfers of step k+1 / k-1 run under the kernels of step k
    struct Xfer {
        cudaStream_t h2d = nullptr, d2h = nullptr;
        float* up_params = nullptr; float* up_state = nullptr; float* dl_state = nullptr;
        cudaEvent_t params_ready = nullptr, params_free = nullptr, state_ready = nullptr, state_free = nullptr, dl_ready = nullptr, dl_free = nullptr, main_mark = nullptr;
        bool params_used = false, state_used = false, dl_used = false;
        // downloads are ENQUEUED lazily: the copy would otherwise start the moment its source is ready -- right behind the kernel, where its PCIe traffic delays
        // the (tiny, host-visible) feature words and completion signals the next launch is waiting for (measured: +0.1 ms per pipelined rollout).  They are handed
        // to the d2h stream once the next rollout has been launched (or at b200l2f_transfers_synchronize), and then overlap that kernel.
        struct Pending { void* dst; const void* src; size_t bytes; cudaEvent_t after; bool releases_dl_staging; };
        std::vector<Pending> pending;
        std::vector<cudaEvent_t> event_pool;
    } xfer;
    // status of the last fused call (b200l2f_last_status): per-environment returns / lengths / done flags the kernels always write, reduction scratch
    float* d_last_returns = nullptr; int* d_last_eplen = nullptr; uint8_t* d_last_done = nullptr; uint8_t* d_nonfinite = nullptr;
    void* d_status = nullptr;        // StatusPartial[STATUS_BLOCKS + 1]
    bool status_valid = false, status_has_episodes = false;
    std::string err;
    int64_t launches = 0;
    const char* last_kernel = "";   // name of the fused kernel the last rollout / collection launched (b200l2f_last_kernel: profiling tools, bench.py)
};

namespace b200l2f {

inline int fail(b200l2f_handle* h, int code, const std::string& msg){
    if(h) h->err = msg; else create_error() = msg;
    return code;
}
#define CU(call) do{ cudaError_t e_ = (call); if(e_ != cudaSuccess){ return fail(h, B200L2F_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } }while(0)
#define LAUNCH_CHECK() do{ h->launches++; cudaError_t e_ = cudaGetLastError(); if(e_ != cudaSuccess){ return fail(h, B200L2F_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e_)); } }while(0)

inline int grid_for(int n, int block){ return (n + block - 1) / block; }

// The bounce buffer is read by asynchronous H2D copies: the host may write it again (or free it) only after the last such copy has finished.
inline int pinned_wait(b200l2f_handle* h){
    if(h->pinned_read_pending){ CU(cudaEventSynchronize(h->pinned_read)); h->pinned_read_pending = false; }
    return B200L2F_OK;
}
inline int ensure_pinned(b200l2f_handle* h, size_t bytes){
    int rc0;
    if((rc0 = pinned_wait(h))) return rc0;
    if(bytes <= h->pinned_bytes) return B200L2F_OK;
    if(h->h_pinned){ cudaFreeHost(h->h_pinned); h->h_pinned = nullptr; h->pinned_bytes = 0; }
    CU(cudaMallocHost(&h->h_pinned, bytes));
    h->pinned_bytes = bytes;
    return B200L2F_OK;
}
inline int ensure_stage(b200l2f_handle* h, size_t bytes){
    if(bytes <= h->stage_bytes) return B200L2F_OK;
    if(h->d_stage){ cudaFree(h->d_stage); h->d_stage = nullptr; h->stage_bytes = 0; }
    CU(cudaMalloc(&h->d_stage, bytes));
    h->stage_bytes = bytes;
    return B200L2F_OK;
}
// is this host pointer page-locked (cudaMallocHost / cudaHostRegister / torch pin_memory)?  Then the DMA engine can read/write it directly.
inline bool is_pinned_host(const void* p){
    cudaPointerAttributes attr;
    if(cudaPointerGetAttributes(&attr, p) != cudaSuccess){ cudaGetLastError(); return false; }
    return attr.type == cudaMemoryTypeHost;
}
// host -> device staging buffer, returns device pointer in *dev.  Pageable memory bounces through the handle's pinned buffer (the caller's
// buffer may be reused as soon as the call returns); page-locked memory is read by the DMA engine directly and asynchronously: the caller must
// not modify it before the stream has passed the copy (b200l2f_synchronize, or any call that returns host results).
inline int upload(b200l2f_handle* h, const void* src, size_t bytes, int memspace, const void** dev){
    if(memspace == B200L2F_DEVICE){ *dev = src; return B200L2F_OK; }
    int rc;
    if((rc = ensure_stage(h, bytes))) return rc;
    if(is_pinned_host(src)){
        CU(cudaMemcpyAsync(h->d_stage, src, bytes, cudaMemcpyHostToDevice, h->stream));
    }
    else{
        if((rc = ensure_pinned(h, bytes))) return rc;         // waits for the previous copy out of the bounce buffer
        std::memcpy(h->h_pinned, src, bytes);
        CU(cudaMemcpyAsync(h->d_stage, h->h_pinned, bytes, cudaMemcpyHostToDevice, h->stream));
        if(!h->pinned_read) CU(cudaEventCreateWithFlags(&h->pinned_read, cudaEventDisableTiming));
        CU(cudaEventRecord(h->pinned_read, h->stream));
        h->pinned_read_pending = true;
    }
    *dev = h->d_stage;
    return B200L2F_OK;
}
// device result buffer: the caller's pointer (device) or the staging buffer (host); finish() copies back
inline int result_buffer(b200l2f_handle* h, void* dst, size_t bytes, int memspace, void** dev, size_t stage_offset = 0){
    if(memspace == B200L2F_DEVICE){ *dev = dst; return B200L2F_OK; }
    int rc;
    if((rc = ensure_stage(h, stage_offset + bytes))) return rc;
    *dev = (char*)h->d_stage + stage_offset;
    return B200L2F_OK;
}
inline int download(b200l2f_handle* h, void* dst, const void* dev, size_t bytes, int memspace){
    if(memspace == B200L2F_DEVICE) return B200L2F_OK;
    int rc;
    if(is_pinned_host(dst)){
        CU(cudaMemcpyAsync(dst, dev, bytes, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        return B200L2F_OK;
    }
    if((rc = ensure_pinned(h, bytes))) return rc;
    CU(cudaMemcpyAsync(h->h_pinned, dev, bytes, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    std::memcpy(dst, h->h_pinned, bytes);
    return B200L2F_OK;
}
inline int transpose(b200l2f_handle* h, const float* in, float* out, int rows, int cols){
    dim3 block(32, 8), grid((cols + 31) / 32, (rows + 31) / 32);
    k_transpose<<<grid, block, 0, h->stream>>>(in, out, rows, cols);
    LAUNCH_CHECK();
    return B200L2F_OK;
}
inline int check_slot(b200l2f_handle* h, int slot){
    if(slot < 0 || slot >= (int)h->d_state.size()) return fail(h, B200L2F_ERR_ARGUMENT, "state slot out of range");
    return B200L2F_OK;
}

template <class F>
int dispatch_spec(b200l2f_handle* h, F&& f){
    switch(h->kind){
        case KIND_DEFAULT: return f(SpecDefault{});
        case KIND_RAPTOR: return f(SpecRaptor{});
        case KIND_TEACHER: return f(SpecTeacher{});
    }
    return fail(h, B200L2F_ERR_UNSUPPORTED, "unknown spec");
}

int refresh_features(b200l2f_handle* h);                                              // engine.cu
int flush_pending_downloads(b200l2f_handle* h);                                      // engine.cu
int enqueue_features(b200l2f_handle* h);
int enqueue_status(b200l2f_handle* h, const float* d_returns, const int* d_eplen, const uint8_t* d_done);   // engine.cu: status reduction behind the fused kernel just launched                                              // engine.cu: feature pass behind the parameter write just enqueued
int prepare_schedule(b200l2f_handle* h, RolloutArgs& a, int cap_in, int* grid, int tile_envs = BLOCK);   // engine.cu; tile_envs = environments per work item

}  // namespace b200l2f
