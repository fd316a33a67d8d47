// raptor_b200/csrc/engine.cu -- C ABI (include/b200_l2f.h) of the B200-native quadrotor rollout engine: lifetime, RNG, parameters, state,
// the vector-API calls, actor loading and the host side of the fused rollout / collection.
//
// The handle (handle.h) owns struct-of-arrays device buffers (parameters [145][n], K state slots [STATE_DIM][n], RNG [n], actor hidden
// state [HD][n]); every entry point enqueues hand-written sm_100a kernels on the handle's stream.  There is no CPU path.  The heavy kernel
// instantiations live in rollout_*.cu / collect*.cu (launch.h).
#include "launch.h"
#include <cmath>
#include "rollout_tc.cuh"   // TcImage, build_tc_image_host

using namespace b200l2f;

namespace b200l2f {
std::string& create_error(){ thread_local std::string e; return e; }
}

namespace {
void nominal_parameters(int spec, float* p);


// nominal parameter values of the supported specifications: rl/environments/l2f/parameters/dynamics/crazyflie.h:10-123,
// parameters/default.h:34-134 (reward, termination, noise, DR ranges of the DR-enabled factory, trajectory), parameters/init/default.h:22-31
void nominal_parameters(int spec, float* p){
    std::memset(p, 0, sizeof(float) * B200L2F_PARAMS_DIM);
    const float pos[4][3] = {{0.028f, -0.028f, 0}, {-0.028f, -0.028f, 0}, {-0.028f, 0.028f, 0}, {0.028f, 0.028f, 0}};
    const float tdir[4] = {-1, +1, -1, +1};
    for(int i = 0; i < 4; i++){
        for(int j = 0; j < 3; j++) p[P_ROTOR_POS + 3 * i + j] = pos[i][j];
        p[P_THRUST_DIR + 3 * i + 2] = 1;
        p[P_TORQUE_DIR + 3 * i + 2] = tdir[i];
        p[P_THRUST_COEF + 3 * i + 0] = (float)0.00352526;
        p[P_THRUST_COEF + 3 * i + 1] = (float)0.01437313;
        p[P_THRUST_COEF + 3 * i + 2] = (float)0.09223048;
        p[P_TORQUE_CONST + i] = (float)4.665e-3;
        p[P_TAU_RISE + i] = (float)0.05545454545454546;
        p[P_TAU_FALL + i] = (float)0.24939393939393945;
    }
    p[P_MASS] = (float)(0.027 + 0.0017 + 0.0003 + 0.0016);
    p[P_GRAVITY + 2] = (float)-9.81;
    p[P_J + 0] = (float)9.416556729130406e-06; p[P_J + 4] = (float)9.644051701582312e-06; p[P_J + 8] = (float)1.745951732253285e-05;
    p[P_JINV + 0] = (float)106195.93007988465; p[P_JINV + 4] = (float)103690.85846314249; p[P_JINV + 8] = (float)57275.35197719487;
    p[P_HOVER] = (float)0.7261389721508553;
    p[P_ACT_MIN] = 0; p[P_ACT_MAX] = 1;
    p[P_DT] = (float)(1.0 / 100.0f);
    p[P_INIT_GUIDANCE] = (float)0.1; p[P_INIT_MAX_POS] = (float)0.5; p[P_INIT_MAX_ANGLE] = (float)1.5707963267948966;
    p[P_INIT_MAX_LINVEL] = 1; p[P_INIT_MAX_ANGVEL] = 1; p[P_INIT_REL_RPM] = 1; p[P_INIT_MIN_RPM] = -1; p[P_INIT_MAX_RPM] = 0;
    p[P_RW_SCALE] = 1; p[P_RW_CONSTANT] = (float)0.5; p[P_RW_TERM_PENALTY] = -100; p[P_RW_POSITION] = 1;
    p[P_RW_ORIENTATION] = (float)0.1; p[P_RW_DACTION] = 1;
    p[P_TERM_ENABLED] = 1; p[P_TERM_POS] = 1; p[P_TERM_LINVEL] = 2; p[P_TERM_ANGVEL] = 35; p[P_TERM_POS_INT] = 10000; p[P_TERM_ORI_INT] = 50000;
    if(spec == B200L2F_SPEC_DEFAULT_DR){
        p[P_DR_T2W_MIN] = (float)1.5; p[P_DR_T2W_MAX] = (float)5.0; p[P_DR_T2I_MIN] = (float)0.001; p[P_DR_T2I_MAX] = (float)0.100;
        p[P_DR_MASS_MIN] = (float)0.02; p[P_DR_MASS_MAX] = (float)5.00; p[P_DR_MASS_SIZE_DEV] = (float)0.1;
        p[P_DR_KQ_MIN] = (float)0.005; p[P_DR_KQ_MAX] = (float)0.05; p[P_DR_DIST_FORCE_MAX] = (float)0.1;
    }
    p[P_TRAJ_MIX0] = (float)0.5; p[P_TRAJ_MIX1] = (float)0.5;
    p[P_LANGEVIN_GAMMA] = 1; p[P_LANGEVIN_OMEGA] = 2; p[P_LANGEVIN_SIGMA] = (float)0.5; p[P_LANGEVIN_ALPHA] = (float)0.01;
}

}  // namespace

namespace b200l2f {

int enqueue_features(b200l2f_handle* h){
    if(!h->features_ready){
        CU(cudaEventCreateWithFlags(&h->features_ready, cudaEventDisableTiming));
        CU(cudaMallocHost(&h->h_features, sizeof(int)));
        CU(cudaMallocHost(&h->h_row0_pinned, sizeof(float) * B200L2F_PARAMS_DIM));
    }
    CU(cudaMemsetAsync(h->d_flags + 1, 0, sizeof(int), h->stream));
    k_param_features<<<grid_for(h->n, 256), 256, 0, h->stream>>>(h->d_params, h->n, h->d_flags + 1);
    LAUNCH_CHECK();
    k_publish_features<<<1, 160, 0, h->stream>>>(h->d_flags + 1, h->d_params, h->n, h->h_features, h->h_row0_pinned);   // page-locked host memory is device-addressable (UVA)
    LAUNCH_CHECK();
    CU(cudaEventRecord(h->features_ready, h->stream));
    h->features_version = h->params_version;
    return B200L2F_OK;
}
int refresh_features(b200l2f_handle* h){
    if(!h->features_dirty) return B200L2F_OK;
    int rc;
    if(h->features_version != h->params_version){ if((rc = enqueue_features(h))) return rc; }   // parameters written by a path that did not enqueue the pass itself
    CU(cudaEventSynchronize(h->features_ready));
    h->features = *h->h_features;
    std::memcpy(h->row0, h->h_row0_pinned, sizeof(float) * B200L2F_PARAMS_DIM);
    h->features_dirty = false;
    return B200L2F_OK;
}


int flush_pending_downloads(b200l2f_handle* h){
    auto& x = h->xfer;
    for(auto& p : x.pending){
        CU(cudaStreamWaitEvent(x.d2h, p.after, 0));
        CU(cudaMemcpyAsync(p.dst, p.src, p.bytes, cudaMemcpyDeviceToHost, x.d2h));
        if(p.releases_dl_staging) CU(cudaEventRecord(x.dl_free, x.d2h));
        else x.event_pool.push_back(p.after);               // the wait has been enqueued: the event object may be re-recorded
    }
    x.pending.clear();
    return B200L2F_OK;
}

constexpr int STATUS_BLOCKS = 256;
int enqueue_status(b200l2f_handle* h, const float* d_returns, const int* d_eplen, const uint8_t* d_done){
    if(!h->d_status){
        CU(cudaMalloc(&h->d_status, sizeof(StatusPartial) * (STATUS_BLOCKS + 1)));
        CU(cudaMalloc(&h->d_nonfinite, (size_t)h->n));
    }
    const int blocks = grid_for(h->n, 256) < STATUS_BLOCKS ? grid_for(h->n, 256) : STATUS_BLOCKS;
    StatusPartial* part = (StatusPartial*)h->d_status;
    k_status_partials<<<blocks, 256, 0, h->stream>>>(h->d_state[0], h->sdim, h->n, d_returns, d_eplen, d_done, h->d_nonfinite, part + 1);
    LAUNCH_CHECK();
    k_status_finish<<<1, 32, 0, h->stream>>>(part + 1, blocks, part);
    LAUNCH_CHECK();
    h->status_valid = true; h->status_has_episodes = d_returns != nullptr;
    return B200L2F_OK;
}

// persistent grid + work queue of the tcgen05 rollout kernels.  When the tiles do not fill an integer number of waves (e.g. 65 536 envs =
// 512 tiles on 444 slots), the rollout is cut into time chunks so that every slot stays busy until the end: makespan 512/444 instead of 2
// tile-times.  Fills a.sched / n_chunks / chunk_steps / acc_*, returns the grid size in *grid.
// Chunk count: 16 (chunks of >= 32 steps).  A model of the launch as ceil(n_tiles c / cap) synchronous rounds predicts c = 13 for 512 tiles on 444
// slots (15 full rounds, 1155 step-times against 1197 for c = 16), but the queue is not round-synchronous -- measured on B200 at T = 1000:
// c = 16 4.433 ms, 13 4.490, 6 4.480, 25 4.482 (profiles/r01_s11_chunk_sweep.log) -- so the measured choice stays.
int prepare_schedule(b200l2f_handle* h, RolloutArgs& a, int cap_in, int* grid, int tile_envs){
    const int n_tiles = grid_for(a.n, tile_envs);
    const int cap = cap_in > 0 ? cap_in : n_tiles;
    int n_chunks = 1;
    const int forced = [](){ const char* e = std::getenv("B200L2F_CHUNKS"); return e ? std::atoi(e) : 0; }();   // test / tuning override
    if(forced > 0) n_chunks = forced;
    else if(n_tiles > cap && n_tiles < 6 * cap && n_tiles % cap != 0 && a.T >= 64) n_chunks = a.T / 32 < 16 ? a.T / 32 : 16;
    if(n_chunks < 1) n_chunks = 1;
    if(std::getenv("B200L2F_VERBOSE")) std::fprintf(stderr, "[b200l2f] schedule: %d tiles on %d slots, T = %d -> %d time chunks\n", n_tiles, cap, a.T, n_chunks);
    a.chunk_steps = a.T > 0 ? (a.T + n_chunks - 1) / n_chunks : 0;
    a.n_chunks = a.T > 0 ? (a.T + a.chunk_steps - 1) / a.chunk_steps : 1;
    const size_t need = 1 + (size_t)n_tiles;
    if(need > h->sched_ints){
        cudaFree(h->d_sched); h->d_sched = nullptr; h->sched_ints = 0;
        CU(cudaMalloc(&h->d_sched, sizeof(int) * need));
        h->sched_ints = need;
    }
    if(!h->d_acc_ret){ CU(cudaMalloc(&h->d_acc_ret, sizeof(float) * (size_t)h->n)); CU(cudaMalloc(&h->d_acc_len, sizeof(int) * (size_t)h->n)); }
    CU(cudaMemsetAsync(h->d_sched, 0, sizeof(int) * need, h->stream));
    a.sched = h->d_sched; a.acc_ret = h->d_acc_ret; a.acc_len = h->d_acc_len;
    *grid = n_tiles < cap ? n_tiles : cap;
    return B200L2F_OK;
}

}  // namespace b200l2f


extern "C" {

const char* b200l2f_last_error(const b200l2f_handle* h){ return h ? h->err.c_str() : create_error().c_str(); }

int b200l2f_create(const b200l2f_config* config, b200l2f_handle** out){
    b200l2f_handle* h = nullptr;
    if(!config || !out) return fail(nullptr, B200L2F_ERR_ARGUMENT, "null argument");
    if(config->struct_size != (int32_t)sizeof(b200l2f_config)) return fail(nullptr, B200L2F_ERR_ARGUMENT, "b200l2f_config.struct_size mismatch");
    if(config->n_envs <= 0) return fail(nullptr, B200L2F_ERR_ARGUMENT, "n_envs must be positive");
    int count = 0;
    if(cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return fail(nullptr, B200L2F_ERR_NO_DEVICE, "no CUDA device visible: this engine has no CPU fallback");
    if(config->device < 0 || config->device >= count) return fail(nullptr, B200L2F_ERR_NO_DEVICE, "device ordinal out of range");
    cudaDeviceProp prop;
    if(cudaGetDeviceProperties(&prop, config->device) != cudaSuccess) return fail(nullptr, B200L2F_ERR_NO_DEVICE, "cudaGetDeviceProperties failed");
    if(prop.major != 10) return fail(nullptr, B200L2F_ERR_NO_DEVICE, std::string("device is sm_") + std::to_string(prop.major) + std::to_string(prop.minor) + ", the engine is built for sm_100a only");
    h = new (std::nothrow) b200l2f_handle();
    if(!h) return fail(nullptr, B200L2F_ERR_ARGUMENT, "out of host memory");
    h->cfg = *config;
    switch(config->spec){
        case B200L2F_SPEC_DEFAULT: h->kind = KIND_DEFAULT; h->dr = false; break;
        case B200L2F_SPEC_DEFAULT_DR: h->kind = KIND_DEFAULT; h->dr = true; break;
        case B200L2F_SPEC_RAPTOR: h->kind = KIND_RAPTOR; h->dr = false; break;
        case B200L2F_SPEC_RAPTOR_DR: h->kind = KIND_RAPTOR; h->dr = true; break;
        case B200L2F_SPEC_TEACHER: h->kind = KIND_TEACHER; h->dr = false; break;
        case B200L2F_SPEC_TEACHER_DR: h->kind = KIND_TEACHER; h->dr = true; break;
        default: delete h; return fail(nullptr, B200L2F_ERR_ARGUMENT, "unknown spec");
    }
    h->H = h->kind == KIND_DEFAULT ? 16 : 1;
    h->obs_dim = h->kind == KIND_DEFAULT ? 82 : (h->kind == KIND_RAPTOR ? 22 : 26);
    h->sdim = state_dim(h->H);
    h->n = config->n_envs;
    const int slots = config->n_state_slots < 2 ? 2 : config->n_state_slots;
    auto bail = [&](const std::string& m, int code){ create_error() = m; b200l2f_destroy(h); return code; };
#define CUC(call) do{ cudaError_t e_ = (call); if(e_ != cudaSuccess) return bail(std::string(#call) + ": " + cudaGetErrorString(e_), B200L2F_ERR_CUDA); }while(0)
    CUC(cudaSetDevice(config->device));
    if(config->stream){ h->stream = (cudaStream_t)config->stream; h->own_stream = false; }
    else{ CUC(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)); h->own_stream = true; }
    const size_t n = (size_t)h->n;
    CUC(cudaMalloc(&h->d_params, sizeof(float) * B200L2F_PARAMS_DIM * n));
    CUC(cudaMalloc(&h->d_env_row, sizeof(float) * B200L2F_PARAMS_DIM));
    for(int s = 0; s < slots; s++){
        float* p = nullptr;
        CUC(cudaMalloc(&p, sizeof(float) * h->sdim * n));
        h->d_state.push_back(p);
        CUC(cudaMemsetAsync(p, 0, sizeof(float) * h->sdim * n, h->stream));
    }
    CUC(cudaMalloc(&h->d_rng, sizeof(uint64_t) * n));
    CUC(cudaMalloc(&h->d_flags, sizeof(int) * 4));
    CUC(cudaMemsetAsync(h->d_flags, 0, sizeof(int) * 4, h->stream));
    CUC(cudaMalloc(&h->d_episode_step, sizeof(int) * n));
    CUC(cudaMalloc(&h->d_episode_return, sizeof(float) * n));
    CUC(cudaMalloc(&h->d_truncated, sizeof(uint8_t) * n));
#undef CUC
    *out = h;
    int rc = b200l2f_initialize_environment(h);
    if(rc == B200L2F_OK) rc = b200l2f_initialize_rng(h, 0, 0);
    if(rc == B200L2F_OK) rc = b200l2f_initial_parameters(h);
    if(rc == B200L2F_OK) rc = b200l2f_collect_reset(h);
    if(rc != B200L2F_OK){ create_error() = h->err; b200l2f_destroy(h); *out = nullptr; return rc; }
    return B200L2F_OK;
}

int b200l2f_destroy(b200l2f_handle* h){
    if(!h) return B200L2F_OK;
    cudaSetDevice(h->cfg.device);
    if(h->stream) cudaStreamSynchronize(h->stream);
    cudaFree(h->d_params); cudaFree(h->d_env_row);
    for(float* p : h->d_state) cudaFree(p);
    cudaFree(h->d_tc_image); cudaFree(h->d_ts_image); cudaFree(h->d_mlp_tc_image); cudaFree(h->d_sched); cudaFree(h->d_acc_ret); cudaFree(h->d_acc_len);
    cudaFree(h->d_rng); cudaFree(h->d_flags); cudaFree(h->d_blob); cudaFree(h->d_hidden); cudaFree(h->d_gru_step);
    cudaFree(h->d_episode_step); cudaFree(h->d_episode_return); cudaFree(h->d_truncated);
    cudaFree(h->d_critic_blob); cudaFree(h->d_critic_tc_image); cudaFree(h->d_colstats);
    cudaFree(h->d_teacher_images); cudaFree(h->d_teacher_blobs); cudaFree(h->d_teacher_offsets);
    cudaFree(h->d_dg_states); cudaFree(h->d_dg_term); cudaFree(h->d_dg_eplen); cudaFree(h->d_dg_offsets); cudaFree(h->d_dg_returns);
    cudaFree(h->d_last_returns); cudaFree(h->d_last_eplen); cudaFree(h->d_last_done); cudaFree(h->d_nonfinite); cudaFree(h->d_status);
    if(h->d_stage) cudaFree(h->d_stage);
    if(h->h_pinned) cudaFreeHost(h->h_pinned);
    if(h->pinned_read) cudaEventDestroy(h->pinned_read);
    if(h->features_ready) cudaEventDestroy(h->features_ready);
    if(h->h_features) cudaFreeHost(h->h_features);
    if(h->h_row0_pinned) cudaFreeHost(h->h_row0_pinned);
    {
        auto& x = h->xfer;
        if(x.h2d){ cudaStreamSynchronize(x.h2d); cudaStreamDestroy(x.h2d); }
        if(x.d2h){ cudaStreamSynchronize(x.d2h); cudaStreamDestroy(x.d2h); }
        cudaFree(x.up_params); cudaFree(x.up_state); cudaFree(x.dl_state);
        for(cudaEvent_t e : {x.params_ready, x.params_free, x.state_ready, x.state_free, x.dl_ready, x.dl_free, x.main_mark}) if(e) cudaEventDestroy(e);
        for(auto& p : x.pending) if(!p.releases_dl_staging) cudaEventDestroy(p.after);
        for(cudaEvent_t e : x.event_pool) cudaEventDestroy(e);
    }
    if(h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return B200L2F_OK;
}
int b200l2f_synchronize(b200l2f_handle* h){ CU(cudaStreamSynchronize(h->stream)); return B200L2F_OK; }
void* b200l2f_stream(b200l2f_handle* h){ return (void*)h->stream; }
const char* b200l2f_last_kernel(const b200l2f_handle* h){ return h ? h->last_kernel : ""; }
int b200l2f_state_dim(const b200l2f_handle* h){ return h->sdim; }
int b200l2f_observation_dim(const b200l2f_handle* h){ return h->obs_dim; }
int b200l2f_action_history_length(const b200l2f_handle* h){ return h->H; }
int b200l2f_n_envs(const b200l2f_handle* h){ return h->n; }
int64_t b200l2f_kernel_launches(const b200l2f_handle* h){ return h->launches; }

// ---- RNG ------------------------------------------------------------------------------------------------------
int b200l2f_initialize_rng(b200l2f_handle* h, uint64_t seed, int32_t warmup){
    CU(cudaSetDevice(h->cfg.device));
    k_init_rng<<<grid_for(h->n, 256), 256, 0, h->stream>>>(h->d_rng, h->n, seed, (uint64_t)h->cfg.first_env_id, warmup);
    LAUNCH_CHECK();
    return B200L2F_OK;
}
int b200l2f_get_rng(b200l2f_handle* h, uint64_t* states, int memspace){
    CU(cudaSetDevice(h->cfg.device));
    if(memspace == B200L2F_DEVICE){ CU(cudaMemcpyAsync(states, h->d_rng, sizeof(uint64_t) * h->n, cudaMemcpyDeviceToDevice, h->stream)); return B200L2F_OK; }
    return download(h, states, h->d_rng, sizeof(uint64_t) * h->n, memspace);
}
int b200l2f_set_rng(b200l2f_handle* h, const uint64_t* states, int memspace){
    CU(cudaSetDevice(h->cfg.device));
    const void* dev; int rc;
    if((rc = upload(h, states, sizeof(uint64_t) * h->n, memspace, &dev))) return rc;
    CU(cudaMemcpyAsync(h->d_rng, dev, sizeof(uint64_t) * h->n, cudaMemcpyDeviceToDevice, h->stream));
    return B200L2F_OK;
}

// ---- environment / parameters -----------------------------------------------------------------------------------
int b200l2f_initialize_environment(b200l2f_handle* h){
    CU(cudaSetDevice(h->cfg.device));
    nominal_parameters(h->cfg.spec, h->h_env_row);
    return b200l2f_set_environment_parameters(h, h->h_env_row);
}
int b200l2f_get_environment_parameters(b200l2f_handle* h, float* row145){
    std::memcpy(row145, h->h_env_row, sizeof(float) * B200L2F_PARAMS_DIM);
    return B200L2F_OK;
}
int b200l2f_set_environment_parameters(b200l2f_handle* h, const float* row145){
    CU(cudaSetDevice(h->cfg.device));
    if(row145 != h->h_env_row) std::memcpy(h->h_env_row, row145, sizeof(float) * B200L2F_PARAMS_DIM);
    CU(cudaStreamSynchronize(h->stream));   // h_env_row is pageable: make the copy synchronous w.r.t. earlier kernels that read d_env_row
    CU(cudaMemcpy(h->d_env_row, h->h_env_row, sizeof(float) * B200L2F_PARAMS_DIM, cudaMemcpyHostToDevice));
    h->params_follow_env_row = false;
    return B200L2F_OK;
}
int b200l2f_initial_parameters(b200l2f_handle* h){
    CU(cudaSetDevice(h->cfg.device));
    k_fill_params<<<grid_for(h->n, 256), 256, 0, h->stream>>>(h->d_params, h->d_env_row, h->n);
    LAUNCH_CHECK();
    h->features_dirty = true; h->params_version++;
    h->params_follow_env_row = true;
    return B200L2F_OK;
}
int b200l2f_sample_initial_parameters(b200l2f_handle* h){
    CU(cudaSetDevice(h->cfg.device));
    if(!h->dr){
        // the reference asserts that all DR ranges are zero when the options are disabled (10_sample_initial_parameters.h:74,92,122,132,166,181,196-199)
        for(int i = P_DR_T2W_MIN; i <= P_DR_DIST_FORCE_MAX; i++){
            if(i == P_DR_ORI_OFFSET || i == P_DR_T2W_MAX) continue;
            if(h->h_env_row[i] != 0.0f) return fail(h, B200L2F_ERR_STATE, "L2f: domain randomization ranges must be 0 when the spec's DR options are disabled");
        }
        k_sample_params<false><<<grid_for(h->n, 256), 256, 0, h->stream>>>(h->d_params, h->d_env_row, h->d_rng, h->n, h->d_flags);
        LAUNCH_CHECK();
    }
    else{
        CU(cudaMemsetAsync(h->d_flags, 0, sizeof(int), h->stream));
        k_sample_params<true><<<grid_for(h->n, 256), 256, 0, h->stream>>>(h->d_params, h->d_env_row, h->d_rng, h->n, h->d_flags);
        LAUNCH_CHECK();
        int flag = 0;
        CU(cudaMemcpyAsync(&flag, h->d_flags, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        if(flag) return fail(h, B200L2F_ERR_STATE, "L2f: invalid domain randomization ranges (see the reference's assert_exit conditions in 10_sample_initial_parameters.h:68-199)");
    }
    h->features_dirty = true; h->params_version++;
    h->params_follow_env_row = true;
    return B200L2F_OK;
}
int b200l2f_get_parameters(b200l2f_handle* h, float* rows, int memspace){
    CU(cudaSetDevice(h->cfg.device));
    const size_t bytes = sizeof(float) * B200L2F_PARAMS_DIM * (size_t)h->n;
    void* dev; int rc;
    if((rc = result_buffer(h, rows, bytes, memspace, &dev))) return rc;
    if((rc = transpose(h, h->d_params, (float*)dev, B200L2F_PARAMS_DIM, h->n))) return rc;
    return download(h, rows, dev, bytes, memspace);
}
int b200l2f_set_parameters(b200l2f_handle* h, const float* rows, int memspace){
    CU(cudaSetDevice(h->cfg.device));
    const void* dev; int rc;
    if((rc = upload(h, rows, sizeof(float) * B200L2F_PARAMS_DIM * (size_t)h->n, memspace, &dev))) return rc;
    h->features_dirty = true; h->params_version++;
    h->params_follow_env_row = false;   // the caller's rows may differ from the nominal row anywhere: the collection kernels must read the columns
    if((rc = transpose(h, (const float*)dev, h->d_params, h->n, B200L2F_PARAMS_DIM))) return rc;
    return enqueue_features(h);
}

// ---- state ------------------------------------------------------------------------------------------------------
int b200l2f_initial_state(b200l2f_handle* h, int slot){
    CU(cudaSetDevice(h->cfg.device));
    int rc; if((rc = check_slot(h, slot))) return rc;
    return dispatch_spec(h, [&](auto spec){
        using Spec = decltype(spec);
        k_init_state<Spec, false><<<grid_for(h->n, BLOCK), BLOCK, 0, h->stream>>>(h->d_params, h->d_state[slot], h->d_rng, h->n);
        LAUNCH_CHECK();
        return (int)B200L2F_OK;
    });
}
int b200l2f_sample_initial_state(b200l2f_handle* h, int slot){
    CU(cudaSetDevice(h->cfg.device));
    int rc; if((rc = check_slot(h, slot))) return rc;
    return dispatch_spec(h, [&](auto spec){
        using Spec = decltype(spec);
        k_init_state<Spec, true><<<grid_for(h->n, BLOCK), BLOCK, 0, h->stream>>>(h->d_params, h->d_state[slot], h->d_rng, h->n);
        LAUNCH_CHECK();
        return (int)B200L2F_OK;
    });
}
int b200l2f_get_state(b200l2f_handle* h, int slot, float* rows, int memspace){
    CU(cudaSetDevice(h->cfg.device));
    int rc; if((rc = check_slot(h, slot))) return rc;
    const size_t bytes = sizeof(float) * h->sdim * (size_t)h->n;
    void* dev;
    if((rc = result_buffer(h, rows, bytes, memspace, &dev))) return rc;
    if((rc = transpose(h, h->d_state[slot], (float*)dev, h->sdim, h->n))) return rc;
    return download(h, rows, dev, bytes, memspace);
}
int b200l2f_set_state(b200l2f_handle* h, int slot, const float* rows, int memspace){
    CU(cudaSetDevice(h->cfg.device));
    int rc; if((rc = check_slot(h, slot))) return rc;
    const void* dev;
    if((rc = upload(h, rows, sizeof(float) * h->sdim * (size_t)h->n, memspace, &dev))) return rc;
    return transpose(h, (const float*)dev, h->d_state[slot], h->n, h->sdim);
}
// ---- asynchronous transfers -----------------------------------------------------------------------------------------
// The synchronous set_* / get_* above serialise copy and compute on one stream (end to end 0.77 of the kernel rate at configs[1], 0.82 scaling
// efficiency on 8 GPUs behind one PCIe root: round-1 verdict).  The *_async twins move the bytes on the handle's own copy streams into / out of
// staging buffers and leave only the device-side transpose on the main stream, in call order:
//   upload    h2d stream : wait(staging free) -> H2D -> record(ready)        main stream : wait(ready) -> transpose into the live buffer -> record(free)
//   download  main stream: wait(staging free) -> transpose -> record(ready)  d2h stream  : wait(ready) -> D2H -> record(free)
// so the upload for rollout k+1 and the download of rollout k-1 run while the kernel of rollout k executes.  Host buffers must be page-locked
// (pageable memory cannot be copied asynchronously; it takes the synchronous path) and must stay untouched until b200l2f_transfers_synchronize.
namespace {
int xfer_init(b200l2f_handle* h){
    auto& x = h->xfer;
    if(x.h2d) return B200L2F_OK;
    CU(cudaStreamCreateWithFlags(&x.h2d, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&x.d2h, cudaStreamNonBlocking));
    for(cudaEvent_t* e : {&x.params_ready, &x.params_free, &x.state_ready, &x.state_free, &x.dl_ready, &x.dl_free, &x.main_mark}) CU(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    return B200L2F_OK;
}
}  // namespace
int b200l2f_set_parameters_async(b200l2f_handle* h, const float* rows_pinned){
    CU(cudaSetDevice(h->cfg.device));
    if(!rows_pinned) return fail(h, B200L2F_ERR_ARGUMENT, "set_parameters_async: null argument");
    if(!is_pinned_host(rows_pinned)) return b200l2f_set_parameters(h, rows_pinned, B200L2F_HOST);
    int rc; if((rc = xfer_init(h))) return rc;
    auto& x = h->xfer;
    const size_t bytes = sizeof(float) * B200L2F_PARAMS_DIM * (size_t)h->n;
    if(!x.up_params) CU(cudaMalloc(&x.up_params, bytes));
    if(x.params_used) CU(cudaStreamWaitEvent(x.h2d, x.params_free, 0));
    CU(cudaMemcpyAsync(x.up_params, rows_pinned, bytes, cudaMemcpyHostToDevice, x.h2d));
    CU(cudaEventRecord(x.params_ready, x.h2d));
    CU(cudaStreamWaitEvent(h->stream, x.params_ready, 0));
    if((rc = transpose(h, x.up_params, h->d_params, h->n, B200L2F_PARAMS_DIM))) return rc;
    CU(cudaEventRecord(x.params_free, h->stream));
    x.params_used = true;
    h->features_dirty = true; h->params_version++;
    h->params_follow_env_row = false;
    return enqueue_features(h);
}
int b200l2f_set_state_async(b200l2f_handle* h, int slot, const float* rows_pinned){
    CU(cudaSetDevice(h->cfg.device));
    int rc; if((rc = check_slot(h, slot))) return rc;
    if(!rows_pinned) return fail(h, B200L2F_ERR_ARGUMENT, "set_state_async: null argument");
    if(!is_pinned_host(rows_pinned)) return b200l2f_set_state(h, slot, rows_pinned, B200L2F_HOST);
    if((rc = xfer_init(h))) return rc;
    auto& x = h->xfer;
    const size_t bytes = sizeof(float) * h->sdim * (size_t)h->n;
    if(!x.up_state) CU(cudaMalloc(&x.up_state, bytes));
    if(x.state_used) CU(cudaStreamWaitEvent(x.h2d, x.state_free, 0));
    CU(cudaMemcpyAsync(x.up_state, rows_pinned, bytes, cudaMemcpyHostToDevice, x.h2d));
    CU(cudaEventRecord(x.state_ready, x.h2d));
    CU(cudaStreamWaitEvent(h->stream, x.state_ready, 0));
    if((rc = transpose(h, x.up_state, h->d_state[slot], h->n, h->sdim))) return rc;
    CU(cudaEventRecord(x.state_free, h->stream));
    x.state_used = true;
    return B200L2F_OK;
}
int b200l2f_get_state_async(b200l2f_handle* h, int slot, float* rows_pinned){
    CU(cudaSetDevice(h->cfg.device));
    int rc; if((rc = check_slot(h, slot))) return rc;
    if(!rows_pinned) return fail(h, B200L2F_ERR_ARGUMENT, "get_state_async: null argument");
    if(!is_pinned_host(rows_pinned)) return b200l2f_get_state(h, slot, rows_pinned, B200L2F_HOST);
    if((rc = xfer_init(h))) return rc;
    auto& x = h->xfer;
    const size_t bytes = sizeof(float) * h->sdim * (size_t)h->n;
    if(!x.dl_state) CU(cudaMalloc(&x.dl_state, bytes));
    if((rc = flush_pending_downloads(h))) return rc;            // an earlier download out of the same staging buffer must be on its way before the buffer is rewritten
    if(x.dl_used) CU(cudaStreamWaitEvent(h->stream, x.dl_free, 0));
    if((rc = transpose(h, h->d_state[slot], x.dl_state, h->sdim, h->n))) return rc;
    CU(cudaEventRecord(x.dl_ready, h->stream));
    x.pending.push_back({rows_pinned, x.dl_state, bytes, x.dl_ready, true});
    x.dl_used = true;
    return B200L2F_OK;
}
int b200l2f_copy_to_host_async(b200l2f_handle* h, void* dst_pinned, const void* src_device, size_t bytes){
    CU(cudaSetDevice(h->cfg.device));
    if(!dst_pinned || !src_device) return fail(h, B200L2F_ERR_ARGUMENT, "copy_to_host_async: null argument");
    int rc; if((rc = xfer_init(h))) return rc;
    auto& x = h->xfer;
    if(!is_pinned_host(dst_pinned)){   // pageable destination: ordinary stream-ordered copy, complete on return
        CU(cudaMemcpyAsync(dst_pinned, src_device, bytes, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        return B200L2F_OK;
    }
    cudaEvent_t ev;
    if(x.event_pool.empty()) CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    else{ ev = x.event_pool.back(); x.event_pool.pop_back(); }
    CU(cudaEventRecord(ev, h->stream));                     // everything enqueued on the main stream so far (the kernel that produced src_device)
    x.pending.push_back({dst_pinned, src_device, bytes, ev, false});
    return B200L2F_OK;
}
int b200l2f_transfers_synchronize(b200l2f_handle* h, int which){
    CU(cudaSetDevice(h->cfg.device));
    auto& x = h->xfer;
    int rc;
    if((which & 2) && (rc = flush_pending_downloads(h))) return rc;
    if((which & 1) && x.h2d) CU(cudaStreamSynchronize(x.h2d));
    if((which & 2) && x.d2h) CU(cudaStreamSynchronize(x.d2h));
    return B200L2F_OK;
}
int b200l2f_copy_state(b200l2f_handle* h, int dst_slot, int src_slot){
    CU(cudaSetDevice(h->cfg.device));
    int rc; if((rc = check_slot(h, dst_slot)) || (rc = check_slot(h, src_slot))) return rc;
    if(dst_slot == src_slot) return B200L2F_OK;
    CU(cudaMemcpyAsync(h->d_state[dst_slot], h->d_state[src_slot], sizeof(float) * h->sdim * (size_t)h->n, cudaMemcpyDeviceToDevice, h->stream));
    return B200L2F_OK;
}

// ---- observe / step / reward / terminated -------------------------------------------------------------------------
int b200l2f_observe(b200l2f_handle* h, int slot, float* observations, int ld, int memspace){
    CU(cudaSetDevice(h->cfg.device));
    int rc; if((rc = check_slot(h, slot))) return rc;
    if(ld < h->obs_dim) return fail(h, B200L2F_ERR_ARGUMENT, "observe: ld < OBSERVATION_DIM");
    const size_t bytes = sizeof(float) * (size_t)ld * h->n;
    void* dev;
    if((rc = result_buffer(h, observations, bytes, memspace, &dev))) return rc;
    if(memspace == B200L2F_HOST && ld != h->obs_dim) CU(cudaMemsetAsync(dev, 0, bytes, h->stream));
    rc = dispatch_spec(h, [&](auto spec){
        using Spec = decltype(spec);
        k_observe<Spec><<<grid_for(h->n, BLOCK), BLOCK, 0, h->stream>>>(h->d_params, h->d_state[slot], h->d_rng, (float*)dev, ld, h->n);
        LAUNCH_CHECK();
        return (int)B200L2F_OK;
    });
    if(rc) return rc;
    if(memspace == B200L2F_HOST && ld != h->obs_dim){
        // preserve the caller's padding columns: copy row by row
        if((rc = ensure_pinned(h, bytes))) return rc;
        CU(cudaMemcpyAsync(h->h_pinned, dev, bytes, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        for(int e = 0; e < h->n; e++) std::memcpy(observations + (size_t)e * ld, (float*)h->h_pinned + (size_t)e * ld, sizeof(float) * h->obs_dim);
        return B200L2F_OK;
    }
    return download(h, observations, dev, bytes, memspace);
}
int b200l2f_step(b200l2f_handle* h, int slot, const float* actions, int next_slot, float* dts, int memspace){
    CU(cudaSetDevice(h->cfg.device));
    int rc; if((rc = check_slot(h, slot)) || (rc = check_slot(h, next_slot))) return rc;
    const size_t act_bytes = sizeof(float) * 4 * (size_t)h->n, dts_bytes = sizeof(float) * (size_t)h->n;
    if(memspace == B200L2F_HOST){ if((rc = ensure_stage(h, act_bytes + dts_bytes))) return rc; }   // reserve before any pointer into the staging buffer is taken
    const void* d_act;
    if((rc = upload(h, actions, act_bytes, memspace, &d_act))) return rc;
    void* d_dts = nullptr;
    if(dts){ if((rc = result_buffer(h, dts, dts_bytes, memspace, &d_dts, act_bytes))) return rc; }
    rc = dispatch_spec(h, [&](auto spec){
        using Spec = decltype(spec);
        auto kern = k_step<Spec>;
        const size_t smem = sizeof(float) * P_DYN_DIM * BLOCK;
        kern<<<grid_for(h->n, BLOCK), BLOCK, smem, h->stream>>>(h->d_params, h->d_state[slot], (const float*)d_act, h->d_state[next_slot], h->d_rng, (float*)d_dts, h->n);
        LAUNCH_CHECK();
        return (int)B200L2F_OK;
    });
    if(rc) return rc;
    if(dts) return download(h, dts, d_dts, dts_bytes, memspace);
    return B200L2F_OK;
}
int b200l2f_reward(b200l2f_handle* h, int slot, const float* actions, int next_slot, float* rewards, int memspace){
    CU(cudaSetDevice(h->cfg.device));
    int rc; if((rc = check_slot(h, slot)) || (rc = check_slot(h, next_slot))) return rc;
    const size_t act_bytes = sizeof(float) * 4 * (size_t)h->n, out_bytes = sizeof(float) * (size_t)h->n;
    if(memspace == B200L2F_HOST){ if((rc = ensure_stage(h, act_bytes + out_bytes))) return rc; }
    const void* d_act;
    if((rc = upload(h, actions, act_bytes, memspace, &d_act))) return rc;
    void* d_out;
    if((rc = result_buffer(h, rewards, out_bytes, memspace, &d_out, act_bytes))) return rc;
    rc = dispatch_spec(h, [&](auto spec){
        using Spec = decltype(spec);
        k_reward<Spec><<<grid_for(h->n, BLOCK), BLOCK, 0, h->stream>>>(h->d_params, h->d_state[slot], (const float*)d_act, h->d_state[next_slot], (float*)d_out, h->n);
        LAUNCH_CHECK();
        return (int)B200L2F_OK;
    });
    if(rc) return rc;
    return download(h, rewards, d_out, out_bytes, memspace);
}
int b200l2f_terminated(b200l2f_handle* h, int slot, uint8_t* flags, int memspace){
    CU(cudaSetDevice(h->cfg.device));
    int rc; if((rc = check_slot(h, slot))) return rc;
    void* d_out;
    if((rc = result_buffer(h, flags, h->n, memspace, &d_out))) return rc;
    rc = dispatch_spec(h, [&](auto spec){
        using Spec = decltype(spec);
        k_terminated<Spec><<<grid_for(h->n, BLOCK), BLOCK, 0, h->stream>>>(h->d_params, h->d_state[slot], (uint8_t*)d_out, h->n);
        LAUNCH_CHECK();
        return (int)B200L2F_OK;
    });
    if(rc) return rc;
    return download(h, flags, d_out, h->n, memspace);
}

// ---- actor ------------------------------------------------------------------------------------------------------
static size_t policy_num_parameters(const b200l2f_policy_desc& d){
    const size_t in = d.input_dim, hd = d.hidden_dim, o = d.output_dim;
    if(d.arch == B200L2F_POLICY_RAPTOR_GRU) return hd * in + hd + 2 * (3 * hd * hd + 3 * hd) + hd + o * hd + o;
    return (d.standardize ? 2 * in : 0) + hd * in + hd + hd * hd + hd + o * hd + o + (d.head == B200L2F_HEAD_PPO_GAUSSIAN ? 4 : 0);
}
int b200l2f_policy_load(b200l2f_handle* h, const b200l2f_policy_desc* desc, const float* blob, size_t n_floats){
    CU(cudaSetDevice(h->cfg.device));
    if(!desc || !blob) return fail(h, B200L2F_ERR_ARGUMENT, "policy_load: null argument");
    if(policy_num_parameters(*desc) != n_floats) return fail(h, B200L2F_ERR_ARGUMENT, "policy_load: blob size does not match the descriptor");
    if(desc->arch == B200L2F_POLICY_RAPTOR_GRU){
        if(!(desc->input_dim == 22 && desc->hidden_dim == 16 && desc->output_dim == 4 && desc->head == B200L2F_HEAD_IDENTITY))
            return fail(h, B200L2F_ERR_UNSUPPORTED, "policy_load: the GRU actor is instantiated for Dense(22->16) -> GRU(16) -> Dense(16->4)");
        if(h->obs_dim < 22) return fail(h, B200L2F_ERR_ARGUMENT, "policy_load: observation narrower than the actor input");
    }
    else if(desc->arch == B200L2F_POLICY_MLP){
        if(desc->hidden_dim != MLP_HD) return fail(h, B200L2F_ERR_UNSUPPORTED, "policy_load: MLP actors are instantiated for hidden_dim 64");
        if(desc->input_dim != h->obs_dim) return fail(h, B200L2F_ERR_ARGUMENT, "policy_load: an MLP actor consumes the full observation of the spec (input_dim must equal OBSERVATION_DIM)");
        const bool ok = (desc->head == B200L2F_HEAD_SQUASH_EVAL && desc->output_dim == 8) || (desc->head != B200L2F_HEAD_SQUASH_EVAL && desc->output_dim == 4);
        if(!ok) return fail(h, B200L2F_ERR_ARGUMENT, "policy_load: output_dim must be 8 for SQUASH_EVAL ([mean, log_std]) and 4 otherwise");
    }
    else return fail(h, B200L2F_ERR_ARGUMENT, "policy_load: unknown architecture");
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(h->d_blob); cudaFree(h->d_hidden); cudaFree(h->d_gru_step);
    h->d_blob = nullptr; h->d_hidden = nullptr; h->d_gru_step = nullptr;
    CU(cudaMalloc(&h->d_blob, sizeof(float) * n_floats));
    CU(cudaMemcpy(h->d_blob, blob, sizeof(float) * n_floats, cudaMemcpyHostToDevice));
    CU(cudaMalloc(&h->d_hidden, sizeof(float) * desc->hidden_dim * (size_t)h->n));
    CU(cudaMalloc(&h->d_gru_step, sizeof(int) * (size_t)h->n));
    if(desc->arch == B200L2F_POLICY_MLP){
        h->pol = *desc; h->blob_floats = n_floats; h->policy_loaded = true;
        cudaFree(h->d_mlp_tc_image); h->d_mlp_tc_image = nullptr;
        {   // tensor-core operand image (H = 1 specs: K1 <= 32; DEFAULT spec, PPO actor: K1 = 88)
            int brc = build_mlp_tc_image(h, desc, blob);
            if(brc) return brc;
        }
        CU(cudaMemsetAsync(h->d_hidden, 0, sizeof(float) * desc->hidden_dim * (size_t)h->n, h->stream));
        CU(cudaMemsetAsync(h->d_gru_step, 0, sizeof(int) * (size_t)h->n, h->stream));
        return B200L2F_OK;
    }
    h->h_image.assign(RaptorImage<22, 16, 4>::SIZE, 0.0f);
    build_raptor_image_host<22, 16, 4>(h->h_image.data(), blob);
    {   // tensor-core operand image: tf32 hi/lo planes in the canonical K-major core-matrix order
        std::vector<float> tcimg(TcImage::SIZE);
        build_tc_image_host(tcimg.data(), blob);
        cudaFree(h->d_tc_image); h->d_tc_image = nullptr;
        CU(cudaMalloc(&h->d_tc_image, TcImage::BYTES));
        CU(cudaMemcpy(h->d_tc_image, tcimg.data(), TcImage::BYTES, cudaMemcpyHostToDevice));
        build_tc_image_host(tcimg.data(), blob, true);   // TMEM-A kernel: gate rows pre-multiplied by the exponent scale of their activation
        cudaFree(h->d_ts_image); h->d_ts_image = nullptr;
        CU(cudaMalloc(&h->d_ts_image, TcImage::BYTES));
        CU(cudaMemcpy(h->d_ts_image, tcimg.data(), TcImage::BYTES, cudaMemcpyHostToDevice));
    }
    {   // tuning knob, measured on B200 (profiles/r01_summary.md): weights staged in shared memory (broadcast LDS.128) are ~3% faster than
        // riding in the launch's constant bank (LDCU + uniform-register FFMA operands), so shared memory is the default
        const char* env = std::getenv("B200L2F_WEIGHTS");
        h->weights_in_constant_bank = env && std::string(env) == "const";
        const char* code = std::getenv("B200L2F_CODE");      // "rolled": compact-code variant (rolled actor k-loops + RK4 stage loop), fits the instruction cache
        h->rolled = code && std::string(code) == "rolled"; }   // measured equal within 2 % on B200 (profiles/): the unrolled body stays the default
    h->pol = *desc; h->blob_floats = n_floats; h->policy_loaded = true;
    if(h->pol.gru_sequence_length <= 0) h->pol.gru_sequence_length = 500;
    return b200l2f_policy_reset(h, nullptr, B200L2F_HOST);
}
static const float* raptor_h0(const b200l2f_handle* h){
    const int in = h->pol.input_dim, hd = h->pol.hidden_dim;
    return h->d_blob + (hd * in + hd + 2 * (3 * hd * hd + 3 * hd));
}
int b200l2f_policy_reset(b200l2f_handle* h, const uint8_t* mask, int memspace){
    CU(cudaSetDevice(h->cfg.device));
    if(!h->policy_loaded) return fail(h, B200L2F_ERR_STATE, "policy_reset: no policy loaded");
    if(h->pol.arch == B200L2F_POLICY_MLP) return B200L2F_OK;   // stateless actor
    const void* d_mask = nullptr; int rc;
    if(mask){ if((rc = upload(h, mask, h->n, memspace, &d_mask))) return rc; }
    k_policy_reset<16><<<grid_for(h->n, 256), 256, 0, h->stream>>>(h->d_hidden, h->d_gru_step, raptor_h0(h), (const uint8_t*)d_mask, h->n);
    LAUNCH_CHECK();
    return B200L2F_OK;
}
int b200l2f_policy_evaluate_step(b200l2f_handle* h, const float* observations, int ld, float* actions, int no_auto_reset, int memspace){
    CU(cudaSetDevice(h->cfg.device));
    if(!h->policy_loaded) return fail(h, B200L2F_ERR_STATE, "policy_evaluate_step: no policy loaded");
    if(ld < h->pol.input_dim) return fail(h, B200L2F_ERR_ARGUMENT, "policy_evaluate_step: ld < input_dim");
    const size_t obs_bytes = sizeof(float) * (size_t)ld * h->n, act_bytes = sizeof(float) * 4 * (size_t)h->n;
    const void* d_obs; void* d_act; int rc;
    if(memspace == B200L2F_HOST){ if((rc = ensure_stage(h, obs_bytes + act_bytes))) return rc; }
    if((rc = upload(h, observations, obs_bytes, memspace, &d_obs))) return rc;
    if((rc = result_buffer(h, actions, act_bytes, memspace, &d_act, obs_bytes))) return rc;
    if(h->pol.arch == B200L2F_POLICY_MLP){
        if((rc = launch_mlp_step(h, (const float*)d_obs, ld, (float*)d_act))) return rc;
        return download(h, actions, d_act, act_bytes, memspace);
    }
    const size_t smem = sizeof(float) * RaptorImage<22, 16, 4>::SIZE;
    if(h->cfg.flags & B200L2F_FLAG_ACCURATE_MATH)
        k_raptor_step<22, 16, 4, false><<<grid_for(h->n, BLOCK), BLOCK, smem, h->stream>>>(h->d_blob, (const float*)d_obs, ld, h->d_hidden, h->d_gru_step, h->pol.gru_sequence_length, no_auto_reset, (float*)d_act, h->n);
    else
        k_raptor_step<22, 16, 4, true><<<grid_for(h->n, BLOCK), BLOCK, smem, h->stream>>>(h->d_blob, (const float*)d_obs, ld, h->d_hidden, h->d_gru_step, h->pol.gru_sequence_length, no_auto_reset, (float*)d_act, h->n);
    LAUNCH_CHECK();
    return download(h, actions, d_act, act_bytes, memspace);
}
int b200l2f_policy_get_hidden(b200l2f_handle* h, float* hidden, int32_t* gru_step, int memspace){
    CU(cudaSetDevice(h->cfg.device));
    if(!h->policy_loaded) return fail(h, B200L2F_ERR_STATE, "policy_get_hidden: no policy loaded");
    int rc;
    if(hidden){
        const size_t bytes = sizeof(float) * h->pol.hidden_dim * (size_t)h->n;
        void* dev;
        if((rc = result_buffer(h, hidden, bytes, memspace, &dev))) return rc;
        if((rc = transpose(h, h->d_hidden, (float*)dev, h->pol.hidden_dim, h->n))) return rc;
        if((rc = download(h, hidden, dev, bytes, memspace))) return rc;
    }
    if(gru_step){
        if(memspace == B200L2F_DEVICE) CU(cudaMemcpyAsync(gru_step, h->d_gru_step, sizeof(int) * h->n, cudaMemcpyDeviceToDevice, h->stream));
        else if((rc = download(h, gru_step, h->d_gru_step, sizeof(int) * h->n, memspace))) return rc;
    }
    return B200L2F_OK;
}
int b200l2f_policy_set_hidden(b200l2f_handle* h, const float* hidden, const int32_t* gru_step, int memspace){
    CU(cudaSetDevice(h->cfg.device));
    if(!h->policy_loaded) return fail(h, B200L2F_ERR_STATE, "policy_set_hidden: no policy loaded");
    int rc; const void* dev;
    if(hidden){
        if((rc = upload(h, hidden, sizeof(float) * h->pol.hidden_dim * (size_t)h->n, memspace, &dev))) return rc;
        if((rc = transpose(h, (const float*)dev, h->d_hidden, h->n, h->pol.hidden_dim))) return rc;
    }
    if(gru_step){
        if((rc = upload(h, gru_step, sizeof(int) * h->n, memspace, &dev))) return rc;
        CU(cudaMemcpyAsync(h->d_gru_step, dev, sizeof(int) * h->n, cudaMemcpyDeviceToDevice, h->stream));
    }
    return B200L2F_OK;
}

// ---- fused rollout ------------------------------------------------------------------------------------------------
int b200l2f_rollout(b200l2f_handle* h, int32_t n_steps, int32_t no_auto_reset, const b200l2f_rollout_out* out){
    CU(cudaSetDevice(h->cfg.device));
    if(!h->policy_loaded) return fail(h, B200L2F_ERR_STATE, "rollout: no policy loaded");
    if(n_steps < 0) return fail(h, B200L2F_ERR_ARGUMENT, "rollout: n_steps < 0");
    if(h->pol.arch == B200L2F_POLICY_RAPTOR_GRU && h->kind == KIND_TEACHER) return fail(h, B200L2F_ERR_UNSUPPORTED, "rollout: the Raptor actor is paired with the DEFAULT and RAPTOR specs");
    if(h->pol.arch == B200L2F_POLICY_MLP && h->pol.head == B200L2F_HEAD_PPO_GAUSSIAN) return fail(h, B200L2F_ERR_UNSUPPORTED, "rollout: stochastic PPO actors are driven by b200l2f_collect");
    int rc;
    if((rc = refresh_features(h))) return rc;
    RolloutArgs a{};
    a.params = h->d_params; a.state = h->d_state[0]; a.rng = h->d_rng; a.hidden = h->d_hidden; a.gru_step = h->d_gru_step; a.blob = h->d_blob;
    a.n = h->n; a.T = n_steps; a.no_auto_reset = no_auto_reset; a.seq_len = h->pol.gru_sequence_length; a.state_stride = 1;
    std::memcpy(a.row0, h->row0, sizeof(a.row0));
    // outputs: device pointers are used directly, host pointers get a slice of the staging buffer
    const size_t n = (size_t)h->n, T = (size_t)n_steps;
    struct Slice { void** kernel_ptr; void* user; size_t bytes; size_t offset; };
    std::vector<Slice> slices;
    size_t total = 0;
    const int ms = out ? out->memspace : B200L2F_DEVICE;
    auto add = [&](void** kp, void* user, size_t bytes){
        if(!user) return;
        const size_t off = (total + 255) / 256 * 256;
        slices.push_back({kp, user, bytes, off});
        total = off + bytes;
    };
    if(out){
        if(out->states){
            if(out->state_stride <= 0) return fail(h, B200L2F_ERR_ARGUMENT, "rollout: states requested with state_stride <= 0");
            a.state_stride = out->state_stride;
            add((void**)&a.out_states, out->states, sizeof(float) * (T / out->state_stride + 1) * n * h->sdim);
        }
        add((void**)&a.out_obs, out->observations, sizeof(float) * T * n * (h->pol.arch == B200L2F_POLICY_MLP ? h->obs_dim : 22));
        add((void**)&a.out_actions, out->actions, sizeof(float) * T * n * 4);
        add((void**)&a.out_rewards, out->rewards, sizeof(float) * T * n);
        add((void**)&a.out_term, out->terminated, T * n);
        add((void**)&a.out_returns, out->returns, sizeof(float) * n);
        add((void**)&a.out_eplen, out->episode_length, sizeof(int) * n);
    }
    if(ms == B200L2F_HOST && total){ if((rc = ensure_stage(h, total))) return rc; }
    for(auto& s : slices) *s.kernel_ptr = (ms == B200L2F_HOST) ? (void*)((char*)h->d_stage + s.offset) : s.user;
    // the per-environment episode summary is always produced (12 bytes per environment and launch): b200l2f_last_status reduces it
    if(!h->d_last_returns){
        CU(cudaMalloc(&h->d_last_returns, sizeof(float) * n)); CU(cudaMalloc(&h->d_last_eplen, sizeof(int) * n)); CU(cudaMalloc(&h->d_last_done, n));
    }
    if(!a.out_returns) a.out_returns = h->d_last_returns;
    if(!a.out_eplen) a.out_eplen = h->d_last_eplen;
    a.out_done = h->d_last_done;
    const bool noise = (h->features & 1) != 0;
    const bool fast = !(h->cfg.flags & B200L2F_FLAG_ACCURATE_MATH);
    const bool constw = h->weights_in_constant_bank;
    const bool tensor_cores = h->pol.gemm == B200L2F_GEMM_TCGEN05_3XTF32;
    // every vehicle thrusts along body z with diagonal inertia (true for all reference vehicles; B200L2F_DYNAMICS=general forces the full matrices)
    const bool allow_axial = [](){ const char* e = std::getenv("B200L2F_DYNAMICS"); return !(e && std::string(e) == "general"); }();
    const bool axial = allow_axial && (h->features & 4) == 0;
    const bool uniform = (h->features & 2) == 0;
    // observation / action noise: carried by the tcgen05 TMEM-A kernels (MUFU Box-Muller) when the MDP constants, which include the noise
    // standard deviations, are uniform across environments; otherwise by the CUDA-core kernels
    const bool noise_on_tc = !noise || uniform;
    if(h->pol.arch == B200L2F_POLICY_MLP){
        // tcgen05 path: H = 1 specs, default math flags
        if(tensor_cores && fast && noise_on_tc && h->d_mlp_tc_image && h->kind != KIND_DEFAULT) rc = launch_mlp_ts(h, a, uniform, axial, noise);
        else rc = launch_mlp_fp32(h, a);
    }
    else if(tensor_cores && !(noise && !(fast && uniform))){
        // A operand in TMEM ("TS" MMAs, 63 KB smem + 128 TMEM columns per CTA -> 3 CTAs/SM) is the default; B200L2F_A=smem selects the
        // shared-memory-A variant (2 CTAs/SM, no noise variant).  Measured: 9.4e9 vs 6.7e9 env-steps/s at 1M envs (profiles/r01_exp8_*).
        static const bool a_in_tmem = [](){ const char* e = std::getenv("B200L2F_A"); return !(e && std::string(e) == "smem"); }();
        static const bool g1_tc = [](){ const char* e = std::getenv("B200L2F_G1"); return !(e && std::string(e) == "cuda"); }();   // tuning knob, default: dense 1 on tcgen05 too (+3.5 % measured)
        // two environments per thread on the packed fp32 pipe (rollout_x2.cuh; foundation-policy spec, default math, uniform MDP constants, axial vehicles,
        // no noise): 27 % fewer instructions per environment step, but 255 registers leave two warps per scheduler and it measures 0.8x of the
        // one-environment kernel (profiles/r02_exp1_two_envs_per_thread.md) -- opt-in with B200L2F_X2=1
        const bool x2 = [](){ const char* e = std::getenv("B200L2F_X2"); return e && e[0] == '1'; }();
        if(a_in_tmem && fast && x2 && uniform && axial && !noise && h->kind == KIND_RAPTOR) rc = launch_raptor_x2(h, a);
        else if(a_in_tmem && fast) rc = launch_raptor_ts(h, a, uniform, axial, noise);
        else if(noise) rc = launch_raptor_fp32(h, a, noise, fast, constw, h->rolled);
        else rc = launch_raptor_tc(h, a, fast, uniform, g1_tc);
    }
    else rc = launch_raptor_fp32(h, a, noise, fast, constw, h->rolled);
    if(rc) return rc;
    if((rc = enqueue_status(h, a.out_returns, a.out_eplen, a.out_done))) return rc;
    if((rc = flush_pending_downloads(h))) return rc;        // asynchronous downloads of the previous results start now, under this kernel
    if(ms == B200L2F_HOST && total){
        if((rc = ensure_pinned(h, total))) return rc;
        CU(cudaMemcpyAsync(h->h_pinned, h->d_stage, total, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        for(auto& s : slices) std::memcpy(s.user, (char*)h->h_pinned + s.offset, s.bytes);
    }
    return B200L2F_OK;
}

// ---- status of the last fused call -------------------------------------------------------------------------------------
int b200l2f_last_status(b200l2f_handle* h, b200l2f_status* out, uint8_t* nonfinite_flags, int memspace){
    CU(cudaSetDevice(h->cfg.device));
    if(!out) return fail(h, B200L2F_ERR_ARGUMENT, "last_status: null argument");
    if(!h->status_valid) return fail(h, B200L2F_ERR_STATE, "last_status: no rollout / collection has run on this handle yet");
    StatusPartial t;
    CU(cudaMemcpyAsync(&t, h->d_status, sizeof(t), cudaMemcpyDeviceToHost, h->stream));
    if(nonfinite_flags){
        if(memspace == B200L2F_DEVICE) CU(cudaMemcpyAsync(nonfinite_flags, h->d_nonfinite, (size_t)h->n, cudaMemcpyDeviceToDevice, h->stream));
        else CU(cudaMemcpyAsync(nonfinite_flags, h->d_nonfinite, (size_t)h->n, cudaMemcpyDeviceToHost, h->stream));
    }
    CU(cudaStreamSynchronize(h->stream));
    std::memset(out, 0, sizeof(*out));
    const double n = (double)h->n;
    out->n_envs = h->n; out->n_nonfinite = (int64_t)t.nonfinite; out->has_episodes = h->status_has_episodes ? 1 : 0;
    if(h->status_has_episodes){
        out->n_terminated = (int64_t)t.terminated;
        out->returns_mean = t.ret / n; out->episode_length_mean = t.len / n;
        const double vr = t.ret2 / n - out->returns_mean * out->returns_mean, vl = t.len2 / n - out->episode_length_mean * out->episode_length_mean;
        out->returns_std = std::sqrt(vr > 0 ? vr : 0); out->episode_length_std = std::sqrt(vl > 0 ? vl : 0);      // max(0, E[x^2] - E[x]^2) as operations_generic.h:210-212
        out->share_terminated = (double)t.terminated / n;
    }
    return B200L2F_OK;
}

// ---- PPO collection -------------------------------------------------------------------------------------------------
int b200l2f_collect_reset(b200l2f_handle* h){
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaMemsetAsync(h->d_episode_step, 0, sizeof(int) * h->n, h->stream));
    CU(cudaMemsetAsync(h->d_episode_return, 0, sizeof(float) * h->n, h->stream));
    CU(cudaMemsetAsync(h->d_truncated, 1, h->n, h->stream));
    return B200L2F_OK;
}
int b200l2f_collect(b200l2f_handle* h, int32_t n_steps, int32_t episode_step_limit, float* dataset, int memspace){
    CU(cudaSetDevice(h->cfg.device));
    if(!h->policy_loaded || h->pol.arch != B200L2F_POLICY_MLP || h->pol.head != B200L2F_HEAD_PPO_GAUSSIAN)
        return fail(h, B200L2F_ERR_STATE, "collect: load an MLP actor with the PPO_GAUSSIAN head first");
    if(n_steps < 0 || !dataset) return fail(h, B200L2F_ERR_ARGUMENT, "collect: bad arguments");
    const int D = h->obs_dim + 15;
    const size_t bytes = sizeof(float) * (size_t)(n_steps + 1) * h->n * D;
    void* dev; int rc;
    if((rc = result_buffer(h, dataset, bytes, memspace, &dev))) return rc;
    if(memspace == B200L2F_HOST) CU(cudaMemsetAsync(dev, 0, bytes, h->stream));
    CU(cudaMemsetAsync(h->d_flags, 0, sizeof(int), h->stream));
    CollectArgs a{};
    a.params = h->d_params; a.env_row = h->d_env_row; a.state = h->d_state[0]; a.rng = h->d_rng; a.blob = h->d_blob; a.has_std = h->pol.standardize;
    a.episode_step = h->d_episode_step; a.episode_return = h->d_episode_return; a.truncated = h->d_truncated; a.dataset = (float*)dev;
    a.n = h->n; a.T = n_steps; a.step_limit = episode_step_limit; a.error_flag = h->d_flags;
    a.bulk_rows = ((uintptr_t)dev % 16 == 0 && h->n % 4 == 0 && !std::getenv("B200L2F_NO_BULK_ROWS")) ? 1 : 0;
    std::memcpy(a.row, h->h_env_row, sizeof(a.row));
    const bool follow = h->params_follow_env_row;
    // axial vehicles (see b200l2f_rollout): decided from the nominal row when the columns follow it (domain randomisation keeps the property)
    const bool allow_axial = [](){ const char* e = std::getenv("B200L2F_DYNAMICS"); return !(e && std::string(e) == "general"); }();
    bool row_axial = allow_axial;
    for(int r = 0; r < 4; r++) if(a.row[P_THRUST_DIR + 3 * r] != 0.0f || a.row[P_THRUST_DIR + 3 * r + 1] != 0.0f || a.row[P_THRUST_DIR + 3 * r + 2] != 1.0f) row_axial = false;
    for(int i = 0; i < 9; i++) if(i % 4 != 0 && (a.row[P_J + i] != 0.0f || a.row[P_JINV + i] != 0.0f)) row_axial = false;
    const bool tensor_cores = h->pol.gemm == B200L2F_GEMM_TCGEN05_3XTF32 && h->d_mlp_tc_image && !(h->cfg.flags & B200L2F_FLAG_ACCURATE_MATH);
    rc = tensor_cores ? launch_collect_ts(h, a, follow, row_axial) : launch_collect_fp32(h, a);
    if(rc) return rc;
    if((rc = enqueue_status(h, nullptr, nullptr, nullptr))) return rc;
    if(!follow){ h->features_dirty = true; h->params_version++; }   // resets rewrite parameter columns (when they follow the row, the variant-selecting features cannot change)
    if((rc = download(h, dataset, dev, bytes, memspace))) return rc;
    if(h->dr){
        int flag = 0;
        CU(cudaMemcpyAsync(&flag, h->d_flags, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        if(flag) return fail(h, B200L2F_ERR_STATE, "L2f: invalid domain randomization ranges (reset inside collect)");
    }
    return B200L2F_OK;
}

}  // extern "C"
