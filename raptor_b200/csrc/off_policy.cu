// raptor_b200/csrc/off_policy.cu -- host side of b200l2f_off_policy_steps + instantiations of k_off_policy / k_off_policy_ts, and the accessors of
// the runner bookkeeping (episode_step / episode_return / truncated) shared with b200l2f_collect.
#include <vector>

#include "launch.h"
#include "offpolicy_tc.cuh"

namespace b200l2f {
namespace {

int launch_off_policy_fp32(b200l2f_handle* h, const OffPolicyArgs& a){
    auto go = [&](auto dr_c) -> int {
        using Spec = SpecCompactCode<SpecTeacher>;
        constexpr int IN = Spec::OBS_DIM;
        auto kern = k_off_policy<Spec, decltype(dr_c)::value>;
        const size_t smem = sizeof(float) * (MlpImg<IN, 8>::SIZE + (size_t)P_DYN_DIM * BLOCK + (size_t)(MLP_HD + IN) * BLOCK);
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid_for(a.c.n, BLOCK), BLOCK, smem, h->stream>>>(a);
        LAUNCH_CHECK();
        return (int)B200L2F_OK;
    };
    return h->dr ? go(std::true_type{}) : go(std::false_type{});
}

int launch_off_policy_ts(b200l2f_handle* h, const OffPolicyArgs& a, bool follow, bool row_axial){
    auto go2 = [&](auto dr_c, auto follow_c, auto axial_c) -> int {
        using Spec = SpecCompactCode<SpecTeacher>;
        using SM = OffPolicyTsSmem<Spec::OBS_DIM>;
        auto kern = k_off_policy_ts<Spec, decltype(dr_c)::value, decltype(follow_c)::value, decltype(axial_c)::value>;
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM::TOTAL));
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        int sms = 0;
        CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->cfg.device));
        if(!h->d_sched){ CU(cudaMalloc(&h->d_sched, sizeof(int) * 64)); h->sched_ints = 64; }
        CU(cudaMemsetAsync(h->d_sched, 0, sizeof(int), h->stream));
        const int n_tiles = grid_for(a.c.n, BLOCK);
        const int grid = n_tiles < 2 * sms ? n_tiles : 2 * sms;    // 256 TMEM columns per CTA -> 2 CTAs/SM, persistent tile loop
        kern<<<grid, BLOCK, SM::TOTAL, h->stream>>>(a, h->d_mlp_tc_image, h->d_sched);
        LAUNCH_CHECK();
        return (int)B200L2F_OK;
    };
    auto go = [&](auto dr_c) -> int {
        if(follow) return row_axial ? go2(dr_c, std::true_type{}, std::true_type{}) : go2(dr_c, std::true_type{}, std::false_type{});
        return go2(dr_c, std::false_type{}, std::false_type{});
    };
    return h->dr ? go(std::true_type{}) : go(std::false_type{});
}

}  // namespace
}  // namespace b200l2f

using namespace b200l2f;

namespace {
// temporary device copies of HOST-side rings / batches: freed on every exit path of the call that made them
struct DeviceTemps {
    std::vector<void*> ptrs;
    cudaError_t alloc(void** dev, size_t bytes){ cudaError_t e = cudaMalloc(dev, bytes); if(e == cudaSuccess) ptrs.push_back(*dev); return e; }
    ~DeviceTemps(){ for(void* p : ptrs) cudaFree(p); }
};
}  // namespace

extern "C" {

int b200l2f_runner_get_state(b200l2f_handle* h, int32_t* episode_step, float* episode_return, uint8_t* truncated, int memspace){
    if(!h) return fail(h, B200L2F_ERR_ARGUMENT, "runner_get_state: null handle");
    CU(cudaSetDevice(h->cfg.device));
    const cudaMemcpyKind k = memspace == B200L2F_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if(episode_step) CU(cudaMemcpyAsync(episode_step, h->d_episode_step, sizeof(int32_t) * h->n, k, h->stream));
    if(episode_return) CU(cudaMemcpyAsync(episode_return, h->d_episode_return, sizeof(float) * h->n, k, h->stream));
    if(truncated) CU(cudaMemcpyAsync(truncated, h->d_truncated, h->n, k, h->stream));
    if(memspace == B200L2F_HOST) CU(cudaStreamSynchronize(h->stream));
    return B200L2F_OK;
}
int b200l2f_runner_set_state(b200l2f_handle* h, const int32_t* episode_step, const float* episode_return, const uint8_t* truncated, int memspace){
    if(!h) return fail(h, B200L2F_ERR_ARGUMENT, "runner_set_state: null handle");
    CU(cudaSetDevice(h->cfg.device));
    const cudaMemcpyKind k = memspace == B200L2F_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if(episode_step) CU(cudaMemcpyAsync(h->d_episode_step, episode_step, sizeof(int32_t) * h->n, k, h->stream));
    if(episode_return) CU(cudaMemcpyAsync(h->d_episode_return, episode_return, sizeof(float) * h->n, k, h->stream));
    if(truncated) CU(cudaMemcpyAsync(h->d_truncated, truncated, h->n, k, h->stream));
    if(memspace == B200L2F_HOST) CU(cudaStreamSynchronize(h->stream));   // pageable sources must not be reused before the copy has read them
    return B200L2F_OK;
}

int b200l2f_off_policy_steps(b200l2f_handle* h, int32_t n_steps, int32_t episode_step_limit, int32_t sample_parameters, const b200l2f_replay_buffers* rb){
    if(!h) return fail(h, B200L2F_ERR_ARGUMENT, "off_policy_steps: null handle");
    CU(cudaSetDevice(h->cfg.device));
    if(!h->policy_loaded || h->pol.arch != B200L2F_POLICY_MLP || h->pol.output_dim != 8 || h->pol.standardize)
        return fail(h, B200L2F_ERR_STATE, "off_policy_steps: load the SAC actor first (MLP, output_dim 8 = [mean, log_std], no standardize layer)");
    if(h->kind != KIND_TEACHER) return fail(h, B200L2F_ERR_UNSUPPORTED, "off_policy_steps: instantiated for the pre-training environment (spec TEACHER / TEACHER_DR)");
    if(!rb || n_steps < 0 || rb->capacity < 1 || !rb->data || !rb->episode_start || !rb->position || !rb->full || !rb->current_episode_start)
        return fail(h, B200L2F_ERR_ARGUMENT, "off_policy_steps: bad arguments");
    const int D = 2 * h->obs_dim + 7;
    const size_t n = (size_t)h->n, rows = n * (size_t)rb->capacity;
    // host buffers: temporary device copies of the rings and their bookkeeping (tests / small runners); device buffers are used in place
    struct Part { void* user; size_t bytes; void* dev; };
    Part parts[5] = {{rb->data, sizeof(float) * rows * D, nullptr}, {rb->episode_start, sizeof(int32_t) * rows, nullptr}, {rb->position, sizeof(int32_t) * n, nullptr},
                     {rb->full, n, nullptr}, {rb->current_episode_start, sizeof(int32_t) * n, nullptr}};
    const bool host = rb->memspace == B200L2F_HOST;
    DeviceTemps temps;
    for(auto& p : parts){
        if(!host){ p.dev = p.user; continue; }
        cudaError_t e = temps.alloc(&p.dev, p.bytes);
        if(e == cudaSuccess) e = cudaMemcpyAsync(p.dev, p.user, p.bytes, cudaMemcpyHostToDevice, h->stream);
        if(e != cudaSuccess) return fail(h, B200L2F_ERR_CUDA, std::string("off_policy_steps: staging the replay buffers: ") + cudaGetErrorString(e));
    }
    if(host) CU(cudaStreamSynchronize(h->stream));   // pageable sources: the caller may reuse them as soon as the call returns
    CU(cudaMemsetAsync(h->d_flags, 0, sizeof(int), h->stream));
    OffPolicyArgs oa{};
    CollectArgs& a = oa.c;
    a.params = h->d_params; a.env_row = h->d_env_row; a.state = h->d_state[0]; a.rng = h->d_rng; a.blob = h->d_blob; a.has_std = 0;
    a.episode_step = h->d_episode_step; a.episode_return = h->d_episode_return; a.truncated = h->d_truncated; a.dataset = nullptr;
    a.n = h->n; a.T = n_steps; a.step_limit = episode_step_limit; a.error_flag = h->d_flags;
    std::memcpy(a.row, h->h_env_row, sizeof(a.row));
    oa.replay = (float*)parts[0].dev; oa.episode_start = (int*)parts[1].dev; oa.position = (int*)parts[2].dev; oa.full = (uint8_t*)parts[3].dev;
    oa.current_episode_start = (int*)parts[4].dev; oa.capacity = rb->capacity; oa.sample_parameters = sample_parameters ? 1 : 0;
    const bool follow = h->params_follow_env_row;
    const bool allow_axial = [](){ const char* e = std::getenv("B200L2F_DYNAMICS"); return !(e && std::string(e) == "general"); }();
    bool row_axial = allow_axial;
    for(int r = 0; r < 4; r++) if(a.row[P_THRUST_DIR + 3 * r] != 0.0f || a.row[P_THRUST_DIR + 3 * r + 1] != 0.0f || a.row[P_THRUST_DIR + 3 * r + 2] != 1.0f) row_axial = false;
    for(int i = 0; i < 9; i++) if(i % 4 != 0 && (a.row[P_J + i] != 0.0f || a.row[P_JINV + i] != 0.0f)) row_axial = false;
    const bool tensor_cores = h->pol.gemm == B200L2F_GEMM_TCGEN05_3XTF32 && h->d_mlp_tc_image && !(h->cfg.flags & B200L2F_FLAG_ACCURATE_MATH);
    int rc = tensor_cores ? launch_off_policy_ts(h, oa, follow, row_axial) : launch_off_policy_fp32(h, oa);
    if(rc) return rc;
    if(!follow && sample_parameters){ h->features_dirty = true; h->params_version++; }
    if(host){
        for(auto& p : parts) CU(cudaMemcpyAsync(p.user, p.dev, p.bytes, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    if(sample_parameters && h->dr){
        int flag = 0;
        CU(cudaMemcpyAsync(&flag, h->d_flags, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        if(flag) return fail(h, B200L2F_ERR_STATE, "L2f: invalid domain randomization ranges (reset inside off_policy_steps)");
    }
    return B200L2F_OK;
}

// bp == nullptr: SEQUENCE_LENGTH 1 with the MLP SAC parameters (k_gather_batch); otherwise the general walk (k_gather_batch_sequential)
static int gather_common(b200l2f_handle* h, const b200l2f_replay_buffers* rb, const b200l2f_batch_parameters* bp, int32_t max_episode_length, int32_t env_begin, int32_t env_count,
                         uint64_t* rng_states, const b200l2f_batch* out){
    if(!h) return fail(h, B200L2F_ERR_ARGUMENT, "gather_batch: null handle");
    CU(cudaSetDevice(h->cfg.device));
    if(!rb || !out || !rng_states || !rb->data || !rb->position || !rb->full || rb->capacity < 1 || out->batch_size < 0 || !out->observations_actions || !out->rewards || !out->terminated)
        return fail(h, B200L2F_ERR_ARGUMENT, "gather_batch: bad arguments");
    if(rb->memspace != out->memspace) return fail(h, B200L2F_ERR_ARGUMENT, "gather_batch: replay buffers and batch must live in the same memory space");
    if(env_begin < 0 || env_count < 1 || env_begin + env_count > h->n) return fail(h, B200L2F_ERR_ARGUMENT, "gather_batch: environment range outside the handle");
    const int B = out->batch_size, OBS = h->obs_dim, D = 2 * OBS + 7, W = OBS + 4;
    const int L = bp ? bp->sequence_length : 1, P = L + 1;
    if(bp){
        if(L < 1 || !rb->episode_start || !out->reset || !out->next_reset || !out->final_step_mask || !out->next_final_step_mask)
            return fail(h, B200L2F_ERR_ARGUMENT, "gather_batch_sequential: sequence_length >= 1, episode_start and all four masks are required");
        if(bp->always_sample_from_initial_state && rb->capacity < max_episode_length)
            return fail(h, B200L2F_ERR_ARGUMENT, "gather_batch_sequential: sampling from the initial state needs capacity >= max_episode_length");   // operations_generic.h:258
    }
    if(B == 0) return B200L2F_OK;
    const size_t n = (size_t)h->n;
    const bool host = rb->memspace == B200L2F_HOST;
    // host memory: temporary device copies (tests / small runners); device memory is used in place
    struct Part { const void* user_in; void* user_out; size_t bytes; void* dev; };
    Part parts[12] = {
        {rb->data, nullptr, sizeof(float) * n * rb->capacity * D, nullptr}, {rb->position, nullptr, sizeof(int32_t) * n, nullptr}, {rb->full, nullptr, n, nullptr},
        {rng_states, rng_states, sizeof(uint64_t) * B, nullptr},
        {nullptr, out->observations_actions, sizeof(float) * P * B * W, nullptr}, {nullptr, out->rewards, sizeof(float) * L * B, nullptr}, {nullptr, out->terminated, (size_t)L * B, nullptr},
        {nullptr, out->reset, (size_t)L * B, nullptr}, {nullptr, out->next_reset, (size_t)P * B, nullptr}, {nullptr, out->final_step_mask, (size_t)L * B, nullptr},
        {nullptr, out->next_final_step_mask, (size_t)P * B, nullptr},
        {bp ? rb->episode_start : nullptr, nullptr, bp ? sizeof(int32_t) * n * rb->capacity : 0, nullptr}};
    Part extra[2] = {{nullptr, out->env_index, sizeof(int32_t) * B, nullptr}, {nullptr, out->sample_index, sizeof(int32_t) * B, nullptr}};
    std::vector<Part*> all;
    for(auto& p : parts) all.push_back(&p);
    for(auto& p : extra) all.push_back(&p);
    DeviceTemps temps;
    for(auto* p : all){
        const void* any = p->user_in ? p->user_in : p->user_out;
        if(!any || !p->bytes) continue;
        if(!host){ p->dev = const_cast<void*>(any); continue; }
        cudaError_t e = temps.alloc(&p->dev, p->bytes);
        if(e == cudaSuccess && p->user_in) e = cudaMemcpyAsync(p->dev, p->user_in, p->bytes, cudaMemcpyHostToDevice, h->stream);
        else if(e == cudaSuccess) e = cudaMemsetAsync(p->dev, 0, p->bytes, h->stream);   // outputs: defined contents even when the call fails before writing them (initcheck)
        if(e != cudaSuccess) return fail(h, B200L2F_ERR_CUDA, std::string("gather_batch: staging: ") + cudaGetErrorString(e));
    }
    if(host) CU(cudaStreamSynchronize(h->stream));
    CU(cudaMemsetAsync(h->d_flags, 0, sizeof(int), h->stream));
    GatherArgs a{};
    a.replay = (const float*)parts[0].dev; a.position = (const int*)parts[1].dev; a.full = (const uint8_t*)parts[2].dev;
    a.obs_dim = OBS; a.capacity = rb->capacity; a.max_episode_length = max_episode_length; a.env_begin = env_begin; a.env_count = env_count; a.batch = B;
    a.rng = (uint64_t*)parts[3].dev; a.observations_actions = (float*)parts[4].dev; a.rewards = (float*)parts[5].dev; a.terminated = (uint8_t*)parts[6].dev;
    a.reset = (uint8_t*)parts[7].dev; a.next_reset = (uint8_t*)parts[8].dev; a.final_step_mask = (uint8_t*)parts[9].dev; a.next_final_step_mask = (uint8_t*)parts[10].dev;
    a.env_index = (int*)extra[0].dev; a.sample_index = (int*)extra[1].dev; a.error_flag = h->d_flags;
    if(bp){
        GatherSeqArgs q{};
        q.g = a; q.episode_start = (const int*)parts[11].dev; q.L = L; q.include_first = bp->include_first_step_in_targets; q.always_initial = bp->always_sample_from_initial_state;
        q.random_len = bp->random_seq_length; q.enable_nominal = bp->enable_nominal_sequence_length_probability; q.nominal_probability = bp->nominal_sequence_length_probability;
        k_gather_batch_sequential<<<grid_for(B * 32, 256), 256, 0, h->stream>>>(q);
    }
    else k_gather_batch<<<grid_for(B * 32, 256), 256, 0, h->stream>>>(a);
    h->launches++;
    cudaError_t le = cudaGetLastError();
    if(le != cudaSuccess) return fail(h, B200L2F_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(le));
    if(host){
        for(auto* p : all){
            if(!p->dev || !p->user_out) continue;
            CU(cudaMemcpyAsync(p->user_out, p->dev, p->bytes, cudaMemcpyDeviceToHost, h->stream));
        }
    }
    int flag = 0;
    CU(cudaMemcpyAsync(&flag, h->d_flags, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if(flag) return fail(h, B200L2F_ERR_STATE, "gather_batch: Replay buffer requires at least one element");
    return B200L2F_OK;
}

int b200l2f_gather_batch(b200l2f_handle* h, const b200l2f_replay_buffers* rb, int32_t max_episode_length, int32_t env_begin, int32_t env_count, uint64_t* rng_states,
                         const b200l2f_batch* out){
    return gather_common(h, rb, nullptr, max_episode_length, env_begin, env_count, rng_states, out);
}
int b200l2f_gather_batch_sequential(b200l2f_handle* h, const b200l2f_replay_buffers* rb, const b200l2f_batch_parameters* parameters, int32_t max_episode_length, int32_t env_begin,
                                    int32_t env_count, uint64_t* rng_states, const b200l2f_batch* out){
    if(!parameters) return fail(h, B200L2F_ERR_ARGUMENT, "gather_batch_sequential: null parameters");
    return gather_common(h, rb, parameters, max_episode_length, env_begin, env_count, rng_states, out);
}

}  // extern "C"
