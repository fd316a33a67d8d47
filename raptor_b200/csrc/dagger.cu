// raptor_b200/csrc/dagger.cu -- C ABI of the foundation-policy DAgger data path (include/b200_l2f.h, "DAgger"): b200l2f_teachers_load and
// b200l2f_dagger_gather = the reference's gather_epoch (src/foundation_policy/post_training/helper.h:112-123) for all teachers at once.
#include "launch.h"
#include "dagger.cuh"

using namespace b200l2f;

extern "C" {

int b200l2f_teachers_load(b200l2f_handle* h, int32_t n_teachers, int32_t episodes_per_teacher, const float* blobs, const float* position_offsets, int32_t gemm){
    CU(cudaSetDevice(h->cfg.device));
    if(h->kind != KIND_RAPTOR) return fail(h, B200L2F_ERR_UNSUPPORTED, "teachers_load: the DAgger path runs on the RAPTOR specs (post-training environment)");
    if(n_teachers <= 0 || episodes_per_teacher <= 0 || !blobs) return fail(h, B200L2F_ERR_ARGUMENT, "teachers_load: bad arguments");
    if((int64_t)n_teachers * episodes_per_teacher != h->n) return fail(h, B200L2F_ERR_ARGUMENT, "teachers_load: n_teachers * episodes_per_teacher must equal n_envs (environment e runs under teacher e / episodes_per_teacher)");
    using I = MlpTcImage<26, 8>;
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(h->d_teacher_images); cudaFree(h->d_teacher_blobs); cudaFree(h->d_teacher_offsets);
    h->d_teacher_images = h->d_teacher_blobs = h->d_teacher_offsets = nullptr;
    std::vector<float> images((size_t)n_teachers * I::SIZE);
    for(int t = 0; t < n_teachers; t++) build_mlp_tc_image_host<26, 8>(images.data() + (size_t)t * I::SIZE, blobs + (size_t)t * DAGGER_TEACHER_BLOB, false, false);
    std::vector<float> off((size_t)n_teachers * 3, 0.0f);
    if(position_offsets) std::memcpy(off.data(), position_offsets, sizeof(float) * off.size());
    CU(cudaMalloc(&h->d_teacher_images, sizeof(float) * images.size()));
    CU(cudaMemcpy(h->d_teacher_images, images.data(), sizeof(float) * images.size(), cudaMemcpyHostToDevice));
    CU(cudaMalloc(&h->d_teacher_blobs, sizeof(float) * (size_t)n_teachers * DAGGER_TEACHER_BLOB));
    CU(cudaMemcpy(h->d_teacher_blobs, blobs, sizeof(float) * (size_t)n_teachers * DAGGER_TEACHER_BLOB, cudaMemcpyHostToDevice));
    CU(cudaMalloc(&h->d_teacher_offsets, sizeof(float) * off.size()));
    CU(cudaMemcpy(h->d_teacher_offsets, off.data(), sizeof(float) * off.size(), cudaMemcpyHostToDevice));
    h->n_teachers = n_teachers; h->episodes_per_teacher = episodes_per_teacher; h->teacher_gemm = gemm;
    return B200L2F_OK;
}

int b200l2f_dagger_gather(b200l2f_handle* h, int32_t n_steps, int32_t no_auto_reset, const b200l2f_dagger_out* out, int64_t* rows_added){
    CU(cudaSetDevice(h->cfg.device));
    if(!h->d_teacher_images) return fail(h, B200L2F_ERR_STATE, "dagger_gather: no teachers loaded (b200l2f_teachers_load)");
    if(!h->policy_loaded || h->pol.arch != B200L2F_POLICY_RAPTOR_GRU) return fail(h, B200L2F_ERR_STATE, "dagger_gather: load the student (Raptor GRU actor) first");
    if(n_steps < 1 || !out || !rows_added || !out->input_student || !out->output_target || !out->truncated || !out->reset || !out->episode_start)
        return fail(h, B200L2F_ERR_ARGUMENT, "dagger_gather: bad arguments");
    int rc;
    if((rc = refresh_features(h))) return rc;
    if(h->features & 1) return fail(h, B200L2F_ERR_UNSUPPORTED, "dagger_gather: the dataset observations are noise-free (post_training/config.h:41-48); set the observation / action noise to 0");
    const size_t n = (size_t)h->n, T = (size_t)n_steps;
    // dataset row offsets (exclusive scan of the episode lengths) and rows_added are 32-bit: the total, at most n * T rows, must fit
    if((int64_t)h->n * (int64_t)n_steps >= ((int64_t)1 << 31)) return fail(h, B200L2F_ERR_ARGUMENT, "dagger_gather: n_envs * n_steps must stay below 2^31 dataset rows");
    // ---- 1. sample_trajectories (helper.h:6-41): the student's closed-loop rollout, states and termination flags recorded step-major
    const size_t need_states = (T + 1) * n * h->sdim;
    if(need_states > h->dg_state_floats){
        cudaFree(h->d_dg_states); cudaFree(h->d_dg_term); h->d_dg_states = nullptr; h->d_dg_term = nullptr; h->dg_state_floats = 0;
        CU(cudaMalloc(&h->d_dg_states, sizeof(float) * need_states));
        CU(cudaMalloc(&h->d_dg_term, T * n));
        h->dg_state_floats = need_states;
    }
    if(!h->d_dg_eplen){
        CU(cudaMalloc(&h->d_dg_eplen, sizeof(int) * n));
        CU(cudaMalloc(&h->d_dg_offsets, sizeof(int) * (n + 1)));
        CU(cudaMalloc(&h->d_dg_returns, sizeof(float) * n));
    }
    b200l2f_rollout_out ro{};
    ro.memspace = B200L2F_DEVICE; ro.state_stride = 1; ro.states = h->d_dg_states; ro.terminated = h->d_dg_term;
    ro.returns = h->d_dg_returns; ro.episode_length = h->d_dg_eplen;
    if((rc = b200l2f_rollout(h, n_steps, no_auto_reset, &ro))) return rc;
    // ---- 2. first dataset row of every episode
    k_exclusive_scan_i32<<<1, 1024, 0, h->stream>>>(h->d_dg_eplen, h->d_dg_offsets, h->n);
    LAUNCH_CHECK();
    int rows = 0;
    CU(cudaMemcpyAsync(&rows, h->d_dg_offsets + n, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if((int64_t)rows > out->capacity_rows) return fail(h, B200L2F_ERR_ARGUMENT, "dagger_gather: Dataset size exceeded (helper.h:84)");
    // ---- 3. add_to_dataset (helper.h:43-110)
    const int ms = out->memspace;
    struct Slice { void** kernel_ptr; void* user; size_t bytes; size_t offset; };
    DaggerArgs a{};
    std::vector<Slice> slices; size_t total = 0;
    auto add = [&](void** kp, void* user, size_t bytes){
        const size_t off = (total + 255) / 256 * 256;
        slices.push_back({kp, user, bytes, off});
        total = off + bytes;
    };
    add((void**)&a.input_student, out->input_student, sizeof(float) * 22 * (size_t)rows);
    add((void**)&a.output_target, out->output_target, sizeof(float) * 4 * (size_t)rows);
    add((void**)&a.truncated, out->truncated, (size_t)rows);
    add((void**)&a.reset, out->reset, (size_t)rows);
    add((void**)&a.episode_start, out->episode_start, sizeof(int) * n);
    if(ms == B200L2F_HOST){ if((rc = ensure_stage(h, total))) return rc; }
    for(auto& s : slices) *s.kernel_ptr = (ms == B200L2F_HOST) ? (void*)((char*)h->d_stage + s.offset) : s.user;
    const bool tensor_cores = h->teacher_gemm == B200L2F_GEMM_TCGEN05_3XTF32 && !(h->cfg.flags & B200L2F_FLAG_ACCURATE_MATH);
    a.params = h->d_params; a.states = h->d_dg_states; a.terminated = h->d_dg_term; a.offsets = h->d_dg_offsets;
    a.teacher_weights = tensor_cores ? h->d_teacher_images : h->d_teacher_blobs; a.position_offsets = h->d_teacher_offsets;
    a.n = h->n; a.T = n_steps; a.n_teachers = h->n_teachers; a.episodes_per_teacher = h->episodes_per_teacher;
    if(!h->d_sched){ CU(cudaMalloc(&h->d_sched, sizeof(int) * 64)); h->sched_ints = 64; }
    CU(cudaMemsetAsync(h->d_sched, 0, sizeof(int), h->stream));
    a.sched = h->d_sched;
    int sms = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->cfg.device));
    const int grid = h->n_teachers < 2 * sms ? h->n_teachers : 2 * sms;
    if(tensor_cores){
        auto kern = k_dagger_relabel<true>;
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DaggerSmem::TOTAL_TC));
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        kern<<<grid, BLOCK, DaggerSmem::TOTAL_TC, h->stream>>>(a);
    }
    else{
        auto kern = k_dagger_relabel<false>;
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DaggerSmem::TOTAL_FP32));
        kern<<<grid, BLOCK, DaggerSmem::TOTAL_FP32, h->stream>>>(a);
    }
    LAUNCH_CHECK();
    if(ms == B200L2F_HOST && total){
        if((rc = ensure_pinned(h, total))) return rc;
        CU(cudaMemcpyAsync(h->h_pinned, h->d_stage, total, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        for(auto& s : slices) std::memcpy(s.user, (char*)h->h_pinned + s.offset, s.bytes);
    }
    // Result of sample_trajectories (rl/utils/evaluation/evaluation.h:48-62): per-episode returns and lengths
    if(out->returns){
        if(ms == B200L2F_DEVICE) CU(cudaMemcpyAsync(out->returns, h->d_dg_returns, sizeof(float) * n, cudaMemcpyDeviceToDevice, h->stream));
        else if((rc = download(h, out->returns, h->d_dg_returns, sizeof(float) * n, ms))) return rc;
    }
    if(out->episode_length){
        if(ms == B200L2F_DEVICE) CU(cudaMemcpyAsync(out->episode_length, h->d_dg_eplen, sizeof(int) * n, cudaMemcpyDeviceToDevice, h->stream));
        else if((rc = download(h, out->episode_length, h->d_dg_eplen, sizeof(int) * n, ms))) return rc;
    }
    *rows_added = rows;
    return B200L2F_OK;
}

}  // extern "C"
