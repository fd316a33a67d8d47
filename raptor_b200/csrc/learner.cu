// raptor_b200/csrc/learner.cu -- C ABI of the learner feed (include/b200_l2f.h, "PPO learner feed"): critic values, generalized advantage
// estimation and the running observation normalizer on the dataset b200l2f_collect wrote, without the data leaving the GPU.
#include "launch.h"
#include "learner.cuh"
#include <cmath>

using namespace b200l2f;

namespace {

size_t dataset_bytes(const b200l2f_handle* h, int T){ return sizeof(float) * (size_t)(T + 1) * h->n * (h->obs_dim + 15); }

// host datasets are staged whole: upload -> kernels -> download
int stage_dataset(b200l2f_handle* h, float* dataset, int T, int memspace, float** dev){
    if(memspace == B200L2F_DEVICE){ *dev = dataset; return B200L2F_OK; }
    const void* d; int rc;
    if((rc = upload(h, dataset, dataset_bytes(h, T), memspace, &d))) return rc;
    *dev = (float*)d;
    return B200L2F_OK;
}
int check_feed(b200l2f_handle* h, int T, const float* dataset, const char* who){
    if(T < 1 || !dataset) return fail(h, B200L2F_ERR_ARGUMENT, std::string(who) + ": bad arguments");
    return B200L2F_OK;
}

template <int IN>
int launch_values(b200l2f_handle* h, FeedArgs a, bool gae){
    // the tcgen05 kernel is instantiated for the H = 1 specs (the observation is one K <= 32 operand); the DEFAULT spec's 82-wide rows run on CUDA cores
    constexpr bool HAS_TS = IN <= 32;
    const bool tensor_cores = HAS_TS && h->critic_gemm == B200L2F_GEMM_TCGEN05_3XTF32 && h->d_critic_tc_image && !(h->cfg.flags & B200L2F_FLAG_ACCURATE_MATH);
    int sms = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->cfg.device));
    if constexpr(HAS_TS) if(tensor_cores){
        using SM = FeedSmem<IN>;
        auto go = [&](auto kern) -> int {
            CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM::TOTAL));
            CU(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
            if(!h->d_sched){ CU(cudaMalloc(&h->d_sched, sizeof(int) * 64)); h->sched_ints = 64; }
            CU(cudaMemsetAsync(h->d_sched, 0, sizeof(int), h->stream));
            a.sched = h->d_sched;
            const int n_tiles = grid_for(a.n, BLOCK);
            const int grid = n_tiles < 2 * sms ? n_tiles : 2 * sms;     // ~100 KB smem + 256 TMEM columns per CTA -> 2 CTAs/SM, persistent tile loop
            kern<<<grid, BLOCK, SM::TOTAL, h->stream>>>(a, h->d_critic_tc_image);
            LAUNCH_CHECK();
            return (int)B200L2F_OK;
        };
        return gae ? go(k_values_ts<IN, true>) : go(k_values_ts<IN, false>);
    }
    auto kern = k_values<IN>;
    const size_t smem = sizeof(float) * (MlpImg<IN, 4>::SIZE + (size_t)(IN > MLP_HD ? IN : MLP_HD) * BLOCK);
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const size_t rows = (size_t)(a.T + 1) * a.n;
    const size_t want = (rows + BLOCK - 1) / BLOCK;
    const int grid = (int)(want < (size_t)8 * sms ? want : (size_t)8 * sms);
    kern<<<grid, BLOCK, smem, h->stream>>>(a);
    LAUNCH_CHECK();
    if(gae){
        k_gae<<<grid_for(a.n, 128), 128, 0, h->stream>>>(a, IN);
        LAUNCH_CHECK();
    }
    return B200L2F_OK;
}

int values_impl(b200l2f_handle* h, int T, float gamma, float lambda, int ignore_termination, float* dataset, int memspace, bool gae, const char* who){
    CU(cudaSetDevice(h->cfg.device));
    int rc;
    if((rc = check_feed(h, T, dataset, who))) return rc;
    if(!h->critic_loaded) return fail(h, B200L2F_ERR_STATE, std::string(who) + ": no critic loaded (b200l2f_critic_load)");
    float* dev;
    if((rc = stage_dataset(h, dataset, T, memspace, &dev))) return rc;
    FeedArgs a{};
    a.dataset = dev; a.n = h->n; a.T = T; a.gamma = gamma; a.lambda = lambda; a.ignore_termination = ignore_termination;
    a.blob = h->d_critic_blob; a.has_std = h->critic_std;
    rc = h->obs_dim == 22 ? launch_values<22>(h, a, gae) : h->obs_dim == 26 ? launch_values<26>(h, a, gae) : launch_values<82>(h, a, gae);
    if(rc) return rc;
    return download(h, dataset, dev, dataset_bytes(h, T), memspace);
}

}  // namespace

extern "C" {

int b200l2f_critic_load(b200l2f_handle* h, const b200l2f_policy_desc* desc, const float* blob, size_t n_floats){
    CU(cudaSetDevice(h->cfg.device));
    if(!desc || !blob) return fail(h, B200L2F_ERR_ARGUMENT, "critic_load: null argument");
    if(desc->arch != B200L2F_POLICY_MLP || desc->hidden_dim != MLP_HD || desc->output_dim != 1 || desc->head != B200L2F_HEAD_IDENTITY || desc->input_dim != h->obs_dim)
        return fail(h, B200L2F_ERR_ARGUMENT, "critic_load: the critic is [standardize ->] Dense(OBS, 64) -> Dense(64, 64) -> Dense(64, 1), head IDENTITY");
    const int in = desc->input_dim, hd = MLP_HD;
    const size_t want = (size_t)(desc->standardize ? 2 * in : 0) + hd * in + hd + hd * hd + hd + hd + 1;
    if(n_floats != want) return fail(h, B200L2F_ERR_ARGUMENT, "critic_load: blob size does not match the descriptor");
    // pad the single output row to the 4-row instantiations the actor kernels already use (rows 1..3 are zero)
    std::vector<float> padded(want + 3 * hd + 3, 0.0f);
    const size_t head = want - hd - 1;                    // everything before W3
    std::memcpy(padded.data(), blob, sizeof(float) * head);
    std::memcpy(padded.data() + head, blob + head, sizeof(float) * hd);               // W3 row 0
    padded[head + 4 * hd] = blob[head + hd];                                           // b3[0]
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(h->d_critic_blob); cudaFree(h->d_critic_tc_image); h->d_critic_blob = nullptr; h->d_critic_tc_image = nullptr;
    CU(cudaMalloc(&h->d_critic_blob, sizeof(float) * padded.size()));
    CU(cudaMemcpy(h->d_critic_blob, padded.data(), sizeof(float) * padded.size(), cudaMemcpyHostToDevice));
    auto build = [&](auto in_c) -> int {
        constexpr int IN = decltype(in_c)::value;
        std::vector<float> img(MlpTcImage<IN, 4>::SIZE);
        build_mlp_tc_image_host<IN, 4>(img.data(), padded.data(), desc->standardize != 0, false);
        CU(cudaMalloc(&h->d_critic_tc_image, MlpTcImage<IN, 4>::BYTES));
        CU(cudaMemcpy(h->d_critic_tc_image, img.data(), MlpTcImage<IN, 4>::BYTES, cudaMemcpyHostToDevice));
        return (int)B200L2F_OK;
    };
    int rc = in == 22 ? build(std::integral_constant<int, 22>{}) : in == 26 ? build(std::integral_constant<int, 26>{}) : (int)B200L2F_OK;   // OBS 82 (DEFAULT spec): CUDA-core kernel, no operand image
    if(rc) return rc;
    h->critic_loaded = true; h->critic_std = desc->standardize; h->critic_gemm = desc->gemm;
    return B200L2F_OK;
}

int b200l2f_evaluate_values(b200l2f_handle* h, int32_t n_steps, float* dataset, int memspace){
    return values_impl(h, n_steps, 0.0f, 0.0f, 0, dataset, memspace, false, "evaluate_values");
}
int b200l2f_values_and_advantages(b200l2f_handle* h, int32_t n_steps, float gamma, float lambda, int ignore_termination, float* dataset, int memspace){
    return values_impl(h, n_steps, gamma, lambda, ignore_termination, dataset, memspace, true, "values_and_advantages");
}
int b200l2f_estimate_generalized_advantages(b200l2f_handle* h, int32_t n_steps, float gamma, float lambda, int ignore_termination, float* dataset, int memspace){
    CU(cudaSetDevice(h->cfg.device));
    int rc;
    if((rc = check_feed(h, n_steps, dataset, "estimate_generalized_advantages"))) return rc;
    float* dev;
    if((rc = stage_dataset(h, dataset, n_steps, memspace, &dev))) return rc;
    FeedArgs a{};
    a.dataset = dev; a.n = h->n; a.T = n_steps; a.gamma = gamma; a.lambda = lambda; a.ignore_termination = ignore_termination;
    k_gae<<<grid_for(a.n, 128), 128, 0, h->stream>>>(a, h->obs_dim);
    LAUNCH_CHECK();
    return download(h, dataset, dev, dataset_bytes(h, n_steps), memspace);
}

int b200l2f_normalizer_update(b200l2f_handle* h, int32_t n_steps, const float* dataset, int memspace, float* mean_io, float* std_io, int32_t* age_io){
    CU(cudaSetDevice(h->cfg.device));
    int rc;
    if((rc = check_feed(h, n_steps, dataset, "normalizer_update"))) return rc;
    if(!mean_io || !std_io || !age_io) return fail(h, B200L2F_ERR_ARGUMENT, "normalizer_update: null argument");
    const size_t rows = (size_t)n_steps * h->n;
    if(rows < 2) return fail(h, B200L2F_ERR_ARGUMENT, "normalizer_update: needs more than one row");
    float* dev;
    if((rc = stage_dataset(h, const_cast<float*>(dataset), n_steps, memspace, &dev))) return rc;
    const int obs = h->obs_dim, D = obs + 15;
    int sms = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->cfg.device));
    const size_t want = (rows + BLOCK - 1) / BLOCK;
    const int grid = (int)(want < (size_t)8 * sms ? want : (size_t)8 * sms);
    const size_t need = (size_t)grid * obs + 2 * obs;
    if(need > h->colstats_doubles){
        cudaFree(h->d_colstats); h->d_colstats = nullptr; h->colstats_doubles = 0;
        CU(cudaMalloc(&h->d_colstats, sizeof(double) * need));
        h->colstats_doubles = need;
    }
    double* partials = h->d_colstats; double* d_mean = partials + (size_t)grid * obs; double* d_ss = d_mean + obs;
    const size_t smem = sizeof(float) * (size_t)BLOCK * D;
    CU(cudaFuncSetAttribute(k_column_partials, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // 49.7 KB for the DEFAULT spec's 97-float rows
    k_column_partials<<<grid, BLOCK, smem, h->stream>>>(dev, rows, obs, nullptr, partials);
    LAUNCH_CHECK();
    k_column_finish<<<1, 128, 0, h->stream>>>(partials, grid, obs, 1.0 / (double)rows, d_mean);
    LAUNCH_CHECK();
    k_column_partials<<<grid, BLOCK, smem, h->stream>>>(dev, rows, obs, d_mean, partials);
    LAUNCH_CHECK();
    k_column_finish<<<1, 128, 0, h->stream>>>(partials, grid, obs, 1.0, d_ss);
    LAUNCH_CHECK();
    double host[2 * 128];
    CU(cudaMemcpyAsync(host, d_mean, sizeof(double) * 2 * obs, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    // rl::components::running_normalizer update (operations_generic.h:33-48): age++, mean += (data_mean - mean) / age, same for std
    *age_io += 1;
    for(int c = 0; c < obs; c++){
        const float data_mean = (float)host[c];
        const float acc = (float)host[obs + c];
        const float data_std = acc < 1e-6 ? 0.0f : std::sqrt((float)(host[obs + c] / (double)(rows - 1)));   // containers/matrix/operations_generic.h:685-690
        mean_io[c] = mean_io[c] + (data_mean - mean_io[c]) / (float)(*age_io);
        std_io[c] = std_io[c] + (data_std - std_io[c]) / (float)(*age_io);
    }
    return B200L2F_OK;
}

}  // extern "C"
