// raptor_b200/csrc/dagger.cuh -- the foundation-policy DAgger data path (src/foundation_policy/post_training/helper.h:43-123, gather_epoch =
// sample_trajectories + add_to_dataset) for ALL teachers in one pass:
//
//   student rollout        the fused Raptor rollout kernel (rollout_tc.cuh) with state snapshots and termination flags recorded step-major
//   k_exclusive_scan_i32   episode lengths -> first dataset row of every episode (the reference's running current_index, helper.h:55,81)
//   k_dagger_relabel<TC>   one CTA per teacher at a time (persistent, atomic teacher counter): the teacher's weight image arrives by ONE TMA
//                          bulk copy, then the teacher's rows (its episodes, up to and including the first terminated step) are processed in
//                          tiles of 128: thread = dataset row -> (episode, step) -> recorded state -> teacher observation (26) and student
//                          observation (22, position minus the teacher's steady-state offset) -> flags -> the teacher's MLP 26-64-64-8 on
//                          tcgen05 (3xTF32, mlp_tc.cuh) -> tanh(mean) = action target.  Time and episodes of one teacher form the M dimension
//                          of the GEMMs: "1000 teachers x their own parameters" is per-CTA weights + per-row parameters.
//                          TC = false: the same with the MLP on fp32 CUDA cores (mlp.cuh).
// The dataset part consumes no random numbers: it requires observation noise off (the reference's post-training setup, post_training/config.h:41-48).
#pragma once
#include "mlp_tc.cuh"

namespace b200l2f {

struct DaggerArgs {
    const float* params;            // [145][n]
    const float* states;            // [T + 1][n][STATE_DIM] snapshots before every step (rows 0 .. T-1 are used)
    const uint8_t* terminated;      // [T][n]
    const int* offsets;             // [n + 1] exclusive scan of the episode lengths
    const float* teacher_weights;   // TC: [n_teachers][MlpTcImage<26, 8>::SIZE]; else [n_teachers][blob]
    const float* position_offsets;  // [n_teachers][3]
    int n, T, n_teachers, episodes_per_teacher;
    float* input_student;           // [rows][22]
    float* output_target;           // [rows][4]
    uint8_t* truncated; uint8_t* reset;   // [rows]
    int* episode_start;             // [n]
    int* sched;
};

constexpr int DAGGER_TEACHER_BLOB = 64 * 26 + 64 + 64 * 64 + 64 + 8 * 64 + 8;

// out[i] = sum of in[0 .. i), out[n] = total; one CTA of 1024 threads
__global__ void __launch_bounds__(1024) k_exclusive_scan_i32(const int* __restrict__ in, int* __restrict__ out, int n){
    __shared__ int warp_sums[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (n + 1023) / 1024;
    const int begin = min(n, tid * per), end = min(n, begin + per);
    int local = 0;
    for(int i = begin; i < end; i++) local += in[i];
    int incl = local;
#pragma unroll
    for(int d = 1; d < 32; d <<= 1){ const int v = __shfl_up_sync(0xffffffffu, incl, d); if(lane >= d) incl += v; }
    if(lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if(warp == 0){
        int w = warp_sums[lane];
#pragma unroll
        for(int d = 1; d < 32; d <<= 1){ const int v = __shfl_up_sync(0xffffffffu, w, d); if(lane >= d) w += v; }
        warp_sums[lane] = w;
    }
    __syncthreads();
    int run = incl - local + (warp > 0 ? warp_sums[warp - 1] : 0);
    for(int i = begin; i < end; i++){ out[i] = run; run += in[i]; }
    if(tid == 1023) out[n] = run;
}

struct DaggerSmem {
    using I = MlpTcImage<26, 8>;
    static constexpr int B = 0;
    static constexpr int BAR = (I::BYTES + 127) / 128 * 128;
    static constexpr int TOTAL_TC = BAR + 64;
    static constexpr int TOTAL_FP32 = (MlpImg<26, 8>::SIZE + MLP_HD * BLOCK) * 4;
};

template <bool TC>
__global__ void __launch_bounds__(BLOCK, 2) k_dagger_relabel(const __grid_constant__ DaggerArgs a){
    using Student = SpecRaptor; using Teacher = SpecTeacher;
    constexpr int SD = Student::STATE_DIM;
    extern __shared__ __align__(1024) unsigned char smraw[];
    const int tid = threadIdx.x;
    TsCtx c{};
    uint64_t* bar_tma = nullptr; uint32_t tma_phase = 0;
    float* img = reinterpret_cast<float*>(smraw);                              // fp32 twin: k-major image
    float* scr = img + MlpImg<26, 8>::SIZE + tid;                              //            this thread's scratch column
    if constexpr(TC){
        c = mlp_ts_prologue_at<26, 8>(smraw, DaggerSmem::B, DaggerSmem::BAR, nullptr);   // barriers + TMEM, no image yet
        bar_tma = reinterpret_cast<uint64_t*>(smraw + DaggerSmem::BAR);
    }
    const size_t n = (size_t)a.n;
    __shared__ int s_item;
    for(;;){
        if(tid == 0) s_item = atomicAdd(a.sched, 1);
        __syncthreads();
        const int teacher = s_item;
        __syncthreads();                                                        // also: every thread is done with the previous teacher's image
        if(teacher >= a.n_teachers) break;
        if constexpr(TC){
            if(tid == 0){
                tc::fence_async_smem();
                tc::mbar_expect_tx(bar_tma, DaggerSmem::I::BYTES);
                tc::tma_load_1d(smraw + DaggerSmem::B, a.teacher_weights + (size_t)teacher * DaggerSmem::I::SIZE, DaggerSmem::I::BYTES, bar_tma);
            }
            tc::mbar_wait(bar_tma, tma_phase); tma_phase ^= 1;
        }
        else{
            stage_mlp_image<26, 8>(img, a.teacher_weights + (size_t)teacher * DAGGER_TEACHER_BLOB, false, false);
            __syncthreads();
        }
        const int e0 = teacher * a.episodes_per_teacher, e1 = min(a.n, e0 + a.episodes_per_teacher);
        const int row0 = a.offsets[e0], row1 = a.offsets[e1];
        const float off_x = a.position_offsets[teacher * 3], off_y = a.position_offsets[teacher * 3 + 1], off_z = a.position_offsets[teacher * 3 + 2];
        for(int base = row0; base < row1; base += BLOCK){
            const int R = base + tid;
            const bool active = R < row1;
            float obs_t[26];
#pragma unroll
            for(int i = 0; i < 26; i++) obs_t[i] = 0.0f;
            if(active){
                int lo = e0, hi = e1 - 1;                                       // episode with offsets[e] <= R < offsets[e + 1]
                while(lo < hi){ const int mid = (lo + hi + 1) >> 1; if(a.offsets[mid] <= R) lo = mid; else hi = mid - 1; }
                const int e = lo, step = R - a.offsets[e];
                const float* row = a.states + ((size_t)step * n + e) * SD;
                ParamsGlobal p{a.params + e, n};
                uint64_t rng_unused = 0;
                {
                    EnvState<Teacher> st;
                    load_state(st, row, 1);
                    observe_regs<Teacher, false>(st, p, rng_unused, obs_t);
                }
                float obs_s[22];
                {
                    EnvState<Student> st;
                    load_state(st, row, 1);
                    observe_regs<Student, false>(st, p, rng_unused, obs_s);
                }
                obs_s[0] -= off_x; obs_s[1] -= off_y; obs_s[2] -= off_z;        // helper.h:66-69
                float* in_row = a.input_student + (size_t)R * 22;
#pragma unroll
                for(int i = 0; i < 22; i++) in_row[i] = obs_s[i];
                const bool term = a.terminated[(size_t)step * n + e] != 0;
                a.truncated[R] = (term || step == a.T - 1) ? 1 : 0;             // helper.h:70
                a.reset[R] = 1;                                                 // helper.h:52,72-75: the flag is never cleared
                if(step == 0) a.episode_start[e] = R;                           // helper.h:55
            }
            float o[8];
            if constexpr(TC) mlp_forward_ts<26, 8>(c, obs_t, o);
            else{
#pragma unroll
                for(int i = 0; i < 26; i++) scr[i * BLOCK] = obs_t[i];
                mlp_forward<26, 8>(img, scr, BLOCK, o);
            }
            if(active) *reinterpret_cast<float4*>(a.output_target + (size_t)R * 4) = make_float4(tanhf(o[0]), tanhf(o[1]), tanhf(o[2]), tanhf(o[3]));
        }
    }
    if constexpr(TC) mlp_ts_epilogue(c);
}

}  // namespace b200l2f
