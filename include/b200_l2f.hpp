// include/b200_l2f.hpp -- header-only C++17 shim in the rl-tools idiom over the C ABI (b200_l2f.h).
//
// rl-tools selects implementations by overloading free functions on the device tag (rl_tools/operations/cpu.h:15-17); this header adds a
// device tag `b200::devices::B200` and vector handle types so that host code written against
//   rl_tools::init / initial_parameters / sample_initial_parameters / initial_state / sample_initial_state / observe / step / reward /
//   terminated                                  (rl_tools/rl/environments/l2f/operations_generic.h:43-176)
//   rl_tools::reset / evaluate_step             (rl_tools/nn_models/sequential/operations_generic.h:63-66,321-325)
//   rl_tools::evaluate                          (rl_tools/rl/utils/evaluation/operations_generic.h:93-214)
// keeps its shape: same names, same argument order (device first), caller-owned objects, explicit malloc/free, no exceptions --
// errors terminate through `assert_exit` exactly like rl_tools::utils::assert_exit (rl_tools/utils/assert/operations_cpu.h).
// Everything forwards to libb200l2f.so; there is no CPU implementation behind it.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "b200_l2f.h"

namespace b200 {
namespace devices {
    struct B200 {            // device tag (one CUDA device ordinal)
        using index_t = size_t;
        int ordinal = 0;
    };
}
namespace utils {
    inline void assert_exit(const devices::B200&, bool condition, const char* message){
        if(!condition){ std::fprintf(stderr, "%s\n", message); std::exit(1); }
    }
}
namespace l2f {
    // compile-time environment specifications (the reference's Specification instantiations on the path)
    template <int T_SPEC_ID, int T_OBSERVATION_DIM, int T_ACTION_HISTORY_LENGTH>
    struct Specification {
        static constexpr int SPEC_ID = T_SPEC_ID;
        static constexpr int OBSERVATION_DIM = T_OBSERVATION_DIM;
        static constexpr int ACTION_HISTORY_LENGTH = T_ACTION_HISTORY_LENGTH;
        static constexpr int ACTION_DIM = 4;
        static constexpr int STATE_DIM = B200L2F_STATE_DIM(T_ACTION_HISTORY_LENGTH);
        static constexpr int EPISODE_STEP_LIMIT = 500;
    };
    using DefaultSpecification = Specification<B200L2F_SPEC_DEFAULT, 82, 16>;          // l2f::Specification<float, size_t>
    using DefaultDRSpecification = Specification<B200L2F_SPEC_DEFAULT_DR, 82, 16>;
    using RaptorSpecification = Specification<B200L2F_SPEC_RAPTOR, 22, 1>;              // foundation-policy post-training environment
    using RaptorDRSpecification = Specification<B200L2F_SPEC_RAPTOR_DR, 22, 1>;
    using TeacherSpecification = Specification<B200L2F_SPEC_TEACHER, 26, 1>;            // foundation-policy pre-training environment
    using TeacherDRSpecification = Specification<B200L2F_SPEC_TEACHER_DR, 26, 1>;

    namespace vector {
        template <typename T_SPEC, size_t T_N>
        struct Environment {
            using SPEC = T_SPEC;
            static constexpr size_t N_ENVIRONMENTS = T_N;
            static constexpr int OBSERVATION_DIM = SPEC::OBSERVATION_DIM;
            static constexpr int ACTION_DIM = SPEC::ACTION_DIM;
            b200l2f_handle* handle = nullptr;
            int64_t first_env_id = 0;     // global id of environment 0 (multi-GPU sharding)
            int next_slot = 0;
        };
        template <size_t T_N> struct Parameters { };    // live in the environment's device buffers; tokens keep the reference's call shapes
        template <size_t T_N> struct Rng { uint64_t seed = 0; };
        template <size_t T_N> struct State { int slot = -1; };
        // row-major host matrices the callers own (what rl_tools::Matrix<Specification<T, TI, N, COLS>> holds)
        template <size_t ROWS, size_t COLS> struct Matrix {
            std::vector<float> data = std::vector<float>(ROWS * COLS, 0.0f);
            float& operator()(size_t r, size_t c){ return data[r * COLS + c]; }
            const float& operator()(size_t r, size_t c) const { return data[r * COLS + c]; }
        };
    }
}

// ---- free functions, device first (rl-tools idiom) -------------------------------------------------------------------------------
namespace detail {
    template <typename ENV>
    inline void check(const devices::B200& device, ENV& env, int rc){
        if(rc != B200L2F_OK) utils::assert_exit(device, false, b200l2f_last_error(env.handle));
    }
    // rl_tools::init(device, rng, seed) only records the seed; the per-environment streams are created on the environment's GPU on first use
    template <typename ENV, typename RNG>
    inline void bind_rng(const devices::B200& device, ENV& env, RNG& rng){
        if(rng.seed != UINT64_MAX){ check(device, env, b200l2f_initialize_rng(env.handle, rng.seed, 0)); rng.seed = UINT64_MAX; }
    }
}
template <typename SPEC, size_t N>
void malloc(devices::B200& device, l2f::vector::Environment<SPEC, N>& env){
    b200l2f_config c{};
    c.struct_size = (int32_t)sizeof(b200l2f_config); c.spec = SPEC::SPEC_ID; c.n_envs = (int32_t)N; c.device = device.ordinal;
    c.first_env_id = env.first_env_id; c.n_state_slots = 4; c.flags = 0; c.stream = nullptr;
    const int rc = b200l2f_create(&c, &env.handle);
    if(rc != B200L2F_OK) utils::assert_exit(device, false, b200l2f_last_error(nullptr));
}
template <typename SPEC, size_t N>
void free(devices::B200&, l2f::vector::Environment<SPEC, N>& env){ b200l2f_destroy(env.handle); env.handle = nullptr; }
template <typename SPEC, size_t N>
void malloc(devices::B200& device, l2f::vector::Environment<SPEC, N>& env, l2f::vector::State<N>& state){
    utils::assert_exit(device, env.next_slot < 4, "b200::malloc: out of state slots");
    state.slot = env.next_slot++;
}
template <size_t N>
void init(devices::B200&, l2f::vector::Rng<N>& rng, uint64_t seed){ rng.seed = seed; }                                       // rl_tools::init(device, rng, seed)
template <typename SPEC, size_t N>
void init(devices::B200& device, l2f::vector::Environment<SPEC, N>& env){ detail::check(device, env, b200l2f_initialize_environment(env.handle)); }
template <typename SPEC, size_t N>
void initial_parameters(devices::B200& device, l2f::vector::Environment<SPEC, N>& env, l2f::vector::Parameters<N>&){ detail::check(device, env, b200l2f_initial_parameters(env.handle)); }
template <typename SPEC, size_t N>
void sample_initial_parameters(devices::B200& device, l2f::vector::Environment<SPEC, N>& env, l2f::vector::Parameters<N>&, l2f::vector::Rng<N>& rng){
    detail::bind_rng(device, env, rng);
    detail::check(device, env, b200l2f_sample_initial_parameters(env.handle));
}
template <typename SPEC, size_t N>
void initial_state(devices::B200& device, l2f::vector::Environment<SPEC, N>& env, l2f::vector::Parameters<N>&, l2f::vector::State<N>& state){
    detail::check(device, env, b200l2f_initial_state(env.handle, state.slot));
}
template <typename SPEC, size_t N>
void sample_initial_state(devices::B200& device, l2f::vector::Environment<SPEC, N>& env, l2f::vector::Parameters<N>&, l2f::vector::State<N>& state, l2f::vector::Rng<N>& rng){
    detail::bind_rng(device, env, rng);
    detail::check(device, env, b200l2f_sample_initial_state(env.handle, state.slot));
}
template <typename SPEC, size_t N, size_t COLS>
void observe(devices::B200& device, l2f::vector::Environment<SPEC, N>& env, l2f::vector::Parameters<N>&, const l2f::vector::State<N>& state,
             l2f::vector::Matrix<N, COLS>& observation, l2f::vector::Rng<N>& rng){
    static_assert(COLS >= (size_t)SPEC::OBSERVATION_DIM, "observation matrix narrower than OBSERVATION_DIM");
    detail::bind_rng(device, env, rng);
    detail::check(device, env, b200l2f_observe(env.handle, state.slot, observation.data.data(), (int)COLS, B200L2F_HOST));
}
template <typename SPEC, size_t N>
float step(devices::B200& device, l2f::vector::Environment<SPEC, N>& env, l2f::vector::Parameters<N>&, const l2f::vector::State<N>& state,
           const l2f::vector::Matrix<N, 4>& action, l2f::vector::State<N>& next_state, l2f::vector::Rng<N>& rng){
    detail::bind_rng(device, env, rng);
    std::vector<float> dts(N);
    detail::check(device, env, b200l2f_step(env.handle, state.slot, action.data.data(), next_state.slot, dts.data(), B200L2F_HOST));
    return dts[N - 1];
}
template <typename SPEC, size_t N>
void reward(devices::B200& device, l2f::vector::Environment<SPEC, N>& env, l2f::vector::Parameters<N>&, const l2f::vector::State<N>& state,
            const l2f::vector::Matrix<N, 4>& action, const l2f::vector::State<N>& next_state, l2f::vector::Matrix<N, 1>& rewards, l2f::vector::Rng<N>&){
    detail::check(device, env, b200l2f_reward(env.handle, state.slot, action.data.data(), next_state.slot, rewards.data.data(), B200L2F_HOST));
}
template <typename SPEC, size_t N>
void terminated(devices::B200& device, l2f::vector::Environment<SPEC, N>& env, l2f::vector::Parameters<N>&, const l2f::vector::State<N>& state,
                std::vector<uint8_t>& flags, l2f::vector::Rng<N>&){
    flags.resize(N);
    detail::check(device, env, b200l2f_terminated(env.handle, state.slot, flags.data(), B200L2F_HOST));
}
template <typename SPEC, size_t N>
void copy(devices::B200& device, l2f::vector::Environment<SPEC, N>& env, const l2f::vector::State<N>& source, l2f::vector::State<N>& target){   // state = next_state
    detail::check(device, env, b200l2f_copy_state(env.handle, target.slot, source.slot));
}
template <typename SPEC, size_t N>
void get(devices::B200& device, l2f::vector::Environment<SPEC, N>& env, const l2f::vector::State<N>& state, l2f::vector::Matrix<N, SPEC::STATE_DIM>& rows){
    detail::check(device, env, b200l2f_get_state(env.handle, state.slot, rows.data.data(), B200L2F_HOST));
}

// ---- actor (foundation_policy.Raptor) --------------------------------------------------------------------------------------------
namespace policy {
    struct Raptor { const float* blob = nullptr; size_t n_floats = 2084; bool tensor_cores = true; };   // Dense 22-16 / GRU 16 / Dense 16-4
}
template <typename SPEC, size_t N>
void malloc(devices::B200& device, l2f::vector::Environment<SPEC, N>& env, const policy::Raptor& p){
    b200l2f_policy_desc d{B200L2F_POLICY_RAPTOR_GRU, 22, 16, 4, 0, B200L2F_HEAD_IDENTITY, 500, p.tensor_cores ? B200L2F_GEMM_TCGEN05_3XTF32 : B200L2F_GEMM_FP32_CUDA_CORES};
    detail::check(device, env, b200l2f_policy_load(env.handle, &d, p.blob, p.n_floats));
}
template <typename SPEC, size_t N>
void reset(devices::B200& device, l2f::vector::Environment<SPEC, N>& env, const policy::Raptor&){ detail::check(device, env, b200l2f_policy_reset(env.handle, nullptr, B200L2F_HOST)); }
template <typename SPEC, size_t N, size_t COLS>
void evaluate_step(devices::B200& device, l2f::vector::Environment<SPEC, N>& env, const policy::Raptor&, const l2f::vector::Matrix<N, COLS>& observation,
                   l2f::vector::Matrix<N, 4>& action){
    static_assert(COLS >= 22, "the Raptor actor consumes the first 22 observation columns");
    detail::check(device, env, b200l2f_policy_evaluate_step(env.handle, observation.data.data(), (int)COLS, action.data.data(), 0, B200L2F_HOST));
}
// rl_tools::evaluate replacement: T closed-loop steps in ONE fused kernel launch; returns / episode lengths in the reference's semantics
template <size_t N>
struct EvaluationResult { std::vector<float> returns = std::vector<float>(N); std::vector<int32_t> episode_length = std::vector<int32_t>(N); };
template <typename SPEC, size_t N>
void evaluate(devices::B200& device, l2f::vector::Environment<SPEC, N>& env, const policy::Raptor& p, EvaluationResult<N>& result, int step_limit){
    reset(device, env, p);
    b200l2f_rollout_out out{};
    out.memspace = B200L2F_HOST; out.returns = result.returns.data(); out.episode_length = result.episode_length.data();
    detail::check(device, env, b200l2f_rollout(env.handle, step_limit, 0, &out));
}
}  // namespace b200
