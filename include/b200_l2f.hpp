// include/b200_l2f.hpp -- header-only C++17 shim in the rl-tools idiom over the C ABI (b200_l2f.h).
//
// rl-tools selects implementations by overloading free functions on the device tag (rl_tools/operations/cpu.h:15-17); this header adds a
// device tag `b200::devices::B200` and vector handle types so that host code written against
//   rl_tools::init / initial_parameters / sample_initial_parameters / initial_state / sample_initial_state / observe / step / reward /
//   terminated                                  (rl_tools/rl/environments/l2f/operations_generic.h:43-176)
//   rl_tools::reset / evaluate_step             (rl_tools/nn_models/sequential/operations_generic.h:63-66,321-325)
//   rl_tools::evaluate                          (rl_tools/rl/utils/evaluation/operations_generic.h:93-214)
//   rl_tools::collect / evaluate (critic) / estimate_generalized_advantages / update (running normalizer)
//                                               (rl/components/on_policy_runner/operations_generic.h:99-131, rl/algorithms/ppo/operations_generic.h:54-89,
//                                                rl/components/running_normalizer/operations_generic.h:27-49: the PPO loop step's data path)
//   rl_tools::step / gather_batch               (rl/components/off_policy_runner/operations_generic.h:215-238,240-434: SAC teacher collection)
//   gather_epoch                                (src/foundation_policy/post_training/helper.h:6-123: DAgger data path)
//   rl_tools::json / from_json, checkpoint load (rl/environments/l2f/operations_cpu.h:139-824; rl/loop/steps/checkpoint/operations_cpu.h:56-160)
// keeps its shape: same names, same argument order (device first), caller-owned objects, explicit malloc/free, no exceptions --
// errors terminate through `assert_exit` exactly like rl_tools::utils::assert_exit (rl_tools/utils/assert/operations_cpu.h).
// Everything forwards to libb200l2f.so; there is no CPU implementation behind it.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <string>
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

// the environment's nominal parameter row [B200L2F_PARAMS_DIM] (env.parameters of the reference: what initial_parameters copies and the DR sampler starts from)
template <typename SPEC, size_t N>
void get_environment_parameters(devices::B200& device, l2f::vector::Environment<SPEC, N>& env, float* row){ detail::check(device, env, b200l2f_get_environment_parameters(env.handle, row)); }
template <typename SPEC, size_t N>
void set_environment_parameters(devices::B200& device, l2f::vector::Environment<SPEC, N>& env, const float* row){ detail::check(device, env, b200l2f_set_environment_parameters(env.handle, row)); }

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
// ====================================================================================================================================
// Training-side callers: the reference's runner / learner-feed / checkpoint free functions over the same handle.
// ====================================================================================================================================
namespace policy {
    // [standardize ->] Dense(IN,64,ReLU) -> Dense(64,64,ReLU) -> Dense(64,OUT) (rl_tools/nn_models/mlp/network.h:15-51); blob in the order of b200_l2f.h:
    // [mean[IN] precision[IN]] W1 b1 W2 b2 W3 b3 [log_std[4]].  head: IDENTITY (critic, deterministic actors), SQUASH_EVAL (SAC actor, OUT = 8),
    // PPO_GAUSSIAN (mlp_unconditional_stddev, sampled in the runner's epilogue)
    struct MLP {
        std::vector<float> blob; int input_dim = 0, output_dim = 4; bool standardize = false; int head = B200L2F_HEAD_IDENTITY; bool tensor_cores = true;
        b200l2f_policy_desc desc() const { return b200l2f_policy_desc{B200L2F_POLICY_MLP, input_dim, 64, output_dim, standardize ? 1 : 0, head, 0, tensor_cores ? B200L2F_GEMM_TCGEN05_3XTF32 : B200L2F_GEMM_FP32_CUDA_CORES}; }
    };
    // an actor read from a checkpoint file (either format), ready for malloc(device, env, actor)
    struct Checkpoint { b200l2f_policy_desc desc{}; std::vector<float> blob; std::string name; };
}
template <typename SPEC, size_t N>
void malloc(devices::B200& device, l2f::vector::Environment<SPEC, N>& env, const policy::MLP& p){         // actor
    const b200l2f_policy_desc d = p.desc();
    detail::check(device, env, b200l2f_policy_load(env.handle, &d, p.blob.data(), p.blob.size()));
}
template <typename SPEC, size_t N>
void malloc(devices::B200& device, l2f::vector::Environment<SPEC, N>& env, const policy::Checkpoint& p){
    detail::check(device, env, b200l2f_policy_load(env.handle, &p.desc, p.blob.data(), p.blob.size()));
}
// rl::loop::steps::checkpoint files: `checkpoint.h` (code export) or `checkpoint.h5`; what the reference does by compiling the header in
// (post_training/load_actor.cpp) or rl_tools::load(device, actor, HighFive::Group)
inline void load(devices::B200& device, const std::string& path, policy::Checkpoint& out, const char* root = nullptr){
    std::ifstream f(path, std::ios::binary);
    utils::assert_exit(device, (bool)f, ("b200::load: cannot open " + path).c_str());
    const std::string bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    b200l2f_checkpoint* c = nullptr; size_t n = 0;
    int rc = b200l2f_checkpoint_parse(bytes.data(), bytes.size(), &c);       // (the status first: argument evaluation order is unspecified)
    if(rc != B200L2F_OK) utils::assert_exit(device, false, b200l2f_last_error(nullptr));
    rc = b200l2f_checkpoint_policy(c, root, &out.desc, nullptr, 0, &n);
    if(rc == B200L2F_OK){ out.blob.resize(n); rc = b200l2f_checkpoint_policy(c, root, &out.desc, out.blob.data(), n, &n); }
    const char* name = b200l2f_checkpoint_string(c, "rl_tools::checkpoint::meta::name");
    out.name = name ? name : "";
    const std::string err = rc == B200L2F_OK ? "" : b200l2f_last_error(nullptr);
    b200l2f_checkpoint_free(c);
    utils::assert_exit(device, rc == B200L2F_OK, err.c_str());
}
// rl_tools::json(device, env, parameters) / from_json on one flat parameter row [B200L2F_PARAMS_DIM] (the reference's exact text)
inline std::string json(devices::B200& device, const float* parameters_row){
    size_t n = 0;
    b200l2f_parameters_to_json(nullptr, parameters_row, nullptr, 0, &n);     // size query: reports "buffer too small" and the required length
    std::string text(n + 1, '\0');
    const int rc = b200l2f_parameters_to_json(nullptr, parameters_row, &text[0], n + 1, &n);
    if(rc != B200L2F_OK) utils::assert_exit(device, false, b200l2f_last_error(nullptr));
    text.resize(n);
    return text;
}
inline void from_json(devices::B200& device, const std::string& text, float* parameters_row_io){
    const int rc = b200l2f_parameters_from_json(nullptr, text.c_str(), parameters_row_io);
    if(rc != B200L2F_OK) utils::assert_exit(device, false, b200l2f_last_error(nullptr));
}

// ---- PPO: on-policy runner + dataset (rl/components/on_policy_runner/on_policy_runner.h:42-64,73-111) -------------------------------------------
namespace on_policy_runner {
    template <typename SPEC, size_t N, size_t T_STEPS_PER_ENV>
    struct Dataset {          // data [(STEPS + 1) * N][OBS + 15]: obs | actions_mean[4] | actions[4] | log_prob | reward | terminated | truncated | value | advantage | target_value
        static constexpr size_t STEPS_PER_ENV = T_STEPS_PER_ENV, STEPS_TOTAL = T_STEPS_PER_ENV * N, DATA_DIM = SPEC::OBSERVATION_DIM + 15;
        static constexpr size_t OBSERVATIONS = 0, ACTIONS_MEAN = SPEC::OBSERVATION_DIM, ACTIONS = ACTIONS_MEAN + 4, ACTION_LOG_PROBS = ACTIONS + 4, REWARDS = ACTION_LOG_PROBS + 1,
                                TERMINATED = REWARDS + 1, TRUNCATED = TERMINATED + 1, ALL_VALUES = TRUNCATED + 1, ADVANTAGES = ALL_VALUES + 1, TARGET_VALUES = ADVANTAGES + 1;
        std::vector<float> data = std::vector<float>((T_STEPS_PER_ENV + 1) * N * DATA_DIM, 0.0f);
        float& operator()(size_t row, size_t col){ return data[row * DATA_DIM + col]; }
        const float& operator()(size_t row, size_t col) const { return data[row * DATA_DIM + col]; }
    };
    template <typename SPEC, size_t N>
    struct Runner { l2f::vector::Environment<SPEC, N>* env = nullptr; int step_limit = SPEC::EPISODE_STEP_LIMIT; size_t step = 0; };
}
// rl_tools::init(device, runner, envs, parameters, rng) (operations_generic.h:65-75): every environment starts truncated, so the first collect re-samples
template <typename SPEC, size_t N>
void init(devices::B200& device, on_policy_runner::Runner<SPEC, N>& runner, l2f::vector::Environment<SPEC, N>& env, l2f::vector::Rng<N>& rng){
    runner.env = &env; runner.step = 0;
    detail::bind_rng(device, env, rng);
    detail::check(device, env, b200l2f_initial_parameters(env.handle));
    detail::check(device, env, b200l2f_initial_state(env.handle, 0));
    detail::check(device, env, b200l2f_collect_reset(env.handle));
}
// rl_tools::collect(device, dataset, runner, actor, actor_buffers, rng) (operations_generic.h:99-131): STEPS_PER_ENV steps of all environments in ONE launch;
// the actor is the one loaded with malloc(device, env, policy::MLP{head = PPO_GAUSSIAN})
template <typename SPEC, size_t N, size_t STEPS>
void collect(devices::B200& device, on_policy_runner::Dataset<SPEC, N, STEPS>& dataset, on_policy_runner::Runner<SPEC, N>& runner){
    detail::check(device, *runner.env, b200l2f_collect(runner.env->handle, (int32_t)STEPS, runner.step_limit, dataset.data.data(), B200L2F_HOST));
    runner.step += STEPS;
}
namespace ppo { struct Parameters { float GAMMA = 0.99f, LAMBDA = 0.95f; bool IGNORE_TERMINATION = false; }; }   // rl/algorithms/ppo/ppo.h:15-16,33
template <typename SPEC, size_t N>
void malloc_critic(devices::B200& device, l2f::vector::Environment<SPEC, N>& env, const policy::MLP& critic){
    const b200l2f_policy_desc d = critic.desc();
    detail::check(device, env, b200l2f_critic_load(env.handle, &d, critic.blob.data(), critic.blob.size()));
}
// evaluate(device, critic, dataset.all_observations_privileged, dataset.all_values, ...) (rl/algorithms/ppo/loop/core/operations_generic.h:112-116)
template <typename SPEC, size_t N, size_t STEPS>
void evaluate_values(devices::B200& device, on_policy_runner::Runner<SPEC, N>& runner, on_policy_runner::Dataset<SPEC, N, STEPS>& dataset){
    detail::check(device, *runner.env, b200l2f_evaluate_values(runner.env->handle, (int32_t)STEPS, dataset.data.data(), B200L2F_HOST));
}
// rl_tools::estimate_generalized_advantages(device, dataset, ppo_parameters) (rl/algorithms/ppo/operations_generic.h:54-89)
template <typename SPEC, size_t N, size_t STEPS>
void estimate_generalized_advantages(devices::B200& device, on_policy_runner::Runner<SPEC, N>& runner, on_policy_runner::Dataset<SPEC, N, STEPS>& dataset, const ppo::Parameters& p = {}){
    detail::check(device, *runner.env, b200l2f_estimate_generalized_advantages(runner.env->handle, (int32_t)STEPS, p.GAMMA, p.LAMBDA, p.IGNORE_TERMINATION ? 1 : 0, dataset.data.data(), B200L2F_HOST));
}
template <size_t DIM> struct RunningNormalizer { std::vector<float> mean = std::vector<float>(DIM, 0.0f), std = std::vector<float>(DIM, 1.0f); int32_t age = 0; };   // running_normalizer.h
// rl_tools::update(device, normalizer, dataset.observations) (rl/components/running_normalizer/operations_generic.h:27-49)
template <typename SPEC, size_t N, size_t STEPS>
void update(devices::B200& device, on_policy_runner::Runner<SPEC, N>& runner, RunningNormalizer<SPEC::OBSERVATION_DIM>& normalizer, const on_policy_runner::Dataset<SPEC, N, STEPS>& dataset){
    detail::check(device, *runner.env, b200l2f_normalizer_update(runner.env->handle, (int32_t)STEPS, dataset.data.data(), B200L2F_HOST, normalizer.mean.data(), normalizer.std.data(), &normalizer.age));
}

// ---- SAC teachers: off-policy runner with per-environment replay rings (rl/components/off_policy_runner/off_policy_runner.h, replay_buffer.h:37-58) ---
namespace off_policy_runner {
    template <typename SPEC, size_t N, size_t T_CAPACITY>
    struct Runner {
        static constexpr size_t CAPACITY = T_CAPACITY, DATA_DIM = 2 * SPEC::OBSERVATION_DIM + 7;
        l2f::vector::Environment<SPEC, N>* env = nullptr;
        int step_limit = SPEC::EPISODE_STEP_LIMIT; bool sample_parameters = true;
        std::vector<float> data = std::vector<float>(N * T_CAPACITY * DATA_DIM, 0.0f);     // [N][CAPACITY][obs | action[4] | reward | next_obs | terminated | truncated]
        std::vector<int32_t> episode_start = std::vector<int32_t>(N * T_CAPACITY, 0), position = std::vector<int32_t>(N, 0), current_episode_start = std::vector<int32_t>(N, 0);
        std::vector<uint8_t> full = std::vector<uint8_t>(N, 0);
        b200l2f_replay_buffers buffers(){ return b200l2f_replay_buffers{B200L2F_HOST, (int32_t)T_CAPACITY, data.data(), episode_start.data(), position.data(), full.data(), current_episode_start.data()}; }
    };
    // batch sampling parameters: the reference's SequentialBatchParameters (off_policy_runner.h:78-85), defaults as there
    template <size_t T_SEQUENCE_LENGTH>
    struct SequentialBatchParameters {
        static constexpr bool INCLUDE_FIRST_STEP_IN_TARGETS = T_SEQUENCE_LENGTH > 1;
        static constexpr bool ALWAYS_SAMPLE_FROM_INITIAL_STATE = T_SEQUENCE_LENGTH > 1;
        static constexpr bool RANDOM_SEQ_LENGTH = T_SEQUENCE_LENGTH > 1;
        static constexpr bool ENABLE_NOMINAL_SEQUENCE_LENGTH_PROBABILITY = true;
        static constexpr float NOMINAL_SEQUENCE_LENGTH_PROBABILITY = 0.5f;
    };
    template <typename SPEC, size_t T_BATCH_SIZE, size_t T_SEQUENCE_LENGTH = 1, typename T_PARAMETERS = SequentialBatchParameters<T_SEQUENCE_LENGTH>>
    struct SequentialBatch {          // off_policy_runner.h:96-141; [PADDED_SEQUENCE_LENGTH] tensors are the reference's `*_base` ones
        using PARAMETERS = T_PARAMETERS;
        static constexpr size_t BATCH_SIZE = T_BATCH_SIZE, SEQUENCE_LENGTH = T_SEQUENCE_LENGTH, PADDED_SEQUENCE_LENGTH = T_SEQUENCE_LENGTH + 1, DIM = SPEC::OBSERVATION_DIM + 4;
        std::vector<float> observations_actions = std::vector<float>(PADDED_SEQUENCE_LENGTH * T_BATCH_SIZE * DIM, 0.0f), rewards = std::vector<float>(SEQUENCE_LENGTH * T_BATCH_SIZE, 0.0f);
        std::vector<uint8_t> terminated = std::vector<uint8_t>(SEQUENCE_LENGTH * T_BATCH_SIZE, 0), reset = std::vector<uint8_t>(SEQUENCE_LENGTH * T_BATCH_SIZE, 0),
                             final_step_mask = std::vector<uint8_t>(SEQUENCE_LENGTH * T_BATCH_SIZE, 0), next_reset = std::vector<uint8_t>(PADDED_SEQUENCE_LENGTH * T_BATCH_SIZE, 0),
                             next_final_step_mask = std::vector<uint8_t>(PADDED_SEQUENCE_LENGTH * T_BATCH_SIZE, 0);
        std::vector<uint64_t> rng = std::vector<uint64_t>(T_BATCH_SIZE, 0);   // one stream per batch sample (operations_cuda.h:36-60); seed them once
    };
}
template <typename SPEC, size_t N, size_t CAPACITY>
void init(devices::B200& device, off_policy_runner::Runner<SPEC, N, CAPACITY>& runner, l2f::vector::Environment<SPEC, N>& env, l2f::vector::Rng<N>& rng){
    runner.env = &env;
    detail::bind_rng(device, env, rng);
    detail::check(device, env, b200l2f_initial_parameters(env.handle));
    detail::check(device, env, b200l2f_initial_state(env.handle, 0));
    detail::check(device, env, b200l2f_collect_reset(env.handle));
}
// rl_tools::step(device, runner, actor, actor_buffers, rng) (operations_generic.h:215-238) x n_steps in ONE launch
template <typename SPEC, size_t N, size_t CAPACITY>
void step(devices::B200& device, off_policy_runner::Runner<SPEC, N, CAPACITY>& runner, int n_steps = 1){
    const b200l2f_replay_buffers rb = runner.buffers();
    detail::check(device, *runner.env, b200l2f_off_policy_steps(runner.env->handle, n_steps, runner.step_limit, runner.sample_parameters ? 1 : 0, &rb));
}
// rl_tools::gather_batch(device, runner, batch, rng) (operations_generic.h:240-434), any SEQUENCE_LENGTH; the batch's PARAMETERS select the sampling
template <typename SPEC, size_t N, size_t CAPACITY, size_t BATCH, size_t SEQUENCE_LENGTH, typename PARAMETERS>
void gather_batch(devices::B200& device, off_policy_runner::Runner<SPEC, N, CAPACITY>& runner, off_policy_runner::SequentialBatch<SPEC, BATCH, SEQUENCE_LENGTH, PARAMETERS>& batch, int env_begin = 0, int env_count = (int)N){
    const b200l2f_replay_buffers rb = runner.buffers();
    b200l2f_batch out{};
    out.memspace = B200L2F_HOST; out.batch_size = (int32_t)BATCH; out.observations_actions = batch.observations_actions.data(); out.rewards = batch.rewards.data(); out.terminated = batch.terminated.data();
    out.reset = batch.reset.data(); out.next_reset = batch.next_reset.data(); out.final_step_mask = batch.final_step_mask.data(); out.next_final_step_mask = batch.next_final_step_mask.data();
    const b200l2f_batch_parameters bp{(int32_t)SEQUENCE_LENGTH, PARAMETERS::INCLUDE_FIRST_STEP_IN_TARGETS, PARAMETERS::ALWAYS_SAMPLE_FROM_INITIAL_STATE, PARAMETERS::RANDOM_SEQ_LENGTH,
                                      PARAMETERS::ENABLE_NOMINAL_SEQUENCE_LENGTH_PROBABILITY, PARAMETERS::NOMINAL_SEQUENCE_LENGTH_PROBABILITY};
    detail::check(device, *runner.env, b200l2f_gather_batch_sequential(runner.env->handle, &rb, &bp, SPEC::EPISODE_STEP_LIMIT, env_begin, env_count, batch.rng.data(), &out));   // MAX_EPISODE_LENGTH = ENVIRONMENT::EPISODE_STEP_LIMIT (off_policy_runner.h:41), not the runner's own step limit
}

// ---- DAgger (src/foundation_policy/post_training/helper.h): gather_epoch for all teachers in one call ----------------------------------------------
namespace dagger {
    template <size_t N, size_t T_STEPS>
    struct Dataset {
        std::vector<float> input_student = std::vector<float>(N * T_STEPS * 22, 0.0f), output_target = std::vector<float>(N * T_STEPS * 4, 0.0f);
        std::vector<uint8_t> truncated = std::vector<uint8_t>(N * T_STEPS, 0), reset = std::vector<uint8_t>(N * T_STEPS, 0);
        std::vector<int32_t> episode_start = std::vector<int32_t>(N, 0);
        int64_t rows = 0;
    };
}
// teachers: [n_teachers] SAC actors (MLP 26-64-64-8) + steady-state position offsets [n_teachers][3] (or nullptr); environment e belongs to teacher e / episodes_per_teacher
template <typename SPEC, size_t N>
void malloc_teachers(devices::B200& device, l2f::vector::Environment<SPEC, N>& env, int n_teachers, int episodes_per_teacher, const float* blobs, const float* position_offsets, bool tensor_cores = true){
    detail::check(device, env, b200l2f_teachers_load(env.handle, n_teachers, episodes_per_teacher, blobs, position_offsets, tensor_cores ? B200L2F_GEMM_TCGEN05_3XTF32 : B200L2F_GEMM_FP32_CUDA_CORES));
}
template <typename SPEC, size_t N, size_t STEPS>
void gather_epoch(devices::B200& device, l2f::vector::Environment<SPEC, N>& env, dagger::Dataset<N, STEPS>& dataset){
    b200l2f_dagger_out out{};
    out.memspace = B200L2F_HOST; out.capacity_rows = (int64_t)(N * STEPS); out.input_student = dataset.input_student.data(); out.output_target = dataset.output_target.data();
    out.truncated = dataset.truncated.data(); out.reset = dataset.reset.data(); out.episode_start = dataset.episode_start.data();
    detail::check(device, env, b200l2f_dagger_gather(env.handle, (int32_t)STEPS, 0, &out, &dataset.rows));
}
}  // namespace b200
