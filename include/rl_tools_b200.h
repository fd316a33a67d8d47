// include/rl_tools_b200.h -- the reference-side binding of the B200 engine, compiled against the REAL rl-tools headers.
//
// rl-tools selects implementations by overloading free functions on the device type (INC/operations/cpu.h:15-17, INC/devices/cuda.h).
// This header adds a device tag `rl_tools::devices::B200` and, for it, overloads of the l2f environment operations
// (L2F/operations_generic.h:38-176), of the actor's reset / evaluate_step (INC/nn_models/sequential/operations_generic.h:63-66,321-325)
// and of the batched evaluation (INC/rl/utils/evaluation/operations_generic.h:93-214) that forward to the C ABI in b200_l2f.h.
// The environment operations of the reference take ONE environment per call; the B200 overloads take a `vector::Environment<ENVIRONMENT, N>`
// (N environments of the reference's own ENVIRONMENT type, owned by the engine) and `rl_tools::Matrix` containers with N rows, i.e. the
// shapes the reference's vectorised callers already hold (`rl::utils::evaluation` keeps N environments / states / observation rows).
// Argument order and meaning follow the reference: device first, environment, parameters, state, ..., rng last.
//
// Usage (see tests/cpp/rl_tools_binding.cpp, built against /root/reference/rl-tools/include by tests/test_rl_tools_binding.py):
//     #include <rl_tools/operations/cpu.h>
//     #include <rl_tools/rl/environments/l2f/operations_generic.h>
//     #include <rl_tools/nn_models/sequential/operations_generic.h>     // + the layer operations the model needs
//     #include <rl_tools_b200.h>
// Link with -lb200l2f.  No CUDA headers are needed on the reference side.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "b200_l2f.h"

RL_TOOLS_NAMESPACE_WRAPPER_START
namespace rl_tools{
    namespace devices{
        // the device tag.  SPEC::MATH / RANDOM / LOGGING of the host side are the CPU ones: everything that runs on the tag's own hardware is behind the C ABI
        struct B200{
            using index_t = size_t;
            using SPEC = devices::DefaultCPUSpecification;
            int ordinal = 0;              // CUDA device ordinal
            int64_t first_env_id = 0;     // global id of local environment 0 (RNG streams are keyed by global id: results do not depend on the sharding)
            int32_t flags = 0;            // B200L2F_FLAG_*
        };
    }
    namespace rl::environments::l2f::b200{
        inline void check(int rc, b200l2f_handle* h, const char* what){   // utils::assert_exit behaviour: message + exit (10_sample_initial_parameters.h:68-199)
            if(rc != B200L2F_OK){ std::fprintf(stderr, "%s: %s\n", what, b200l2f_last_error(h)); std::exit(1); }
        }
        // which engine specification a reference ENVIRONMENT type is: by its action-history length, observation width, trajectory option and DR option
        // are the domain-randomisation options of a reference parameter type switched on?  (walks ParametersTrajectory -> ParametersDomainRandomization, L2F/multirotor.h:179-208)
        template <typename P> struct dr_enabled{ static constexpr bool value = false; };
        template <typename DR_SPEC> struct dr_enabled<ParametersDomainRandomization<DR_SPEC>>{
            using O = typename DR_SPEC::DOMAIN_RANDOMIZATION_OPTIONS;
            static constexpr bool value = O::THRUST_TO_WEIGHT || O::MASS || O::TORQUE_TO_INERTIA || O::MASS_SIZE_DEVIATION || O::ROTOR_TORQUE_CONSTANT || O::DISTURBANCE_FORCE || O::ROTOR_TIME_CONSTANT;
            static_assert(!value || (O::THRUST_TO_WEIGHT && O::MASS && O::TORQUE_TO_INERTIA && O::MASS_SIZE_DEVIATION && O::ROTOR_TORQUE_CONSTANT && O::DISTURBANCE_FORCE && O::ROTOR_TIME_CONSTANT),
                          "the engine instantiates domain randomisation with all options on (DEFAULT_DOMAIN_RANDOMIZATION_OPTIONS<true>) or all off");
        };
        template <typename T_SPEC> struct dr_enabled<ParametersTrajectory<T_SPEC>>: dr_enabled<typename T_SPEC::NEXT_COMPONENT>{ };
        template <typename ENVIRONMENT>
        constexpr int spec_id(){
            using STATIC = typename ENVIRONMENT::SPEC::STATIC_PARAMETERS;
            constexpr auto H = STATIC::ACTION_HISTORY_LENGTH;
            constexpr auto OBS = ENVIRONMENT::Observation::DIM;
            constexpr bool DR = dr_enabled<typename ENVIRONMENT::Parameters>::value;
            static_assert((H == 16 && OBS == 82) || (H == 1 && OBS == 22) || (H == 1 && OBS == 26),
                          "the engine instantiates the DEFAULT (H 16, OBS 82), RAPTOR (H 1, OBS 22) and TEACHER (H 1, OBS 26) specifications");
            return H == 16 ? (DR ? B200L2F_SPEC_DEFAULT_DR : B200L2F_SPEC_DEFAULT) : OBS == 22 ? (DR ? B200L2F_SPEC_RAPTOR_DR : B200L2F_SPEC_RAPTOR) : (DR ? B200L2F_SPEC_TEACHER_DR : B200L2F_SPEC_TEACHER);
        }
        // reference parameter struct -> the engine's flat row (layout documented in b200_l2f.h, mirrors L2F/multirotor.h:23-140)
        template <typename P>
        void flatten(const P& p, float* o){
            int k = 0;
            for(int i = 0; i < 4; i++) for(int j = 0; j < 3; j++) o[k++] = p.dynamics.rotor_positions[i][j];
            for(int i = 0; i < 4; i++) for(int j = 0; j < 3; j++) o[k++] = p.dynamics.rotor_thrust_directions[i][j];
            for(int i = 0; i < 4; i++) for(int j = 0; j < 3; j++) o[k++] = p.dynamics.rotor_torque_directions[i][j];
            for(int i = 0; i < 4; i++) for(int j = 0; j < 3; j++) o[k++] = p.dynamics.rotor_thrust_coefficients[i][j];
            for(int i = 0; i < 4; i++) o[k++] = p.dynamics.rotor_torque_constants[i];
            for(int i = 0; i < 4; i++) o[k++] = p.dynamics.rotor_time_constants_rising[i];
            for(int i = 0; i < 4; i++) o[k++] = p.dynamics.rotor_time_constants_falling[i];
            o[k++] = p.dynamics.mass;
            for(int i = 0; i < 3; i++) o[k++] = p.dynamics.gravity[i];
            for(int i = 0; i < 3; i++) for(int j = 0; j < 3; j++) o[k++] = p.dynamics.J[i][j];
            for(int i = 0; i < 3; i++) for(int j = 0; j < 3; j++) o[k++] = p.dynamics.J_inv[i][j];
            o[k++] = p.dynamics.hovering_throttle_relative; o[k++] = p.dynamics.action_limit.min; o[k++] = p.dynamics.action_limit.max;
            o[k++] = p.integration.dt;
            o[k++] = p.mdp.init.guidance; o[k++] = p.mdp.init.max_position; o[k++] = p.mdp.init.max_angle; o[k++] = p.mdp.init.max_linear_velocity;
            o[k++] = p.mdp.init.max_angular_velocity; o[k++] = p.mdp.init.relative_rpm ? 1.0f : 0.0f; o[k++] = p.mdp.init.min_rpm; o[k++] = p.mdp.init.max_rpm;
            const auto& r = p.mdp.reward;
            o[k++] = r.non_negative ? 1.0f : 0.0f; o[k++] = r.scale; o[k++] = r.constant; o[k++] = r.termination_penalty; o[k++] = r.position; o[k++] = r.position_clip;
            o[k++] = r.orientation; o[k++] = r.linear_velocity; o[k++] = r.angular_velocity; o[k++] = r.linear_acceleration; o[k++] = r.angular_acceleration;
            o[k++] = r.action; o[k++] = r.d_action; o[k++] = r.position_error_integral;
            const auto& on = p.mdp.observation_noise;
            o[k++] = on.position; o[k++] = on.orientation; o[k++] = on.linear_velocity; o[k++] = on.angular_velocity; o[k++] = on.imu_acceleration;
            o[k++] = p.mdp.action_noise.normalized_rpm;
            const auto& t = p.mdp.termination;
            o[k++] = t.enabled ? 1.0f : 0.0f; o[k++] = t.position_threshold; o[k++] = t.linear_velocity_threshold; o[k++] = t.angular_velocity_threshold;
            o[k++] = t.position_integral_threshold; o[k++] = t.orientation_integral_threshold;
            o[k++] = p.disturbances.random_force.mean; o[k++] = p.disturbances.random_force.std; o[k++] = p.disturbances.random_torque.mean; o[k++] = p.disturbances.random_torque.std;
            const auto& d = p.domain_randomization;
            o[k++] = d.thrust_to_weight_min; o[k++] = d.thrust_to_weight_max; o[k++] = d.torque_to_inertia_min; o[k++] = d.torque_to_inertia_max;
            o[k++] = d.mass_min; o[k++] = d.mass_max; o[k++] = d.mass_size_deviation;
            o[k++] = d.rotor_time_constant_rising_min; o[k++] = d.rotor_time_constant_rising_max; o[k++] = d.rotor_time_constant_falling_min; o[k++] = d.rotor_time_constant_falling_max;
            o[k++] = d.rotor_torque_constant_min; o[k++] = d.rotor_torque_constant_max; o[k++] = d.orientation_offset_angle_max; o[k++] = d.disturbance_force_max;
            o[k++] = p.trajectory.mixture[0]; o[k++] = p.trajectory.mixture[1];
            o[k++] = p.trajectory.langevin.gamma; o[k++] = p.trajectory.langevin.omega; o[k++] = p.trajectory.langevin.sigma; o[k++] = p.trajectory.langevin.alpha;
            if(k != B200L2F_PARAMS_DIM){ std::fprintf(stderr, "rl_tools_b200: parameter row has %d entries\n", k); std::exit(1); }
        }
        namespace vector{
            // N environments of the reference's ENVIRONMENT type (rl::environments::Multirotor<SPEC>).  `parameters` is the reference's env.parameters member:
            // init() fills it from SPEC::STATIC_PARAMETERS::PARAMETER_VALUES like the reference and hands it to the engine as the nominal row
            template <typename T_ENVIRONMENT, size_t T_N>
            struct Environment{
                using ENVIRONMENT = T_ENVIRONMENT;
                using T = typename ENVIRONMENT::T;
                static constexpr size_t N = T_N;
                static constexpr size_t OBSERVATION_DIM = ENVIRONMENT::Observation::DIM;
                static constexpr size_t ACTION_DIM = ENVIRONMENT::ACTION_DIM;
                typename ENVIRONMENT::Parameters parameters;
                b200l2f_handle* handle = nullptr;
            };
            template <size_t N> struct Parameters{ };                         // the per-environment parameter rows live in HBM (b200l2f_get/set_parameters)
            template <size_t N> struct State{ int slot = 0; };                // a state buffer of the engine: slot 0 = `state`, slot 1 = `next_state` of step()
            struct Rng{ uint64_t seed = 0; };                                 // one xorshift64 stream per environment, state 0xAAAAAAAA + seed + global id
            // the Raptor actor (Dense -> GRU -> Dense) resident on the device, loaded from a reference model instance by copy()
            template <size_t N> struct Policy{ b200l2f_handle* handle = nullptr; };
        }
        // weights of a reference sequential Dense-GRU-Dense model in the order b200l2f_policy_load takes them (= the order checkpoint.h stores them)
        // (parameter containers are rl_tools::Tensor of rank 1 or 2, INC/nn/parameters/parameters.h)
        template <typename DEVICE, typename TENSOR_SPEC>
        void append(DEVICE& device, const Tensor<TENSOR_SPEC>& t, std::vector<float>& blob){
            using TI = typename DEVICE::index_t;
            using SHAPE = typename TENSOR_SPEC::SHAPE;
            static_assert(SHAPE::LENGTH == 1 || SHAPE::LENGTH == 2);
            if constexpr(SHAPE::LENGTH == 1){ for(TI i = 0; i < SHAPE::template GET<0>; i++) blob.push_back(get(device, t, i)); }
            else{ for(TI r = 0; r < SHAPE::template GET<0>; r++) for(TI c = 0; c < SHAPE::template GET<1>; c++) blob.push_back(get(device, t, r, c)); }
        }
        template <typename DEVICE, typename MATRIX_SPEC>                      // dense layers keep rl_tools::Matrix containers (INC/nn/layers/dense/layer.h:94-104)
        void append(DEVICE&, const Matrix<MATRIX_SPEC>& m, std::vector<float>& blob){
            for(typename DEVICE::index_t r = 0; r < MATRIX_SPEC::ROWS; r++) for(typename DEVICE::index_t c = 0; c < MATRIX_SPEC::COLS; c++) blob.push_back(get(m, r, c));
        }
        template <typename MATRIX_SPEC> constexpr int dim0(const Matrix<MATRIX_SPEC>&){ return (int)MATRIX_SPEC::ROWS; }
        template <typename MATRIX_SPEC> constexpr int dim1(const Matrix<MATRIX_SPEC>&){ return (int)MATRIX_SPEC::COLS; }
        template <typename TENSOR_SPEC> constexpr int dim0(const Tensor<TENSOR_SPEC>&){ return (int)TENSOR_SPEC::SHAPE::template GET<0>; }
        template <typename TENSOR_SPEC> constexpr int dim1(const Tensor<TENSOR_SPEC>&){ return (int)TENSOR_SPEC::SHAPE::template GET<1>; }
    }
    // ---- lifetime: rl_tools::malloc / free / init (L2F/operations_generic.h:38-46)
    template <typename ENVIRONMENT, size_t N>
    void malloc(devices::B200& device, rl::environments::l2f::b200::vector::Environment<ENVIRONMENT, N>& env){
        b200l2f_config c{};
        c.struct_size = (int32_t)sizeof(b200l2f_config); c.spec = rl::environments::l2f::b200::spec_id<ENVIRONMENT>(); c.n_envs = (int32_t)N; c.device = device.ordinal;
        c.first_env_id = device.first_env_id; c.n_state_slots = 2; c.flags = device.flags; c.stream = nullptr;
        const int rc = b200l2f_create(&c, &env.handle);
        rl::environments::l2f::b200::check(rc, nullptr, "rl_tools::malloc(devices::B200)");
    }
    template <typename ENVIRONMENT, size_t N>
    void free(devices::B200&, rl::environments::l2f::b200::vector::Environment<ENVIRONMENT, N>& env){ b200l2f_destroy(env.handle); env.handle = nullptr; }
    template <typename ENVIRONMENT, size_t N>
    void init(devices::B200&, rl::environments::l2f::b200::vector::Environment<ENVIRONMENT, N>& env){
        env.parameters = ENVIRONMENT::SPEC::STATIC_PARAMETERS::PARAMETER_VALUES;                          // as rl_tools::init, operations_generic.h:43-46
        float row[B200L2F_PARAMS_DIM];
        rl::environments::l2f::b200::flatten(env.parameters, row);
        rl::environments::l2f::b200::check(b200l2f_set_environment_parameters(env.handle, row), env.handle, "rl_tools::init(devices::B200)");
    }
    template <typename ENVIRONMENT, size_t N>
    void init(devices::B200&, rl::environments::l2f::b200::vector::Environment<ENVIRONMENT, N>& env, rl::environments::l2f::b200::vector::Rng& rng, uint64_t seed, int warmup = 0){
        rng.seed = seed;
        rl::environments::l2f::b200::check(b200l2f_initialize_rng(env.handle, seed, warmup), env.handle, "rl_tools::init(rng)");
    }
    // ---- parameters / states (operations_generic.h:69-86)
    template <typename ENVIRONMENT, size_t N>
    void initial_parameters(devices::B200&, rl::environments::l2f::b200::vector::Environment<ENVIRONMENT, N>& env, rl::environments::l2f::b200::vector::Parameters<N>&){
        rl::environments::l2f::b200::check(b200l2f_initial_parameters(env.handle), env.handle, "rl_tools::initial_parameters");
    }
    template <typename ENVIRONMENT, size_t N>
    void sample_initial_parameters(devices::B200&, rl::environments::l2f::b200::vector::Environment<ENVIRONMENT, N>& env, rl::environments::l2f::b200::vector::Parameters<N>&, rl::environments::l2f::b200::vector::Rng&){
        rl::environments::l2f::b200::check(b200l2f_sample_initial_parameters(env.handle), env.handle, "rl_tools::sample_initial_parameters");
    }
    template <typename ENVIRONMENT, size_t N>
    void initial_state(devices::B200&, rl::environments::l2f::b200::vector::Environment<ENVIRONMENT, N>& env, rl::environments::l2f::b200::vector::Parameters<N>&, rl::environments::l2f::b200::vector::State<N>& state){
        rl::environments::l2f::b200::check(b200l2f_initial_state(env.handle, state.slot), env.handle, "rl_tools::initial_state");
    }
    template <typename ENVIRONMENT, size_t N>
    void sample_initial_state(devices::B200&, rl::environments::l2f::b200::vector::Environment<ENVIRONMENT, N>& env, rl::environments::l2f::b200::vector::Parameters<N>&, rl::environments::l2f::b200::vector::State<N>& state,
                              rl::environments::l2f::b200::vector::Rng&){
        rl::environments::l2f::b200::check(b200l2f_sample_initial_state(env.handle, state.slot), env.handle, "rl_tools::sample_initial_state");
    }
    // ---- observe / step / reward / terminated (operations_generic.h:87-176): Matrix containers with one row per environment
    template <typename ENVIRONMENT, size_t N, typename OBS_SPEC>
    void observe(devices::B200&, rl::environments::l2f::b200::vector::Environment<ENVIRONMENT, N>& env, rl::environments::l2f::b200::vector::Parameters<N>&, const rl::environments::l2f::b200::vector::State<N>& state,
                 Matrix<OBS_SPEC>& observation, rl::environments::l2f::b200::vector::Rng&){
        static_assert(OBS_SPEC::ROWS == N && OBS_SPEC::COLS == ENVIRONMENT::Observation::DIM);
        static_assert(OBS_SPEC::COL_PITCH == 1, "row-major observation matrix");
        rl::environments::l2f::b200::check(b200l2f_observe(env.handle, state.slot, observation._data, (int)OBS_SPEC::ROW_PITCH, B200L2F_HOST), env.handle, "rl_tools::observe");
    }
    template <typename ENVIRONMENT, size_t N, typename ACTION_SPEC>
    typename ENVIRONMENT::T step(devices::B200&, rl::environments::l2f::b200::vector::Environment<ENVIRONMENT, N>& env, rl::environments::l2f::b200::vector::Parameters<N>&,
                                 const rl::environments::l2f::b200::vector::State<N>& state, const Matrix<ACTION_SPEC>& action, rl::environments::l2f::b200::vector::State<N>& next_state,
                                 rl::environments::l2f::b200::vector::Rng&){
        static_assert(ACTION_SPEC::ROWS == N && ACTION_SPEC::COLS == ENVIRONMENT::ACTION_DIM);
        static_assert(ACTION_SPEC::COL_PITCH == 1 && ACTION_SPEC::ROW_PITCH == ENVIRONMENT::ACTION_DIM, "dense row-major action matrix");
        rl::environments::l2f::b200::check(b200l2f_step(env.handle, state.slot, action._data, next_state.slot, nullptr, B200L2F_HOST), env.handle, "rl_tools::step");
        return env.parameters.integration.dt;                                                               // operations_generic.h:129
    }
    template <typename ENVIRONMENT, size_t N, typename ACTION_SPEC, typename REWARD_SPEC>
    void reward(devices::B200&, rl::environments::l2f::b200::vector::Environment<ENVIRONMENT, N>& env, rl::environments::l2f::b200::vector::Parameters<N>&, const rl::environments::l2f::b200::vector::State<N>& state,
                const Matrix<ACTION_SPEC>& action, const rl::environments::l2f::b200::vector::State<N>& next_state, Matrix<REWARD_SPEC>& rewards, rl::environments::l2f::b200::vector::Rng&){
        static_assert(REWARD_SPEC::ROWS * REWARD_SPEC::COLS == N && REWARD_SPEC::COL_PITCH == 1);
        rl::environments::l2f::b200::check(b200l2f_reward(env.handle, state.slot, action._data, next_state.slot, rewards._data, B200L2F_HOST), env.handle, "rl_tools::reward");
    }
    template <typename ENVIRONMENT, size_t N>
    void terminated(devices::B200&, rl::environments::l2f::b200::vector::Environment<ENVIRONMENT, N>& env, rl::environments::l2f::b200::vector::Parameters<N>&, const rl::environments::l2f::b200::vector::State<N>& state,
                    bool* flags, rl::environments::l2f::b200::vector::Rng&){
        uint8_t raw[N];
        rl::environments::l2f::b200::check(b200l2f_terminated(env.handle, state.slot, raw, B200L2F_HOST), env.handle, "rl_tools::terminated");
        for(size_t i = 0; i < N; i++) flags[i] = raw[i] != 0;
    }
    template <typename ENVIRONMENT, size_t N>
    void copy(devices::B200&, rl::environments::l2f::b200::vector::Environment<ENVIRONMENT, N>& env, const rl::environments::l2f::b200::vector::State<N>& source, rl::environments::l2f::b200::vector::State<N>& target){
        rl::environments::l2f::b200::check(b200l2f_copy_state(env.handle, target.slot, source.slot), env.handle, "rl_tools::copy(state)");   // state = next_state
    }
    // ---- the actor: rl_tools::copy(source device -> B200) of a reference Dense-GRU-Dense sequential model, reset, evaluate_step
    template <typename SOURCE_DEVICE, typename MODEL, typename ENVIRONMENT, size_t N>
    void copy(SOURCE_DEVICE& source_device, devices::B200&, const MODEL& model, rl::environments::l2f::b200::vector::Environment<ENVIRONMENT, N>& env, rl::environments::l2f::b200::vector::Policy<N>& policy){
        namespace b = rl::environments::l2f::b200;
        std::vector<float> blob;
        const auto& dense_in = model.content; const auto& gru = model.next_module.content; const auto& dense_out = model.next_module.next_module.content;
        b::append(source_device, dense_in.weights.parameters, blob); b::append(source_device, dense_in.biases.parameters, blob);
        b::append(source_device, gru.weights_input.parameters, blob); b::append(source_device, gru.biases_input.parameters, blob);
        b::append(source_device, gru.weights_hidden.parameters, blob); b::append(source_device, gru.biases_hidden.parameters, blob);
        b::append(source_device, gru.initial_hidden_state.parameters, blob);
        b::append(source_device, dense_out.weights.parameters, blob); b::append(source_device, dense_out.biases.parameters, blob);
        b200l2f_policy_desc d{};
        d.arch = B200L2F_POLICY_RAPTOR_GRU;
        d.input_dim = b::dim1(dense_in.weights.parameters); d.hidden_dim = b::dim0(dense_in.weights.parameters); d.output_dim = b::dim0(dense_out.weights.parameters);
        d.head = B200L2F_HEAD_IDENTITY; d.gemm = B200L2F_GEMM_TCGEN05_3XTF32;
        d.gru_sequence_length = (int32_t)MODEL::INPUT_SHAPE::template GET<0>;     // SEQUENCE_LENGTH: the auto-reset period of gru/operations_generic.h:76-86
        policy.handle = env.handle;
        b::check(b200l2f_policy_load(env.handle, &d, blob.data(), blob.size()), env.handle, "rl_tools::copy(model -> devices::B200)");
    }
    template <size_t N>
    void reset(devices::B200&, rl::environments::l2f::b200::vector::Policy<N>& policy){                   // nn_models/sequential/operations_generic.h:63-66
        rl::environments::l2f::b200::check(b200l2f_policy_reset(policy.handle, nullptr, B200L2F_HOST), policy.handle, "rl_tools::reset(policy)");
    }
    template <size_t N, typename INPUT_SPEC, typename OUTPUT_SPEC>
    void evaluate_step(devices::B200&, rl::environments::l2f::b200::vector::Policy<N>& policy, const Matrix<INPUT_SPEC>& input, Matrix<OUTPUT_SPEC>& output, bool no_auto_reset = false){   // :321-325
        static_assert(INPUT_SPEC::ROWS == N && OUTPUT_SPEC::ROWS == N && INPUT_SPEC::COL_PITCH == 1 && OUTPUT_SPEC::COL_PITCH == 1 && OUTPUT_SPEC::ROW_PITCH == OUTPUT_SPEC::COLS);
        rl::environments::l2f::b200::check(b200l2f_policy_evaluate_step(policy.handle, input._data, (int)INPUT_SPEC::ROW_PITCH, output._data, no_auto_reset ? 1 : 0, B200L2F_HOST), policy.handle, "rl_tools::evaluate_step");
    }
    // ---- the loop owner: the body of rl_tools::evaluate (rl/utils/evaluation/operations_generic.h:138-189) as ONE fused launch; fills the reference's
    // ---- Result fields returns / episode_length per episode and their mean / std (:201-213)
    template <typename ENVIRONMENT, size_t N, typename RESULT>
    void evaluate(devices::B200&, rl::environments::l2f::b200::vector::Environment<ENVIRONMENT, N>& env, rl::environments::l2f::b200::vector::Policy<N>& policy, RESULT& results, size_t step_limit, bool no_auto_reset = false){
        namespace b = rl::environments::l2f::b200;
        static_assert(RESULT::SPEC::N_EPISODES == N, "one episode per environment");
        using T = typename RESULT::T;
        float returns[N]; int32_t lengths[N];
        b200l2f_rollout_out out{};
        out.memspace = B200L2F_HOST; out.returns = returns; out.episode_length = lengths;
        b::check(b200l2f_policy_reset(policy.handle, nullptr, B200L2F_HOST), env.handle, "rl_tools::evaluate: reset");
        b::check(b200l2f_rollout(env.handle, (int32_t)step_limit, no_auto_reset ? 1 : 0, &out), env.handle, "rl_tools::evaluate");
        for(size_t i = 0; i < N; i++){ results.returns[i] = returns[i]; results.episode_length[i] = (typename RESULT::TI)lengths[i]; }
        b200l2f_status st;                                                                                  // reduced on the device behind the fused kernel
        b::check(b200l2f_last_status(env.handle, &st, nullptr, B200L2F_HOST), env.handle, "rl_tools::evaluate: status");
        results.returns_mean = (T)st.returns_mean; results.returns_std = (T)st.returns_std;
        results.episode_length_mean = (T)st.episode_length_mean; results.episode_length_std = (T)st.episode_length_std;
        results.num_terminated = (typename RESULT::TI)st.n_terminated; results.share_terminated = (T)st.share_terminated;
        if(st.n_nonfinite != 0) std::fprintf(stderr, "rl_tools::evaluate(devices::B200): %lld of %lld environments ended with a non-finite state\n", (long long)st.n_nonfinite, (long long)st.n_envs);
    }
}
RL_TOOLS_NAMESPACE_WRAPPER_END
