// oracle/ref_l2f.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Thin extern "C" driver around the UNMODIFIED reference (rl-tools headers under
// /root/reference/rl-tools/include, env factories under .../src/foundation_policy and the
// Raptor checkpoint header extracted from /root/reference/data/raptor-policy-checkpoint.tar.gz).
// It is compiled by oracle/Makefile into oracle/_ref/libl2f_ref.so *from the sources where they
// lie* (never copied into this repository) and is used only by tests/, by
// __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.
//
// Everything in here is glue: flat float32 rows <-> the reference's Parameters / State structs,
// and loops that call the reference's own free functions
//   rl_tools::sample_initial_parameters / sample_initial_state / initial_state / observe / step /
//   reward / terminated                   (rl/environments/l2f/operations_generic.h:69-176)
//   rl_tools::reset / evaluate_step       (nn_models/sequential/operations_generic.h:63-66,321-325)
// in the order of rl_tools::evaluate      (rl/utils/evaluation/operations_generic.h:118-189).
//
// RNG contract (see DESIGN.md "RNG"): the engine is the reference's Generic xorshift64
// (random/operations_generic.h:16-31, uniform :52-58, Box-Muller normal :59-71) plugged into the
// reference's CPU device type, with the reference CPU device's "std == 0 returns the mean
// without consuming the stream" rule (random/operations_cpu.h:39-41) for the normal
// distribution.  One stream per environment, seeded base_seed + global_env_id
// (precedent: rl/components/on_policy_runner/operations_cpu.h:36-44).

#include <cstdint>
#include <cstring>
#include <cstdio>
#include <thread>
#include <vector>
#include <chrono>

#include <rl_tools/operations/cpu.h>

// ---- RNG plug-in: must be declared before the l2f / nn operations are parsed (qualified calls
// ---- inside templates only see overloads declared before the template definition).
RL_TOOLS_NAMESPACE_WRAPPER_START
namespace rl_tools{
    namespace devices::random{
        struct B200Contract: devices::random::Generic<devices::math::CPU>{
            static constexpr Type TYPE = Type::random;
        };
    }
    namespace random::normal_distribution{
        template<typename T, typename RNG>
        T sample(const devices::random::B200Contract& dev, T mean, T std, RNG& rng){
            if(std == 0){ // reference CPU-device rule, random/operations_cpu.h:39-41
                return mean;
            }
            return sample(static_cast<const devices::random::Generic<devices::math::CPU>&>(dev), mean, std, rng);
        }
    }
}
RL_TOOLS_NAMESPACE_WRAPPER_END

#include <rl_tools/nn/layers/standardize/operations_generic.h>
#include <rl_tools/nn/layers/dense/operations_generic.h>
#include <rl_tools/nn/layers/sample_and_squash/operations_generic.h>
#include <rl_tools/nn/layers/gru/operations_generic.h>
#include <rl_tools/nn_models/mlp/operations_generic.h>
#include <rl_tools/nn_models/mlp_unconditional_stddev/operations_generic.h>
#include <rl_tools/nn_models/sequential/operations_generic.h>

#include <rl_tools/rl/environments/l2f/operations_generic.h>

namespace rlt = rl_tools;

using DEV_SPEC = rlt::devices::cpu::Specification<rlt::devices::math::CPU, rlt::devices::random::B200Contract, rlt::devices::logging::CPU>;
using DEVICE = rlt::devices::CPU<DEV_SPEC>;
using RNG = rlt::devices::random::B200Contract::ENGINE<>;
using T = float;
using TI = typename DEVICE::index_t;
static_assert(sizeof(RNG) == 8, "xorshift64 engine");

// the reference's own environment factories for the foundation policy (Raptor) training
#include "pre_training/options.h"
struct OPTIONS_POST_TRAINING: OPTIONS_PRE_TRAINING{ // post_training/config.h:1-7 (values restated; config.h itself drags in the trainer)
    static constexpr bool OBSERVE_THRASH_MARKOV = false;
    static constexpr bool MOTOR_DELAY = true;
    static constexpr bool ACTION_HISTORY = true;
    static constexpr TI ACTION_HISTORY_LENGTH = 1;
    static constexpr bool OBSERVATION_NOISE = true;
};
#include "post_training/environment.h"

// the Raptor checkpoint (weights + known-answer test vectors), extracted by oracle/Makefile
#include "checkpoint.h"

// ------------------------------------------------------------------------------------------------
// Environment specs. ids are shared with include/b200_l2f.h (B200L2F_SPEC_*)
//   0 DEFAULT : l2f::Specification<float,size_t>, H=16, OBS 82, LANGEVIN off, DR options off   (parameters/default.h:29-176)
//   1 DEFAULT_DR : same state/observation, DEFAULT_DOMAIN_RANDOMIZATION_OPTIONS<true>           (parameters/default.h:16-26)
//   2 RAPTOR  : foundation-policy post-training env, H=1, OBS 22, LANGEVIN on                   (src/foundation_policy/post_training/environment.h:13-46)
//   3 TEACHER : foundation-policy pre-training env,  H=1, OBS 26, LANGEVIN on                   (src/foundation_policy/pre_training/environment.h:58-90)
//   4 RAPTOR_DR: RAPTOR state/observation/trajectory options but DR options on (used to draw per-env
//               dynamics directly in float32; the reference draws them off-line in double and ships JSON,
//               src/foundation_policy/pre_training/sample_dynamics_parameters.cpp:48-64)
// ------------------------------------------------------------------------------------------------
namespace l2f = rlt::rl::environments::l2f;
using ENV_DEFAULT = rlt::rl::environments::Multirotor<l2f::Specification<T, TI>>;
using FACTORY_DR = l2f::parameters::DEFAULT_PARAMETERS_FACTORY<T, TI, l2f::parameters::DEFAULT_DOMAIN_RANDOMIZATION_OPTIONS<true>>;
using ENV_DEFAULT_DR = rlt::rl::environments::Multirotor<l2f::Specification<T, TI, FACTORY_DR::STATIC_PARAMETERS>>;
using ENV_RAPTOR = typename builder::ENVIRONMENT_FACTORY_POST_TRAINING<DEVICE, T, TI, OPTIONS_POST_TRAINING>::ENVIRONMENT;
using ENV_TEACHER = typename builder::ENVIRONMENT_FACTORY<DEVICE, T, TI, OPTIONS_PRE_TRAINING>::ENVIRONMENT;

// RAPTOR_DR / TEACHER_DR: same static parameters but a parameter type whose DR options are enabled
template <typename BASE_STATIC, typename PARAMS_DR>
struct STATIC_WITH_DR: BASE_STATIC{
    using PARAMETERS = PARAMS_DR;
    static constexpr PARAMS_DR make(){
        PARAMS_DR p{};
        constexpr auto b = BASE_STATIC::PARAMETER_VALUES;
        p.dynamics = b.dynamics; p.integration = b.integration; p.mdp = b.mdp;
        p.disturbances = b.disturbances; p.domain_randomization = b.domain_randomization; p.trajectory = b.trajectory;
        return p;
    }
    static constexpr PARAMS_DR PARAMETER_VALUES = make();
};
struct TRAJ_LANGEVIN{ static constexpr bool LANGEVIN = true; };
using PARAMS_SPEC_BASE = l2f::ParametersBaseSpecification<T, TI, 4, l2f::parameters::reward_functions::Squared<T>>;
using PARAMS_LANGEVIN_DR = l2f::ParametersTrajectory<l2f::ParametersTrajectorySpecification<T, TI, TRAJ_LANGEVIN,
      l2f::ParametersDomainRandomization<l2f::ParametersDomainRandomizationSpecification<T, TI, l2f::parameters::DEFAULT_DOMAIN_RANDOMIZATION_OPTIONS<true>,
      l2f::ParametersDisturbances<l2f::ParametersSpecification<T, TI, l2f::ParametersBase<PARAMS_SPEC_BASE>>>>>>>;
using ENV_RAPTOR_DR = rlt::rl::environments::Multirotor<l2f::Specification<T, TI, STATIC_WITH_DR<typename ENV_RAPTOR::SPEC::STATIC_PARAMETERS, PARAMS_LANGEVIN_DR>>>;
using ENV_TEACHER_DR = rlt::rl::environments::Multirotor<l2f::Specification<T, TI, STATIC_WITH_DR<typename ENV_TEACHER::SPEC::STATIC_PARAMETERS, PARAMS_LANGEVIN_DR>>>;

static_assert(ENV_DEFAULT::Observation::DIM == 82);
static_assert(ENV_RAPTOR::Observation::DIM == 22);
static_assert(ENV_TEACHER::Observation::DIM == 26);

// ------------------------------------------------------------------------------------------------
// Flat layouts (shared with include/b200_l2f.h; offsets are asserted in tests)
// ------------------------------------------------------------------------------------------------
constexpr int PARAMS_DIM = 145;
constexpr int state_dim(int H){ return 44 + 4 * H; }

template <typename P>
static void flatten_parameters(const P& p, float* o){
    int k = 0;
    for(int i=0;i<4;i++) for(int j=0;j<3;j++) o[k++] = p.dynamics.rotor_positions[i][j];
    for(int i=0;i<4;i++) for(int j=0;j<3;j++) o[k++] = p.dynamics.rotor_thrust_directions[i][j];
    for(int i=0;i<4;i++) for(int j=0;j<3;j++) o[k++] = p.dynamics.rotor_torque_directions[i][j];
    for(int i=0;i<4;i++) for(int j=0;j<3;j++) o[k++] = p.dynamics.rotor_thrust_coefficients[i][j];
    for(int i=0;i<4;i++) o[k++] = p.dynamics.rotor_torque_constants[i];
    for(int i=0;i<4;i++) o[k++] = p.dynamics.rotor_time_constants_rising[i];
    for(int i=0;i<4;i++) o[k++] = p.dynamics.rotor_time_constants_falling[i];
    o[k++] = p.dynamics.mass;
    for(int i=0;i<3;i++) o[k++] = p.dynamics.gravity[i];
    for(int i=0;i<3;i++) for(int j=0;j<3;j++) o[k++] = p.dynamics.J[i][j];
    for(int i=0;i<3;i++) for(int j=0;j<3;j++) o[k++] = p.dynamics.J_inv[i][j];
    o[k++] = p.dynamics.hovering_throttle_relative;
    o[k++] = p.dynamics.action_limit.min;
    o[k++] = p.dynamics.action_limit.max;
    o[k++] = p.integration.dt;                               // 85
    o[k++] = p.mdp.init.guidance;                            // 86
    o[k++] = p.mdp.init.max_position;
    o[k++] = p.mdp.init.max_angle;
    o[k++] = p.mdp.init.max_linear_velocity;
    o[k++] = p.mdp.init.max_angular_velocity;
    o[k++] = p.mdp.init.relative_rpm ? 1.0f : 0.0f;
    o[k++] = p.mdp.init.min_rpm;
    o[k++] = p.mdp.init.max_rpm;                             // 93
    o[k++] = p.mdp.reward.non_negative ? 1.0f : 0.0f;        // 94
    o[k++] = p.mdp.reward.scale;
    o[k++] = p.mdp.reward.constant;
    o[k++] = p.mdp.reward.termination_penalty;
    o[k++] = p.mdp.reward.position;
    o[k++] = p.mdp.reward.position_clip;
    o[k++] = p.mdp.reward.orientation;
    o[k++] = p.mdp.reward.linear_velocity;
    o[k++] = p.mdp.reward.angular_velocity;
    o[k++] = p.mdp.reward.linear_acceleration;
    o[k++] = p.mdp.reward.angular_acceleration;
    o[k++] = p.mdp.reward.action;
    o[k++] = p.mdp.reward.d_action;
    o[k++] = p.mdp.reward.position_error_integral;           // 107
    o[k++] = p.mdp.observation_noise.position;               // 108
    o[k++] = p.mdp.observation_noise.orientation;
    o[k++] = p.mdp.observation_noise.linear_velocity;
    o[k++] = p.mdp.observation_noise.angular_velocity;
    o[k++] = p.mdp.observation_noise.imu_acceleration;       // 112
    o[k++] = p.mdp.action_noise.normalized_rpm;              // 113
    o[k++] = p.mdp.termination.enabled ? 1.0f : 0.0f;        // 114
    o[k++] = p.mdp.termination.position_threshold;
    o[k++] = p.mdp.termination.linear_velocity_threshold;
    o[k++] = p.mdp.termination.angular_velocity_threshold;
    o[k++] = p.mdp.termination.position_integral_threshold;
    o[k++] = p.mdp.termination.orientation_integral_threshold; // 119
    o[k++] = p.disturbances.random_force.mean;               // 120
    o[k++] = p.disturbances.random_force.std;
    o[k++] = p.disturbances.random_torque.mean;
    o[k++] = p.disturbances.random_torque.std;               // 123
    const auto& d = p.domain_randomization;                  // 124
    o[k++] = d.thrust_to_weight_min; o[k++] = d.thrust_to_weight_max;
    o[k++] = d.torque_to_inertia_min; o[k++] = d.torque_to_inertia_max;
    o[k++] = d.mass_min; o[k++] = d.mass_max; o[k++] = d.mass_size_deviation;
    o[k++] = d.rotor_time_constant_rising_min; o[k++] = d.rotor_time_constant_rising_max;
    o[k++] = d.rotor_time_constant_falling_min; o[k++] = d.rotor_time_constant_falling_max;
    o[k++] = d.rotor_torque_constant_min; o[k++] = d.rotor_torque_constant_max;
    o[k++] = d.orientation_offset_angle_max; o[k++] = d.disturbance_force_max; // 138
    o[k++] = p.trajectory.mixture[0]; o[k++] = p.trajectory.mixture[1];       // 139
    o[k++] = p.trajectory.langevin.gamma; o[k++] = p.trajectory.langevin.omega;
    o[k++] = p.trajectory.langevin.sigma; o[k++] = p.trajectory.langevin.alpha; // 144
    if(k != PARAMS_DIM){ std::fprintf(stderr, "flatten_parameters: %d\n", k); std::abort(); }
}
template <typename P>
static void unflatten_parameters(const float* o, P& p){
    int k = 0;
    for(int i=0;i<4;i++) for(int j=0;j<3;j++) p.dynamics.rotor_positions[i][j] = o[k++];
    for(int i=0;i<4;i++) for(int j=0;j<3;j++) p.dynamics.rotor_thrust_directions[i][j] = o[k++];
    for(int i=0;i<4;i++) for(int j=0;j<3;j++) p.dynamics.rotor_torque_directions[i][j] = o[k++];
    for(int i=0;i<4;i++) for(int j=0;j<3;j++) p.dynamics.rotor_thrust_coefficients[i][j] = o[k++];
    for(int i=0;i<4;i++) p.dynamics.rotor_torque_constants[i] = o[k++];
    for(int i=0;i<4;i++) p.dynamics.rotor_time_constants_rising[i] = o[k++];
    for(int i=0;i<4;i++) p.dynamics.rotor_time_constants_falling[i] = o[k++];
    p.dynamics.mass = o[k++];
    for(int i=0;i<3;i++) p.dynamics.gravity[i] = o[k++];
    for(int i=0;i<3;i++) for(int j=0;j<3;j++) p.dynamics.J[i][j] = o[k++];
    for(int i=0;i<3;i++) for(int j=0;j<3;j++) p.dynamics.J_inv[i][j] = o[k++];
    p.dynamics.hovering_throttle_relative = o[k++];
    p.dynamics.action_limit.min = o[k++];
    p.dynamics.action_limit.max = o[k++];
    p.integration.dt = o[k++];
    p.mdp.init.guidance = o[k++];
    p.mdp.init.max_position = o[k++];
    p.mdp.init.max_angle = o[k++];
    p.mdp.init.max_linear_velocity = o[k++];
    p.mdp.init.max_angular_velocity = o[k++];
    p.mdp.init.relative_rpm = o[k++] != 0;
    p.mdp.init.min_rpm = o[k++];
    p.mdp.init.max_rpm = o[k++];
    p.mdp.reward.non_negative = o[k++] != 0;
    p.mdp.reward.scale = o[k++];
    p.mdp.reward.constant = o[k++];
    p.mdp.reward.termination_penalty = o[k++];
    p.mdp.reward.position = o[k++];
    p.mdp.reward.position_clip = o[k++];
    p.mdp.reward.orientation = o[k++];
    p.mdp.reward.linear_velocity = o[k++];
    p.mdp.reward.angular_velocity = o[k++];
    p.mdp.reward.linear_acceleration = o[k++];
    p.mdp.reward.angular_acceleration = o[k++];
    p.mdp.reward.action = o[k++];
    p.mdp.reward.d_action = o[k++];
    p.mdp.reward.position_error_integral = o[k++];
    p.mdp.observation_noise.position = o[k++];
    p.mdp.observation_noise.orientation = o[k++];
    p.mdp.observation_noise.linear_velocity = o[k++];
    p.mdp.observation_noise.angular_velocity = o[k++];
    p.mdp.observation_noise.imu_acceleration = o[k++];
    p.mdp.action_noise.normalized_rpm = o[k++];
    p.mdp.termination.enabled = o[k++] != 0;
    p.mdp.termination.position_threshold = o[k++];
    p.mdp.termination.linear_velocity_threshold = o[k++];
    p.mdp.termination.angular_velocity_threshold = o[k++];
    p.mdp.termination.position_integral_threshold = o[k++];
    p.mdp.termination.orientation_integral_threshold = o[k++];
    p.disturbances.random_force.mean = o[k++];
    p.disturbances.random_force.std = o[k++];
    p.disturbances.random_torque.mean = o[k++];
    p.disturbances.random_torque.std = o[k++];
    auto& d = p.domain_randomization;
    d.thrust_to_weight_min = o[k++]; d.thrust_to_weight_max = o[k++];
    d.torque_to_inertia_min = o[k++]; d.torque_to_inertia_max = o[k++];
    d.mass_min = o[k++]; d.mass_max = o[k++]; d.mass_size_deviation = o[k++];
    d.rotor_time_constant_rising_min = o[k++]; d.rotor_time_constant_rising_max = o[k++];
    d.rotor_time_constant_falling_min = o[k++]; d.rotor_time_constant_falling_max = o[k++];
    d.rotor_torque_constant_min = o[k++]; d.rotor_torque_constant_max = o[k++];
    d.orientation_offset_angle_max = o[k++]; d.disturbance_force_max = o[k++];
    p.trajectory.mixture[0] = o[k++]; p.trajectory.mixture[1] = o[k++];
    p.trajectory.langevin.gamma = o[k++]; p.trajectory.langevin.omega = o[k++];
    p.trajectory.langevin.sigma = o[k++]; p.trajectory.langevin.alpha = o[k++];
}

template <typename S>
static void flatten_state(const S& s, float* o){
    constexpr int H = S::HISTORY_LENGTH;
    int k = 0;
    for(int i=0;i<3;i++) o[k++] = s.position[i];
    for(int i=0;i<4;i++) o[k++] = s.orientation[i];
    for(int i=0;i<3;i++) o[k++] = s.linear_velocity[i];
    for(int i=0;i<3;i++) o[k++] = s.angular_velocity[i];            // 13
    for(int i=0;i<4;i++) o[k++] = s.last_action[i];                 // 17
    for(int i=0;i<3;i++) o[k++] = s.angular_velocity_history[0][i]; // 20
    for(int i=0;i<3;i++) o[k++] = s.force[i];
    for(int i=0;i<3;i++) o[k++] = s.torque[i];                      // 26
    for(int i=0;i<4;i++) o[k++] = s.rpm[i];                         // 30
    o[k++] = (float)s.current_step;                                 // 31
    for(int h=0;h<H;h++) for(int i=0;i<4;i++) o[k++] = s.action_history[h][i];
    o[k++] = (float)(int)s.trajectory.type;
    for(int i=0;i<3;i++) o[k++] = s.trajectory.langevin.position[i];
    for(int i=0;i<3;i++) o[k++] = s.trajectory.langevin.velocity[i];
    for(int i=0;i<3;i++) o[k++] = s.trajectory.langevin.position_raw[i];
    for(int i=0;i<3;i++) o[k++] = s.trajectory.langevin.velocity_raw[i];
    if(k != state_dim(H)){ std::fprintf(stderr, "flatten_state: %d\n", k); std::abort(); }
}
template <typename S>
static void unflatten_state(const float* o, S& s){
    constexpr int H = S::HISTORY_LENGTH;
    int k = 0;
    for(int i=0;i<3;i++) s.position[i] = o[k++];
    for(int i=0;i<4;i++) s.orientation[i] = o[k++];
    for(int i=0;i<3;i++) s.linear_velocity[i] = o[k++];
    for(int i=0;i<3;i++) s.angular_velocity[i] = o[k++];
    for(int i=0;i<4;i++) s.last_action[i] = o[k++];
    for(int i=0;i<3;i++) s.angular_velocity_history[0][i] = o[k++];
    for(int i=0;i<3;i++) s.force[i] = o[k++];
    for(int i=0;i<3;i++) s.torque[i] = o[k++];
    for(int i=0;i<4;i++) s.rpm[i] = o[k++];
    s.current_step = (TI)o[k++];
    for(int h=0;h<H;h++) for(int i=0;i<4;i++) s.action_history[h][i] = o[k++];
    s.trajectory.type = (l2f::TrajectoryType)(int)o[k++];
    for(int i=0;i<3;i++) s.trajectory.langevin.position[i] = o[k++];
    for(int i=0;i<3;i++) s.trajectory.langevin.velocity[i] = o[k++];
    for(int i=0;i<3;i++) s.trajectory.langevin.position_raw[i] = o[k++];
    for(int i=0;i<3;i++) s.trajectory.langevin.velocity_raw[i] = o[k++];
}
template <typename S>
static void zero_state(S& s){ // the reference leaves trajectory fields uninitialised when LANGEVIN is off
    std::memset((void*)&s, 0, sizeof(S));
}

// ------------------------------------------------------------------------------------------------
// per-spec implementations
// ------------------------------------------------------------------------------------------------
template <typename ENV>
struct Impl{
    using State = typename ENV::State;
    using Parameters = typename ENV::Parameters;
    static constexpr int H = State::HISTORY_LENGTH;
    static constexpr int OBS = ENV::Observation::DIM;
    static void nominal_parameters(float* p){
        ENV env; DEVICE device;
        rlt::init(device, env);
        Parameters params;
        rlt::initial_parameters(device, env, params);
        flatten_parameters(params, p);
    }
    static void sample_initial_parameters(const float* env_p, uint64_t* rng_state, float* out){
        ENV env; DEVICE device; RNG rng; rng.state = *rng_state;
        unflatten_parameters(env_p, env.parameters);
        Parameters params;
        rlt::sample_initial_parameters(device, env, params, rng);
        flatten_parameters(params, out);
        *rng_state = rng.state;
    }
    static void initial_state(const float* p, float* s){
        ENV env; DEVICE device; Parameters params; State state; zero_state(state);
        unflatten_parameters(p, params);
        rlt::initial_state(device, env, params, state);
        flatten_state(state, s);
    }
    static void sample_initial_state(const float* p, uint64_t* rng_state, float* s){
        ENV env; DEVICE device; RNG rng; rng.state = *rng_state; Parameters params; State state; zero_state(state);
        unflatten_parameters(p, params);
        rlt::sample_initial_state(device, env, params, state, rng);
        flatten_state(state, s);
        *rng_state = rng.state;
    }
    static void observe(const float* p, const float* s, uint64_t* rng_state, float* obs){
        ENV env; DEVICE device; RNG rng; rng.state = *rng_state; Parameters params; State state; zero_state(state);
        unflatten_parameters(p, params); unflatten_state(s, state);
        rlt::Matrix<rlt::matrix::Specification<T, TI, 1, OBS, false>> o;
        rlt::observe(device, env, params, state, typename ENV::Observation{}, o, rng);
        for(int i=0;i<OBS;i++) obs[i] = rlt::get(o, 0, i);
        *rng_state = rng.state;
    }
    static float step(const float* p, const float* s, const float* a, uint64_t* rng_state, float* s_next){
        ENV env; DEVICE device; RNG rng; rng.state = *rng_state; Parameters params; State state, next; zero_state(state); zero_state(next);
        unflatten_parameters(p, params); unflatten_state(s, state);
        rlt::Matrix<rlt::matrix::Specification<T, TI, 1, 4, false>> action;
        for(int i=0;i<4;i++) rlt::set(action, 0, i, a[i]);
        T dt = rlt::step(device, env, params, state, action, next, rng);
        flatten_state(next, s_next);
        *rng_state = rng.state;
        return dt;
    }
    static float reward(const float* p, const float* s, const float* a, const float* s_next, uint64_t* rng_state){
        ENV env; DEVICE device; RNG rng; rng.state = *rng_state; Parameters params; State state, next; zero_state(state); zero_state(next);
        unflatten_parameters(p, params); unflatten_state(s, state); unflatten_state(s_next, next);
        rlt::Matrix<rlt::matrix::Specification<T, TI, 1, 4, false>> action;
        for(int i=0;i<4;i++) rlt::set(action, 0, i, a[i]);
        T r = rlt::reward(device, env, params, state, action, next, rng);
        *rng_state = rng.state;
        return r;
    }
    static int terminated(const float* p, const float* s){
        ENV env; DEVICE device; RNG rng; rng.state = 0; Parameters params; State state; zero_state(state);
        unflatten_parameters(p, params); unflatten_state(s, state);
        return rlt::terminated(device, env, params, state, rng) ? 1 : 0;
    }
};

// ------------------------------------------------------------------------------------------------
// Raptor policy (the reference's own checkpoint model, batch size 1 per environment)
// ------------------------------------------------------------------------------------------------
using ACTOR = rlt::checkpoint::actor::TYPE::template CHANGE_BATCH_SIZE<TI, 1>::template CHANGE_SEQUENCE_LENGTH<TI, 1>;
struct PolicyInstance{
    ACTOR::template State<false> state;
    ACTOR::template Buffer<false> buffer;
};
static void policy_reset(DEVICE& device, PolicyInstance& pi, RNG& rng){
    rlt::reset(device, rlt::checkpoint::actor::module, pi.state, rng);
}
template <bool NO_AUTO_RESET>
static void policy_evaluate_step(DEVICE& device, PolicyInstance& pi, const float* obs22, float* act4, RNG& rng){
    rlt::Tensor<rlt::tensor::Specification<T, TI, rlt::tensor::Shape<TI, 1, 22>, false>> input;
    rlt::Tensor<rlt::tensor::Specification<T, TI, rlt::tensor::Shape<TI, 1, 4>, false>> output;
    for(TI i=0;i<22;i++) rlt::set(device, input, obs22[i], 0, i);
    if constexpr(NO_AUTO_RESET){
        rlt::Mode<rlt::nn::layers::gru::NoAutoResetMode<rlt::mode::Evaluation<>>> mode;
        rlt::evaluate_step(device, rlt::checkpoint::actor::module, input, pi.state, output, pi.buffer, rng, mode);
    }
    else{
        rlt::Mode<rlt::mode::Evaluation<>> mode;
        rlt::evaluate_step(device, rlt::checkpoint::actor::module, input, pi.state, output, pi.buffer, rng, mode);
    }
    for(TI i=0;i<4;i++) act4[i] = rlt::get(device, output, 0, i);
}
static float* policy_hidden(PolicyInstance& pi){ // h [1,16] of the GRU layer
    return rlt::data(pi.state.content_state.next_content_state.state.state);
}
static TI& policy_step_counter(PolicyInstance& pi){
    return *rlt::data(pi.state.content_state.next_content_state.state.step);
}

// ------------------------------------------------------------------------------------------------
// closed-loop rollout in the order of rl_tools::evaluate (rl/utils/evaluation/operations_generic.h:138-189),
// per-environment RNG streams; no termination skip when flags&1 (plain README loop, R/README.md:94-99)
// ------------------------------------------------------------------------------------------------
struct RolloutOut{
    float* states;      // [T+1, N, STATE_DIM] or null
    float* observations;// [T, N, OBS] or null
    float* actions;     // [T, N, 4] or null
    float* rewards;     // [T, N] or null
    unsigned char* terminated; // [T, N] or null
    float* hidden;      // [N, 16] final hidden or null
};
template <typename ENV>
static void rollout_range(int n0, int n1, int N, int T_steps, const float* params, float* states_io, uint64_t* rng_states, float* hidden_io, int* gru_step_io, int no_auto_reset, RolloutOut out){
    using I = Impl<ENV>;
    constexpr int SD = state_dim(I::H);
    constexpr int OBS = I::OBS;
    DEVICE device;
    for(int n=n0;n<n1;n++){
        ENV env; rlt::init(device, env);
        typename ENV::Parameters p; unflatten_parameters(params + (size_t)n * PARAMS_DIM, p);
        typename ENV::State state, next; zero_state(state); zero_state(next);
        unflatten_state(states_io + (size_t)n * SD, state);
        RNG rng; rng.state = rng_states[n];
        PolicyInstance pi;
        policy_reset(device, pi, rng);
        if(hidden_io){ std::memcpy(policy_hidden(pi), hidden_io + (size_t)n*16, 16*sizeof(float)); }
        if(gru_step_io){ policy_step_counter(pi) = gru_step_io[n]; }
        rlt::Matrix<rlt::matrix::Specification<T, TI, 1, OBS, false>> o;
        rlt::Matrix<rlt::matrix::Specification<T, TI, 1, 4, false>> action;
        for(int t=0;t<T_steps;t++){
            if(out.states) flatten_state(state, out.states + ((size_t)t * N + n) * SD);
            rlt::observe(device, env, p, state, typename ENV::Observation{}, o, rng);
            float obs[OBS]; for(int i=0;i<OBS;i++) obs[i] = rlt::get(o, 0, i);
            if(out.observations) std::memcpy(out.observations + ((size_t)t * N + n) * OBS, obs, sizeof(obs));
            float a[4];
            if(no_auto_reset) policy_evaluate_step<true>(device, pi, obs, a, rng);
            else policy_evaluate_step<false>(device, pi, obs, a, rng);
            for(int i=0;i<4;i++) rlt::set(action, 0, i, a[i]);
            if(out.actions) std::memcpy(out.actions + ((size_t)t * N + n) * 4, a, sizeof(a));
            rlt::step(device, env, p, state, action, next, rng);
            T r = rlt::reward(device, env, p, state, action, next, rng);
            bool term = rlt::terminated(device, env, p, next, rng);
            if(out.rewards) out.rewards[(size_t)t * N + n] = r;
            if(out.terminated) out.terminated[(size_t)t * N + n] = term ? 1 : 0;
            state = next;
        }
        if(out.states) flatten_state(state, out.states + ((size_t)T_steps * N + n) * SD);
        flatten_state(state, states_io + (size_t)n * SD);
        rng_states[n] = rng.state;
        if(hidden_io) std::memcpy(hidden_io + (size_t)n*16, policy_hidden(pi), 16*sizeof(float));
        if(gru_step_io) gru_step_io[n] = (int)policy_step_counter(pi);
        if(out.hidden) std::memcpy(out.hidden + (size_t)n*16, policy_hidden(pi), 16*sizeof(float));
    }
}
template <typename ENV>
static void rollout(int N, int T_steps, int threads, const float* params, float* states_io, uint64_t* rng_states, float* hidden_io, int* gru_step_io, int no_auto_reset, RolloutOut out){
    if(threads <= 1){ rollout_range<ENV>(0, N, N, T_steps, params, states_io, rng_states, hidden_io, gru_step_io, no_auto_reset, out); return; }
    std::vector<std::thread> pool;
    for(int t=0;t<threads;t++){
        int n0 = (int)((long long)N * t / threads), n1 = (int)((long long)N * (t+1) / threads);
        pool.emplace_back([=](){ rollout_range<ENV>(n0, n1, N, T_steps, params, states_io, rng_states, hidden_io, gru_step_io, no_auto_reset, out); });
    }
    for(auto& th: pool) th.join();
}

#define DISPATCH(spec, CALL) \
    switch(spec){ \
        case 0: { using E = ENV_DEFAULT;    CALL; } break; \
        case 1: { using E = ENV_DEFAULT_DR; CALL; } break; \
        case 2: { using E = ENV_RAPTOR;     CALL; } break; \
        case 3: { using E = ENV_TEACHER;    CALL; } break; \
        case 4: { using E = ENV_RAPTOR_DR;  CALL; } break; \
        case 5: { using E = ENV_TEACHER_DR; CALL; } break; \
        default: std::fprintf(stderr, "ref_l2f: bad spec %d\n", spec); std::abort(); \
    }

extern "C" {
int ref_params_dim(){ return PARAMS_DIM; }
int ref_state_dim(int spec){ int r = 0; DISPATCH(spec, r = state_dim(Impl<E>::H)); return r; }
int ref_observation_dim(int spec){ int r = 0; DISPATCH(spec, r = Impl<E>::OBS); return r; }
int ref_action_history_length(int spec){ int r = 0; DISPATCH(spec, r = Impl<E>::H); return r; }
uint64_t ref_rng_init(uint64_t seed){ DEVICE device; RNG rng; rlt::init(device, rng, (TI)seed); return rng.state; }
float ref_rng_uniform(uint64_t* state, float lo, float hi){ DEVICE device; RNG rng; rng.state = *state; float v = rlt::random::uniform_real_distribution(device.random, lo, hi, rng); *state = rng.state; return v; }
float ref_rng_normal(uint64_t* state, float mean, float std){ DEVICE device; RNG rng; rng.state = *state; float v = rlt::random::normal_distribution::sample(device.random, mean, std, rng); *state = rng.state; return v; }
void ref_nominal_parameters(int spec, float* p){ DISPATCH(spec, Impl<E>::nominal_parameters(p)); }
void ref_sample_initial_parameters(int spec, const float* env_p, uint64_t* rng, float* out){ DISPATCH(spec, Impl<E>::sample_initial_parameters(env_p, rng, out)); }
void ref_initial_state(int spec, const float* p, float* s){ DISPATCH(spec, Impl<E>::initial_state(p, s)); }
void ref_sample_initial_state(int spec, const float* p, uint64_t* rng, float* s){ DISPATCH(spec, Impl<E>::sample_initial_state(p, rng, s)); }
void ref_observe(int spec, const float* p, const float* s, uint64_t* rng, float* obs){ DISPATCH(spec, Impl<E>::observe(p, s, rng, obs)); }
float ref_step(int spec, const float* p, const float* s, const float* a, uint64_t* rng, float* s_next){ float r = 0; DISPATCH(spec, r = Impl<E>::step(p, s, a, rng, s_next)); return r; }
float ref_reward(int spec, const float* p, const float* s, const float* a, const float* s_next, uint64_t* rng){ float r = 0; DISPATCH(spec, r = Impl<E>::reward(p, s, a, s_next, rng)); return r; }
int ref_terminated(int spec, const float* p, const float* s){ int r = 0; DISPATCH(spec, r = Impl<E>::terminated(p, s)); return r; }

// ---- Raptor policy ----
int ref_policy_num_parameters(){ return 16*22 + 16 + 48*16 + 48 + 48*16 + 48 + 16 + 4*16 + 4; }
// blob layout (include/b200_l2f.h, B200L2F_POLICY_RAPTOR_GRU): W1[16][22] b1[16] W_ih[48][16] b_ih[48] W_hh[48][16] b_hh[48] h0[16] W2[4][16] b2[4]
void ref_policy_export(float* blob){
    namespace a = rlt::checkpoint::actor;
    int k = 0;
    auto cp = [&](const unsigned char* mem, int n){ std::memcpy(blob + k, mem, n * sizeof(float)); k += n; };
    cp(a::layer_0::weights::parameters_memory::memory, 16*22);
    cp(a::layer_0::biases::parameters_memory::memory, 16);
    cp(a::layer_1::weights_input::parameters_memory::memory, 48*16);
    cp(a::layer_1::biases_input::parameters_memory::memory, 48);
    cp(a::layer_1::weights_hidden::parameters_memory::memory, 48*16);
    cp(a::layer_1::biases_hidden::parameters_memory::memory, 48);
    cp(a::layer_1::initial_hidden_state::parameters_memory::memory, 16);
    cp(a::layer_2::weights::parameters_memory::memory, 4*16);
    cp(a::layer_2::biases::parameters_memory::memory, 4);
}
void ref_policy_kat_shape(int* seq, int* batch){ *seq = 500; *batch = 2; }
void ref_policy_kat_export(float* input /*[500,2,22]*/, float* output /*[500,2,4]*/){
    std::memcpy(input, rlt::checkpoint::example::input::memory, sizeof(float) * 500*2*22);
    std::memcpy(output, rlt::checkpoint::example::output::memory, sizeof(float) * 500*2*4);
}
// replays the checkpoint's own known-answer test exactly like inference/applications/l2f/c_backend.h:54-83
float ref_policy_kat(float* max_abs){
    DEVICE device; RNG rng; rng.state = 1;
    double acc = 0; float mx = 0; size_t cnt = 0;
    for(int b=0;b<2;b++){
        PolicyInstance pi; policy_reset(device, pi, rng);
        for(int t=0;t<500;t++){
            const float* in = (const float*)rlt::checkpoint::example::input::memory + ((size_t)t*2 + b)*22;
            const float* ref = (const float*)rlt::checkpoint::example::output::memory + ((size_t)t*2 + b)*4;
            float a[4];
            policy_evaluate_step<false>(device, pi, in, a, rng);
            for(int i=0;i<4;i++){ float d = std::fabs(a[i]-ref[i]); acc += d; if(d>mx) mx = d; cnt++; }
        }
    }
    if(max_abs) *max_abs = mx;
    return (float)(acc / cnt);
}
// stateless single evaluate_step: h [N,16] and step counters [N] in/out
void ref_policy_evaluate_step(int N, const float* obs /*[N,22]*/, float* hidden /*[N,16]*/, int* gru_step /*[N]*/, int no_auto_reset, float* actions /*[N,4]*/){
    DEVICE device; RNG rng; rng.state = 1;
    for(int n=0;n<N;n++){
        PolicyInstance pi; policy_reset(device, pi, rng);
        std::memcpy(policy_hidden(pi), hidden + (size_t)n*16, 16*sizeof(float));
        policy_step_counter(pi) = gru_step[n];
        if(no_auto_reset) policy_evaluate_step<true>(device, pi, obs + (size_t)n*22, actions + (size_t)n*4, rng);
        else policy_evaluate_step<false>(device, pi, obs + (size_t)n*22, actions + (size_t)n*4, rng);
        std::memcpy(hidden + (size_t)n*16, policy_hidden(pi), 16*sizeof(float));
        gru_step[n] = (int)policy_step_counter(pi);
    }
}
void ref_policy_initial_hidden(float* h16){
    std::memcpy(h16, rlt::checkpoint::actor::layer_1::initial_hidden_state::parameters_memory::memory, 16*sizeof(float));
}

// ---- closed-loop rollout with the Raptor policy ----
void ref_rollout(int spec, int N, int T_steps, int threads, const float* params, float* states_io, uint64_t* rng_states, float* hidden_io, int* gru_step_io, int no_auto_reset,
                 float* out_states, float* out_observations, float* out_actions, float* out_rewards, unsigned char* out_terminated){
    RolloutOut out{out_states, out_observations, out_actions, out_rewards, out_terminated, nullptr};
    DISPATCH(spec, rollout<E>(N, T_steps, threads, params, states_io, rng_states, hidden_io, gru_step_io, no_auto_reset, out));
}
int ref_hardware_threads(){ return (int)std::thread::hardware_concurrency(); }
const char* ref_checkpoint_name(){ return rlt::checkpoint::meta::name; }
}

// ================================================================================================
// PPO data path: MLP actor / critic forward, rl_tools::collect's per-environment prologue / epilogue and
// rl_tools::estimate_generalized_advantages, all the reference's own templates.  Sizes that the reference fixes at
// compile time (N_ENVIRONMENTS, STEPS_PER_ENV, STEP_LIMIT) are fixed here at REF_PPO_N / REF_PPO_T / REF_PPO_STEP_LIMIT; the
// tests run the port and the GPU engine at exactly these sizes.
// ================================================================================================
#include <rl_tools/rl/components/on_policy_runner/operations_generic.h>
#include <rl_tools/rl/components/running_normalizer/operations_generic.h>
#include <rl_tools/rl/algorithms/ppo/ppo.h>
#include <rl_tools/rl/algorithms/ppo/operations_generic.h>

namespace ref_ppo{
    constexpr TI N = 32, STEPS = 48, STEP_LIMIT = 20;
    using CAP = rlt::nn::capability::Forward<true>;
    template <TI IN, TI OUT, TI BATCH>
    struct Net{ // rl/algorithms/ppo/loop/core/config.h:48-76 (actor and critic have the same structure: standardize -> mlp_unconditional_stddev)
        using INPUT_SHAPE = rlt::tensor::Shape<TI, 1, BATCH, IN>;
        using STANDARDIZATION_LAYER = rlt::nn::layers::standardize::BindConfiguration<rlt::nn::layers::standardize::Configuration<T, TI>>;
        using CONFIG = rlt::nn_models::mlp::Configuration<T, TI, OUT, 3, 64, rlt::nn::activation_functions::ActivationFunction::RELU, rlt::nn::activation_functions::IDENTITY>;
        using TYPE = rlt::nn_models::mlp_unconditional_stddev::BindConfiguration<CONFIG>;
        template <typename T_CONTENT, typename T_NEXT_MODULE = rlt::nn_models::sequential::OutputModule>
        using Module = typename rlt::nn_models::sequential::Module<T_CONTENT, T_NEXT_MODULE>;
        using MODEL = rlt::nn_models::sequential::Build<CAP, Module<STANDARDIZATION_LAYER, Module<TYPE>>, INPUT_SHAPE>;
    };
    // blob order of include/b200_l2f.h: [mean[in] precision[in]] W1 b1 W2 b2 W3 b3 [log_std[out]]
    template <typename MODEL, TI IN, TI OUT>
    static void load(DEVICE& device, MODEL& model, const float* blob, int has_std, int has_log_std){
        auto& stdz = rlt::get_first_layer(model);
        auto& mlp = rlt::get_last_layer(model);
        const float* b = blob;
        for(TI i = 0; i < IN; i++){
            rlt::set(stdz.mean.parameters, 0, i, has_std ? b[i] : 0.0f);
            rlt::set(stdz.precision.parameters, 0, i, has_std ? b[IN + i] : 1.0f);
        }
        if(has_std) b += 2 * IN;
        auto put = [&](auto& layer, TI out, TI in){
            for(TI o = 0; o < out; o++) for(TI i = 0; i < in; i++) rlt::set(layer.weights.parameters, o, i, *b++);
            for(TI o = 0; o < out; o++) rlt::set(layer.biases.parameters, 0, o, *b++);
        };
        put(mlp.input_layer, 64, IN);
        put(mlp.hidden_layers[0], 64, 64);
        put(mlp.output_layer, OUT, 64);
        for(TI o = 0; o < OUT; o++) rlt::set(mlp.log_std.parameters, 0, o, has_log_std ? b[o] : 0.0f);
    }
    template <TI IN, TI OUT>
    static void mlp_evaluate(const float* blob, int has_std, int n_rows, const float* in, int ld_in, float* out, int ld_out){
        using NET = Net<IN, OUT, 1>;
        DEVICE device; RNG rng; rng.state = 0;
        typename NET::MODEL model;
        typename NET::MODEL::template Buffer<> buffer;
        rlt::malloc(device, model); rlt::malloc(device, buffer);
        load<typename NET::MODEL, IN, OUT>(device, model, blob, has_std, 0);
        rlt::Tensor<rlt::tensor::Specification<T, TI, typename NET::INPUT_SHAPE>> x;
        rlt::Tensor<rlt::tensor::Specification<T, TI, rlt::tensor::Shape<TI, 1, 1, OUT>>> y;
        rlt::malloc(device, x); rlt::malloc(device, y);
        for(int r = 0; r < n_rows; r++){
            for(TI i = 0; i < IN; i++) rlt::set(device, x, in[(size_t)r * ld_in + i], 0, 0, i);
            rlt::evaluate(device, model, x, y, buffer, rng);
            for(TI o = 0; o < OUT; o++) out[(size_t)r * ld_out + o] = rlt::get(device, y, 0, 0, o);
        }
        rlt::free(device, x); rlt::free(device, y); rlt::free(device, model); rlt::free(device, buffer);
    }

    template <typename ENV>
    struct RunnerSpec{
        using OPR_SPEC = rlt::rl::components::on_policy_runner::Specification<T, TI, ENV, N, STEP_LIMIT>;
        using RUNNER = rlt::rl::components::OnPolicyRunner<OPR_SPEC>;
        using DATASET_SPEC = rlt::rl::components::on_policy_runner::DatasetSpecification<OPR_SPEC, STEPS>;
        using DATASET = rlt::rl::components::on_policy_runner::Dataset<DATASET_SPEC>;
    };
    // rl_tools::collect (on_policy_runner/operations_generic.h:99-131) with ONE RNG STREAM PER ENVIRONMENT: the loop below is the reference's
    // step loop, its per-environment prologue / epilogue are called with that environment's stream (the reference's multi-threaded collect does
    // the same, on_policy_runner/operations_cpu.h:36-44).
    template <typename ENV>
    static void collect(const float* actor_blob, int has_std, const float* env_params, float* params_io, float* states_io, uint64_t* rng_states,
                        int* episode_step_io, float* episode_return_io, unsigned char* truncated_io, float* dataset_out){
        using RS = RunnerSpec<ENV>;
        constexpr TI OBS = ENV::Observation::DIM;
        using NET = Net<OBS, 4, N>;
        DEVICE device;
        typename RS::RUNNER runner; typename RS::DATASET dataset;
        rlt::malloc(device, runner); rlt::malloc(device, dataset);
        typename NET::MODEL actor; typename NET::MODEL::template Buffer<> buffer;
        rlt::malloc(device, actor); rlt::malloc(device, buffer);
        load<typename NET::MODEL, OBS, 4>(device, actor, actor_blob, has_std, 1);
        const int SD = state_dim(Impl<ENV>::H);
        static ENV envs[N]; static typename ENV::Parameters params[N];
        for(TI e = 0; e < N; e++){
            rlt::init(device, envs[e]);
            unflatten_parameters(env_params, envs[e].parameters);
            unflatten_parameters(params_io + e * PARAMS_DIM, params[e]);
        }
        RNG rng0; rng0.state = 0;
        rlt::init(device, runner, envs, params, rng0);
        for(TI e = 0; e < N; e++){
            auto& st = rlt::get(runner.states, 0, e); zero_state(st); unflatten_state(states_io + e * SD, st);
            rlt::set(runner.episode_step, 0, e, (TI)episode_step_io[e]);
            rlt::set(runner.episode_return, 0, e, episode_return_io[e]);
            rlt::set(runner.truncated, 0, e, truncated_io[e] != 0);
        }
        std::vector<RNG> rngs(N);
        for(TI e = 0; e < N; e++) rngs[e].state = rng_states[e];
        rlt::set_all(device, dataset.data, 0);
        for(TI step_i = 0; step_i < STEPS; step_i++){
            auto actions_mean            = rlt::view(device, dataset.actions_mean               , rlt::matrix::ViewSpec<N, 4>()  , step_i*N, 0);
            auto actions                 = rlt::view(device, dataset.actions                    , rlt::matrix::ViewSpec<N, 4>()  , step_i*N, 0);
            auto observations_privileged = rlt::view(device, dataset.all_observations_privileged, rlt::matrix::ViewSpec<N, ENV::ObservationPrivileged::DIM>(), step_i*N, 0);
            auto observations            = rlt::view(device, dataset.observations               , rlt::matrix::ViewSpec<N, OBS>(), step_i*N, 0);
            for(TI e = 0; e < N; e++) rlt::rl::components::on_policy_runner::per_env::prologue(device, observations_privileged, observations, runner, rngs[e], e);
            typename NET::MODEL::template State<> actor_state;
            rlt::Mode<rlt::mode::Rollout<>> mode;
            auto observations_tensor = rlt::to_tensor(device, observations);
            auto actions_mean_tensor = rlt::to_tensor(device, actions_mean);
            rlt::evaluate_step(device, actor, observations_tensor, actor_state, actions_mean_tensor, buffer, rng0, mode);
            auto& last_layer = rlt::get_last_layer(actor);
            for(TI e = 0; e < N; e++) rlt::rl::components::on_policy_runner::per_env::epilogue(device, dataset, runner, actions_mean, actions, last_layer.log_std.parameters, rngs[e], step_i * N + e, e);
        }
        for(TI e = 0; e < N; e++){
            auto& env = rlt::get(runner.environments, 0, e);
            auto& state = rlt::get(runner.states, 0, e);
            auto& parameters = rlt::get(runner.env_parameters, 0, e);
            auto observation = rlt::row(device, dataset.all_observations_privileged, STEPS * N + e);
            rlt::observe(device, env, parameters, state, typename ENV::ObservationPrivileged{}, observation, rngs[e]);
        }
        constexpr TI D = RS::DATASET::DATA_DIM;
        for(TI r = 0; r < (STEPS + 1) * N; r++) for(TI c = 0; c < D; c++) dataset_out[r * D + c] = rlt::get(dataset.data, r, c);
        for(TI e = 0; e < N; e++){
            flatten_parameters(rlt::get(runner.env_parameters, 0, e), params_io + e * PARAMS_DIM);
            flatten_state(rlt::get(runner.states, 0, e), states_io + e * SD);
            rng_states[e] = rngs[e].state;
            episode_step_io[e] = (int)rlt::get(runner.episode_step, 0, e);
            episode_return_io[e] = rlt::get(runner.episode_return, 0, e);
            truncated_io[e] = rlt::get(runner.truncated, 0, e) ? 1 : 0;
        }
        rlt::free(device, runner); rlt::free(device, dataset); rlt::free(device, actor); rlt::free(device, buffer);
    }
    template <bool T_IGNORE_TERMINATION>
    struct PPOParameters: rlt::rl::algorithms::ppo::DefaultParameters<T, TI, 64>{ static constexpr bool IGNORE_TERMINATION = T_IGNORE_TERMINATION; };
    template <typename ENV>
    static void gae(float* data, int ignore_termination){
        using RS = RunnerSpec<ENV>;
        DEVICE device;
        typename RS::DATASET dataset;
        rlt::malloc(device, dataset);
        constexpr TI D = RS::DATASET::DATA_DIM;
        for(TI r = 0; r < (STEPS + 1) * N; r++) for(TI c = 0; c < D; c++) rlt::set(dataset.data, r, c, data[r * D + c]);
        if(ignore_termination) rlt::estimate_generalized_advantages(device, dataset, PPOParameters<true>{});
        else rlt::estimate_generalized_advantages(device, dataset, PPOParameters<false>{});
        for(TI r = 0; r < (STEPS + 1) * N; r++) for(TI c = 0; c < D; c++) data[r * D + c] = rlt::get(dataset.data, r, c);
        rlt::free(device, dataset);
    }
    // rl::components::running_normalizer update (operations_generic.h:27-49) on the dataset's observation columns
    template <typename ENV>
    static void normalizer_update(const float* data, float* mean_io, float* std_io, int* age_io){
        using RS = RunnerSpec<ENV>;
        constexpr TI OBS = ENV::Observation::DIM, D = RS::DATASET::DATA_DIM;
        DEVICE device;
        rlt::rl::components::RunningNormalizer<rlt::rl::components::running_normalizer::Specification<T, TI, OBS>> normalizer;
        rlt::malloc(device, normalizer);
        normalizer.age = *age_io;
        for(TI i = 0; i < OBS; i++){ rlt::set(normalizer.mean, 0, i, mean_io[i]); rlt::set(normalizer.std, 0, i, std_io[i]); }
        rlt::Matrix<rlt::matrix::Specification<T, TI, STEPS * N, OBS>> obs;
        rlt::malloc(device, obs);
        for(TI r = 0; r < STEPS * N; r++) for(TI c = 0; c < OBS; c++) rlt::set(obs, r, c, data[r * D + c]);
        rlt::update(device, normalizer, obs);
        *age_io = (int)normalizer.age;
        for(TI i = 0; i < OBS; i++){ mean_io[i] = rlt::get(normalizer.mean, 0, i); std_io[i] = rlt::get(normalizer.std, 0, i); }
        rlt::free(device, obs); rlt::free(device, normalizer);
    }
}

extern "C" {
void ref_ppo_sizes(int* n, int* steps, int* step_limit){ *n = ref_ppo::N; *steps = ref_ppo::STEPS; *step_limit = ref_ppo::STEP_LIMIT; }
float ref_ppo_gamma(){ return ref_ppo::PPOParameters<false>::GAMMA; }
float ref_ppo_lambda(){ return ref_ppo::PPOParameters<false>::LAMBDA; }
// standardize -> Dense(in,64,ReLU) -> Dense(64,64,ReLU) -> Dense(64,out): in/out pairs on the path (PPO actor 22/26 -> 4, critic 22/26 -> 1, SAC teacher 26 -> 8)
int ref_mlp_evaluate(int in_dim, int out_dim, const float* blob, int has_std, int n_rows, const float* in, int ld_in, float* out, int ld_out){
    if(in_dim == 22 && out_dim == 4) ref_ppo::mlp_evaluate<22, 4>(blob, has_std, n_rows, in, ld_in, out, ld_out);
    else if(in_dim == 26 && out_dim == 4) ref_ppo::mlp_evaluate<26, 4>(blob, has_std, n_rows, in, ld_in, out, ld_out);
    else if(in_dim == 22 && out_dim == 1) ref_ppo::mlp_evaluate<22, 1>(blob, has_std, n_rows, in, ld_in, out, ld_out);
    else if(in_dim == 26 && out_dim == 1) ref_ppo::mlp_evaluate<26, 1>(blob, has_std, n_rows, in, ld_in, out, ld_out);
    else if(in_dim == 26 && out_dim == 8) ref_ppo::mlp_evaluate<26, 8>(blob, has_std, n_rows, in, ld_in, out, ld_out);
    else if(in_dim == 82 && out_dim == 4) ref_ppo::mlp_evaluate<82, 4>(blob, has_std, n_rows, in, ld_in, out, ld_out);
    else if(in_dim == 82 && out_dim == 1) ref_ppo::mlp_evaluate<82, 1>(blob, has_std, n_rows, in, ld_in, out, ld_out);
    else return 1;
    return 0;
}
void ref_collect(int spec, const float* actor_blob, int has_std, const float* env_params, float* params_io, float* states_io, uint64_t* rng_states,
                 int* episode_step_io, float* episode_return_io, unsigned char* truncated_io, float* dataset){
    switch(spec){
        case 0: ref_ppo::collect<ENV_DEFAULT>(actor_blob, has_std, env_params, params_io, states_io, rng_states, episode_step_io, episode_return_io, truncated_io, dataset); break;   // the PPO zoo's env (rl/zoo/l2f/ppo.h)
        case 1: ref_ppo::collect<ENV_DEFAULT_DR>(actor_blob, has_std, env_params, params_io, states_io, rng_states, episode_step_io, episode_return_io, truncated_io, dataset); break;
        case 2: ref_ppo::collect<ENV_RAPTOR>(actor_blob, has_std, env_params, params_io, states_io, rng_states, episode_step_io, episode_return_io, truncated_io, dataset); break;
        case 4: ref_ppo::collect<ENV_RAPTOR_DR>(actor_blob, has_std, env_params, params_io, states_io, rng_states, episode_step_io, episode_return_io, truncated_io, dataset); break;
        case 5: ref_ppo::collect<ENV_TEACHER_DR>(actor_blob, has_std, env_params, params_io, states_io, rng_states, episode_step_io, episode_return_io, truncated_io, dataset); break;
        default: std::fprintf(stderr, "ref_collect: spec %d not instantiated\n", spec); std::abort();
    }
}
void ref_gae(int spec, float* dataset, int ignore_termination){
    switch(spec){
        case 0: case 1: ref_ppo::gae<ENV_DEFAULT>(dataset, ignore_termination); break;
        case 2: case 4: ref_ppo::gae<ENV_RAPTOR>(dataset, ignore_termination); break;
        case 3: case 5: ref_ppo::gae<ENV_TEACHER>(dataset, ignore_termination); break;
        default: std::fprintf(stderr, "ref_gae: spec %d not instantiated\n", spec); std::abort();
    }
}
void ref_normalizer_update(int spec, const float* dataset, float* mean_io, float* std_io, int* age_io){
    switch(spec){
        case 0: case 1: ref_ppo::normalizer_update<ENV_DEFAULT>(dataset, mean_io, std_io, age_io); break;
        case 2: case 4: ref_ppo::normalizer_update<ENV_RAPTOR>(dataset, mean_io, std_io, age_io); break;
        case 3: case 5: ref_ppo::normalizer_update<ENV_TEACHER>(dataset, mean_io, std_io, age_io); break;
        default: std::fprintf(stderr, "ref_normalizer_update: spec %d not instantiated\n", spec); std::abort();
    }
}
}

// ================================================================================================
// Foundation-policy DAgger data path: the reference's own add_to_dataset (src/foundation_policy/post_training/helper.h:43-110), called on
// rollout data recorded elsewhere (the engine / the port): teacher + student observations of every recorded state up to the first termination,
// episode compaction, truncated / reset flags and the teacher's action targets (MLP 26-64-64-8 + sample_and_squash in Evaluation mode).
// helper.h is included where it lies; its sample_trajectories is the reference's rollout (rl_tools::evaluate) and stays uninstantiated here.
// ================================================================================================
#include <iostream>
#include <sstream>
#include <rl_tools/rl/environments/l2f/operations_cpu.h>
#include <rl_tools/rl/utils/evaluation/operations_generic.h>
#include "post_training/helper.h"

namespace ref_dagger{
    constexpr TI N_EPISODES = 10;
    constexpr TI STEP_LIMIT = ENV_RAPTOR::EPISODE_STEP_LIMIT;
    constexpr TI DATASET_SIZE = N_EPISODES * STEP_LIMIT;
    static_assert(STEP_LIMIT == 500, "post_training/environment.h:14");
    using CAP = rlt::nn::capability::Forward<true>;
    struct Teacher{ // rl/algorithms/sac/loop/core/approximators_mlp.h:14-37 with pre_training/config.h:27-29 (3 layers, hidden 64, ReLU)
        using INPUT_SHAPE = rlt::tensor::Shape<TI, 1, 32, ENV_TEACHER::Observation::DIM>;
        using MLP_CONFIG = rlt::nn_models::mlp::Configuration<T, TI, 2 * 4, 3, 64, rlt::nn::activation_functions::ActivationFunction::RELU, rlt::nn::activation_functions::IDENTITY>;
        using MLP = rlt::nn_models::mlp::BindConfiguration<MLP_CONFIG>;
        using SAMPLE_AND_SQUASH_CONFIG = rlt::nn::layers::sample_and_squash::Configuration<T, TI, rlt::nn::layers::sample_and_squash::DefaultParameters<T>>;
        using SAMPLE_AND_SQUASH = rlt::nn::layers::sample_and_squash::BindConfiguration<SAMPLE_AND_SQUASH_CONFIG>;
        template <typename T_CONTENT, typename T_NEXT_MODULE = rlt::nn_models::sequential::OutputModule>
        using Module = typename rlt::nn_models::sequential::Module<T_CONTENT, T_NEXT_MODULE>;
        using MODEL = rlt::nn_models::sequential::Build<CAP, Module<MLP, Module<SAMPLE_AND_SQUASH>>, INPUT_SHAPE>;
    };
    using RESULT_SPEC = rlt::rl::utils::evaluation::Specification<T, TI, ENV_RAPTOR, N_EPISODES, STEP_LIMIT>;
    using DATA = rlt::rl::utils::evaluation::Data<rlt::rl::utils::evaluation::DataSpecification<RESULT_SPEC>>;
}

extern "C" {
void ref_dagger_sizes(int* n_episodes, int* step_limit){ *n_episodes = ref_dagger::N_EPISODES; *step_limit = ref_dagger::STEP_LIMIT; }
// params [10][145], states [10][500][48], terminated [10][500] (u8), teacher_blob: W1[64][26] b1 W2[64][64] b2 W3[8][64] b3, offset[3]
// -> episode_start [5000] (i32, the reference stores the running episode number at each episode's first row), input_student [5000][22],
//    output_target [5000][4], truncated / reset [5000] (u8); returns the number of rows added
int ref_dagger_add_to_dataset(const float* params, const float* states, const unsigned char* terminated, const float* teacher_blob, const float* offset,
                              int* episode_start, float* input_student, float* output_target, unsigned char* truncated_out, unsigned char* reset_out){
    using namespace ref_dagger;
    DEVICE device; RNG rng; rng.state = 0x1234;
    DATA data; rlt::malloc(device, data);
    const int SD = state_dim(1);
    for(TI e = 0; e < N_EPISODES; e++){
        ENV_RAPTOR::Parameters p; unflatten_parameters(params + e * PARAMS_DIM, p);
        rlt::set(device, data.parameters, p, e);
        for(TI t = 0; t < STEP_LIMIT; t++){
            ENV_RAPTOR::State s; zero_state(s); unflatten_state(states + ((size_t)e * STEP_LIMIT + t) * SD, s);
            rlt::set(device, data.states, s, e, t);
            rlt::set(device, data.terminated, terminated[e * STEP_LIMIT + t] != 0, e, t);
        }
    }
    Teacher::MODEL teacher; rlt::malloc(device, teacher);
    {
        auto& mlp = rlt::get_first_layer(teacher);
        const float* b = teacher_blob;
        auto put = [&](auto& layer, TI out, TI in){
            for(TI o = 0; o < out; o++) for(TI i = 0; i < in; i++) rlt::set(layer.weights.parameters, o, i, *b++);
            for(TI o = 0; o < out; o++) rlt::set(layer.biases.parameters, 0, o, *b++);
        };
        put(mlp.input_layer, 64, 26); put(mlp.hidden_layers[0], 64, 64); put(mlp.output_layer, 8, 64);
    }
    TeacherMeta<T> meta; for(int i = 0; i < 3; i++) meta.steady_state_position_offset[i] = offset[i];
    rlt::Tensor<rlt::tensor::Specification<TI, TI, rlt::tensor::Shape<TI, DATASET_SIZE>>> ds_start;
    rlt::Tensor<rlt::tensor::Specification<T, TI, rlt::tensor::Shape<TI, DATASET_SIZE, 22>>> ds_in;
    rlt::Tensor<rlt::tensor::Specification<T, TI, rlt::tensor::Shape<TI, DATASET_SIZE, 4>>> ds_out;
    rlt::Tensor<rlt::tensor::Specification<bool, TI, rlt::tensor::Shape<TI, DATASET_SIZE>>> ds_trunc, ds_reset;
    rlt::malloc(device, ds_start); rlt::malloc(device, ds_in); rlt::malloc(device, ds_out); rlt::malloc(device, ds_trunc); rlt::malloc(device, ds_reset);
    rlt::set_all(device, ds_start, (TI)0); rlt::set_all(device, ds_in, (T)0); rlt::set_all(device, ds_out, (T)0); rlt::set_all(device, ds_trunc, false); rlt::set_all(device, ds_reset, false);
    TI current_episode = 0, current_index = 0;
    TI added = add_to_dataset<ENV_RAPTOR, ENV_TEACHER::Observation, ENV_RAPTOR::Observation, true>(device, data, teacher, meta, ds_start, ds_in, ds_out, ds_trunc, ds_reset, current_episode, current_index, rng);
    for(TI r = 0; r < DATASET_SIZE; r++){
        episode_start[r] = (int)rlt::get(device, ds_start, r);
        for(TI c = 0; c < 22; c++) input_student[r * 22 + c] = rlt::get(device, ds_in, r, c);
        for(TI c = 0; c < 4; c++) output_target[r * 4 + c] = rlt::get(device, ds_out, r, c);
        truncated_out[r] = rlt::get(device, ds_trunc, r) ? 1 : 0;
        reset_out[r] = rlt::get(device, ds_reset, r) ? 1 : 0;
    }
    rlt::free(device, data); rlt::free(device, teacher); rlt::free(device, ds_start); rlt::free(device, ds_in); rlt::free(device, ds_out); rlt::free(device, ds_trunc); rlt::free(device, ds_reset);
    return (int)added;
}
}

// ================================================================================================
// Parameter / state JSON wire format: the reference's own json(...) / from_json(...) (rl/environments/l2f/operations_cpu.h:139-411, 412-560,
// 565-824; nlohmann::json comes from the image's cudnn_frontend/thirdparty tree, oracle/Makefile)
// ================================================================================================
#ifdef RL_TOOLS_ENABLE_JSON
template <typename ENV>
static int params_to_json(const float* row, char* buf, int cap){
    DEVICE device; ENV env; rlt::init(device, env);
    typename ENV::Parameters p; unflatten_parameters(row, p);
    std::string s = rlt::json(device, env, p);
    if((int)s.size() + 1 > cap) return -(int)s.size() - 1;
    std::memcpy(buf, s.c_str(), s.size() + 1);
    return (int)s.size();
}
template <typename ENV>
static void params_from_json(const char* json, float* row_io){
    DEVICE device; ENV env; rlt::init(device, env);
    typename ENV::Parameters p; unflatten_parameters(row_io, p);
    rlt::from_json(device, env, std::string(json), p);
    flatten_parameters(p, row_io);
}
template <typename ENV>
static int state_to_json(const float* prow, const float* srow, char* buf, int cap){
    DEVICE device; ENV env; rlt::init(device, env);
    typename ENV::Parameters p; unflatten_parameters(prow, p);
    typename ENV::State s; zero_state(s); unflatten_state(srow, s);
    std::string js = rlt::json(device, env, p, s);
    if((int)js.size() + 1 > cap) return -(int)js.size() - 1;
    std::memcpy(buf, js.c_str(), js.size() + 1);
    return (int)js.size();
}
template <typename ENV>
static void state_from_json(const float* prow, const char* json, float* srow_io){
    DEVICE device; ENV env; rlt::init(device, env);
    typename ENV::Parameters p; unflatten_parameters(prow, p);
    typename ENV::State s; zero_state(s); unflatten_state(srow_io, s);
    rlt::from_json(device, env, p, std::string(json), s);
    flatten_state(s, srow_io);
}
extern "C" {
int ref_json_available(){ return 1; }
int ref_parameters_to_json(int spec, const float* row, char* buf, int cap){ int r = 0; DISPATCH(spec, r = params_to_json<E>(row, buf, cap)); return r; }
void ref_parameters_from_json(int spec, const char* json, float* row_io){ DISPATCH(spec, params_from_json<E>(json, row_io)); }
int ref_state_to_json(int spec, const float* prow, const float* srow, char* buf, int cap){ int r = 0; DISPATCH(spec, r = state_to_json<E>(prow, srow, buf, cap)); return r; }
void ref_state_from_json(int spec, const float* prow, const char* json, float* srow_io){ DISPATCH(spec, state_from_json<E>(prow, json, srow_io)); }
}
#else
extern "C" { int ref_json_available(){ return 0; } }
#endif

// ================================================================================================
// Checkpoint code export: the reference's own rl_tools::save_code (containers/{matrix,tensor}/persist_code.h, nn/layers/*/persist_code.h,
// nn_models/{mlp,mlp_unconditional_stddev,sequential}/persist_code.h) on the three actor shapes of the path, assembled the way
// rl::loop::steps::checkpoint::save_code does (rl/loop/steps/checkpoint/operations_cpu.h:56-84: actor, example input / output, meta).
// Produces the fixtures of tests/golden/checkpoints/ and pins the engine's reader (raptor_b200/csrc/checkpoint_io.cu).
// ================================================================================================
#include <rl_tools/containers/matrix/persist_code.h>
#include <rl_tools/containers/tensor/persist_code.h>
#include <rl_tools/nn/parameters/persist_code.h>
#include <rl_tools/nn/layers/dense/persist_code.h>
#include <rl_tools/nn/layers/gru/persist_code.h>
#include <rl_tools/nn/layers/standardize/persist_code.h>
#include <rl_tools/nn/layers/sample_and_squash/persist_code.h>
#include <rl_tools/nn_models/mlp/persist_code.h>
#include <rl_tools/nn_models/mlp_unconditional_stddev/persist_code.h>
#include <rl_tools/nn_models/sequential/persist_code.h>

namespace ref_export{
    template <typename MODEL>
    static std::string with_example(DEVICE& device, MODEL& model, const char* name, uint64_t seed){
        RNG rng; rng.state = seed;
        std::stringstream ss;
        ss << rlt::save_code(device, model, std::string("rl_tools::checkpoint::actor"), true);
        rlt::Tensor<rlt::tensor::Specification<T, TI, typename MODEL::INPUT_SHAPE>> input;
        rlt::Tensor<rlt::tensor::Specification<T, TI, typename MODEL::OUTPUT_SHAPE>> output;
        typename MODEL::template Buffer<> buffer;
        rlt::malloc(device, input); rlt::malloc(device, output); rlt::malloc(device, buffer);
        rlt::randn(device, input, rng);
        rlt::Mode<rlt::mode::Evaluation<>> mode;
        rlt::evaluate(device, model, input, output, buffer, rng, mode);
        ss << "\n" << rlt::save_code(device, input, std::string("rl_tools::checkpoint::example::input"), true);
        ss << "\n" << rlt::save_code(device, output, std::string("rl_tools::checkpoint::example::output"), true);
        ss << "\n" << "namespace rl_tools::checkpoint::meta{";
        ss << "\n" << "   " << "char name[] = \"" << name << "\";";
        ss << "\n" << "   " << "char commit_hash[] = \"" << "fixture" << "\";";
        ss << "\n" << "}";
        rlt::free(device, input); rlt::free(device, output); rlt::free(device, buffer);
        return ss.str();
    }
    static int emit(const std::string& s, char* buf, int cap){
        if((int)s.size() + 1 > cap) return -(int)s.size() - 1;
        std::memcpy(buf, s.c_str(), s.size() + 1);
        return (int)s.size();
    }
}
extern "C" {
// (the Dense-GRU-Dense export is the Raptor checkpoint file itself: its tensors are `const`, which save_code cannot take)
// kind 1: SAC teacher MLP 26-64-64-8 + sample_and_squash from `blob` (6408 floats);
// kind 2: PPO actor standardize -> mlp_unconditional_stddev 22-64-64-4 from `blob` ([mean precision] W1 b1 W2 b2 W3 b3 log_std[4]).
// returns the text length, or -(needed capacity) if cap is too small
int ref_save_code(int kind, const float* blob, int has_std, const char* name, char* buf, int cap){
    DEVICE device;
    if(kind == 1){
        ref_dagger::Teacher::MODEL teacher; rlt::malloc(device, teacher);
        auto& mlp = rlt::get_first_layer(teacher);
        const float* b = blob;
        auto put = [&](auto& layer, TI out, TI in){
            for(TI o = 0; o < out; o++) for(TI i = 0; i < in; i++) rlt::set(layer.weights.parameters, o, i, *b++);
            for(TI o = 0; o < out; o++) rlt::set(layer.biases.parameters, 0, o, *b++);
        };
        put(mlp.input_layer, 64, 26); put(mlp.hidden_layers[0], 64, 64); put(mlp.output_layer, 8, 64);
        int r = ref_export::emit(ref_export::with_example(device, teacher, name, 0xC0DE + 1), buf, cap);
        rlt::free(device, teacher);
        return r;
    }
    if(kind == 2){
        using NET = ref_ppo::Net<22, 4, 8>;
        typename NET::MODEL model; rlt::malloc(device, model);
        ref_ppo::load<typename NET::MODEL, 22, 4>(device, model, blob, has_std, 1);
        int r = ref_export::emit(ref_export::with_example(device, model, name, 0xC0DE + 2), buf, cap);
        rlt::free(device, model);
        return r;
    }
    return 0;
}
}

// ================================================================================================
// Off-policy runner (SAC teacher data collection): the reference's OffPolicyRunner, its own prologue_per_env / epilogue_per_env
// (rl/components/off_policy_runner/operations_generic_per_env.h:8-110) and replay buffer `add`, driven with ONE RNG STREAM PER ENVIRONMENT
// exactly as the reference's CUDA kernels drive them (operations_cuda.h:62-106: `get(rng.states, 0, env_i)`); the interlude is the
// reference's evaluate_step of the SAC actor (MLP 26-64-64-8 + sample_and_squash) in Mode<Rollout> (operations_generic.h:190-204), per
// environment with that environment's stream.
// ================================================================================================
#include <rl_tools/rl/components/replay_buffer/operations_generic.h>
#include <rl_tools/rl/components/off_policy_runner/operations_generic.h>

namespace ref_opr{
    constexpr TI N = 16, STEPS = 60, STEP_LIMIT = 20, CAPACITY = 48;   // the ring wraps (60 > 48); episodes end by the step limit and by termination
    using CAP = rlt::nn::capability::Forward<true>;
    struct Actor{ // rl/algorithms/sac/loop/core/approximators_mlp.h:14-37
        using INPUT_SHAPE = rlt::tensor::Shape<TI, 1, 1, ENV_TEACHER::Observation::DIM>;
        using MLP_CONFIG = rlt::nn_models::mlp::Configuration<T, TI, 2 * 4, 3, 64, rlt::nn::activation_functions::ActivationFunction::RELU, rlt::nn::activation_functions::IDENTITY>;
        using MLP = rlt::nn_models::mlp::BindConfiguration<MLP_CONFIG>;
        using SAMPLE_AND_SQUASH_CONFIG = rlt::nn::layers::sample_and_squash::Configuration<T, TI, rlt::nn::layers::sample_and_squash::DefaultParameters<T>>;
        using SAMPLE_AND_SQUASH = rlt::nn::layers::sample_and_squash::BindConfiguration<SAMPLE_AND_SQUASH_CONFIG>;
        template <typename T_CONTENT, typename T_NEXT_MODULE = rlt::nn_models::sequential::OutputModule>
        using Module = typename rlt::nn_models::sequential::Module<T_CONTENT, T_NEXT_MODULE>;
        using MODEL = rlt::nn_models::sequential::Build<CAP, Module<MLP, Module<SAMPLE_AND_SQUASH>>, INPUT_SHAPE>;
    };
    template <typename ENV, bool SAMPLE, TI CAP = CAPACITY>
    struct RunnerSpec{
        struct PARAMETERS: rlt::rl::components::off_policy_runner::ParametersDefault<T, TI>{
            static constexpr TI N_ENVIRONMENTS = N;
            static constexpr bool ASYMMETRIC_OBSERVATIONS = false;
            static constexpr TI REPLAY_BUFFER_CAPACITY = CAP;
            static constexpr TI EPISODE_STEP_LIMIT = STEP_LIMIT;
            static constexpr bool SAMPLE_PARAMETERS = SAMPLE;
        };
        using POLICIES = rlt::utils::Tuple<TI, Actor::MODEL>;
        using SPEC = rlt::rl::components::off_policy_runner::Specification<T, TI, ENV, POLICIES, PARAMETERS, true>;
        using RUNNER = rlt::rl::components::OffPolicyRunner<SPEC>;
    };
    template <typename ENV, bool SAMPLE>
    static void run(const float* actor_blob, const float* env_params, float* params_io, float* states_io, uint64_t* rng_states,
                    int* episode_step_io, float* episode_return_io, unsigned char* truncated_io,
                    float* replay, int* episode_start, int* position_io, unsigned char* full_io, int* current_episode_start_io, float* states_out, float* next_states_out){
        using RS = RunnerSpec<ENV, SAMPLE>;
        using RUNNER = typename RS::RUNNER;
        constexpr TI OBS = ENV::Observation::DIM;
        constexpr TI D = RUNNER::REPLAY_BUFFER_TYPE::DATA_COLS;
        static_assert(D == 2 * OBS + 7, "symmetric replay row");
        DEVICE device;
        auto* runner_ptr = new RUNNER(); RUNNER& runner = *runner_ptr;
        rlt::malloc(device, runner);
        rlt::init(device, runner);
        Actor::MODEL actor; Actor::MODEL::template Buffer<> buffer; Actor::MODEL::template State<> actor_state;
        rlt::malloc(device, actor); rlt::malloc(device, buffer); rlt::malloc(device, actor_state);
        {
            auto& mlp = rlt::get_first_layer(actor);
            const float* b = actor_blob;
            auto put = [&](auto& layer, TI out, TI in){
                for(TI o = 0; o < out; o++) for(TI i = 0; i < in; i++) rlt::set(layer.weights.parameters, o, i, *b++);
                for(TI o = 0; o < out; o++) rlt::set(layer.biases.parameters, 0, o, *b++);
            };
            put(mlp.input_layer, 64, OBS); put(mlp.hidden_layers[0], 64, 64); put(mlp.output_layer, 8, 64);
        }
        const int SD = state_dim(Impl<ENV>::H);
        std::vector<RNG> rngs(N);
        for(TI e = 0; e < N; e++){
            auto& env = rlt::get(runner.envs, 0, e);
            unflatten_parameters(env_params, env.parameters);
            unflatten_parameters(params_io + e * PARAMS_DIM, rlt::get(runner.env_parameters, 0, e));
            auto& st = rlt::get(runner.states, 0, e); zero_state(st); unflatten_state(states_io + e * SD, st);
            rlt::set(runner.episode_step, 0, e, (TI)episode_step_io[e]);
            rlt::set(runner.episode_return, 0, e, episode_return_io[e]);
            rlt::set(runner.truncated, 0, e, truncated_io[e] != 0);
            auto& rb = rlt::get(runner.replay_buffers, 0, e);
            for(TI r = 0; r < CAPACITY; r++){
                for(TI c = 0; c < D; c++) rlt::set(rb.data, r, c, replay[(e * CAPACITY + r) * D + c]);
                rlt::set(device, rb.episode_start, (TI)episode_start[e * CAPACITY + r], r);
            }
            rb.position = position_io[e]; rb.full = full_io[e] != 0; rb.current_episode_start = current_episode_start_io[e];
            rngs[e].state = rng_states[e];
        }
        for(TI step_i = 0; step_i < STEPS; step_i++){
            for(TI e = 0; e < N; e++) rlt::rl::components::off_policy_runner::prologue_per_env(device, runner, rngs[e], e);
            for(TI e = 0; e < N; e++){
                auto obs_row = rlt::row(device, runner.buffers.observations, e);
                auto act_row = rlt::row(device, runner.buffers.actions, e);
                auto obs_tensor = rlt::to_tensor(device, obs_row);
                auto act_tensor = rlt::to_tensor(device, act_row);
                rlt::Mode<rlt::mode::Rollout<>> mode;
                rlt::evaluate_step(device, actor, obs_tensor, actor_state, act_tensor, buffer, rngs[e], mode);
            }
            for(TI e = 0; e < N; e++) rlt::rl::components::off_policy_runner::epilogue_per_env(device, runner, actor, rngs[e], e);
        }
        for(TI e = 0; e < N; e++){
            flatten_parameters(rlt::get(runner.env_parameters, 0, e), params_io + e * PARAMS_DIM);
            flatten_state(rlt::get(runner.states, 0, e), states_io + e * SD);
            rng_states[e] = rngs[e].state;
            episode_step_io[e] = (int)rlt::get(runner.episode_step, 0, e);
            episode_return_io[e] = rlt::get(runner.episode_return, 0, e);
            truncated_io[e] = rlt::get(runner.truncated, 0, e) ? 1 : 0;
            auto& rb = rlt::get(runner.replay_buffers, 0, e);
            for(TI r = 0; r < CAPACITY; r++){
                for(TI c = 0; c < D; c++) replay[(e * CAPACITY + r) * D + c] = rlt::get(rb.data, r, c);
                episode_start[e * CAPACITY + r] = (int)rlt::get(device, rb.episode_start, r);
                if(states_out) flatten_state(rlt::get(rb.states, r, 0), states_out + ((size_t)e * CAPACITY + r) * SD);
                if(next_states_out) flatten_state(rlt::get(rb.next_states, r, 0), next_states_out + ((size_t)e * CAPACITY + r) * SD);
            }
            position_io[e] = (int)rb.position; full_io[e] = rb.full ? 1 : 0; current_episode_start_io[e] = (int)rb.current_episode_start;
        }
        rlt::free(device, runner); rlt::free(device, actor); rlt::free(device, buffer); rlt::free(device, actor_state);
        delete runner_ptr;
    }
}
namespace ref_opr{
    constexpr TI BATCH = 64;
    struct BATCH_PARAMETERS{   // pre_training/config.h:52-58
        static constexpr bool INCLUDE_FIRST_STEP_IN_TARGETS = false;
        static constexpr bool ALWAYS_SAMPLE_FROM_INITIAL_STATE = false;
        static constexpr bool RANDOM_SEQ_LENGTH = false;
        static constexpr bool ENABLE_NOMINAL_SEQUENCE_LENGTH_PROBABILITY = true;
        static constexpr T NOMINAL_SEQUENCE_LENGTH_PROBABILITY = 0.1;
    };
    // the reference's gather_batch_step (operations_generic.h:240-420) on its own SequentialBatch with the MLP SAC settings (SEQUENCE_LENGTH 1,
    // pre_training/config.h:52-58), one RNG stream per batch sample and the environment drawn from that stream first (operations_cuda.h:36-60)
    template <typename ENV>
    static void gather(const float* replay, const int* position_in, const unsigned char* full_in, uint64_t* rng_states, float* observations_actions, float* rewards,
                       unsigned char* terminated, unsigned char* reset, unsigned char* next_reset, unsigned char* final_step_mask, unsigned char* next_final_step_mask){
        using RS = RunnerSpec<ENV, true>;
        using RUNNER = typename RS::RUNNER;
        using BATCH_SPEC = rlt::rl::components::off_policy_runner::SequentialBatchSpecification<typename RS::SPEC, 1, BATCH, BATCH_PARAMETERS, true>;
        constexpr TI OBS = ENV::Observation::DIM, D = RUNNER::REPLAY_BUFFER_TYPE::DATA_COLS, W = OBS + 4;
        DEVICE device;
        auto* runner_ptr = new RUNNER(); RUNNER& runner = *runner_ptr;
        rlt::malloc(device, runner);
        rlt::init(device, runner);
        rlt::rl::components::off_policy_runner::SequentialBatch<BATCH_SPEC> batch;
        rlt::malloc(device, batch);
        for(TI e = 0; e < N; e++){
            auto& rb = rlt::get(runner.replay_buffers, 0, e);
            for(TI r = 0; r < CAPACITY; r++) for(TI c = 0; c < D; c++) rlt::set(rb.data, r, c, replay[(e * CAPACITY + r) * D + c]);
            rb.position = position_in[e]; rb.full = full_in[e] != 0;
        }
        for(TI b = 0; b < BATCH; b++){
            RNG rng; rng.state = rng_states[b];
            TI env_i = rlt::random::uniform_int_distribution(typename DEVICE::SPEC::RANDOM(), (TI)0, (TI)(N - 1), rng);
            auto& rb = rlt::get(runner.replay_buffers, 0, env_i);
            rlt::gather_batch_step<false>(device, runner, rb, batch, b, rng);
            rng_states[b] = rng.state;
        }
        for(TI s = 0; s < 2; s++) for(TI b = 0; b < BATCH; b++){
            for(TI c = 0; c < W; c++) observations_actions[(s * BATCH + b) * W + c] = rlt::get(device, batch.observations_actions_base, s, b, c);
            next_reset[s * BATCH + b] = rlt::get(device, batch.next_reset_base, s, b, 0) ? 1 : 0;
            next_final_step_mask[s * BATCH + b] = rlt::get(device, batch.next_final_step_mask_base, s, b, 0) ? 1 : 0;
        }
        for(TI b = 0; b < BATCH; b++){
            rewards[b] = rlt::get(device, batch.rewards, 0, b, 0);
            terminated[b] = rlt::get(device, batch.terminated, 0, b, 0) ? 1 : 0;
            reset[b] = rlt::get(device, batch.reset, 0, b, 0) ? 1 : 0;
            final_step_mask[b] = rlt::get(device, batch.final_step_mask, 0, b, 0) ? 1 : 0;
        }
        rlt::free(device, batch); rlt::free(device, runner);
        delete runner_ptr;
    }
}
namespace ref_opr{
    // ---- any SEQUENCE_LENGTH: the reference's gather_batch_step on its own SequentialBatch, for a handful of compile-time parameter sets
    // (off_policy_runner.h:78-85 defaults for SEQUENCE_LENGTH > 1 and variations of every switch)
    template <bool FIRST, bool INITIAL, bool RANDOM, bool ENABLE, int PROB_PERCENT>
    struct SeqParameters{
        static constexpr bool INCLUDE_FIRST_STEP_IN_TARGETS = FIRST;
        static constexpr bool ALWAYS_SAMPLE_FROM_INITIAL_STATE = INITIAL;
        static constexpr bool RANDOM_SEQ_LENGTH = RANDOM;
        static constexpr bool ENABLE_NOMINAL_SEQUENCE_LENGTH_PROBABILITY = ENABLE;
        static constexpr T NOMINAL_SEQUENCE_LENGTH_PROBABILITY = (T)PROB_PERCENT / (T)100;
    };
    template <TI L, typename PARAMS, TI CAP>
    static void gather_sequential(const float* replay, const int* episode_start, const int* position_in, const unsigned char* full_in, uint64_t* rng_states,
                                  float* observations_actions, float* rewards, unsigned char* terminated, unsigned char* reset, unsigned char* next_reset_base,
                                  unsigned char* final_step_mask, unsigned char* next_final_step_mask_base){
        using ENV = ENV_TEACHER;
        using RS = RunnerSpec<ENV, true, CAP>;
        using RUNNER = typename RS::RUNNER;
        constexpr TI CAPACITY = CAP;
        using BATCH_SPEC = rlt::rl::components::off_policy_runner::SequentialBatchSpecification<typename RS::SPEC, L, BATCH, PARAMS, true>;
        constexpr TI OBS = ENV::Observation::DIM, D = RUNNER::REPLAY_BUFFER_TYPE::DATA_COLS, W = OBS + 4, P = L + 1;
        DEVICE device;
        auto* runner_ptr = new RUNNER(); RUNNER& runner = *runner_ptr;
        rlt::malloc(device, runner);
        rlt::init(device, runner);
        rlt::rl::components::off_policy_runner::SequentialBatch<BATCH_SPEC> batch;
        rlt::malloc(device, batch);
        rlt::set_all(device, batch.observations_actions_base, 0); rlt::set_all(device, batch.rewards, 0); rlt::set_all(device, batch.terminated, false);
        for(TI e = 0; e < N; e++){
            auto& rb = rlt::get(runner.replay_buffers, 0, e);
            for(TI r = 0; r < CAPACITY; r++){
                for(TI c = 0; c < D; c++) rlt::set(rb.data, r, c, replay[(e * CAPACITY + r) * D + c]);
                rlt::set(device, rb.episode_start, (TI)episode_start[e * CAPACITY + r], r);
            }
            rb.position = position_in[e]; rb.full = full_in[e] != 0;
        }
        for(TI b = 0; b < BATCH; b++){
            RNG rng; rng.state = rng_states[b];
            TI env_i = rlt::random::uniform_int_distribution(typename DEVICE::SPEC::RANDOM(), (TI)0, (TI)(N - 1), rng);
            auto& rb = rlt::get(runner.replay_buffers, 0, env_i);
            rlt::gather_batch_step<false>(device, runner, rb, batch, b, rng);
            rng_states[b] = rng.state;
        }
        for(TI s = 0; s < P; s++) for(TI b = 0; b < BATCH; b++){
            for(TI c = 0; c < W; c++) observations_actions[(s * BATCH + b) * W + c] = rlt::get(device, batch.observations_actions_base, s, b, c);
            next_reset_base[s * BATCH + b] = rlt::get(device, batch.next_reset_base, s, b, 0) ? 1 : 0;
            next_final_step_mask_base[s * BATCH + b] = rlt::get(device, batch.next_final_step_mask_base, s, b, 0) ? 1 : 0;
        }
        for(TI s = 0; s < L; s++) for(TI b = 0; b < BATCH; b++){
            rewards[s * BATCH + b] = rlt::get(device, batch.rewards, s, b, 0);
            terminated[s * BATCH + b] = rlt::get(device, batch.terminated, s, b, 0) ? 1 : 0;
            reset[s * BATCH + b] = rlt::get(device, batch.reset, s, b, 0) ? 1 : 0;
            final_step_mask[s * BATCH + b] = rlt::get(device, batch.final_step_mask, s, b, 0) ? 1 : 0;
        }
        rlt::free(device, batch); rlt::free(device, runner);
        delete runner_ptr;
    }
    // sampling from the initial state needs CAPACITY >= MAX_EPISODE_LENGTH = ENVIRONMENT::EPISODE_STEP_LIMIT (operations_generic.h:258; 500 here): rings of 640 rows
    struct SeqConfig{ int L, first, initial, random, enable, prob_percent, capacity; };
    static const SeqConfig SEQ_CONFIGS[] = {{8, 1, 0, 0, 1, 50, 48}, {8, 1, 1, 1, 1, 50, 640}, {8, 0, 0, 1, 0, 50, 48}, {24, 1, 1, 1, 1, 10, 640}, {2, 1, 1, 1, 1, 50, 640}, {8, 0, 1, 0, 1, 50, 640}};
}
extern "C" {
int ref_gather_batch_sequential_configs(){ return (int)(sizeof(ref_opr::SEQ_CONFIGS) / sizeof(ref_opr::SEQ_CONFIGS[0])); }
void ref_gather_batch_sequential_config(int config, int* L, int* first, int* initial, int* random, int* enable, float* probability, int* capacity){
    const auto& c = ref_opr::SEQ_CONFIGS[config];
    *capacity = c.capacity;
    *L = c.L; *first = c.first; *initial = c.initial; *random = c.random; *enable = c.enable; *probability = (float)((T)c.prob_percent / (T)100);
}
int ref_gather_batch_sequential(int config, const float* replay, const int* episode_start, const int* position, const unsigned char* full, uint64_t* rng_states,
                                float* observations_actions, float* rewards, unsigned char* terminated, unsigned char* reset, unsigned char* next_reset_base,
                                unsigned char* final_step_mask, unsigned char* next_final_step_mask_base){
#define SEQ_ARGS replay, episode_start, position, full, rng_states, observations_actions, rewards, terminated, reset, next_reset_base, final_step_mask, next_final_step_mask_base
    using namespace ref_opr;
    switch(config){
        case 0: gather_sequential<8, SeqParameters<true, false, false, true, 50>, 48>(SEQ_ARGS); return 0;
        case 1: gather_sequential<8, SeqParameters<true, true, true, true, 50>, 640>(SEQ_ARGS); return 0;
        case 2: gather_sequential<8, SeqParameters<false, false, true, false, 50>, 48>(SEQ_ARGS); return 0;
        case 3: gather_sequential<24, SeqParameters<true, true, true, true, 10>, 640>(SEQ_ARGS); return 0;
        case 4: gather_sequential<2, SeqParameters<true, true, true, true, 50>, 640>(SEQ_ARGS); return 0;
        case 5: gather_sequential<8, SeqParameters<false, true, false, true, 50>, 640>(SEQ_ARGS); return 0;
    }
#undef SEQ_ARGS
    return 1;
}
}
extern "C" {
int ref_gather_batch_size(){ return ref_opr::BATCH; }
int ref_gather_batch_max_episode_length(){ return (int)ENV_TEACHER::EPISODE_STEP_LIMIT; }
void ref_gather_batch(const float* replay, const int* position, const unsigned char* full, uint64_t* rng_states, float* observations_actions, float* rewards,
                      unsigned char* terminated, unsigned char* reset, unsigned char* next_reset, unsigned char* final_step_mask, unsigned char* next_final_step_mask){
    ref_opr::gather<ENV_TEACHER>(replay, position, full, rng_states, observations_actions, rewards, terminated, reset, next_reset, final_step_mask, next_final_step_mask);
}
void ref_off_policy_sizes(int* n, int* steps, int* step_limit, int* capacity){ *n = ref_opr::N; *steps = ref_opr::STEPS; *step_limit = ref_opr::STEP_LIMIT; *capacity = ref_opr::CAPACITY; }
// spec: 3 (TEACHER) or 5 (TEACHER_DR); sample_parameters = OffPolicyRunner PARAMETERS::SAMPLE_PARAMETERS
int ref_off_policy_steps(int spec, int sample_parameters, const float* actor_blob, const float* env_params, float* params_io, float* states_io, uint64_t* rng_states,
                         int* episode_step_io, float* episode_return_io, unsigned char* truncated_io,
                         float* replay, int* episode_start, int* position_io, unsigned char* full_io, int* current_episode_start_io, float* states_out, float* next_states_out){
#define OPR_ARGS actor_blob, env_params, params_io, states_io, rng_states, episode_step_io, episode_return_io, truncated_io, replay, episode_start, position_io, full_io, current_episode_start_io, states_out, next_states_out
    if(spec == 3){ if(sample_parameters) ref_opr::run<ENV_TEACHER, true>(OPR_ARGS); else ref_opr::run<ENV_TEACHER, false>(OPR_ARGS); return 0; }
    if(spec == 5){ if(sample_parameters) ref_opr::run<ENV_TEACHER_DR, true>(OPR_ARGS); else ref_opr::run<ENV_TEACHER_DR, false>(OPR_ARGS); return 0; }
#undef OPR_ARGS
    return 1;
}
}
