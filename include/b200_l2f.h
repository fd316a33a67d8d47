/* include/b200_l2f.h -- C ABI of the B200-native vectorised quadrotor rollout engine.
 *
 * Drop-in boundary for ONE hot path of rl-tools/raptor: the l2f environment's vectorised
 * step/observe (+reward/terminated/samplers) fused with the actor forward.  Every entry point cites
 * the reference interface it replaces.  Shorthands:
 *   L2F/ = rl-tools/include/rl_tools/rl/environments/l2f/
 *   INC/ = rl-tools/include/rl_tools/
 *   R/   = the raptor repository root (README.md = the Python `l2f` / `foundation_policy` surface)
 *
 * Conventions (following the reference's only extern "C" precedent,
 * INC/inference/applications/l2f/c_interface.h:12-31 and INC/inference/executor/c_interface.h:10-44):
 *   - plain pointers and sizes only, float32 data, no C++/torch types, no exceptions across the ABI;
 *   - every function returns an int status (0 = B200L2F_OK); b200l2f_last_error() gives the message;
 *   - the opaque handle owns all device memory; one handle per GPU; a handle is not thread-safe,
 *     different handles are independent;
 *   - pointer arguments carry a memory-space tag (B200L2F_HOST / B200L2F_DEVICE).  Host RESULTS are complete when
 *     the call returns.  Host INPUTS in pageable memory are copied before the call returns (the buffer may be reused
 *     at once); host inputs in page-locked memory (cudaMallocHost, torch pin_memory) are read by the DMA engine
 *     asynchronously: do not modify them before the stream has passed the copy (b200l2f_synchronize, or any call that
 *     returns host results).  With device pointers the work is only enqueued on the handle's stream
 *     (b200l2f_stream / b200l2f_synchronize);
 *   - there is NO CPU fallback: creating a handle without a usable sm_100 device fails.
 *
 * Flat layouts (float32, row-major [n_envs, DIM] at the boundary; struct-of-arrays [DIM][n_envs] in HBM):
 *
 *   parameters row, B200L2F_PARAMS_DIM = 145  (L2F/multirotor.h:23-140, composed at L2F/parameters/default.h:136-149)
 *     0  rotor_positions[4][3]          60 mass                      94  reward.non_negative (0/1)      120 disturbances.random_force.mean
 *     12 rotor_thrust_directions[4][3]  61 gravity[3]                95  reward.scale                   121 disturbances.random_force.std
 *     24 rotor_torque_directions[4][3]  64 J[3][3]                   96  reward.constant                122 disturbances.random_torque.mean
 *     36 rotor_thrust_coefficients[4][3]73 J_inv[3][3]               97  reward.termination_penalty     123 disturbances.random_torque.std
 *     48 rotor_torque_constants[4]      82 hovering_throttle_relative 98 reward.position                124..138 domain_randomization (15, struct order)
 *     52 rotor_time_constants_rising[4] 83 action_limit.min          99  reward.position_clip           139 trajectory.mixture[2]
 *     56 rotor_time_constants_falling[4]84 action_limit.max          100 reward.orientation             141 langevin.gamma
 *                                       85 integration.dt            101 reward.linear_velocity         142 langevin.omega
 *     86 init.guidance                  90 init.max_angular_velocity 102 reward.angular_velocity        143 langevin.sigma
 *     87 init.max_position              91 init.relative_rpm (0/1)   103 reward.linear_acceleration     144 langevin.alpha
 *     88 init.max_angle                 92 init.min_rpm              104 reward.angular_acceleration
 *     89 init.max_linear_velocity       93 init.max_rpm              105 reward.action  106 reward.d_action  107 reward.position_error_integral
 *     108..112 observation_noise {position, orientation, linear_velocity, angular_velocity, imu_acceleration}
 *     113 action_noise.normalized_rpm   114 termination.enabled (0/1) 115..119 termination thresholds {position, linear_velocity,
 *                                                                              angular_velocity, position_integral, orientation_integral}
 *
 *   state row, DIM = 44 + 4*H  (H = action history length; L2F/multirotor.h:574-725)
 *     0 position[3]  3 orientation[4] (w,x,y,z)  7 linear_velocity[3]  10 angular_velocity[3]  13 last_action[4]
 *     17 angular_velocity_history[1][3]  20 force[3]  23 torque[3]  26 rpm[4]  30 current_step (integer value)
 *     31 action_history[H][4]   31+4H trajectory.type (0 POSITION, 1 LANGEVIN)
 *     32+4H langevin {position[3], velocity[3], position_raw[3], velocity_raw[3]}
 */
#ifndef B200_L2F_H
#define B200_L2F_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200L2F_PARAMS_DIM 145
#define B200L2F_ACTION_DIM 4
#define B200L2F_STATE_DIM(H) (44 + 4 * (H))

/* status codes */
enum { B200L2F_OK = 0, B200L2F_ERR_ARGUMENT = 1, B200L2F_ERR_CUDA = 2, B200L2F_ERR_NO_DEVICE = 3, B200L2F_ERR_STATE = 4, B200L2F_ERR_UNSUPPORTED = 5 };
/* memory space of pointer arguments */
enum { B200L2F_HOST = 0, B200L2F_DEVICE = 1 };
/* b200l2f_config.flags */
enum { B200L2F_FLAG_ACCURATE_MATH = 1 /* expf/tanhf/IEEE division in the actor instead of MUFU ex2/rcp (parity debugging) */ };

/* Environment specifications = the reference's compile-time Specification instantiations that are on the path.
 *   DEFAULT   l2f::Specification<float,size_t>: H=16, OBS 82, no Langevin target           (L2F/parameters/default.h:29-176)
 *   RAPTOR    foundation-policy post-training env: H=1, OBS 22, Langevin target            (src/foundation_policy/post_training/environment.h:13-46)
 *   TEACHER   foundation-policy pre-training env:  H=1, OBS 26, Langevin target            (src/foundation_policy/pre_training/environment.h:58-90)
 *   *_DR      same state/observation with DEFAULT_DOMAIN_RANDOMIZATION_OPTIONS<true>         (L2F/parameters/default.h:16-26)
 *             Note: the reference's nominal DR ranges leave rotor_time_constant_{rising,falling} at 0, and sample_initial_parameters asserts on a
 *             zero range (10_sample_initial_parameters.h:176-189); as in the reference, a *_DR handle therefore needs its DR ranges set through
 *             b200l2f_set_environment_parameters (row entries 124..138) before b200l2f_sample_initial_parameters / a collection with resets --
 *             e.g. the foundation-policy ranges of sample_dynamics_parameters.cpp:48-64 -- otherwise those calls return B200L2F_ERR_STATE.            */
enum { B200L2F_SPEC_DEFAULT = 0, B200L2F_SPEC_DEFAULT_DR = 1, B200L2F_SPEC_RAPTOR = 2, B200L2F_SPEC_TEACHER = 3, B200L2F_SPEC_RAPTOR_DR = 4, B200L2F_SPEC_TEACHER_DR = 5 };

/* Actor architectures.
 *   RAPTOR_GRU  Dense(in->hid, ReLU) -> GRU(hid) -> Dense(hid->out)          (checkpoint.h:40-185; post_training/config.h:62-68)
 *               blob: W1[hid][in] b1[hid] W_ih[3hid][hid] b_ih[3hid] W_hh[3hid][hid] b_hh[3hid] h0[hid] W2[out][hid] b2[out]
 *   MLP         [standardize(mean,precision)] -> Dense(in->hid,ReLU) -> Dense(hid->hid,ReLU) -> Dense(hid->out)   (INC/nn_models/mlp/network.h:15-51)
 *               blob: [mean[in] precision[in]] W1[hid][in] b1[hid] W2[hid][hid] b2[hid] W3[out][hid] b3[out] [log_std[4] if head == PPO_GAUSSIAN]
 * Heads.
 *   IDENTITY      action = network output
 *   SQUASH_EVAL   out = [mean, log_std]; action = tanh(mean)      (sample_and_squash in Mode<Evaluation>, INC/nn/layers/sample_and_squash/operations_generic.h:148-194)
 *   PPO_GAUSSIAN  action ~ N(mean, exp(log_std)), log-prob summed  (INC/rl/components/on_policy_runner/operations_generic_per_env.h:43-58) */
enum { B200L2F_POLICY_RAPTOR_GRU = 0, B200L2F_POLICY_MLP = 1 };
enum { B200L2F_HEAD_IDENTITY = 0, B200L2F_HEAD_SQUASH_EVAL = 1, B200L2F_HEAD_PPO_GAUSSIAN = 2 };
/* which kernels compute the policy GEMMs */
enum { B200L2F_GEMM_FP32_CUDA_CORES = 0, B200L2F_GEMM_TCGEN05_3XTF32 = 1 };

typedef struct b200l2f_handle b200l2f_handle;

typedef struct {
    int32_t struct_size;   /* sizeof(b200l2f_config), for ABI evolution */
    int32_t spec;          /* B200L2F_SPEC_* */
    int32_t n_envs;        /* environments owned by this handle (this GPU's shard) */
    int32_t device;        /* CUDA device ordinal */
    int64_t first_env_id;  /* global id of local environment 0: RNG streams are keyed by global id so results do not depend on the number of GPUs */
    int32_t n_state_slots; /* state buffers (slot 0 = `state`, slot 1 = `next_state` of the reference's step signature); >= 2 */
    int32_t flags;         /* B200L2F_FLAG_* */
    void*   stream;        /* cudaStream_t to enqueue on, or NULL: the handle creates its own non-blocking stream */
} b200l2f_config;

typedef struct {
    int32_t arch, input_dim, hidden_dim, output_dim, standardize, head;
    int32_t gru_sequence_length; /* GRU auto-reset period (SEQUENCE_LENGTH, INC/nn/layers/gru/operations_generic.h:80,403); Raptor: 500 */
    int32_t gemm;                /* B200L2F_GEMM_* */
} b200l2f_policy_desc;

/* outputs of the fused rollout; any pointer may be NULL. All row-major, step-major: [T, n_envs, ...]. */
typedef struct {
    int32_t memspace;        /* B200L2F_HOST / B200L2F_DEVICE for all pointers below */
    int32_t state_stride;    /* record a state snapshot every `state_stride` steps (0 = never): states[T/stride + 1, n_envs, STATE_DIM] */
    float*   states;
    float*   observations;   /* [T, n_envs, policy input_dim] */
    float*   actions;        /* [T, n_envs, 4] */
    float*   rewards;        /* [T, n_envs] */
    uint8_t* terminated;     /* [T, n_envs] */
    float*   returns;        /* [n_envs] sum of rewards until (and including) the first termination, rl_tools::evaluate semantics */
    int32_t* episode_length; /* [n_envs] steps until (and including) the first termination */
} b200l2f_rollout_out;

/* ---- lifetime ------------------------------------------------------------------------------- */
int  b200l2f_create(const b200l2f_config* config, b200l2f_handle** out);
int  b200l2f_destroy(b200l2f_handle* h);
const char* b200l2f_last_error(const b200l2f_handle* h); /* h may be NULL: error of the last failed create on this thread */
int  b200l2f_synchronize(b200l2f_handle* h);
void* b200l2f_stream(b200l2f_handle* h);                 /* the cudaStream_t all work is enqueued on */
const char* b200l2f_last_kernel(const b200l2f_handle* h);      /* name of the fused kernel the last b200l2f_rollout / b200l2f_collect launched (profiling tools; "" before the first) */
int  b200l2f_state_dim(const b200l2f_handle* h);
int  b200l2f_observation_dim(const b200l2f_handle* h);
int  b200l2f_action_history_length(const b200l2f_handle* h);
int  b200l2f_n_envs(const b200l2f_handle* h);
int64_t b200l2f_kernel_launches(const b200l2f_handle* h); /* number of engine kernels launched so far on this handle */

/* ---- RNG: vector.initialize_rng(device, rng, seed)  (R/README.md:58; rl_tools::init INC/random/operations_generic.h:16-18).
 * One xorshift64 stream per environment, state = 0xAAAAAAAA + seed + global_env_id; `warmup` extra engine
 * advances decorrelate neighbouring seeds (the reference warms its RNG up the same way, pre_training/config.h RNG_PARAMS_WARMUP_STEPS). */
int b200l2f_initialize_rng(b200l2f_handle* h, uint64_t seed, int32_t warmup);
int b200l2f_get_rng(b200l2f_handle* h, uint64_t* states, int memspace);
int b200l2f_set_rng(b200l2f_handle* h, const uint64_t* states, int memspace);

/* ---- environment: vector.initialize_environment (R/README.md:59; rl_tools::init L2F/operations_generic.h:43-46) */
int b200l2f_initialize_environment(b200l2f_handle* h);                                   /* env.parameters = nominal values of the spec */
int b200l2f_get_environment_parameters(b200l2f_handle* h, float* row145);                /* host pointer */
int b200l2f_set_environment_parameters(b200l2f_handle* h, const float* row145);          /* host pointer; e.g. to install DR ranges */

/* ---- parameters: rl_tools::initial_parameters / sample_initial_parameters (L2F/operations_generic.h:69-78,
 * L2F/operations_generic/10_sample_initial_parameters.h:20-206); vector.sample_initial_parameters (R/README.md:60) */
int b200l2f_initial_parameters(b200l2f_handle* h);
int b200l2f_sample_initial_parameters(b200l2f_handle* h);
int b200l2f_get_parameters(b200l2f_handle* h, float* rows, int memspace);                /* [n_envs, 145] */
int b200l2f_set_parameters(b200l2f_handle* h, const float* rows, int memspace);

/* ---- state: rl_tools::initial_state / sample_initial_state (L2F/operations_generic.h:79-86,
 * L2F/operations_generic/20_initial_state.h, 30_sample_initial_state.h); vector.sample_initial_state (R/README.md:61) */
int b200l2f_initial_state(b200l2f_handle* h, int slot);
int b200l2f_sample_initial_state(b200l2f_handle* h, int slot);
int b200l2f_get_state(b200l2f_handle* h, int slot, float* rows, int memspace);           /* [n_envs, STATE_DIM] */
int b200l2f_set_state(b200l2f_handle* h, int slot, const float* rows, int memspace);
int b200l2f_copy_state(b200l2f_handle* h, int dst_slot, int src_slot);                   /* state.assign(next_state), R/README.md:99 */

/* ---- asynchronous twins of set_parameters / set_state / get_state for callers that feed one rollout after another from host memory (the reference
 * has no counterpart: its CPU environments live in host memory, rl_tools::copy(device_cpu, device_gpu, ...) is synchronous).  The bytes move on the
 * handle's own copy streams through staging buffers; only the device-side transpose is ordered on the main stream, in call order -- so an upload
 * issued after rollout k is enqueued overlaps that kernel and takes effect before rollout k+1, and a download issued after rollout k leaves the host
 * free to enqueue rollout k+1 at once.  Host buffers must be page-locked (cudaHostAlloc / cudaHostRegister / torch pin_memory; pageable memory falls
 * back to the synchronous call) and must not be touched until b200l2f_transfers_synchronize(h, which) returns (which: 1 = uploads, 2 = downloads, 3 = both).
 * copy_to_host_async: any device buffer the main stream produced (e.g. the `returns` of b200l2f_rollout) -> page-locked host memory, same ordering. */
int b200l2f_set_parameters_async(b200l2f_handle* h, const float* rows_pinned);           /* [n_envs, 145] */
int b200l2f_set_state_async(b200l2f_handle* h, int slot, const float* rows_pinned);      /* [n_envs, STATE_DIM] */
int b200l2f_get_state_async(b200l2f_handle* h, int slot, float* rows_pinned);
int b200l2f_copy_to_host_async(b200l2f_handle* h, void* dst_pinned, const void* src_device, size_t bytes);
int b200l2f_transfers_synchronize(b200l2f_handle* h, int which);

/* ---- rl_tools::observe (L2F/operations_generic.h:87-92, L2F/operations_generic/40_observe.h); vector.observe (R/README.md:96).
 * observations: [n_envs, ld] with ld >= OBSERVATION_DIM */
int b200l2f_observe(b200l2f_handle* h, int slot, float* observations, int ld, int memspace);
/* ---- rl_tools::step (L2F/operations_generic.h:94-130); vector.step (R/README.md:98). actions [n_envs, 4]; dts [n_envs] or NULL */
int b200l2f_step(b200l2f_handle* h, int slot, const float* actions, int next_slot, float* dts, int memspace);
/* ---- n_steps x rl_tools::step under ONE held action (host pointer, 4 floats) for every environment, in place on `slot`, state in registers across the
 * steps: the loop of the reference's GPU benchmark (RT/src/rl/environments/l2f/cuda/benchmark.cu:111-120: action 0, step, state = next_state).
 * Default-math arithmetic of the fused kernels; needs noise-free parameters with uniform MDP constants. */
int b200l2f_step_repeated(b200l2f_handle* h, int slot, const float* action4, int32_t n_steps);
/* ---- rl_tools::reward / terminated (L2F/operations_generic.h:142-176, L2F/parameters/reward_functions/squared/operations_generic.h:100-129) */
int b200l2f_reward(b200l2f_handle* h, int slot, const float* actions, int next_slot, float* rewards, int memspace);
int b200l2f_terminated(b200l2f_handle* h, int slot, uint8_t* flags, int memspace);

/* ---- actor: foundation_policy.Raptor().reset() / .evaluate_step(obs[:, :22]) (R/README.md:19-24,94-97);
 * rl_tools::reset / evaluate_step (INC/nn_models/sequential/operations_generic.h:63-66,321-325) */
int b200l2f_policy_load(b200l2f_handle* h, const b200l2f_policy_desc* desc, const float* blob, size_t n_floats); /* host blob */
int b200l2f_policy_reset(b200l2f_handle* h, const uint8_t* mask, int memspace);          /* mask[n_envs] or NULL = all (mode::sequential::ResetMask) */
int b200l2f_policy_evaluate_step(b200l2f_handle* h, const float* observations, int ld, float* actions, int no_auto_reset, int memspace);
int b200l2f_policy_get_hidden(b200l2f_handle* h, float* hidden, int32_t* gru_step, int memspace); /* [n_envs, hidden_dim], [n_envs] */
int b200l2f_policy_set_hidden(b200l2f_handle* h, const float* hidden, const int32_t* gru_step, int memspace);

/* ---- the fused hot path: T closed-loop steps of observe -> actor -> step -> reward -> terminated for every environment in ONE
 * persistent kernel launch, state / hidden state / RNG resident on chip; replaces the loop body of rl_tools::evaluate
 * (INC/rl/utils/evaluation/operations_generic.h:138-189) and of the README loop (R/README.md:94-99). Operates in place on slot 0. */
int b200l2f_rollout(b200l2f_handle* h, int32_t n_steps, int32_t no_auto_reset, const b200l2f_rollout_out* out);

/* ---- PPO collection with on-device auto-reset and trajectory write-back: replaces rl_tools::collect
 * (INC/rl/components/on_policy_runner/operations_generic.h:99-131, operations_generic_per_env.h:8-75).
 * dataset: [(T+1)*n_envs, OBS+15] rows = step*n_envs + env, columns obs | actions_mean[4] | actions[4] | log_prob | reward |
 * terminated | truncated | value | advantage | target_value (on_policy_runner.h:42-64); the last n_envs rows hold the final observations.
 * All specs: RAPTOR / TEACHER (H = 1; tcgen05 or CUDA-core actor) and DEFAULT (the PPO zoo's environment, INC/rl/zoo/l2f/ppo.h: H = 16 action
 * history, OBS 82, 97-float rows; CUDA-core actor, as for the learner feed below). */
int b200l2f_collect_reset(b200l2f_handle* h);                                             /* runner init: truncated = true, episode_step/return = 0 (operations_generic.h:65-75) */
int b200l2f_collect(b200l2f_handle* h, int32_t n_steps, int32_t episode_step_limit, float* dataset, int memspace);

/* ---- status of the last b200l2f_rollout / b200l2f_collect on this handle, reduced on the device right behind the fused kernel:
 *   n_nonfinite          environments whose state holds a NaN or Inf afterwards (the reference's per-state check: L2F/operations_generic/05_state_is_nan.h;
 *                        nonfinite_flags [n_envs] receives the per-environment flags, or NULL)
 *   rollout only (has_episodes = 1): the aggregates of rl::utils::evaluation::Result (INC/rl/utils/evaluation/operations_generic.h:201-213) over the
 *   n_envs episodes -- returns / episode length mean and std (population std, max(0, E[x^2] - E[x]^2)), num_terminated, share_terminated */
typedef struct {
    int64_t n_envs, n_nonfinite, n_terminated;
    int32_t has_episodes, reserved;
    double returns_mean, returns_std, episode_length_mean, episode_length_std, share_terminated;
} b200l2f_status;
int b200l2f_last_status(b200l2f_handle* h, b200l2f_status* out, uint8_t* nonfinite_flags, int memspace);

/* ---- multi-GPU: environments shard by global id (b200l2f_config.first_env_id) and the rollout path has NO collective.  The one optional call:
 * all-gather of equally sized trajectory slabs (e.g. the dataset b200l2f_collect wrote) across the ranks of the caller's NCCL communicator, enqueued on the
 * handle's stream behind the kernel that produced the slab.  nccl_comm: ncclComm_t (one rank per GPU / process); send: count_per_rank floats on this device;
 * recv: n_ranks * count_per_rank floats on this device, rank-major; n_ranks_out: ncclCommCount, or NULL.  NCCL is bound at run time (dlopen libnccl.so.2,
 * B200L2F_NCCL_LIB overrides the name): B200L2F_ERR_UNSUPPORTED when it is not in the process / on the library path. */
int b200l2f_allgather_trajectories(b200l2f_handle* h, void* nccl_comm, const float* send, float* recv, size_t count_per_rank, int32_t* n_ranks_out);

/* ---- PPO learner feed: what the reference's loop step does between collect and train on the dataset above
 * (INC/rl/algorithms/ppo/loop/core/operations_generic.h:104-117), without the data leaving the GPU.
 * critic_load: the value network [standardize ->] Dense(OBS,64,ReLU) -> Dense(64,64,ReLU) -> Dense(64,1) (loop/core/config.h:62-76), blob in the
 *   MLP order above with output_dim 1, head IDENTITY; desc->gemm selects tcgen05 (3xTF32) or fp32 CUDA cores.
 * evaluate_values: `evaluate(device, critic, all_observations_privileged, all_values, ...)` (:112-116) over all (T+1) n rows -> column OBS+12.
 * estimate_generalized_advantages: INC/rl/algorithms/ppo/operations_generic.h:54-89 on the value column -> columns OBS+13 (advantage), OBS+14
 *   (target_value); gamma / lambda / ignore_termination = PPO_PARAMETERS::GAMMA / LAMBDA / IGNORE_TERMINATION (ppo.h:15-16,33).
 * values_and_advantages: both in ONE backward pass over time (the dataset is read once).
 * normalizer_update: rl::components::running_normalizer `update` (INC/rl/components/running_normalizer/operations_generic.h:27-49) with the
 *   observation block [0, T n) x [0, OBS) as data; mean_io / std_io [OBS] and age_io are host pointers (the normalizer lives with the caller,
 *   who passes mean and 1 / std to the standardize layers: set_statistics, INC/nn/layers/standardize/operations_generic.h). */
int b200l2f_critic_load(b200l2f_handle* h, const b200l2f_policy_desc* desc, const float* blob, size_t n_floats); /* host blob */
int b200l2f_evaluate_values(b200l2f_handle* h, int32_t n_steps, float* dataset, int memspace);
int b200l2f_estimate_generalized_advantages(b200l2f_handle* h, int32_t n_steps, float gamma, float lambda, int ignore_termination, float* dataset, int memspace);
int b200l2f_values_and_advantages(b200l2f_handle* h, int32_t n_steps, float gamma, float lambda, int ignore_termination, float* dataset, int memspace);
int b200l2f_normalizer_update(b200l2f_handle* h, int32_t n_steps, const float* dataset, int memspace, float* mean_io, float* std_io, int32_t* age_io);

/* ---- Foundation-policy DAgger data path: gather_epoch = sample_trajectories + add_to_dataset
 * (src/foundation_policy/post_training/helper.h:6-41,43-110,112-123; driven per teacher by post_training/main.cpp:261-309) for ALL teachers in one
 * call.  Handle spec RAPTOR / RAPTOR_DR (the post-training environment); environment e is an episode of teacher e / episodes_per_teacher and runs
 * with the parameters the caller installed for it (b200l2f_set_parameters: teacher_parameters[i], main.cpp:96,205-207); the student is the
 * loaded Raptor GRU actor; initial states are whatever slot 0 holds (b200l2f_sample_initial_state).
 * teachers_load: blobs [n_teachers][6408] = W1[64][26] b1[64] W2[64][64] b2[64] W3[8][64] b3[8] (SAC actor MLP, rl/algorithms/sac/loop/core/
 *   approximators_mlp.h:14-37; evaluated with sample_and_squash in Evaluation mode: action = tanh(mean)); position_offsets [n_teachers][3] =
 *   TeacherMeta::steady_state_position_offset (helper.h:1-4, main.cpp:214-235) or NULL for zeros.  Host pointers.
 * dagger_gather: rolls the student out for n_steps (= ENVIRONMENT::EPISODE_STEP_LIMIT, 500), then for every episode up to and including its
 *   first terminated step appends: input_student [rows][22] (student observation, position minus the teacher's offset), output_target [rows][4]
 *   (the teacher's action for the teacher observation of the same state), truncated / reset [rows], episode_start [n_envs] (first row of each
 *   episode).  Rows are ordered by environment, then step.  *rows_added = rows (<= capacity_rows, else B200L2F_ERR_ARGUMENT "Dataset size
 *   exceeded").  returns / episode_length [n_envs] (optional) are the Result of sample_trajectories.  Observation noise must be off. */
typedef struct {
    int32_t memspace;        /* B200L2F_HOST / B200L2F_DEVICE for all pointers below */
    int32_t reserved;
    int64_t capacity_rows;   /* rows the row-indexed buffers can hold; n_envs * n_steps always suffices */
    float*   input_student;
    float*   output_target;
    uint8_t* truncated;
    uint8_t* reset;
    int32_t* episode_start;
    float*   returns;        /* may be NULL */
    int32_t* episode_length; /* may be NULL */
} b200l2f_dagger_out;
int b200l2f_teachers_load(b200l2f_handle* h, int32_t n_teachers, int32_t episodes_per_teacher, const float* blobs, const float* position_offsets, int32_t gemm);
int b200l2f_dagger_gather(b200l2f_handle* h, int32_t n_steps, int32_t no_auto_reset, const b200l2f_dagger_out* out, int64_t* rows_added);

/* ---- off-policy runner steps (SAC teacher data collection of the foundation-policy pre-training): rl::components::off_policy_runner `step`
 * = prologue + interlude + epilogue (INC/rl/components/off_policy_runner/operations_generic.h:215-238, operations_generic_per_env.h:8-110; the
 * reference's CUDA precedent: operations_cuda.h:62-106) with the replay buffer `add` (INC/rl/components/replay_buffer/operations_generic.h:54-79),
 * n_steps runner steps in one launch.  Handle spec TEACHER / TEACHER_DR; the loaded actor is the SAC MLP OBS-64-64-8 ([mean, log_std], either SQUASH
 * head): the runner always evaluates it in Mode<Rollout>, action = tanh(mean + N(0,1) exp(clamp(log_std, -20, 2))).  Every environment owns one
 * ring = the reference's ReplayBuffer::data matrix [capacity][2*OBS + 7], columns obs | action[4] | reward | next_obs | terminated | truncated
 * (replay_buffer.h:37-58; symmetric observations), plus episode_start[capacity] and position / full / current_episode_start.  A fresh runner is
 * all-zero buffers + b200l2f_collect_reset (truncated = true).  episode_step_limit = PARAMETERS::EPISODE_STEP_LIMIT (truncated = terminated or
 * episode_step == limit); sample_parameters = PARAMETERS::SAMPLE_PARAMETERS (re-sample the dynamics on every reset).  The states / next_states
 * of ReplayBufferWithStates (used by recalculate_rewards only) are not stored.  Buffers are updated in place (device) or staged (host).
 * runner_get_state / runner_set_state: the per-environment bookkeeping shared with b200l2f_collect; any pointer may be NULL. */
typedef struct {
    int32_t memspace;               /* B200L2F_HOST / B200L2F_DEVICE for all pointers below */
    int32_t capacity;               /* REPLAY_BUFFER_CAPACITY: rows per environment */
    float*   data;                  /* [n_envs][capacity][2*OBS + 7] */
    int32_t* episode_start;         /* [n_envs][capacity] */
    int32_t* position;              /* [n_envs] */
    uint8_t* full;                  /* [n_envs] */
    int32_t* current_episode_start; /* [n_envs] */
} b200l2f_replay_buffers;
int b200l2f_off_policy_steps(b200l2f_handle* h, int32_t n_steps, int32_t episode_step_limit, int32_t sample_parameters, const b200l2f_replay_buffers* rb);
/* gather_batch for SEQUENCE_LENGTH = 1 (the MLP SAC configuration, pre_training/config.h:19,52-58): the learner-side read of the rings,
 * INC/rl/components/off_policy_runner/operations_generic.h:240-434 with one RNG stream per batch sample (operations_cuda.h:36-60): sample b draws its
 * environment env_begin + uniform_int(0, env_count - 1) (env_begin / env_count select one runner's group of environments, e.g. one teacher), then its
 * ring offset; max_episode_length = ENVIRONMENT::EPISODE_STEP_LIMIT (full rings are read from position + max_episode_length on, :300-303).
 * Outputs in the SequentialBatch layout (off_policy_runner.h:96-141): observations_actions [2][B][OBS + 4] (step 0 = obs | action, step 1 = next_obs | 0),
 * rewards / terminated [B], masks as the reference fills them for this configuration.  rng_states [B] in/out, memory space of the batch (= of the rings). */
typedef struct {
    int32_t memspace;                /* B200L2F_HOST / B200L2F_DEVICE for all pointers below and for rng_states */
    int32_t batch_size;
    float*   observations_actions;   /* [2][B][OBS + 4] */
    float*   rewards;                /* [B] */
    uint8_t* terminated;             /* [B] */
    uint8_t* reset;                  /* [B]    = 1, may be NULL */
    uint8_t* next_reset;             /* [2][B] = 1, may be NULL */
    uint8_t* final_step_mask;        /* [B]    = 1, may be NULL */
    uint8_t* next_final_step_mask;   /* [2][B] = {0, 1}, may be NULL */
    int32_t* env_index;              /* [B] which ring, may be NULL */
    int32_t* sample_index;           /* [B] which row, may be NULL */
} b200l2f_batch;
int b200l2f_gather_batch(b200l2f_handle* h, const b200l2f_replay_buffers* rb, int32_t max_episode_length, int32_t env_begin, int32_t env_count, uint64_t* rng_states,
                         const b200l2f_batch* out);
/* gather_batch for any SEQUENCE_LENGTH (recurrent SAC): INC/rl/components/off_policy_runner/operations_generic.h:240-434 with the batch parameters of
 * off_policy_runner.h:78-85 as run-time values (b200l2f_batch_parameters_default = the reference's defaults for SEQUENCE_LENGTH > 1).  One RNG stream per batch
 * sample as above.  With L = sequence_length the batch tensors have the SequentialBatch shapes (off_policy_runner.h:96-141): observations_actions [L+1][B][OBS+4],
 * rewards / terminated / reset / final_step_mask [L][B], next_reset / next_final_step_mask [L+1][B] (the `*_base` tensors; the reference's `next_*` views start
 * at row include_first_step_in_targets ? 0 : 1, operations_generic.h:87-92).  All four masks and rb->episode_start are required; env_index / sample_index report
 * the environment and the first ring row of each sample.  rewards / terminated of padding steps, which the reference never writes, are 0. */
typedef struct {
    int32_t sequence_length;                              /* SEQUENCE_LENGTH >= 1 */
    int32_t include_first_step_in_targets;                /* INCLUDE_FIRST_STEP_IN_TARGETS */
    int32_t always_sample_from_initial_state;             /* ALWAYS_SAMPLE_FROM_INITIAL_STATE (needs capacity >= max_episode_length, :258) */
    int32_t random_seq_length;                            /* RANDOM_SEQ_LENGTH */
    int32_t enable_nominal_sequence_length_probability;   /* ENABLE_NOMINAL_SEQUENCE_LENGTH_PROBABILITY */
    float   nominal_sequence_length_probability;          /* NOMINAL_SEQUENCE_LENGTH_PROBABILITY */
} b200l2f_batch_parameters;
int b200l2f_gather_batch_sequential(b200l2f_handle* h, const b200l2f_replay_buffers* rb, const b200l2f_batch_parameters* parameters, int32_t max_episode_length, int32_t env_begin,
                                    int32_t env_count, uint64_t* rng_states, const b200l2f_batch* out);
int b200l2f_runner_get_state(b200l2f_handle* h, int32_t* episode_step, float* episode_return, uint8_t* truncated, int memspace);
int b200l2f_runner_set_state(b200l2f_handle* h, const int32_t* episode_step, const float* episode_return, const uint8_t* truncated, int memspace);

/* ---- parameter / state JSON wire format: rl_tools::json / from_json (L2F/operations_cpu.h:139-411 parameters, :412-560 state, :565-824 import).
 * Host functions on the flat rows above; to_json writes the reference's exact text (key order, separators, std::to_string formatting), *length
 * receives the size without the terminator (call with buf = NULL to query it); from_json reads every key the reference reads (a missing or
 * mistyped key is an error naming its path, like nlohmann's exception), narrows doubles to float, ignores unknown keys and leaves the row
 * untouched on error.  The dynamics_parameters / <id>.json files of the foundation-policy data set (sample_dynamics_parameters.cpp:84-111,
 * loaded by post_training/main.cpp:202-207) are parameters_from_json inputs.  parameters_* accept h = NULL (no handle state is used). */
int b200l2f_parameters_to_json(b200l2f_handle* h, const float* row145, char* buf, size_t capacity, size_t* length);
int b200l2f_parameters_from_json(b200l2f_handle* h, const char* json, float* row145_io);
int b200l2f_state_to_json(b200l2f_handle* h, const float* state_row, char* buf, size_t capacity, size_t* length);
int b200l2f_state_from_json(b200l2f_handle* h, const char* json, float* state_row_io);

/* ---- rl-tools checkpoint code export (`checkpoint.h`): what rl::loop::steps::checkpoint::save_code writes
 * (INC/rl/loop/steps/checkpoint/operations_cpu.h:56-118; body by rl_tools::save_code of the actor, containers/{matrix,tensor}/persist_code.h)
 * and what the reference consumes by compiling it in (src/foundation_policy/post_training/load_actor.cpp, INC/inference/applications/l2f/c_backend.h).
 * Host functions, no handle, no GPU; errors through b200l2f_last_error(NULL).
 * parse:   reads the header text: every `memory[]` byte list becomes a float tensor named by its namespace path
 *          ("rl_tools::checkpoint::actor::layer_1::weights_input", "rl_tools::checkpoint::example::input", ...), row padding removed.
 * tensor:  i-th tensor in file order: path, rank, dims, data (owned by the checkpoint object).
 * string:  `char name[] = "..."` values, e.g. "rl_tools::checkpoint::meta::name" / "::commit_hash"; NULL if absent.  string_count / string_at
 *          enumerate them (sorted by path).
 * policy:  recognises the actor under `root` (NULL = "rl_tools::checkpoint::actor") and assembles desc + blob for b200l2f_policy_load /
 *          b200l2f_critic_load / b200l2f_teachers_load in the blob orders documented above: Dense(ReLU)-GRU-Dense (Raptor) or
 *          [Standardize] MLP(3 layers, ReLU) [SampleAndSquash | log_std]; anything else -> B200L2F_ERR_UNSUPPORTED.  blob may be NULL to query
 *          *n_floats.  desc->gemm is preset to TCGEN05_3XTF32, desc->gru_sequence_length to the export's SEQUENCE_LENGTH (500 for Raptor).
 * parse_h5: the HDF5 twin `checkpoint.h5` (rl::loop::steps::checkpoint::save, INC/rl/loop/steps/checkpoint/operations_cpu.h:119-160; written by
 *          rl_tools::save(device, actor, HighFive::Group): nn_models/sequential/persist.h:14-21, nn/layers/{dense,gru,standardize,sample_and_squash}/persist.h,
 *          nn/parameters/persist.h:10-13, containers/{matrix,tensor}/persist.h) from a memory image of the file, read by the engine's own HDF5 reader
 *          (no libhdf5): superblock 0/1, version-1 object headers, symbol-table groups, contiguous / compact float32 / float64 datasets, string
 *          attributes; anything else is an error that names it.  Datasets appear under the code export's paths
 *          ("/actor/layers/1/weights_input/parameters" -> "rl_tools::checkpoint::actor::layer_1::weights_input"), string attributes under
 *          "<path>::<attribute>" ("rl_tools::checkpoint::actor::layer_0::activation_function", "rl_tools::checkpoint::actor::meta"), the actor
 *          group's checkpoint_name also as "rl_tools::checkpoint::meta::name"; tensor / string / policy work as for the code export.  The file
 *          does not record SEQUENCE_LENGTH: desc->gru_sequence_length is 0 (= the engine default 500).  `parse` forwards here when the buffer
 *          starts with the HDF5 signature. */
typedef struct b200l2f_checkpoint b200l2f_checkpoint;
int b200l2f_checkpoint_parse(const char* text, size_t length, b200l2f_checkpoint** out);
int b200l2f_checkpoint_parse_h5(const void* bytes, size_t length, b200l2f_checkpoint** out);
int b200l2f_checkpoint_free(b200l2f_checkpoint* c);
int b200l2f_checkpoint_tensor_count(const b200l2f_checkpoint* c);
int b200l2f_checkpoint_tensor(const b200l2f_checkpoint* c, int index, const char** path, int32_t* rank, const int64_t** dims, const float** data);
const char* b200l2f_checkpoint_string(const b200l2f_checkpoint* c, const char* path);
int b200l2f_checkpoint_string_count(const b200l2f_checkpoint* c);
int b200l2f_checkpoint_string_at(const b200l2f_checkpoint* c, int index, const char** path, const char** value);
int b200l2f_checkpoint_policy(const b200l2f_checkpoint* c, const char* root, b200l2f_policy_desc* desc, float* blob, size_t capacity, size_t* n_floats);

#ifdef __cplusplus
}
#endif
#endif /* B200_L2F_H */
