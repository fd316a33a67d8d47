/* oracle/l2f_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's vectorised quadrotor rollout hot path
 * (rl-tools l2f step/observe/reward/terminated/samplers + the actor forward).  It exists to CHECK
 * the CUDA engine (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg); nothing under
 * raptor_b200/ may include, link or call it.
 *
 * Pinning: oracle/_ref/libl2f_ref.so (the unmodified reference, built by oracle/Makefile from
 * /root/reference) is compared against this file function by function in
 * tests/test_oracle_vs_reference.py, and both are compared against the committed fixtures in
 * tests/golden/ (generated from the reference by tests/golden/generate.py), including the
 * known-answer test that ships inside the Raptor checkpoint.
 *
 * Flat layouts: identical to include/b200_l2f.h (B200L2F_PARAMS_DIM = 145 floats per environment,
 * state = 44 + 4*H floats per environment).
 */
#ifndef L2F_ORACLE_H
#define L2F_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_PARAMS_DIM 145

/* spec ids (same numbering as oracle/ref_l2f.cpp and include/b200_l2f.h) */
enum { ORACLE_SPEC_DEFAULT = 0, ORACLE_SPEC_DEFAULT_DR = 1, ORACLE_SPEC_RAPTOR = 2, ORACLE_SPEC_TEACHER = 3, ORACLE_SPEC_RAPTOR_DR = 4, ORACLE_SPEC_TEACHER_DR = 5 };

int      oracle_params_dim(void);
int      oracle_state_dim(int spec);
int      oracle_observation_dim(int spec);
int      oracle_action_history_length(int spec);

uint64_t oracle_rng_init(uint64_t seed);
float    oracle_rng_uniform(uint64_t* state, float lo, float hi);
float    oracle_rng_normal(uint64_t* state, float mean, float std);

void  oracle_nominal_parameters(int spec, float* p);
void  oracle_sample_initial_parameters(int spec, const float* env_p, uint64_t* rng, float* out);
void  oracle_initial_state(int spec, const float* p, float* s);
void  oracle_sample_initial_state(int spec, const float* p, uint64_t* rng, float* s);
void  oracle_observe(int spec, const float* p, const float* s, uint64_t* rng, float* obs);
float oracle_step(int spec, const float* p, const float* s, const float* a, uint64_t* rng, float* s_next);
float oracle_reward(int spec, const float* p, const float* s, const float* a, const float* s_next, uint64_t* rng);
int   oracle_terminated(int spec, const float* p, const float* s);

/* ---- policies ---- */
enum { ORACLE_POLICY_RAPTOR_GRU = 0, ORACLE_POLICY_MLP = 1 };
enum { ORACLE_HEAD_IDENTITY = 0, ORACLE_HEAD_SQUASH_EVAL = 1, ORACLE_HEAD_PPO_GAUSSIAN = 2, ORACLE_HEAD_SQUASH_SAMPLE = 3 /* sample_and_squash in Mode<Rollout>: tanh(mean + N(0,1) exp(clamp(log_std))) */ };
typedef struct {
    int arch;        /* ORACLE_POLICY_* */
    int input_dim;   /* leading observation columns consumed */
    int hidden_dim;  /* 16 (Raptor) / 64 (MLP actors) */
    int output_dim;  /* 4, or 8 = [mean, log_std] for the SAC actor */
    int standardize; /* MLP only: leading standardize layer (mean, precision) */
    int head;        /* ORACLE_HEAD_* */
    const float* blob;
} oracle_policy_t;
int  oracle_policy_num_parameters(const oracle_policy_t* pol);
/* one evaluate_step for N environments. hidden [N,hidden_dim] and gru_step [N] are used by the GRU
 * arch only.  head==PPO_GAUSSIAN draws from rng[N] (may be NULL for the other heads); out_mean and
 * out_log_prob may be NULL. */
void oracle_policy_evaluate_step(const oracle_policy_t* pol, int N, const float* obs, int obs_ld, float* hidden, int* gru_step,
                                 int no_auto_reset, uint64_t* rng, float* actions, float* out_mean, float* out_log_prob);

/* closed-loop rollout in the order of rl_tools::evaluate; all out_* may be NULL */
void oracle_rollout(int spec, const oracle_policy_t* pol, int N, int T, int threads, const float* params, float* states_io, uint64_t* rng_states,
                    float* hidden_io, int* gru_step_io, int no_auto_reset,
                    float* out_states, float* out_observations, float* out_actions, float* out_rewards, unsigned char* out_terminated);

/* PPO collection in the order of rl_tools::collect (on_policy_runner); dataset row layout documented in l2f_oracle.c */
void oracle_collect(int spec, const oracle_policy_t* pol, int N, int T, int threads, int episode_step_limit, const float* env_params,
                    float* params_io, float* states_io, uint64_t* rng_states, int* episode_step_io, float* episode_return_io, unsigned char* truncated_io,
                    float* dataset, int data_dim);
/* off-policy runner steps (SAC teacher data collection into per-environment replay rings); layout documented in l2f_oracle.c */
void oracle_off_policy_steps(int spec, const oracle_policy_t* pol, int N, int T, int episode_step_limit, int capacity, int sample_parameters, const float* env_params,
                             float* params_io, float* states_io, uint64_t* rng_states, int* episode_step_io, float* episode_return_io, unsigned char* truncated_io,
                             float* replay, int* episode_start, int* position_io, unsigned char* full_io, int* current_episode_start_io,
                             float* states_out, float* next_states_out);
/* gather_batch of the off-policy runner for SEQUENCE_LENGTH = 1; layout documented in l2f_oracle.c */
void oracle_gather_batch(int obs_dim, int capacity, int max_episode_length, int env_begin, int env_count, const float* replay, const int* position, const unsigned char* full,
                         int B, uint64_t* rng_states, float* observations_actions, float* rewards, unsigned char* terminated,
                         unsigned char* reset, unsigned char* next_reset, unsigned char* final_step_mask, unsigned char* next_final_step_mask,
                         int* env_index, int* sample_index_out);
void oracle_gather_batch_sequential(int obs_dim, int capacity, int max_episode_length, int env_begin, int env_count, const float* replay, const int* episode_start,
                                    const int* position, const unsigned char* full, int L, int include_first_step_in_targets, int always_sample_from_initial_state,
                                    int random_seq_length, int enable_nominal, float nominal_probability,
                                    int B, uint64_t* rng_states, float* observations_actions, float* rewards, unsigned char* terminated,
                                    unsigned char* reset, unsigned char* next_reset_base, unsigned char* final_step_mask, unsigned char* next_final_step_mask_base,
                                    int* env_index, int* sample_index_out);
/* learner feed (PPO loop step between collect and train): critic values, GAE, running observation normalizer */
void oracle_evaluate_values(const oracle_policy_t* critic, int N, int T, float* dataset, int data_dim);
void oracle_estimate_generalized_advantages(int N, int T, float* dataset, int data_dim, float gamma, float lambda, int ignore_termination);
void oracle_normalizer_update(int N, int T, const float* dataset, int data_dim, float* mean_io, float* std_io, int* age_io);
/* foundation-policy DAgger data path (post_training/helper.h:43-110): see l2f_oracle.c */
long long oracle_dagger_add_to_dataset(int n, int T, int episodes_per_teacher, const float* params, const float* states, const unsigned char* terminated, uint64_t* rng,
                                       const float* teacher_blobs, const float* offsets, int* episode_start, float* input_student, float* output_target,
                                       unsigned char* truncated_out, unsigned char* reset_out);
int oracle_hardware_threads(void);
#ifdef __cplusplus
}
#endif
#endif
