/* oracle/l2f_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see l2f_oracle.h).
 *
 * Plain-C restatement of the reference's quadrotor rollout hot path.  Every function cites the
 * reference lines it follows.  Shorthands:
 *   L2F/ = /root/reference/rl-tools/include/rl_tools/rl/environments/l2f/
 *   INC/ = /root/reference/rl-tools/include/rl_tools/
 *   SRC/ = /root/reference/rl-tools/src/foundation_policy/
 * Expression trees (association order, float-vs-double promotion of literals) are kept exactly as
 * the reference's C++ evaluates them for T = float, so that this file compiled with
 * -ffp-contract=off is bit-identical to oracle/_ref/libl2f_ref.so built with the same flags
 * (tests/test_oracle_vs_reference.py asserts exact equality).
 */
#include "l2f_oracle.h"
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <stdio.h>
#include <pthread.h>
#include <unistd.h>

/* ---------------------------------------------------------------------------------------------
 * flat parameter offsets (include/b200_l2f.h documents the same table)
 * ------------------------------------------------------------------------------------------- */
enum {
    P_ROTOR_POS = 0, P_THRUST_DIR = 12, P_TORQUE_DIR = 24, P_THRUST_COEF = 36, P_TORQUE_CONST = 48,
    P_TAU_RISE = 52, P_TAU_FALL = 56, P_MASS = 60, P_GRAVITY = 61, P_J = 64, P_JINV = 73, P_HOVER = 82,
    P_ACT_MIN = 83, P_ACT_MAX = 84, P_DT = 85,
    P_INIT_GUIDANCE = 86, P_INIT_MAX_POS = 87, P_INIT_MAX_ANGLE = 88, P_INIT_MAX_LINVEL = 89, P_INIT_MAX_ANGVEL = 90,
    P_INIT_REL_RPM = 91, P_INIT_MIN_RPM = 92, P_INIT_MAX_RPM = 93,
    P_RW_NONNEG = 94, P_RW_SCALE = 95, P_RW_CONSTANT = 96, P_RW_TERM_PENALTY = 97, P_RW_POSITION = 98, P_RW_POSITION_CLIP = 99,
    P_RW_ORIENTATION = 100, P_RW_LINVEL = 101, P_RW_ANGVEL = 102, P_RW_LINACC = 103, P_RW_ANGACC = 104, P_RW_ACTION = 105,
    P_RW_DACTION = 106, P_RW_POS_INTEGRAL = 107,
    P_NOISE_POS = 108, P_NOISE_ORI = 109, P_NOISE_LINVEL = 110, P_NOISE_ANGVEL = 111, P_NOISE_IMU = 112, P_ACTION_NOISE = 113,
    P_TERM_ENABLED = 114, P_TERM_POS = 115, P_TERM_LINVEL = 116, P_TERM_ANGVEL = 117, P_TERM_POS_INT = 118, P_TERM_ORI_INT = 119,
    P_DIST_FORCE_MEAN = 120, P_DIST_FORCE_STD = 121, P_DIST_TORQUE_MEAN = 122, P_DIST_TORQUE_STD = 123,
    P_DR_T2W_MIN = 124, P_DR_T2W_MAX = 125, P_DR_T2I_MIN = 126, P_DR_T2I_MAX = 127, P_DR_MASS_MIN = 128, P_DR_MASS_MAX = 129,
    P_DR_MASS_SIZE_DEV = 130, P_DR_TAU_RISE_MIN = 131, P_DR_TAU_RISE_MAX = 132, P_DR_TAU_FALL_MIN = 133, P_DR_TAU_FALL_MAX = 134,
    P_DR_KQ_MIN = 135, P_DR_KQ_MAX = 136, P_DR_ORI_OFFSET = 137, P_DR_DIST_FORCE_MAX = 138,
    P_TRAJ_MIX0 = 139, P_TRAJ_MIX1 = 140, P_LANGEVIN_GAMMA = 141, P_LANGEVIN_OMEGA = 142, P_LANGEVIN_SIGMA = 143, P_LANGEVIN_ALPHA = 144
};
/* flat state offsets; S_HIST holds H*4 floats, the trajectory block follows */
enum { S_POS = 0, S_ORI = 3, S_LINVEL = 7, S_ANGVEL = 10, S_LAST_ACTION = 13, S_ANGVEL_HIST = 17, S_FORCE = 20, S_TORQUE = 23, S_RPM = 26, S_CURRENT_STEP = 30, S_HIST = 31 };
#define S_TRAJ_TYPE(H) (31 + 4 * (H))
#define S_LANGEVIN(H) (32 + 4 * (H)) /* position[3] velocity[3] position_raw[3] velocity_raw[3] */
#define MAX_STATE_DIM (44 + 4 * 16)
#define MAX_OBS_DIM 82

typedef struct { int H; int langevin; int dr; int obs_layout; int obs_dim; } spec_t;
enum { OBS_DEFAULT = 0, OBS_RAPTOR = 1, OBS_TEACHER = 2 };
/* L2F/parameters/default.h:159-171 (H=16, OBS 82); SRC/post_training/environment.h:13-46 (H=1, OBS 22);
 * SRC/pre_training/environment.h:58-90 (H=1, OBS 26) */
static spec_t get_spec(int spec){
    spec_t s;
    switch(spec){
        case ORACLE_SPEC_DEFAULT:    s.H = 16; s.langevin = 0; s.dr = 0; s.obs_layout = OBS_DEFAULT; s.obs_dim = 82; break;
        case ORACLE_SPEC_DEFAULT_DR: s.H = 16; s.langevin = 0; s.dr = 1; s.obs_layout = OBS_DEFAULT; s.obs_dim = 82; break;
        case ORACLE_SPEC_RAPTOR:     s.H = 1;  s.langevin = 1; s.dr = 0; s.obs_layout = OBS_RAPTOR;  s.obs_dim = 22; break;
        case ORACLE_SPEC_TEACHER:    s.H = 1;  s.langevin = 1; s.dr = 0; s.obs_layout = OBS_TEACHER; s.obs_dim = 26; break;
        case ORACLE_SPEC_RAPTOR_DR:  s.H = 1;  s.langevin = 1; s.dr = 1; s.obs_layout = OBS_RAPTOR;  s.obs_dim = 22; break;
        case ORACLE_SPEC_TEACHER_DR: s.H = 1;  s.langevin = 1; s.dr = 1; s.obs_layout = OBS_TEACHER; s.obs_dim = 26; break;
        default: fprintf(stderr, "l2f_oracle: bad spec %d\n", spec); abort();
    }
    return s;
}
int oracle_params_dim(void){ return ORACLE_PARAMS_DIM; }
int oracle_state_dim(int spec){ return 44 + 4 * get_spec(spec).H; }
int oracle_observation_dim(int spec){ return get_spec(spec).obs_dim; }
int oracle_action_history_length(int spec){ return get_spec(spec).H; }

/* ---------------------------------------------------------------------------------------------
 * RNG: INC/random/operations_generic.h:16-18 (init), :26-31 (xorshift), :52-58 (uniform),
 * :59-71 (Box-Muller normal, cosine branch); std==0 shortcut from INC/random/operations_cpu.h:39-41.
 * MAX_INDEX = SIZE_MAX (INC/devices/cpu.h:31) so (float)MAX_INDEX == 2^64.
 * ------------------------------------------------------------------------------------------- */
static const float RNG_MAX_F = 18446744073709551616.0f;
static const float PI_F = (float)3.141592653589793238462643383279502884L; /* INC/math/operations_generic.h:14 */
uint64_t oracle_rng_init(uint64_t seed){ return 0xAAAAAAAAull + seed; }
static inline void rng_next(uint64_t* s){ *s ^= (*s << 13); *s ^= (*s >> 17); *s ^= (*s << 5); }
static inline float rng_uniform(uint64_t* s, float lo, float hi){
    rng_next(s);
    return ((float)*s / RNG_MAX_F) * (hi - lo) + lo;
}
static inline float rng_normal(uint64_t* s, float mean, float std){
    if(std == 0){ return mean; }
    rng_next(s);
    float u1 = (float)*s / RNG_MAX_F;
    rng_next(s);
    float u2 = (float)*s / RNG_MAX_F;
    float x = (float)sqrt(-2.0 * logf(u1));
    float y = (float)(2.0 * PI_F * u2);
    float z = x * cosf(y);
    return z * std + mean;
}
float oracle_rng_uniform(uint64_t* state, float lo, float hi){ return rng_uniform(state, lo, hi); }
float oracle_rng_normal(uint64_t* state, float mean, float std){ return rng_normal(state, mean, std); }

static inline float clampf(float x, float lo, float hi){ return x < lo ? lo : (x > hi ? hi : x); } /* std::clamp, INC/math/operations_cpu.h:88-90 */

/* ---------------------------------------------------------------------------------------------
 * nominal parameters: L2F/parameters/dynamics/crazyflie.h:10-123, L2F/parameters/default.h:34-134,
 * L2F/parameters/init/default.h:22-31 (init_90_deg)
 * ------------------------------------------------------------------------------------------- */
void oracle_nominal_parameters(int spec, float* p){
    spec_t sp = get_spec(spec);
    memset(p, 0, sizeof(float) * ORACLE_PARAMS_DIM);
    const float pos[4][3] = {{0.028f, -0.028f, 0}, {-0.028f, -0.028f, 0}, {-0.028f, 0.028f, 0}, {0.028f, 0.028f, 0}};
    const float tdir[4] = {-1, +1, -1, +1};
    for(int i = 0; i < 4; i++){
        for(int j = 0; j < 3; j++) p[P_ROTOR_POS + 3*i + j] = pos[i][j];
        p[P_THRUST_DIR + 3*i + 2] = 1;
        p[P_TORQUE_DIR + 3*i + 2] = tdir[i];
        p[P_THRUST_COEF + 3*i + 0] = (float)0.00352526;
        p[P_THRUST_COEF + 3*i + 1] = (float)0.01437313;
        p[P_THRUST_COEF + 3*i + 2] = (float)0.09223048;
        p[P_TORQUE_CONST + i] = (float)4.665e-3;
        p[P_TAU_RISE + i] = (float)0.05545454545454546;
        p[P_TAU_FALL + i] = (float)0.24939393939393945;
    }
    p[P_MASS] = (float)(0.027 + 0.0017 + 0.0003 + 0.0016);
    p[P_GRAVITY + 2] = (float)-9.81;
    p[P_J + 0] = (float)9.416556729130406e-06; p[P_J + 4] = (float)9.644051701582312e-06; p[P_J + 8] = (float)1.745951732253285e-05;
    p[P_JINV + 0] = (float)106195.93007988465; p[P_JINV + 4] = (float)103690.85846314249; p[P_JINV + 8] = (float)57275.35197719487;
    p[P_HOVER] = (float)0.7261389721508553;
    p[P_ACT_MIN] = 0; p[P_ACT_MAX] = 1;
    p[P_DT] = (float)(1.0 / ((float)100));
    p[P_INIT_GUIDANCE] = (float)0.1; p[P_INIT_MAX_POS] = (float)0.5; p[P_INIT_MAX_ANGLE] = (float)1.5707963267948966;
    p[P_INIT_MAX_LINVEL] = 1; p[P_INIT_MAX_ANGVEL] = 1; p[P_INIT_REL_RPM] = 1; p[P_INIT_MIN_RPM] = -1; p[P_INIT_MAX_RPM] = 0;
    p[P_RW_NONNEG] = 0; p[P_RW_SCALE] = 1; p[P_RW_CONSTANT] = (float)0.5; p[P_RW_TERM_PENALTY] = -100; p[P_RW_POSITION] = 1;
    p[P_RW_ORIENTATION] = (float)0.1; p[P_RW_DACTION] = 1;
    p[P_TERM_ENABLED] = 1; p[P_TERM_POS] = 1; p[P_TERM_LINVEL] = 2; p[P_TERM_ANGVEL] = 35; p[P_TERM_POS_INT] = 10000; p[P_TERM_ORI_INT] = 50000;
    if(spec == ORACLE_SPEC_DEFAULT_DR){ /* L2F/parameters/default.h:88-104: only the DR-enabled default factory carries these ranges */
        p[P_DR_T2W_MIN] = (float)1.5; p[P_DR_T2W_MAX] = (float)5.0; p[P_DR_T2I_MIN] = (float)0.001; p[P_DR_T2I_MAX] = (float)0.100;
        p[P_DR_MASS_MIN] = (float)0.02; p[P_DR_MASS_MAX] = (float)5.00; p[P_DR_MASS_SIZE_DEV] = (float)0.1;
        p[P_DR_KQ_MIN] = (float)0.005; p[P_DR_KQ_MAX] = (float)0.05; p[P_DR_DIST_FORCE_MAX] = (float)0.1;
    }
    (void)sp;
    p[P_TRAJ_MIX0] = (float)0.5; p[P_TRAJ_MIX1] = (float)0.5;
    p[P_LANGEVIN_GAMMA] = 1; p[P_LANGEVIN_OMEGA] = 2; p[P_LANGEVIN_SIGMA] = (float)0.5; p[P_LANGEVIN_ALPHA] = (float)0.01;
}

/* ---------------------------------------------------------------------------------------------
 * sample_initial_parameters: L2F/operations_generic/10_sample_initial_parameters.h:20-23 (copy),
 * :35-201 (domain randomisation), :30-34 (size-deviation factor).  All DR options are either all on
 * or all off (L2F/parameters/default.h:16-26).
 * ------------------------------------------------------------------------------------------- */
static void fail(const char* msg){ fprintf(stderr, "l2f_oracle: %s\n", msg); exit(1); } /* utils::assert_exit */
void oracle_sample_initial_parameters(int spec, const float* env_p, uint64_t* rng, float* p){
    spec_t sp = get_spec(spec);
    memmove(p, env_p, sizeof(float) * ORACLE_PARAMS_DIM);
    if(!sp.dr){
        /* the reference asserts that every range is zero when its option is off (:74,92,122,132,166,181,196-199) */
        for(int i = P_DR_T2W_MIN; i <= P_DR_DIST_FORCE_MAX; i++){
            if(i == P_DR_ORI_OFFSET || i == P_DR_T2W_MAX) continue; /* not checked by the reference */
            if(p[i] != 0) fail("domain randomization ranges must be 0 when the options are disabled");
        }
        return;
    }
    float thrust_to_weight_nominal;
    {
        float max_action = p[P_ACT_MAX];
        float max_thrust_nominal = 0;
        for(int r = 0; r < 4; r++){
            max_thrust_nominal += p[P_THRUST_COEF + 3*r + 0] + p[P_THRUST_COEF + 3*r + 1] * max_action + p[P_THRUST_COEF + 3*r + 2] * max_action * max_action;
        }
        float gravity_norm = sqrtf(p[P_GRAVITY] * p[P_GRAVITY] + p[P_GRAVITY+1] * p[P_GRAVITY+1] + p[P_GRAVITY+2] * p[P_GRAVITY+2]);
        thrust_to_weight_nominal = max_thrust_nominal / (p[P_MASS] * gravity_norm);
    }
    if(!(p[P_DR_T2W_MIN] < p[P_DR_T2W_MAX])) fail("thrust_to_weight max should be larger than min");
    if(!(p[P_DR_T2W_MIN] >= 1.5)) fail("thrust_to_weight min should be >= 1.5");
    float thrust_to_weight = rng_uniform(rng, p[P_DR_T2W_MIN], p[P_DR_T2W_MAX]);
    float factor_thrust_to_weight = thrust_to_weight / thrust_to_weight_nominal;

    if(!(p[P_DR_MASS_MIN] < p[P_DR_MASS_MAX])) fail("mass max should be larger than min");
    float relative_size_min = cbrtf(p[P_DR_MASS_MIN]);
    float relative_size_max = cbrtf(p[P_DR_MASS_MAX]);
    float size_new = rng_uniform(rng, relative_size_min, relative_size_max);
    float mass_new = size_new * size_new * size_new;
    mass_new = clampf(mass_new, p[P_DR_MASS_MIN], p[P_DR_MASS_MAX]);
    float scale_relative = cbrtf(mass_new / p[P_MASS]);
    float factor_mass = mass_new / p[P_MASS];
    p[P_MASS] = mass_new;

    float factor_thrust_coefficients = factor_thrust_to_weight * factor_mass;
    for(int r = 0; r < 4; r++) for(int o = 0; o < 3; o++) p[P_THRUST_COEF + 3*r + o] *= factor_thrust_coefficients;

    float torque_to_inertia_factor;
    {
        float gravity_norm = sqrtf(p[P_GRAVITY] * p[P_GRAVITY] + p[P_GRAVITY+1] * p[P_GRAVITY+1] + p[P_GRAVITY+2] * p[P_GRAVITY+2]);
        float max_thrust = thrust_to_weight * p[P_MASS] * gravity_norm / 4;
        float first_rotor_distance_nominal = fabsf(p[P_ROTOR_POS]);
        float max_torque = (float)(first_rotor_distance_nominal * 1.414213562373095 * max_thrust);
        float x_inertia = p[P_J];
        float torque_to_inertia_nominal = max_torque / x_inertia;
        if(!(p[P_DR_T2I_MIN] < p[P_DR_T2I_MAX])) fail("torque_to_inertia max should be larger than min");
        float torque_to_inertia = rng_uniform(rng, p[P_DR_T2I_MIN], p[P_DR_T2I_MAX]);
        torque_to_inertia_factor = torque_to_inertia / torque_to_inertia_nominal;
    }
    if(p[P_DR_MASS_SIZE_DEV] == 0) fail("mass_size_deviation should be != 0");
    float size_factor;
    {   /* _sample_domain_randomization_factor: normal(mean = -range, std = range) -- sic, :31 */
        float range = p[P_DR_MASS_SIZE_DEV];
        float factor = rng_normal(rng, -range, range);
        size_factor = factor < 0 ? 1 / (1 - factor) : 1 + factor;
    }
    float rotor_distance_factor = scale_relative * size_factor;
    {
        float inertia_factor = torque_to_inertia_factor / rotor_distance_factor;
        for(int a = 0; a < 3; a++){
            p[P_J + 4*a] /= inertia_factor;
            p[P_JINV + 4*a] *= inertia_factor;
        }
        for(int r = 0; r < 4; r++) for(int a = 0; a < 3; a++) p[P_ROTOR_POS + 3*r + a] *= rotor_distance_factor;
        float max_rotor_distance = 0;
        for(int r = 0; r < 4; r++){
            const float* rp = p + P_ROTOR_POS + 3*r;
            float d = sqrtf(rp[0]*rp[0] + rp[1]*rp[1] + rp[2]*rp[2]);
            if(d > max_rotor_distance) max_rotor_distance = d;
        }
        p[P_TERM_POS] = max_rotor_distance * 20;
        p[P_INIT_MAX_POS] = max_rotor_distance * 10;
    }
    if(p[P_DR_KQ_MIN] == 0 || p[P_DR_KQ_MAX] == 0) fail("rotor_torque_constant range should be != 0");
    {
        float kq = rng_uniform(rng, p[P_DR_KQ_MIN], p[P_DR_KQ_MAX]);
        for(int r = 0; r < 4; r++) p[P_TORQUE_CONST + r] = kq;
    }
    if(p[P_DR_DIST_FORCE_MAX] == 0) fail("disturbance_force_max should be != 0");
    {
        float surplus = (float)(thrust_to_weight - 1.0);
        if(surplus < 0) surplus = 0;
        float multiple = rng_uniform(rng, 0.0f, surplus * p[P_DR_DIST_FORCE_MAX]);
        float std = multiple * thrust_to_weight * p[P_MASS] / 3;
        p[P_DIST_FORCE_MEAN] = 0;
        p[P_DIST_FORCE_STD] = std;
    }
    if(p[P_DR_TAU_RISE_MIN] == 0 || p[P_DR_TAU_RISE_MAX] == 0 || p[P_DR_TAU_FALL_MIN] == 0 || p[P_DR_TAU_FALL_MAX] == 0) fail("rotor_time_constant ranges should be != 0");
    {
        float rising = rng_uniform(rng, p[P_DR_TAU_RISE_MIN], p[P_DR_TAU_RISE_MAX]);
        float falling = rng_uniform(rng, p[P_DR_TAU_FALL_MIN], p[P_DR_TAU_FALL_MAX]);
        for(int r = 0; r < 4; r++){ p[P_TAU_RISE + r] = rising; p[P_TAU_FALL + r] = falling; }
    }
}

/* ---------------------------------------------------------------------------------------------
 * initial_state: L2F/operations_generic/20_initial_state.h:21-121
 * ------------------------------------------------------------------------------------------- */
void oracle_initial_state(int spec, const float* p, float* s){
    spec_t sp = get_spec(spec);
    memset(s, 0, sizeof(float) * (44 + 4 * sp.H));
    s[S_ORI] = 1;
    for(int i = 0; i < 4; i++) s[S_RPM + i] = p[P_HOVER] * (p[P_ACT_MAX] - p[P_ACT_MIN]) + p[P_ACT_MIN];
    for(int h = 0; h < sp.H; h++) for(int i = 0; i < 4; i++)
        s[S_HIST + 4*h + i] = (s[S_RPM + i] - p[P_ACT_MIN]) / (p[P_ACT_MAX] - p[P_ACT_MIN]) * 2 - 1;
    s[S_CURRENT_STEP] = 0;
    s[S_TRAJ_TYPE(sp.H)] = 0; /* POSITION */
}

/* ---------------------------------------------------------------------------------------------
 * sample_initial_state: L2F/operations_generic/30_sample_initial_state.h:21-40 (orientation),
 * :42-85 (base), :87-94, :104-113, :124-152 (random force), :161-183 (rotors), :185-196 (history),
 * :198-228 (trajectory)
 * ------------------------------------------------------------------------------------------- */
void oracle_sample_initial_state(int spec, const float* p, uint64_t* rng, float* s){
    spec_t sp = get_spec(spec);
    memset(s, 0, sizeof(float) * (44 + 4 * sp.H));
    int guidance = rng_uniform(rng, 0.0f, 1.0f) < p[P_INIT_GUIDANCE];
    if(!guidance){
        for(int i = 0; i < 3; i++) s[S_POS + i] = rng_uniform(rng, -p[P_INIT_MAX_POS], p[P_INIT_MAX_POS]);
    }
    if(p[P_INIT_MAX_ANGLE] > 0 && !guidance){
        float u = rng_uniform(rng, 0.0f, 1.0f);
        float v = rng_uniform(rng, 0.0f, 1.0f);
        float phi = (float)(2.0 * PI_F * u);
        float cos_theta = (float)(1.0 - 2.0 * v);
        float sin_theta = (float)sqrt(1.0 - cos_theta * cos_theta);
        float x = sin_theta * cosf(phi);
        float y = sin_theta * sinf(phi);
        float z = cos_theta;
        float angle = rng_uniform(rng, 0.0f, 1.0f); /* the limit only gates, :31,60 */
        float half = (float)(0.5 * angle);
        float sn = sinf(half);
        s[S_ORI + 0] = cosf(half); s[S_ORI + 1] = x * sn; s[S_ORI + 2] = y * sn; s[S_ORI + 3] = z * sn;
    }
    else{
        s[S_ORI] = 1;
    }
    if(!guidance){
        for(int i = 0; i < 3; i++) s[S_LINVEL + i] = rng_uniform(rng, -p[P_INIT_MAX_LINVEL], p[P_INIT_MAX_LINVEL]);
        for(int i = 0; i < 3; i++) s[S_ANGVEL + i] = rng_uniform(rng, -p[P_INIT_MAX_ANGVEL], p[P_INIT_MAX_ANGVEL]);
    }
    /* last_action = 0; angular velocity history = angular velocity */
    for(int i = 0; i < 3; i++) s[S_ANGVEL_HIST + i] = s[S_ANGVEL + i];
    /* random force / torque */
    for(int i = 0; i < 3; i++) s[S_FORCE + i] = rng_normal(rng, p[P_DIST_FORCE_MEAN], p[P_DIST_FORCE_STD]);
    s[S_TORQUE + 0] = rng_normal(rng, p[P_DIST_TORQUE_MEAN], p[P_DIST_TORQUE_STD]);
    s[S_TORQUE + 1] = rng_normal(rng, p[P_DIST_TORQUE_MEAN], p[P_DIST_TORQUE_STD]);
    s[S_TORQUE + 2] = rng_normal(rng, p[P_DIST_TORQUE_MEAN], p[P_DIST_TORQUE_STD] / 100);
    /* rotors */
    float min_rpm, max_rpm;
    if(p[P_INIT_REL_RPM] != 0){
        min_rpm = (p[P_INIT_MIN_RPM] + 1) / 2 * (p[P_ACT_MAX] - p[P_ACT_MIN]) + p[P_ACT_MIN];
        max_rpm = (p[P_INIT_MAX_RPM] + 1) / 2 * (p[P_ACT_MAX] - p[P_ACT_MIN]) + p[P_ACT_MIN];
    }
    else{
        min_rpm = p[P_INIT_MIN_RPM] < 0 ? p[P_ACT_MIN] : p[P_INIT_MIN_RPM];
        max_rpm = p[P_INIT_MAX_RPM] < 0 ? p[P_ACT_MAX] : p[P_INIT_MAX_RPM];
        if(max_rpm > p[P_ACT_MAX]) max_rpm = p[P_ACT_MAX];
        if(min_rpm > max_rpm) min_rpm = max_rpm;
    }
    for(int i = 0; i < 4; i++) s[S_RPM + i] = rng_uniform(rng, min_rpm, max_rpm);
    s[S_CURRENT_STEP] = 0;
    for(int h = 0; h < sp.H; h++) for(int i = 0; i < 4; i++)
        s[S_HIST + 4*h + i] = (s[S_RPM + i] - p[P_ACT_MIN]) / (p[P_ACT_MAX] - p[P_ACT_MIN]) * 2 - 1;
    if(sp.langevin){
        float threshold = rng_uniform(rng, 0.0f, 1.0f);
        float acc = 0;
        int type = 0;
        for(int t = 0; t < 2; t++){
            acc += p[P_TRAJ_MIX0 + t];
            if(threshold < acc){ type = t; break; }
        }
        s[S_TRAJ_TYPE(sp.H)] = (float)type; /* langevin block already zero */
    }
}

/* ---------------------------------------------------------------------------------------------
 * get_desired_state: L2F/operations_generic/35_get_desired_state.h:17-50
 * ------------------------------------------------------------------------------------------- */
static void desired_state(const spec_t* sp, const float* s, float dpos[3], float dvel[3]){
    if((int)s[S_TRAJ_TYPE(sp->H)] == 1){ /* LANGEVIN */
        const float* l = s + S_LANGEVIN(sp->H);
        for(int i = 0; i < 3; i++){ dpos[i] = l[i]; dvel[i] = l[3 + i]; }
    }
    else{
        for(int i = 0; i < 3; i++){ dpos[i] = 0; dvel[i] = 0; }
    }
}

/* ---------------------------------------------------------------------------------------------
 * observe: L2F/operations_generic/40_observe.h (Position :42-60, RotationMatrix :81-106,
 * LinearVelocity :108-125, AngularVelocityDelayed :221-254, RotorSpeeds :256-268,
 * ActionHistory :270-291, TrajectoryTrackingPosition :387-407, TrajectoryTrackingLinearVelocity :409-429)
 * ------------------------------------------------------------------------------------------- */
static void observe_impl(const spec_t* sp, const float* p, const float* s, uint64_t* rng, float* o){
    int k = 0;
    float dpos[3], dvel[3];
    desired_state(sp, s, dpos, dvel);
    /* position */
    for(int i = 0; i < 3; i++){
        float noise = rng_normal(rng, 0.0f, p[P_NOISE_POS]);
        if(sp->obs_layout == OBS_RAPTOR) o[k++] = s[S_POS + i] + noise;
        else o[k++] = s[S_POS + i] - dpos[i] + noise;
    }
    /* rotation matrix */
    const float* q = s + S_ORI;
    o[k + 0] = (1 - 2*q[2]*q[2] - 2*q[3]*q[3]);
    o[k + 1] = (    2*q[1]*q[2] - 2*q[0]*q[3]);
    o[k + 2] = (    2*q[1]*q[3] + 2*q[0]*q[2]);
    o[k + 3] = (    2*q[1]*q[2] + 2*q[0]*q[3]);
    o[k + 4] = (1 - 2*q[1]*q[1] - 2*q[3]*q[3]);
    o[k + 5] = (    2*q[2]*q[3] - 2*q[0]*q[1]);
    o[k + 6] = (    2*q[1]*q[3] - 2*q[0]*q[2]);
    o[k + 7] = (    2*q[2]*q[3] + 2*q[0]*q[1]);
    o[k + 8] = (1 - 2*q[1]*q[1] - 2*q[2]*q[2]);
    for(int i = 0; i < 9; i++){
        float noise = rng_normal(rng, 0.0f, p[P_NOISE_ORI]);
        o[k + i] += noise;
    }
    k += 9;
    /* linear velocity */
    for(int i = 0; i < 3; i++){
        float noise = rng_normal(rng, 0.0f, p[P_NOISE_LINVEL]);
        if(sp->obs_layout == OBS_RAPTOR) o[k++] = s[S_LINVEL + i] + noise;
        else o[k++] = s[S_LINVEL + i] - dvel[i] + noise;
    }
    /* angular velocity (delay 0) */
    for(int i = 0; i < 3; i++){
        float noise = rng_normal(rng, 0.0f, p[P_NOISE_ANGVEL]);
        o[k++] = s[S_ANGVEL + i] + noise;
    }
    /* action history, most recent first */
    {
        int H = sp->H;
        int hist_len = H; /* default: H entries; raptor/teacher: 1 == H */
        int current = (int)s[S_CURRENT_STEP];
        current = current == 0 ? H - 1 : current - 1;
        for(int step = 0; step < hist_len; step++){
            for(int a = 0; a < 4; a++) o[k++] = s[S_HIST + 4*current + a];
            current = current == 0 ? H - 1 : current - 1;
        }
    }
    if(sp->obs_layout == OBS_TEACHER){
        for(int a = 0; a < 4; a++) o[k++] = (s[S_RPM + a] - p[P_ACT_MIN]) / (p[P_ACT_MAX] - p[P_ACT_MIN]) * 2 - 1;
    }
    if(k != sp->obs_dim){ fprintf(stderr, "observe: %d != %d\n", k, sp->obs_dim); abort(); }
}
void oracle_observe(int spec, const float* p, const float* s, uint64_t* rng, float* obs){
    spec_t sp = get_spec(spec);
    observe_impl(&sp, p, s, rng, obs);
}

/* ---------------------------------------------------------------------------------------------
 * dynamics: L2F/operations_generic/60_dynamics.h:18-72 (base), :87-99 (random force), :100-111 (rotors);
 * helpers L2F/quaternion_helper.h:11-18,22-35; INC/utils/generic/vector_operations.h
 * integrated vector x[17] = pos3 quat4 vel3 omega3 rpm4
 * ------------------------------------------------------------------------------------------- */
enum { X_POS = 0, X_ORI = 3, X_VEL = 7, X_OMEGA = 10, X_RPM = 13, X_DIM = 17 };
static void dynamics(const float* p, const float* x, const float* force, const float* torque_dist, const float* setpoint, float* dx){
    float thrust[3] = {0, 0, 0}, torque[3] = {0, 0, 0};
    for(int r = 0; r < 4; r++){
        float rpm = x[X_RPM + r];
        const float* c = p + P_THRUST_COEF + 3*r;
        float thrust_magnitude = c[0] + c[1] * rpm + c[2] * rpm * rpm;
        float rotor_thrust[3];
        for(int i = 0; i < 3; i++) rotor_thrust[i] = p[P_THRUST_DIR + 3*r + i] * thrust_magnitude;
        for(int i = 0; i < 3; i++) thrust[i] += rotor_thrust[i];
        float sc = thrust_magnitude * p[P_TORQUE_CONST + r];
        for(int i = 0; i < 3; i++) torque[i] += p[P_TORQUE_DIR + 3*r + i] * sc;
        const float* rp = p + P_ROTOR_POS + 3*r;
        torque[0] += rp[1]*rotor_thrust[2] - rp[2]*rotor_thrust[1];
        torque[1] += rp[2]*rotor_thrust[0] - rp[0]*rotor_thrust[2];
        torque[2] += rp[0]*rotor_thrust[1] - rp[1]*rotor_thrust[0];
    }
    for(int i = 0; i < 3; i++) dx[X_POS + i] = x[X_VEL + i];
    const float* q = x + X_ORI; const float* w = x + X_OMEGA;
    dx[X_ORI + 0] = -q[1]*w[0] - q[2]*w[1] - q[3]*w[2];
    dx[X_ORI + 1] =  q[0]*w[0] + q[2]*w[2] - q[3]*w[1];
    dx[X_ORI + 2] =  q[0]*w[1] + q[3]*w[0] - q[1]*w[2];
    dx[X_ORI + 3] =  q[0]*w[2] + q[1]*w[1] - q[2]*w[0];
    for(int i = 0; i < 4; i++) dx[X_ORI + i] *= 0.5f;
    {   /* rotate_vector_by_quaternion(q, thrust) */
        float var[3], out[3];
        var[0] = q[2]*thrust[2] - q[3]*thrust[1];
        var[1] = q[3]*thrust[0] - q[1]*thrust[2];
        var[2] = q[1]*thrust[1] - q[2]*thrust[0];
        for(int i = 0; i < 3; i++) var[i] *= 2;
        out[0] = q[2]*var[2] - q[3]*var[1];
        out[1] = q[3]*var[0] - q[1]*var[2];
        out[2] = q[1]*var[1] - q[2]*var[0];
        for(int i = 0; i < 3; i++) out[i] += var[i] * q[0];
        for(int i = 0; i < 3; i++) out[i] += thrust[i];
        float inv_mass = 1 / p[P_MASS];
        for(int i = 0; i < 3; i++) out[i] *= inv_mass;
        for(int i = 0; i < 3; i++) out[i] += p[P_GRAVITY + i];
        for(int i = 0; i < 3; i++) dx[X_VEL + i] = out[i];
    }
    {
        float v[3], v2[3];
        for(int i = 0; i < 3; i++){ v[i] = 0; for(int j = 0; j < 3; j++) v[i] += p[P_J + 3*i + j] * w[j]; }
        v2[0] = w[1]*v[2] - w[2]*v[1];
        v2[1] = w[2]*v[0] - w[0]*v[2];
        v2[2] = w[0]*v[1] - w[1]*v[0];
        for(int i = 0; i < 3; i++) v[i] = torque[i] - v2[i];
        for(int i = 0; i < 3; i++){ float a = 0; for(int j = 0; j < 3; j++) a += p[P_JINV + 3*i + j] * v[j]; dx[X_OMEGA + i] = a; }
    }
    /* StateRandomForce */
    for(int i = 0; i < 3; i++) dx[X_VEL + i] += force[i] / p[P_MASS];
    {
        float aa[3];
        for(int i = 0; i < 3; i++){ aa[i] = 0; for(int j = 0; j < 3; j++) aa[i] += p[P_JINV + 3*i + j] * torque_dist[j]; }
        for(int i = 0; i < 3; i++) dx[X_OMEGA + i] += aa[i];
    }
    /* StateRotors (not closed form) */
    for(int r = 0; r < 4; r++){
        float tau = setpoint[r] >= x[X_RPM + r] ? p[P_TAU_RISE + r] : p[P_TAU_FALL + r];
        dx[X_RPM + r] = (setpoint[r] - x[X_RPM + r]) * 1 / tau;
    }
}

/* ---------------------------------------------------------------------------------------------
 * step: L2F/operations_generic.h:94-130; rk4: INC/utils/generic/integrators.h:18-50;
 * axpy: L2F/operations_generic/50_state_algebra.h:22-53; post_integration:
 * L2F/operations_generic/70_post_integration.h:20-37,40-48,59-81,85-100,102-111,112-125,127-170
 * ------------------------------------------------------------------------------------------- */
static float step_impl(const spec_t* sp, const float* p, const float* s, const float* a, uint64_t* rng, float* n){
    const int H = sp->H;
    float setpoint[4];
    for(int i = 0; i < 4; i++){
        float half_range = (p[P_ACT_MAX] - p[P_ACT_MIN]) / 2;
        float action_noisy = a[i];
        action_noisy += rng_normal(rng, 0.0f, p[P_ACTION_NOISE]);
        action_noisy = clampf(action_noisy, -1.0f, 1.0f);
        setpoint[i] = action_noisy * half_range + p[P_ACT_MIN] + half_range;
    }
    float x[X_DIM], k1[X_DIM], k2[X_DIM], k3[X_DIM], k4[X_DIM], tmp[X_DIM], xn[X_DIM];
    for(int i = 0; i < 13; i++) x[i] = s[i];
    for(int i = 0; i < 4; i++) x[X_RPM + i] = s[S_RPM + i];
    const float* force = s + S_FORCE; const float* torque = s + S_TORQUE;
    const float dt = p[P_DT];
    dynamics(p, x, force, torque, setpoint, k1);
    for(int i = 0; i < X_DIM; i++){ tmp[i] = x[i]; tmp[i] += (dt / 2) * k1[i]; }
    dynamics(p, tmp, force, torque, setpoint, k2);
    for(int i = 0; i < X_DIM; i++){ tmp[i] = x[i]; tmp[i] += (dt / 2) * k2[i]; }
    dynamics(p, tmp, force, torque, setpoint, k3);
    for(int i = 0; i < X_DIM; i++){ tmp[i] = x[i]; tmp[i] += dt * k3[i]; }
    dynamics(p, tmp, force, torque, setpoint, k4);
    for(int i = 0; i < X_DIM; i++){
        xn[i] = x[i];
        xn[i] += (dt / 6) * k1[i];
        xn[i] += (dt / 3) * k2[i];
        xn[i] += (dt / 3) * k3[i];
        xn[i] += (dt / 6) * k4[i];
    }
    /* next_state = state (all non-integrated fields carried over), then post_integration */
    memmove(n, s, sizeof(float) * (44 + 4 * H));
    for(int i = 0; i < 13; i++) n[i] = xn[i];
    for(int i = 0; i < 4; i++) n[S_RPM + i] = xn[X_RPM + i];
    {
        float norm = 0;
        for(int i = 0; i < 4; i++) norm += n[S_ORI + i] * n[S_ORI + i];
        norm = sqrtf(norm);
        for(int i = 0; i < 4; i++) n[S_ORI + i] /= norm;
        for(int i = 0; i < 3; i++){
            n[S_POS + i] = clampf(n[S_POS + i], -100000.0f, 100000.0f);
            n[S_LINVEL + i] = clampf(n[S_LINVEL + i], -100000.0f, 100000.0f);
            n[S_ANGVEL + i] = clampf(n[S_ANGVEL + i], -100000.0f, 100000.0f);
        }
    }
    for(int i = 0; i < 4; i++) n[S_LAST_ACTION + i] = a[i];
    for(int i = 0; i < 3; i++) n[S_ANGVEL_HIST + i] = n[S_ANGVEL + i];
    for(int i = 0; i < 4; i++) n[S_RPM + i] = clampf(n[S_RPM + i], p[P_ACT_MIN], p[P_ACT_MAX]);
    {
        int cs = (int)s[S_CURRENT_STEP];
        for(int i = 0; i < 4; i++) n[S_HIST + 4*cs + i] = a[i];
        n[S_CURRENT_STEP] = (float)((cs + 1) % H);
    }
    if(sp->langevin && (int)s[S_TRAJ_TYPE(H)] == 1){
        const float gamma = p[P_LANGEVIN_GAMMA], omega = p[P_LANGEVIN_OMEGA], sigma = p[P_LANGEVIN_SIGMA], alpha = p[P_LANGEVIN_ALPHA];
        const float sqrt_dt = sqrtf(dt);
        const float* ls = s + S_LANGEVIN(H);
        float* ln = n + S_LANGEVIN(H);
        for(int d = 0; d < 3; d++){
            const float x_prev = ls[6 + d];
            const float v_prev = ls[9 + d];
            const float dW = sqrt_dt * rng_normal(rng, 0.0f, 1.0f);
            const float v_next = v_prev + (-gamma * v_prev - omega * omega * x_prev) * dt + sigma * dW;
            const float x_next = x_prev + v_next * dt;
            ln[6 + d] = x_next;
            ln[9 + d] = v_next;
            const float v_smooth_prev = ls[3 + d];
            const float v_smooth = alpha * v_next + (1.0f - alpha) * v_smooth_prev;
            const float x_smooth_prev = ls[d];
            const float x_smooth = x_smooth_prev + v_smooth * dt;
            ln[d] = x_smooth;
            ln[3 + d] = v_smooth;
        }
    }
    return dt;
}
float oracle_step(int spec, const float* p, const float* s, const float* a, uint64_t* rng, float* s_next){
    spec_t sp = get_spec(spec);
    float tmp[MAX_STATE_DIM];
    float r = step_impl(&sp, p, s, a, rng, tmp);
    memcpy(s_next, tmp, sizeof(float) * (44 + 4 * sp.H));
    return r;
}

/* ---------------------------------------------------------------------------------------------
 * terminated: L2F/operations_generic.h:142-166
 * ------------------------------------------------------------------------------------------- */
static int terminated_impl(const float* p, const float* s){
    if(p[P_TERM_ENABLED] != 0){
        for(int i = 0; i < 3; i++){
            if(fabsf(s[S_POS + i]) > p[P_TERM_POS] || fabsf(s[S_LINVEL + i]) > p[P_TERM_LINVEL] || fabsf(s[S_ANGVEL + i]) > p[P_TERM_ANGVEL]) return 1;
        }
    }
    return 0;
}
int oracle_terminated(int spec, const float* p, const float* s){ (void)spec; return terminated_impl(p, s); }

/* ---------------------------------------------------------------------------------------------
 * reward (Squared): L2F/parameters/reward_functions/squared/operations_generic.h:13-46,48-58,100-129
 * ------------------------------------------------------------------------------------------- */
static float reward_impl(const spec_t* sp, const float* p, const float* s, const float* a, const float* n){
    float dpos[3], dvel[3];
    desired_state(sp, s, dpos, dvel);
    float orientation_cost = 2 * acosf(1 - fabsf(s[S_ORI + 3]));
    float x = s[S_POS + 0] - dpos[0], y = s[S_POS + 1] - dpos[1], z = s[S_POS + 2] - dpos[2];
    float position_cost = sqrtf(x*x + y*y + z*z);
    if(p[P_RW_POSITION_CLIP] > 0) position_cost = position_cost < p[P_RW_POSITION_CLIP] ? position_cost : p[P_RW_POSITION_CLIP];
    float vx = s[S_LINVEL + 0] - dvel[0], vy = s[S_LINVEL + 1] - dvel[1], vz = s[S_LINVEL + 2] - dvel[2];
    float linear_vel_cost = sqrtf(vx*vx + vy*vy + vz*vz);
    float angular_vel_cost = sqrtf(s[S_ANGVEL + 0] * s[S_ANGVEL + 0] + s[S_ANGVEL + 1] * s[S_ANGVEL + 1] + s[S_ANGVEL + 2] * s[S_ANGVEL + 2]);
    float la[3], aa[3];
    for(int i = 0; i < 3; i++){ la[i] = n[S_LINVEL + i] - s[S_LINVEL + i]; aa[i] = n[S_ANGVEL + i] - s[S_ANGVEL + i]; }
    float linear_acc_cost = sqrtf(la[0]*la[0] + la[1]*la[1] + la[2]*la[2]) / p[P_DT];
    float angular_acc_cost = sqrtf(aa[0]*aa[0] + aa[1]*aa[1] + aa[2]*aa[2]) / p[P_DT];
    float acc = 0;
    for(int i = 0; i < 4; i++){
        float rel = (a[i] + 1.0f) / 2.0f;
        float d = rel - p[P_HOVER];
        acc += d * d;
    }
    float action_cost = sqrtf(acc);
    action_cost *= action_cost;
    float d_action_cost = 0;
    for(int i = 0; i < 4; i++){
        float d = a[i] - s[S_LAST_ACTION + i];
        d_action_cost += d * d;
    }
    d_action_cost = sqrtf(d_action_cost);
    float weighted = 0;
    weighted += p[P_RW_POSITION] * position_cost;
    weighted += p[P_RW_ORIENTATION] * orientation_cost;
    weighted += p[P_RW_LINVEL] * linear_vel_cost;
    weighted += p[P_RW_ANGVEL] * angular_vel_cost;
    weighted += p[P_RW_LINACC] * linear_acc_cost;
    weighted += p[P_RW_ANGACC] * angular_acc_cost;
    weighted += p[P_RW_ACTION] * action_cost;
    weighted += p[P_RW_DACTION] * d_action_cost;
    weighted += p[P_RW_POS_INTEGRAL] * 0.0f;
    int term = terminated_impl(p, n);
    float scaled = p[P_RW_SCALE] * weighted;
    float r;
    if(term){ r = p[P_RW_TERM_PENALTY]; }
    else{
        r = -scaled + p[P_RW_CONSTANT];
        r = (r > 0 || !(p[P_RW_NONNEG] != 0)) ? r : 0;
    }
    return r;
}
float oracle_reward(int spec, const float* p, const float* s, const float* a, const float* s_next, uint64_t* rng){
    (void)rng; spec_t sp = get_spec(spec); return reward_impl(&sp, p, s, a, s_next);
}

/* ---------------------------------------------------------------------------------------------
 * policies.  dense: INC/nn/layers/dense/operations_generic.h:94-108; GRU evaluate_step:
 * INC/nn/layers/gru/operations_generic.h:76-86,343-411 with helper_operations_generic.h:10-71 and the
 * generic matmul INC/containers/matrix/operations_generic.h:848-868; sigmoid/tanh:
 * INC/containers/tensor/operations_generic.h:378-383; standardize: INC/nn/layers/standardize/
 * operations_generic.h:67-84; sample_and_squash (Evaluation): INC/nn/layers/sample_and_squash/
 * operations_generic.h:148-194; PPO action sampling: INC/rl/components/on_policy_runner/
 * operations_generic_per_env.h:43-58 with log_prob INC/random/operations_generic.h:72-81.
 * GRU SEQUENCE_LENGTH = 500 (checkpoint.h layer_1 CONFIG / SRC/post_training/config.h:16).
 * ------------------------------------------------------------------------------------------- */
#define GRU_SEQUENCE_LENGTH 500
int oracle_policy_num_parameters(const oracle_policy_t* pol){
    int in = pol->input_dim, h = pol->hidden_dim, o = pol->output_dim;
    if(pol->arch == ORACLE_POLICY_RAPTOR_GRU) return h*in + h + 3*h*h + 3*h + 3*h*h + 3*h + h + o*h + o;
    return (pol->standardize ? 2*in : 0) + h*in + h + h*h + h + o*h + o + (pol->head == ORACLE_HEAD_PPO_GAUSSIAN ? 4 : 0);
}
static void dense(const float* W, const float* b, int out, int in, const float* x, float* y, int relu){
    for(int o = 0; o < out; o++){
        float acc = b[o];
        for(int i = 0; i < in; i++) acc += W[o*in + i] * x[i];
        y[o] = relu ? (acc > 0 ? acc : 0) : acc; /* math::max(x, 0) */
    }
}
static void raptor_forward(const oracle_policy_t* pol, const float* obs, float* h, int* step, int no_auto_reset, float* action){
    const int IN = pol->input_dim, HD = pol->hidden_dim, OUT = pol->output_dim;
    const float* W1 = pol->blob; const float* b1 = W1 + HD*IN;
    const float* Wih = b1 + HD; const float* bih = Wih + 3*HD*HD;
    const float* Whh = bih + 3*HD; const float* bhh = Whh + 3*HD*HD;
    const float* h0 = bhh + 3*HD; const float* W2 = h0 + HD; const float* b2 = W2 + OUT*HD;
    float x1[64], pre[192], npp[64], hn[64];
    dense(W1, b1, HD, IN, obs, x1, 1);
    if(!no_auto_reset && *step >= GRU_SEQUENCE_LENGTH){ memcpy(h, h0, sizeof(float)*HD); *step = 0; }
    for(int j = 0; j < 3*HD; j++){
        float acc = bhh[j];
        for(int k = 0; k < HD; k++) acc += h[k] * Whh[j*HD + k];
        pre[j] = acc;
    }
    for(int j = 0; j < HD; j++){ npp[j] = pre[2*HD + j]; pre[2*HD + j] = 0; }
    for(int j = 0; j < 3*HD; j++){
        float acc = pre[j] + bih[j];
        for(int k = 0; k < HD; k++) acc += x1[k] * Wih[j*HD + k];
        pre[j] = acc;
    }
    for(int j = 0; j < 2*HD; j++) pre[j] = 1 / (1 + expf(-pre[j]));
    for(int j = 0; j < HD; j++){
        float nn = pre[2*HD + j];
        nn += npp[j] * pre[j];
        nn = tanhf(nn);
        float z = pre[HD + j];
        float out = 1 - z;
        out *= nn;
        out += z * h[j];
        hn[j] = out;
    }
    dense(W2, b2, OUT, HD, hn, action, 0);
    memcpy(h, hn, sizeof(float)*HD);
    int new_step = *step + 1;
    if(!no_auto_reset && new_step >= GRU_SEQUENCE_LENGTH){ new_step = 0; memcpy(h, h0, sizeof(float)*HD); }
    *step = new_step;
}
static float normal_log_prob(float mean, float log_std, float value){
    float neg_log_sqrt_pi = (float)(-0.5 * logf(2 * PI_F));
    float diff = (value - mean);
    float std = expf(log_std);
    float pre_square = diff / std;
    return (float)(neg_log_sqrt_pi - log_std - 0.5 * pre_square * pre_square);
}
static void mlp_forward(const oracle_policy_t* pol, const float* obs, uint64_t* rng, float* action, float* mean_out, float* log_prob_out){
    const int IN = pol->input_dim, HD = pol->hidden_dim, OUT = pol->output_dim;
    const float* b = pol->blob;
    float x0[MAX_OBS_DIM], x1[256], x2[256], y[16];
    if(pol->standardize){
        const float* mean = b; const float* prec = b + IN; b += 2*IN;
        for(int i = 0; i < IN; i++){ float v = obs[i] - mean[i]; if(prec[i] != 0) v *= prec[i]; x0[i] = v; }
    }
    else{ memcpy(x0, obs, sizeof(float)*IN); }
    const float* W1 = b; const float* b1 = W1 + HD*IN; const float* W2 = b1 + HD; const float* b2 = W2 + HD*HD;
    const float* W3 = b2 + HD; const float* b3 = W3 + OUT*HD; const float* log_std = b3 + OUT;
    dense(W1, b1, HD, IN, x0, x1, 1);
    dense(W2, b2, HD, HD, x1, x2, 1);
    dense(W3, b3, OUT, HD, x2, y, 0);
    if(pol->head == ORACLE_HEAD_SQUASH_EVAL){
        for(int i = 0; i < OUT/2; i++) action[i] = tanhf(y[i]); /* Evaluation mode: sample = mean */
    }
    else if(pol->head == ORACLE_HEAD_SQUASH_SAMPLE){
        /* Mode<Rollout> / Default of sample_and_squash evaluate_per_sample (operations_generic.h:148-194): noise ~ N(0, 1) per action
         * dimension, log_std clamped to [-20, 2] (layer.h:31-32), action = tanh(mean + noise * exp(log_std)) */
        for(int i = 0; i < OUT/2; i++){
            float log_std = y[OUT/2 + i];
            float clipped = log_std < -20.0f ? -20.0f : (log_std > 2.0f ? 2.0f : log_std);
            float std = expf(clipped);
            float noise = rng_normal(rng, 0.0f, 1.0f);
            action[i] = tanhf(y[i] + noise * std);
        }
    }
    else if(pol->head == ORACLE_HEAD_PPO_GAUSSIAN){
        float lp = 0;
        for(int i = 0; i < OUT; i++){
            float std = expf(log_std[i]);
            float noisy = rng_normal(rng, y[i], std);
            lp += normal_log_prob(y[i], log_std[i], noisy);
            action[i] = noisy;
            if(mean_out) mean_out[i] = y[i];
        }
        if(log_prob_out) *log_prob_out = lp;
    }
    else{
        for(int i = 0; i < OUT; i++) action[i] = y[i];
    }
}
static void policy_forward(const oracle_policy_t* pol, const float* obs, float* h, int* step, int no_auto_reset, uint64_t* rng, float* action, float* mean_out, float* lp_out){
    if(pol->arch == ORACLE_POLICY_RAPTOR_GRU) raptor_forward(pol, obs, h, step, no_auto_reset, action);
    else mlp_forward(pol, obs, rng, action, mean_out, lp_out);
}
void oracle_policy_evaluate_step(const oracle_policy_t* pol, int N, const float* obs, int obs_ld, float* hidden, int* gru_step, int no_auto_reset,
                                 uint64_t* rng, float* actions, float* out_mean, float* out_log_prob){
    int adim = (pol->head == ORACLE_HEAD_SQUASH_EVAL || pol->head == ORACLE_HEAD_SQUASH_SAMPLE) ? pol->output_dim / 2 : pol->output_dim;
    for(int n = 0; n < N; n++){
        policy_forward(pol, obs + (size_t)n * obs_ld, hidden ? hidden + (size_t)n * pol->hidden_dim : NULL, gru_step ? gru_step + n : NULL, no_auto_reset,
                       rng ? rng + n : NULL, actions + (size_t)n * adim, out_mean ? out_mean + (size_t)n * adim : NULL, out_log_prob ? out_log_prob + n : NULL);
    }
}

/* ---------------------------------------------------------------------------------------------
 * closed-loop rollout, order of INC/rl/utils/evaluation/operations_generic.h:138-189:
 * observe -> evaluate_step -> step -> reward(state, action, next_state) -> terminated(next_state)
 * (no termination skip: every environment keeps stepping, as in the README loop R/README.md:94-99)
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int spec; const oracle_policy_t* pol; int N, T, n0, n1; const float* params; float* states_io; uint64_t* rng; float* hidden_io; int* gru_step_io; int no_auto_reset;
    float* out_states; float* out_obs; float* out_actions; float* out_rewards; unsigned char* out_term;
} rollout_job_t;
static void* rollout_range(void* arg){
    rollout_job_t* j = (rollout_job_t*)arg;
    spec_t sp = get_spec(j->spec);
    const int SD = 44 + 4 * sp.H, OBS = sp.obs_dim, N = j->N;
    const oracle_policy_t* pol = j->pol;
    const int HD = pol->hidden_dim;
    for(int n = j->n0; n < j->n1; n++){
        const float* p = j->params + (size_t)n * ORACLE_PARAMS_DIM;
        float s[MAX_STATE_DIM], nx[MAX_STATE_DIM], obs[MAX_OBS_DIM], a[8], h[64];
        memcpy(s, j->states_io + (size_t)n * SD, sizeof(float) * SD);
        uint64_t rng = j->rng[n];
        int gstep = j->gru_step_io ? j->gru_step_io[n] : 0;
        if(pol->arch == ORACLE_POLICY_RAPTOR_GRU){
            if(j->hidden_io) memcpy(h, j->hidden_io + (size_t)n * HD, sizeof(float) * HD);
            else memcpy(h, pol->blob + (HD*pol->input_dim + HD + 2*(3*HD*HD + 3*HD)), sizeof(float) * HD); /* h0 */
        }
        for(int t = 0; t < j->T; t++){
            if(j->out_states) memcpy(j->out_states + ((size_t)t * N + n) * SD, s, sizeof(float) * SD);
            observe_impl(&sp, p, s, &rng, obs);
            if(j->out_obs) memcpy(j->out_obs + ((size_t)t * N + n) * OBS, obs, sizeof(float) * OBS);
            policy_forward(pol, obs, h, &gstep, j->no_auto_reset, &rng, a, NULL, NULL);
            if(j->out_actions) memcpy(j->out_actions + ((size_t)t * N + n) * 4, a, sizeof(float) * 4);
            step_impl(&sp, p, s, a, &rng, nx);
            float r = reward_impl(&sp, p, s, a, nx);
            int term = terminated_impl(p, nx);
            if(j->out_rewards) j->out_rewards[(size_t)t * N + n] = r;
            if(j->out_term) j->out_term[(size_t)t * N + n] = (unsigned char)term;
            memcpy(s, nx, sizeof(float) * SD);
        }
        if(j->out_states) memcpy(j->out_states + ((size_t)j->T * N + n) * SD, s, sizeof(float) * SD);
        memcpy(j->states_io + (size_t)n * SD, s, sizeof(float) * SD);
        j->rng[n] = rng;
        if(pol->arch == ORACLE_POLICY_RAPTOR_GRU && j->hidden_io) memcpy(j->hidden_io + (size_t)n * HD, h, sizeof(float) * HD);
        if(j->gru_step_io) j->gru_step_io[n] = gstep;
    }
    return NULL;
}
void oracle_rollout(int spec, const oracle_policy_t* pol, int N, int T, int threads, const float* params, float* states_io, uint64_t* rng_states,
                    float* hidden_io, int* gru_step_io, int no_auto_reset,
                    float* out_states, float* out_observations, float* out_actions, float* out_rewards, unsigned char* out_terminated){
    if(threads < 1) threads = 1;
    if(threads > 256) threads = 256;
    rollout_job_t jobs[256]; pthread_t th[256];
    for(int t = 0; t < threads; t++){
        rollout_job_t j = {spec, pol, N, T, (int)((long long)N * t / threads), (int)((long long)N * (t + 1) / threads), params, states_io, rng_states, hidden_io, gru_step_io, no_auto_reset,
                           out_states, out_observations, out_actions, out_rewards, out_terminated};
        jobs[t] = j;
    }
    if(threads == 1){ rollout_range(&jobs[0]); return; }
    for(int t = 0; t < threads; t++) pthread_create(&th[t], NULL, rollout_range, &jobs[t]);
    for(int t = 0; t < threads; t++) pthread_join(th[t], NULL);
}

/* ---------------------------------------------------------------------------------------------
 * PPO collection, order of INC/rl/components/on_policy_runner/operations_generic.h:99-131 with
 * operations_generic_per_env.h:8-75.  Dataset rows (on_policy_runner.h:42-64, operations_generic.h:12-29):
 *   [(T+1)*N, OBS + 15]:  obs[OBS] | actions_mean[4] | actions[4] | log_prob | reward | terminated |
 *                         truncated | value | advantage | target_value     row = step*N + env
 * (value / advantage / target_value columns are left untouched: they belong to the learner); the
 * final observation of every environment goes to rows T*N + env.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int spec; const oracle_policy_t* pol; int N, T, n0, n1, step_limit; const float* env_params; float* params_io; float* states_io; uint64_t* rng;
    int* episode_step; float* episode_return; unsigned char* truncated; float* dataset; int data_dim;
} collect_job_t;
static void* collect_range(void* arg){
    collect_job_t* j = (collect_job_t*)arg;
    spec_t sp = get_spec(j->spec);
    const int SD = 44 + 4 * sp.H, OBS = sp.obs_dim, N = j->N, D = j->data_dim;
    for(int n = j->n0; n < j->n1; n++){
        float* p = j->params_io + (size_t)n * ORACLE_PARAMS_DIM;
        float* s = j->states_io + (size_t)n * SD;
        float nx[MAX_STATE_DIM];
        uint64_t rng = j->rng[n];
        for(int t = 0; t < j->T; t++){
            float* row = j->dataset + ((size_t)t * N + n) * D;
            if(j->truncated[n]){
                j->truncated[n] = 0; j->episode_step[n] = 0; j->episode_return[n] = 0;
                oracle_sample_initial_parameters(j->spec, j->env_params, &rng, p);
                /* the reference's reset leaves the previous episode's Langevin target in place when it draws the POSITION trajectory
                 * (30_sample_initial_state.h:217-219); oracle_sample_initial_state builds a fresh state, so carry the block over here */
                float dead_target[12];
                memcpy(dead_target, s + S_TRAJ_TYPE(sp.H) + 1, sizeof(dead_target));
                oracle_sample_initial_state(j->spec, p, &rng, s);
                if(sp.langevin && (int)s[S_TRAJ_TYPE(sp.H)] == 0) memcpy(s + S_TRAJ_TYPE(sp.H) + 1, dead_target, sizeof(dead_target));
            }
            observe_impl(&sp, p, s, &rng, row);
            float* mean = row + OBS; float* act = row + OBS + 4; float lp = 0;
            policy_forward(j->pol, row, NULL, NULL, 1, &rng, act, mean, &lp);
            row[OBS + 8] = lp;
            step_impl(&sp, p, s, act, &rng, nx);
            int term = terminated_impl(p, nx);
            row[OBS + 10] = (float)term;
            float r = reward_impl(&sp, p, s, act, nx);
            j->episode_return[n] += r;
            row[OBS + 9] = r;
            j->episode_step[n] += 1;
            int trunc = term || (j->step_limit > 0 && j->episode_step[n] >= j->step_limit);
            row[OBS + 11] = (float)trunc;
            j->truncated[n] = (unsigned char)trunc;
            memcpy(s, nx, sizeof(float) * SD);
        }
        observe_impl(&sp, p, s, &rng, j->dataset + ((size_t)j->T * N + n) * D);
        j->rng[n] = rng;
    }
    return NULL;
}
void oracle_collect(int spec, const oracle_policy_t* pol, int N, int T, int threads, int episode_step_limit, const float* env_params,
                    float* params_io, float* states_io, uint64_t* rng_states, int* episode_step_io, float* episode_return_io, unsigned char* truncated_io,
                    float* dataset, int data_dim){
    if(threads < 1) threads = 1;
    if(threads > 256) threads = 256;
    collect_job_t jobs[256]; pthread_t th[256];
    for(int t = 0; t < threads; t++){
        collect_job_t j = {spec, pol, N, T, (int)((long long)N * t / threads), (int)((long long)N * (t + 1) / threads), episode_step_limit, env_params, params_io, states_io, rng_states,
                           episode_step_io, episode_return_io, truncated_io, dataset, data_dim};
        jobs[t] = j;
    }
    if(threads == 1){ collect_range(&jobs[0]); return; }
    for(int t = 0; t < threads; t++) pthread_create(&th[t], NULL, collect_range, &jobs[t]);
    for(int t = 0; t < threads; t++) pthread_join(th[t], NULL);
}

/* ---------------------------------------------------------------------------------------------
 * Off-policy runner step (SAC teachers): rl::components::off_policy_runner `step` = prologue + interlude + epilogue,
 * INC/rl/components/off_policy_runner/operations_generic.h:215-238 with operations_generic_per_env.h:8-58 (prologue_per_env), :60-110
 * (epilogue_per_env) and the replay buffer `add` INC/rl/components/replay_buffer/operations_generic.h:54-79; one RNG stream per
 * environment as in the reference's CUDA kernels (operations_cuda.h:62-106).  Symmetric observations (the teacher env's
 * ObservationPrivileged is its Observation), so a replay row is  obs[OBS] | action[4] | reward | next_obs[OBS] | terminated | truncated
 * (replay_buffer.h:37-58, update_views operations_generic.h:12-22) and every environment owns a ring [capacity][2*OBS + 7]:
 *   replay [N][capacity][2*OBS+7], episode_start [N][capacity], position / full / current_episode_start [N].
 * states_out / next_states_out (ReplayBufferWithStates, operations_generic.h:81-85) [N][capacity][SD] may be NULL.
 * The actor is the SAC MLP with sample_and_squash in Mode<Rollout> (interlude, operations_generic.h:190-204): ORACLE_HEAD_SQUASH_SAMPLE.
 * ------------------------------------------------------------------------------------------- */
void oracle_off_policy_steps(int spec, const oracle_policy_t* pol, int N, int T, int episode_step_limit, int capacity, int sample_parameters, const float* env_params,
                             float* params_io, float* states_io, uint64_t* rng_states, int* episode_step_io, float* episode_return_io, unsigned char* truncated_io,
                             float* replay, int* episode_start, int* position_io, unsigned char* full_io, int* current_episode_start_io,
                             float* states_out, float* next_states_out){
    spec_t sp = get_spec(spec);
    const int SD = 44 + 4 * sp.H, OBS = sp.obs_dim, D = 2 * OBS + 7;
    for(int n = 0; n < N; n++){
        float* p = params_io + (size_t)n * ORACLE_PARAMS_DIM;
        float* s = states_io + (size_t)n * SD;
        float* rb = replay + (size_t)n * capacity * D;
        int* es = episode_start + (size_t)n * capacity;
        float nx[MAX_STATE_DIM], obs[MAX_OBS_DIM], nobs[MAX_OBS_DIM], act[8];
        uint64_t rng = rng_states[n];
        int position = position_io[n], full = full_io[n], current_start = current_episode_start_io[n];
        for(int t = 0; t < T; t++){
            /* ---- prologue_per_env */
            if(truncated_io[n]){
                if(sample_parameters) oracle_sample_initial_parameters(spec, env_params, &rng, p);
                float dead_target[12];   /* see collect_range: the reset keeps the previous Langevin block for POSITION trajectories */
                memcpy(dead_target, s + S_TRAJ_TYPE(sp.H) + 1, sizeof(dead_target));
                oracle_sample_initial_state(spec, p, &rng, s);
                if(sp.langevin && (int)s[S_TRAJ_TYPE(sp.H)] == 0) memcpy(s + S_TRAJ_TYPE(sp.H) + 1, dead_target, sizeof(dead_target));
                episode_step_io[n] = 0; episode_return_io[n] = 0;
                if(full || position > 0){
                    int previous = position == 0 ? capacity - 1 : position - 1;
                    rb[(size_t)previous * D + 2 * OBS + 6] = 1.0f;
                    current_start = position;
                }
            }
            observe_impl(&sp, p, s, &rng, obs);
            /* ---- interlude: evaluate_step in Mode<Rollout> */
            policy_forward(pol, obs, NULL, NULL, 1, &rng, act, NULL, NULL);
            /* ---- epilogue_per_env */
            step_impl(&sp, p, s, act, &rng, nx);
            float r = reward_impl(&sp, p, s, act, nx);
            observe_impl(&sp, p, nx, &rng, nobs);
            int term = terminated_impl(p, nx);
            episode_step_io[n] += 1;
            episode_return_io[n] += r;
            int trunc = term || episode_step_io[n] == episode_step_limit;
            truncated_io[n] = (unsigned char)trunc;
            float* row = rb + (size_t)position * D;
            if(states_out) memcpy(states_out + ((size_t)n * capacity + position) * SD, s, sizeof(float) * SD);
            if(next_states_out) memcpy(next_states_out + ((size_t)n * capacity + position) * SD, nx, sizeof(float) * SD);
            memcpy(row, obs, sizeof(float) * OBS);
            memcpy(row + OBS, act, sizeof(float) * 4);
            row[OBS + 4] = r;
            memcpy(row + OBS + 5, nobs, sizeof(float) * OBS);
            row[2 * OBS + 5] = (float)term;
            row[2 * OBS + 6] = (float)trunc;
            es[position] = current_start;
            position = (position + 1) % capacity;
            if(trunc) current_start = position;
            if(position == 0 && !full) full = 1;
            memcpy(s, nx, sizeof(float) * SD);
        }
        rng_states[n] = rng;
        position_io[n] = position; full_io[n] = (unsigned char)full; current_episode_start_io[n] = current_start;
    }
}

/* gather_batch for SEQUENCE_LENGTH = 1 (the MLP SAC configuration, pre_training/config.h:19,52-58: no random sequence length, not from the initial
 * state): INC/rl/components/off_policy_runner/operations_generic.h:240-420 (gather_batch_step) with the environment drawn per sample as in
 * :423-434 / the CUDA kernel operations_cuda.h:36-60 (one RNG stream per batch sample: env_i = uniform_int(0, N - 1), then the sample offset).
 * uniform_int = next(state) % range (INC/random/operations_generic.h:43-50).  Outputs in the SequentialBatch layout (off_policy_runner.h:96-141):
 *   observations_actions [2][B][OBS + 4]: step 0 = obs | action, step 1 = next_obs | 0;  rewards / terminated [B];
 *   reset [B] = 1, next_reset [2][B] = {1, 1} (base row 0 by :287,303; row 1 through the `next_reset` view, whose offset is 1 when the first step is
 *   not a target, :288 with :86-92), final_step_mask [B] = 1, next_final_step_mask [2][B] = {0, 1}  (any of the masks may be NULL).
 * env_begin / env_count select the runner's environments (one teacher's group); env_index / sample_index [B] (optional) report the draw. */
static uint64_t rng_next_state(uint64_t* s){ rng_next(s); return *s; }
void oracle_gather_batch(int obs_dim, int capacity, int max_episode_length, int env_begin, int env_count, const float* replay, const int* position, const unsigned char* full,
                         int B, uint64_t* rng_states, float* observations_actions, float* rewards, unsigned char* terminated,
                         unsigned char* reset, unsigned char* next_reset, unsigned char* final_step_mask, unsigned char* next_final_step_mask,
                         int* env_index, int* sample_index_out){
    const int OBS = obs_dim, D = 2 * OBS + 7, W = OBS + 4;
    for(int b = 0; b < B; b++){
        uint64_t* rng = rng_states + b;
        const int env = env_begin + (int)(rng_next_state(rng) % (uint64_t)env_count);
        const float* rb = replay + (size_t)env * capacity * D;
        const int is_full = full[env], pos = position[env];
        const uint64_t eligible = is_full ? (uint64_t)capacity : (uint64_t)pos;
        const uint64_t offset = rng_next_state(rng) % eligible;                 /* uniform_int(0, eligible - 1) */
        const int sample = is_full ? (int)(((uint64_t)pos + (uint64_t)max_episode_length + offset) % (uint64_t)capacity) : (int)offset;
        const float* row = rb + (size_t)sample * D;
        float* o0 = observations_actions + (size_t)b * W;
        float* o1 = observations_actions + ((size_t)B + b) * W;
        memcpy(o0, row, sizeof(float) * W);                                      /* obs | action */
        memcpy(o1, row + OBS + 5, sizeof(float) * OBS);                          /* next_obs */
        for(int i = 0; i < 4; i++) o1[OBS + i] = 0;                              /* next action = 0 (:408-410) */
        rewards[b] = row[OBS + 4];
        terminated[b] = row[2 * OBS + 5] != 0;
        if(reset) reset[b] = 1;
        if(next_reset){ next_reset[b] = 1; next_reset[B + b] = 1; }
        if(final_step_mask) final_step_mask[b] = 1;
        if(next_final_step_mask){ next_final_step_mask[b] = 0; next_final_step_mask[B + b] = 1; }
        if(env_index) env_index[b] = env;
        if(sample_index_out) sample_index_out[b] = sample;
    }
}

/* gather_batch for any SEQUENCE_LENGTH (recurrent SAC): INC/rl/components/off_policy_runner/operations_generic.h:240-420 (gather_batch_step) restated
 * line by line, with the batch parameters of off_policy_runner.h:78-85 as run-time values and one RNG stream per batch sample, the environment drawn
 * from it first (:423-434, operations_cuda.h:36-60).  L = SEQUENCE_LENGTH, P = L + 1 = PADDED_SEQUENCE_LENGTH.  Outputs (off_policy_runner.h:96-141):
 *   observations_actions [P][B][OBS + 4]; rewards / terminated / reset / final_step_mask [L][B]; next_reset / next_final_step_mask [P][B] (the
 *   *_base tensors; the reference's `next_*` views start at row include_first_step_in_targets ? 0 : 1, operations_generic.h:87-92).
 * uniform_real(0.0, 1.0) is the double overload: state / (double)MAX_INDEX (random/operations_generic.h:52-58), compared with the float parameter.
 * rewards / terminated of padding steps are not written by the reference (stale memory): this restatement, like the CUDA path, writes zeros there. */
void oracle_gather_batch_sequential(int obs_dim, int capacity, int max_episode_length, int env_begin, int env_count, const float* replay, const int* episode_start,
                                    const int* position, const unsigned char* full, int L, int include_first_step_in_targets, int always_sample_from_initial_state,
                                    int random_seq_length, int enable_nominal, float nominal_probability,
                                    int B, uint64_t* rng_states, float* observations_actions, float* rewards, unsigned char* terminated,
                                    unsigned char* reset, unsigned char* next_reset_base, unsigned char* final_step_mask, unsigned char* next_final_step_mask_base,
                                    int* env_index, int* sample_index_out){
    const int OBS = obs_dim, D = 2 * OBS + 7, W = OBS + 4, P = L + 1;
    const int next_offset = include_first_step_in_targets ? 0 : 1;
    for(int b = 0; b < B; b++){
        uint64_t* rng = rng_states + b;
        const int env = env_begin + (int)(rng_next_state(rng) % (uint64_t)env_count);
        const float* rb = replay + (size_t)env * capacity * D;
        const int* ep_start = episode_start + (size_t)env * capacity;
        const int is_full = full[env], pos = position[env];
        const size_t eligible = is_full ? (always_sample_from_initial_state ? (size_t)(capacity - max_episode_length) : (size_t)capacity) : (size_t)pos;
        const size_t sample_index_max = eligible - 1;
        size_t sample_index = 0;
        size_t current_seq_length = (size_t)L;
        if(random_seq_length){
            /* the reference evaluates `ENABLE && uniform_real(...) < p` first: the draw happens only when ENABLE is set (:266) */
            if(enable_nominal){
                const double u = (double)rng_next_state(rng) / (double)UINT64_MAX;
                if(u < (double)nominal_probability) current_seq_length = (size_t)L;
                else if(L > 1) current_seq_length = 1 + (size_t)(rng_next_state(rng) % (uint64_t)(L - 1));
            }
            else if(L > 1) current_seq_length = 1 + (size_t)(rng_next_state(rng) % (uint64_t)(L - 1));
        }
        size_t current_seq_step = 0;
        int previous_step_truncated = 0, previous_padding_step = 1;
        for(int s = 0; s < L; s++){ final_step_mask[(size_t)s * B + b] = 0; reset[(size_t)s * B + b] = 0; rewards[(size_t)s * B + b] = 0; terminated[(size_t)s * B + b] = 0; }
        for(int s = 0; s < P; s++){ next_final_step_mask_base[(size_t)s * B + b] = 0; next_reset_base[(size_t)s * B + b] = 0; }
        reset[b] = 1;                                              /* :286-288 */
        next_reset_base[b] = 1;
        next_reset_base[(size_t)next_offset * B + b] = 1;
        int first_sample = -1;
        for(int seq = 0; seq < P; seq++){
            if(previous_step_truncated){                           /* :290-299 */
                previous_step_truncated = 0;
                previous_padding_step = 1;
                next_final_step_mask_base[(size_t)seq * B + b] = 1;
                if(seq < P - 1) reset[(size_t)seq * B + b] = 1;
                continue;
            }
            if(previous_padding_step){                             /* :300-321 */
                if(seq < P - 1) reset[(size_t)seq * B + b] = 1;
                next_reset_base[(size_t)seq * B + b] = 1;
                const size_t sample_offset = (size_t)(rng_next_state(rng) % (uint64_t)(sample_index_max + 1));
                if(is_full) sample_index = ((size_t)pos + (size_t)max_episode_length + sample_offset) % (size_t)capacity;
                else sample_index = sample_offset;
                if(always_sample_from_initial_state) sample_index = (size_t)ep_start[sample_index];
                if(first_sample < 0) first_sample = (int)sample_index;
            }
            const float* row = rb + sample_index * D;
            float* oa = observations_actions + ((size_t)seq * B + b) * W;
            memcpy(oa, row, sizeof(float) * W);                    /* obs | action (:324-345) */
            if(seq < P - 1){
                rewards[(size_t)seq * B + b] = row[OBS + 4];
                terminated[(size_t)seq * B + b] = row[2 * OBS + 5] != 0;
            }
            int truncated = row[2 * OBS + 6] != 0;                 /* :352 */
            size_t next_sample_index = sample_index + 1;
            if(is_full) next_sample_index = next_sample_index % (size_t)capacity;
            if(next_sample_index == (size_t)pos) truncated = 1;
            if(seq == P - 2) truncated = 1;
            if(random_seq_length){                                 /* :363-388 */
                if(current_seq_step == current_seq_length - 1){
                    truncated = 1;
                    if(L > 1){
                        if(enable_nominal){
                            const double u = (double)rng_next_state(rng) / (double)UINT64_MAX;
                            if(u < (double)nominal_probability) current_seq_length = (size_t)L;
                            else current_seq_length = 1 + (size_t)(rng_next_state(rng) % (uint64_t)(L - 1));
                        }
                        else current_seq_length = 1 + (size_t)(rng_next_state(rng) % (uint64_t)(L - 1));
                    }
                    else current_seq_length = (size_t)L;
                }
                if(truncated) current_seq_step = 0;
                else current_seq_step += 1;
            }
            if(truncated){                                         /* :393-417 */
                if(seq < P - 1){
                    final_step_mask[(size_t)seq * B + b] = 1;
                    float* nx = observations_actions + ((size_t)(seq + 1) * B + b) * W;
                    memcpy(nx, row + OBS + 5, sizeof(float) * OBS);
                    for(int i = 0; i < 4; i++) nx[OBS + i] = 0;
                }
            }
            sample_index = next_sample_index;
            previous_padding_step = 0;
            previous_step_truncated = truncated;
        }
        if(env_index) env_index[b] = env;
        if(sample_index_out) sample_index_out[b] = first_sample;
    }
}

/* ---------------------------------------------------------------------------------------------
 * learner feed of the PPO loop step (INC/rl/algorithms/ppo/loop/core/operations_generic.h:104-117): critic values over
 * all_observations -> all_values column, generalized advantage estimation, running observation normalizer.
 * dataset rows: [(T+1)*N, D], D = OBS + 15, columns as documented above oracle_collect.
 * ------------------------------------------------------------------------------------------- */
/* evaluate(device, critic, all_observations_privileged, all_values): operations_generic.h:112-116 (critic = standardize -> MLP, output 1) */
void oracle_evaluate_values(const oracle_policy_t* critic, int N, int T, float* dataset, int data_dim){
    const int OBS = data_dim - 15;
    for(size_t r = 0; r < (size_t)(T + 1) * N; r++){
        float* row = dataset + r * data_dim;
        float v = 0;
        mlp_forward(critic, row, NULL, &v, NULL, NULL);
        row[OBS + 12] = v;
    }
}
/* estimate_generalized_advantages: INC/rl/algorithms/ppo/operations_generic.h:54-89 */
void oracle_estimate_generalized_advantages(int N, int T, float* dataset, int data_dim, float gamma, float lambda, int ignore_termination){
    const int OBS = data_dim - 15, D = data_dim;
    for(int env = 0; env < N; env++){
        float previous_value = dataset[((size_t)T * N + env) * D + OBS + 12];
        float previous_advantage = 0;
        for(int fwd = 0; fwd < T; fwd++){
            const int step = T - 1 - fwd;
            float* row = dataset + ((size_t)step * N + env) * D;
            const int terminated = row[OBS + 10] != 0, truncated = row[OBS + 11] != 0;
            const float current_step_value = row[OBS + 12];
            const int terminated_actual = terminated && !ignore_termination;
            const float next_step_value = terminated_actual ? 0 : previous_value;
            float td_error = row[OBS + 9] + gamma * next_step_value - current_step_value;
            if(truncated){
                if(!terminated) td_error = 0;
                previous_advantage = 0;
            }
            const float advantage = lambda * gamma * previous_advantage + td_error;
            row[OBS + 13] = advantage;
            row[OBS + 14] = advantage + current_step_value;
            previous_advantage = advantage;
            previous_value = current_step_value;
        }
    }
}
/* rl::components::running_normalizer update: INC/rl/components/running_normalizer/operations_generic.h:27-49 with the column mean / std of
 * INC/containers/matrix/operations_generic.h:641-691 (sequential fp32 sums, sample std, acc < 1e-6 -> 0), over the T*N observation rows */
void oracle_normalizer_update(int N, int T, const float* dataset, int data_dim, float* mean_io, float* std_io, int* age_io){
    const int OBS = data_dim - 15;
    const size_t rows = (size_t)T * N;
    *age_io += 1;
    for(int c = 0; c < OBS; c++){
        float acc = 0;
        for(size_t r = 0; r < rows; r++) acc += dataset[r * data_dim + c];
        const float data_mean = acc / rows;
        float acc2 = 0;
        for(size_t r = 0; r < rows; r++){ const float diff = dataset[r * data_dim + c] - data_mean; acc2 += diff * diff; }
        const float data_std = acc2 < 1e-6 ? 0 : sqrtf(acc2 / (rows - 1));
        mean_io[c] = mean_io[c] + (data_mean - mean_io[c]) / (*age_io);
        std_io[c] = std_io[c] + (data_std - std_io[c]) / (*age_io);
    }
}

/* ---------------------------------------------------------------------------------------------
 * Foundation-policy DAgger data path: add_to_dataset, src/foundation_policy/post_training/helper.h:43-110, for all teachers at once.
 * The student rollout was recorded step-major (states [T][n][SD], terminated [T][n]); environment e belongs to teacher e / episodes_per_teacher
 * and episodes are appended in environment order (= the reference's teacher loop, post_training/main.cpp:304-309, then its episode loop).
 * Per episode and step, up to and including the first terminated step: teacher observation (26, pre-training layout) and student observation
 * (22, post-training layout, position minus the teacher's steady-state offset) of the recorded state, with the environment's parameters and
 * its RNG stream (draws only when the observation noise is on: teacher observation first, helper.h:64-65); truncated = terminated or last step
 * (:70); reset: the reference's flag is initialised true and never cleared (:52,72-75), so every row carries true; episode_start[episode] = its
 * first row (:55); then the teacher (MLP 26-64-64-8 + sample_and_squash in Evaluation mode: tanh(mean)) labels every
 * row (:92-104).  Returns the number of rows.
 * ------------------------------------------------------------------------------------------- */
long long oracle_dagger_add_to_dataset(int n, int T, int episodes_per_teacher, const float* params, const float* states, const unsigned char* terminated, uint64_t* rng,
                                       const float* teacher_blobs, const float* offsets, int* episode_start, float* input_student, float* output_target,
                                       unsigned char* truncated_out, unsigned char* reset_out){
    spec_t sp_student = get_spec(ORACLE_SPEC_RAPTOR), sp_teacher = get_spec(ORACLE_SPEC_TEACHER);
    const int SD = 44 + 4 * sp_student.H;
    const int BLOB = 64 * 26 + 64 + 64 * 64 + 64 + 8 * 64 + 8;
    long long index = 0;
    for(int e = 0; e < n; e++){
        const int teacher = e / episodes_per_teacher;
        const float* p = params + (size_t)e * ORACLE_PARAMS_DIM;
        oracle_policy_t pol = {ORACLE_POLICY_MLP, 26, 64, 8, 0, ORACLE_HEAD_SQUASH_EVAL, teacher_blobs + (size_t)teacher * BLOB};
        episode_start[e] = (int)index;
        int step;
        for(step = 0; step < T; step++){
            const float* s = states + ((size_t)step * n + e) * SD;
            float obs_teacher[26], obs_student[22];
            observe_impl(&sp_teacher, p, s, &rng[e], obs_teacher);
            observe_impl(&sp_student, p, s, &rng[e], obs_student);
            for(int i = 0; i < 3; i++) obs_student[i] = obs_student[i] - offsets[teacher * 3 + i];
            memcpy(input_student + (size_t)(index + step) * 22, obs_student, sizeof(obs_student));
            const int term = terminated[(size_t)step * n + e] != 0;
            truncated_out[index + step] = (unsigned char)(term || step == T - 1);
            reset_out[index + step] = 1;
            mlp_forward(&pol, obs_teacher, NULL, output_target + (size_t)(index + step) * 4, NULL, NULL);
            if(term){ step++; break; }
        }
        index += step;
    }
    return index;
}
int oracle_hardware_threads(void){ long n = sysconf(_SC_NPROCESSORS_ONLN); return n > 0 ? (int)n : 1; }
