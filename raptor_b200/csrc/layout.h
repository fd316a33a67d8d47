// raptor_b200/csrc/layout.h -- flat parameter / state layouts of the engine (documented in include/b200_l2f.h).
// Reference structs: rl/environments/l2f/multirotor.h:23-140 (parameters), :574-725 (state).
#pragma once

namespace b200l2f {

enum ParamOffset : int {
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
    P_TRAJ_MIX0 = 139, P_TRAJ_MIX1 = 140, P_LANGEVIN_GAMMA = 141, P_LANGEVIN_OMEGA = 142, P_LANGEVIN_SIGMA = 143, P_LANGEVIN_ALPHA = 144,
    PARAMS_DIM = 145,
    P_DYN_DIM = 86  // [0, 86): everything one dynamics evaluation reads (dynamics struct + dt); staged in shared memory by the fused kernels
};

enum StateOffset : int {
    S_POS = 0, S_ORI = 3, S_LINVEL = 7, S_ANGVEL = 10, S_LAST_ACTION = 13, S_ANGVEL_HIST = 17, S_FORCE = 20, S_TORQUE = 23, S_RPM = 26,
    S_CURRENT_STEP = 30, S_HIST = 31
};
__host__ __device__ constexpr int s_traj_type(int H){ return 31 + 4 * H; }
__host__ __device__ constexpr int s_langevin(int H){ return 32 + 4 * H; }  // position[3] velocity[3] position_raw[3] velocity_raw[3]
__host__ __device__ constexpr int state_dim(int H){ return 44 + 4 * H; }

enum ObsLayout : int { OBS_DEFAULT = 0, OBS_RAPTOR = 1, OBS_TEACHER = 2 };

// compile-time description of an environment specification (the reference does the same with template parameter packs)
// T_RNG_OOL: the Box-Muller draw is an out-of-line call (kernels with ~30 draw sites, see rng.cuh); false: inlined (rollout kernels)
template <int T_H, bool T_LANGEVIN, int T_OBS_LAYOUT, bool T_RNG_OOL = false>
struct EnvSpec {
    static constexpr bool RNG_OOL = T_RNG_OOL;
    static constexpr int H = T_H;
    static constexpr bool LANGEVIN = T_LANGEVIN;
    static constexpr int OBS_LAYOUT = T_OBS_LAYOUT;
    static constexpr int STATE_DIM = state_dim(T_H);
    static constexpr int OBS_DIM = 18 + 4 * T_H + (T_OBS_LAYOUT == OBS_TEACHER ? 4 : 0);
};
using SpecDefault = EnvSpec<16, false, OBS_DEFAULT>;  // rl/environments/l2f/parameters/default.h:159-171
using SpecRaptor  = EnvSpec<1, true, OBS_RAPTOR>;     // src/foundation_policy/post_training/environment.h:23-33
using SpecTeacher = EnvSpec<1, true, OBS_TEACHER>;    // src/foundation_policy/pre_training/environment.h:64-75
template <class Spec> using SpecCompactCode = EnvSpec<Spec::H, Spec::LANGEVIN, Spec::OBS_LAYOUT, true>;

}  // namespace b200l2f
