// tests/cpp/rl_tools_binding.cpp -- the README / rl_tools::evaluate loop written with the REFERENCE's own headers and types, on the B200 device tag of
// include/rl_tools_b200.h: rlt::malloc / init / sample_initial_parameters / sample_initial_state / observe / evaluate_step / step / evaluate resolve, by
// overload on rl_tools::devices::B200, to the engine's C ABI.  The environment type is the reference's default l2f::Specification<float, size_t> (the
// README's vector8 case), the actor is the reference's checkpoint model object (rl_tools::checkpoint::actor::module), containers are rl_tools::Matrix.
// Built in the container (needs /root/reference) by tests/test_rl_tools_binding.py; the binary travels to the GPU box, where the -m gpu test runs it
// and compares actions / final states / returns with the golden trajectory generated from the reference (tests/golden/default_8x500.npz).
#include <cstdio>
#include <vector>
#include <rl_tools/operations/cpu.h>
#include <rl_tools/nn/layers/dense/operations_generic.h>
#include <rl_tools/nn/layers/gru/operations_generic.h>
#include <rl_tools/nn_models/sequential/operations_generic.h>
#include <rl_tools/rl/environments/l2f/operations_generic.h>
#include <rl_tools/rl/utils/evaluation/evaluation.h>
#include "checkpoint.h"
#include <rl_tools_b200.h>

namespace rlt = rl_tools;
namespace l2f = rlt::rl::environments::l2f;
namespace vec = l2f::b200::vector;
using T = float;
using TI = size_t;
using DEVICE_CPU = rlt::devices::DefaultCPU;
using ENVIRONMENT = rlt::rl::environments::Multirotor<l2f::Specification<T, TI>>;   // H = 16, OBS 82: the README's environment
constexpr TI N = 8;
constexpr int STEPS = 100;

int main(int argc, char** argv){
    if(argc < 2){ std::fprintf(stderr, "usage: rl_tools_binding <out.bin>\n"); return 2; }
    rlt::devices::B200 device;
    DEVICE_CPU device_cpu;
    vec::Environment<ENVIRONMENT, N> env;
    vec::Parameters<N> parameters;
    vec::State<N> state, next_state; next_state.slot = 1;
    vec::Rng rng;
    vec::Policy<N> policy;
    rlt::Matrix<rlt::matrix::Specification<T, TI, N, ENVIRONMENT::Observation::DIM, false>> observation;
    rlt::Matrix<rlt::matrix::Specification<T, TI, N, ENVIRONMENT::ACTION_DIM, false>> action;

    rlt::malloc(device, env);
    rlt::init(device, env);
    rlt::init(device, env, rng, 0);
    rlt::copy(device_cpu, device, rlt::checkpoint::actor::module, env, policy);
    rlt::sample_initial_parameters(device, env, parameters, rng);
    rlt::sample_initial_state(device, env, parameters, state, rng);
    rlt::reset(device, policy);
    std::vector<float> actions;
    for(int t = 0; t < STEPS; t++){
        rlt::observe(device, env, parameters, state, observation, rng);
        auto policy_input = rlt::view(device_cpu, observation, rlt::matrix::ViewSpec<N, 22>{}, 0, 0);     // observation[:, :22] (R/README.md:96)
        rlt::evaluate_step(device, policy, policy_input, action);
        const T dt = rlt::step(device, env, parameters, state, action, next_state, rng);
        rlt::copy(device, env, next_state, state);
        for(TI i = 0; i < N; i++) for(TI j = 0; j < 4; j++) actions.push_back(rlt::get(action, i, j));
        if(dt < 0.0099f || dt > 0.0101f){ std::fprintf(stderr, "unexpected dt %f\n", dt); return 1; }
    }
    std::vector<float> rows(N * 108);
    if(b200l2f_get_state(env.handle, 0, rows.data(), B200L2F_HOST) != B200L2F_OK) return 1;
    // the fused replacement of rl_tools::evaluate on a fresh copy of the same initial conditions, into the reference's own Result type
    using EVAL_SPEC = rlt::rl::utils::evaluation::Specification<T, TI, ENVIRONMENT, N, STEPS>;
    rlt::rl::utils::evaluation::Result<EVAL_SPEC> result;
    rlt::init(device, env, rng, 0);
    rlt::sample_initial_parameters(device, env, parameters, rng);
    rlt::sample_initial_state(device, env, parameters, state, rng);
    rlt::evaluate(device, env, policy, result, STEPS);
    FILE* f = std::fopen(argv[1], "wb");
    std::fwrite(actions.data(), sizeof(float), actions.size(), f);
    std::fwrite(rows.data(), sizeof(float), rows.size(), f);
    std::fwrite(result.returns, sizeof(float), N, f);
    const float stats[6] = {result.returns_mean, result.returns_std, result.episode_length_mean, result.episode_length_std, (float)result.num_terminated, result.share_terminated};
    std::fwrite(stats, sizeof(float), 6, f);
    std::fclose(f);
    rlt::free(device, env);
    std::printf("ok\n");
    return 0;
}
