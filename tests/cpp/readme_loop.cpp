// tests/cpp/readme_loop.cpp -- the README loop (R/README.md:40-105) written against the header-only C++ shim (include/b200_l2f.hpp)
// in the rl-tools idiom: device first, caller-owned objects, explicit malloc/free.  Writes actions [T][8][4] and the final states
// [8][108] to argv[2]; tests/test_cpp_shim.py compares them with the golden trajectory generated from the reference.
#include <cstdio>
#include <vector>
#include <b200_l2f.hpp>

namespace bl = b200;
using SPEC = bl::l2f::DefaultSpecification;
constexpr size_t N = 8;
constexpr int T = 100;

int main(int argc, char** argv){
    if(argc < 3){ std::fprintf(stderr, "usage: readme_loop <raptor_policy_2084.f32> <out.bin>\n"); return 2; }
    std::vector<float> blob(2084);
    { FILE* f = std::fopen(argv[1], "rb"); if(!f || std::fread(blob.data(), sizeof(float), blob.size(), f) != blob.size()){ std::fprintf(stderr, "cannot read the policy blob\n"); return 2; } std::fclose(f); }
    bl::devices::B200 device;
    bl::l2f::vector::Environment<SPEC, N> env;
    bl::l2f::vector::Parameters<N> params;
    bl::l2f::vector::Rng<N> rng;
    bl::l2f::vector::State<N> state, next_state;
    bl::l2f::vector::Matrix<N, SPEC::OBSERVATION_DIM> observation;
    bl::l2f::vector::Matrix<N, 4> action;
    bl::policy::Raptor policy; policy.blob = blob.data();

    bl::malloc(device, env);
    bl::malloc(device, env, state);
    bl::malloc(device, env, next_state);
    bl::malloc(device, env, policy);
    bl::init(device, rng, 0);
    bl::init(device, env);
    bl::sample_initial_parameters(device, env, params, rng);
    bl::sample_initial_state(device, env, params, state, rng);
    bl::reset(device, env, policy);
    std::vector<float> actions;
    for(int t = 0; t < T; t++){
        bl::observe(device, env, params, state, observation, rng);
        bl::evaluate_step(device, env, policy, observation, action);
        float dt = bl::step(device, env, params, state, action, next_state, rng);
        bl::copy(device, env, next_state, state);
        actions.insert(actions.end(), action.data.begin(), action.data.end());
        if(dt < 0.0099f || dt > 0.0101f){ std::fprintf(stderr, "unexpected dt %f\n", dt); return 1; }
    }
    bl::l2f::vector::Matrix<N, SPEC::STATE_DIM> rows;
    bl::get(device, env, state, rows);
    // the fused replacement of rl_tools::evaluate on a fresh copy of the same initial conditions
    bl::EvaluationResult<N> result;
    bl::init(device, rng, 0);
    bl::sample_initial_parameters(device, env, params, rng);
    bl::sample_initial_state(device, env, params, state, rng);
    bl::evaluate(device, env, policy, result, T);
    FILE* f = std::fopen(argv[2], "wb");
    std::fwrite(actions.data(), sizeof(float), actions.size(), f);
    std::fwrite(rows.data.data(), sizeof(float), rows.data.size(), f);
    std::fwrite(result.returns.data(), sizeof(float), N, f);
    std::fclose(f);
    bl::free(device, env);
    std::printf("ok\n");
    return 0;
}
