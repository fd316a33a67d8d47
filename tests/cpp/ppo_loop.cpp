// tests/cpp/ppo_loop.cpp -- the PPO loop step's data path (collect -> critic values -> GAE -> running normalizer) written against the header-only shim
// (include/b200_l2f.hpp) the way a caller of rl_tools::collect / estimate_generalized_advantages / update would write it, plus the checkpoint and JSON
// helpers.   usage: ppo_loop raptor|default blobs.f32 out.f32  |  ppo_loop checkpoint FILE  |  ppo_loop json row.f32 [out.json [row_again.f32]]   (the last two: no GPU)
// blobs.f32 = actor (standardize + MLP + log_std) followed by the critic (standardize + MLP, 1 output); out.f32 = dataset | normalizer mean | std.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "b200_l2f.hpp"

namespace rlt = b200;

template <typename SPEC>
int run(const char* blobs_path, const char* out_path){
    constexpr size_t N = 200, STEPS = 24;
    constexpr int OBS = SPEC::OBSERVATION_DIM;
    rlt::devices::B200 device;
    rlt::l2f::vector::Environment<SPEC, N> env;
    rlt::l2f::vector::Rng<N> rng;
    rlt::malloc(device, env);
    rlt::init(device, env);
    rlt::init(device, rng, 77);
    float row[B200L2F_PARAMS_DIM];
    rlt::get_environment_parameters(device, env, row);
    const float dr[15] = {1.5f, 5.0f, 40.f, 1200.f, 0.02f, 5.0f, 0.1f, 0.03f, 0.10f, 0.03f, 0.30f, 0.005f, 0.05f, 0.0f, 0.3f};   // sample_dynamics_parameters.cpp:48-64
    std::memcpy(row + 124, dr, sizeof(dr));
    rlt::set_environment_parameters(device, env, row);
    // what every parameter file goes through (6 decimals, std::to_string: the text does not round-trip bit-exactly, so the row is not replaced here)
    float again[B200L2F_PARAMS_DIM];
    std::memcpy(again, row, sizeof(row));
    rlt::from_json(device, rlt::json(device, row), again);
    rlt::utils::assert_exit(device, again[124] == row[124] && again[125] == row[125], "parameter JSON: domain-randomisation range lost");

    const size_t actor_floats = 2 * OBS + 64 * OBS + 64 + 64 * 64 + 64 + 4 * 64 + 4 + 4, critic_floats = 2 * OBS + 64 * OBS + 64 + 64 * 64 + 64 + 64 + 1;
    std::vector<float> blobs(actor_floats + critic_floats);
    std::ifstream f(blobs_path, std::ios::binary);
    f.read((char*)blobs.data(), sizeof(float) * blobs.size());
    rlt::utils::assert_exit(device, (size_t)f.gcount() == sizeof(float) * blobs.size(), "blobs file too short");
    rlt::policy::MLP actor, critic;
    actor.blob.assign(blobs.begin(), blobs.begin() + actor_floats); actor.input_dim = OBS; actor.output_dim = 4; actor.standardize = true; actor.head = B200L2F_HEAD_PPO_GAUSSIAN;
    critic.blob.assign(blobs.begin() + actor_floats, blobs.end()); critic.input_dim = OBS; critic.output_dim = 1; critic.standardize = true;
    rlt::malloc(device, env, actor);
    rlt::malloc_critic(device, env, critic);

    rlt::on_policy_runner::Runner<SPEC, N> runner;
    runner.step_limit = 9;
    using DATASET = rlt::on_policy_runner::Dataset<SPEC, N, STEPS>;
    DATASET dataset;
    rlt::RunningNormalizer<OBS> normalizer;
    rlt::init(device, runner, env, rng);
    for(int iteration = 0; iteration < 2; iteration++){          // the second pass continues from the runner's carried-over state
        rlt::collect(device, dataset, runner);
        rlt::evaluate_values(device, runner, dataset);
        rlt::estimate_generalized_advantages(device, runner, dataset, rlt::ppo::Parameters{});
        rlt::update(device, runner, normalizer, dataset);
    }
    double truncated = 0;
    for(size_t r = 0; r < DATASET::STEPS_TOTAL; r++) truncated += dataset(r, DATASET::TRUNCATED);
    std::printf("steps %zu truncated %.0f normalizer age %d mean[0] %.6f\n", runner.step, truncated, normalizer.age, normalizer.mean[0]);
    std::ofstream o(out_path, std::ios::binary);
    o.write((const char*)dataset.data.data(), sizeof(float) * dataset.data.size());
    o.write((const char*)normalizer.mean.data(), sizeof(float) * OBS);
    o.write((const char*)normalizer.std.data(), sizeof(float) * OBS);
    rlt::free(device, env);
    return 0;
}

// SAC-teacher data path (off-policy runner steps into the replay rings, then the learner-side batches): `ppo_loop sac actor.f32 out.f32`
// out.f32 = rings | batch (SEQUENCE_LENGTH 1) | batch (SEQUENCE_LENGTH 6, random lengths, from any row), every tensor as float
struct SacSequenceParameters {   // rings of 64 rows < EPISODE_STEP_LIMIT: sampling from the episode's first row is not available (operations_generic.h:258)
    static constexpr bool INCLUDE_FIRST_STEP_IN_TARGETS = true, ALWAYS_SAMPLE_FROM_INITIAL_STATE = false, RANDOM_SEQ_LENGTH = true, ENABLE_NOMINAL_SEQUENCE_LENGTH_PROBABILITY = true;
    static constexpr float NOMINAL_SEQUENCE_LENGTH_PROBABILITY = 0.25f;
};
int run_sac(const char* actor_path, const char* out_path){
    using SPEC = rlt::l2f::TeacherDRSpecification;
    constexpr size_t N = 64, CAPACITY = 64, BATCH = 32;
    constexpr int OBS = SPEC::OBSERVATION_DIM;
    rlt::devices::B200 device;
    rlt::l2f::vector::Environment<SPEC, N> env;
    rlt::l2f::vector::Rng<N> rng;
    rlt::malloc(device, env);
    rlt::init(device, env);
    rlt::init(device, rng, 91);
    float row[B200L2F_PARAMS_DIM];
    rlt::get_environment_parameters(device, env, row);
    const float dr[15] = {1.5f, 5.0f, 40.f, 1200.f, 0.02f, 5.0f, 0.1f, 0.03f, 0.10f, 0.03f, 0.30f, 0.005f, 0.05f, 0.0f, 0.3f};
    std::memcpy(row + 124, dr, sizeof(dr));
    rlt::set_environment_parameters(device, env, row);
    rlt::policy::MLP actor;
    actor.blob.resize(64 * OBS + 64 + 64 * 64 + 64 + 8 * 64 + 8); actor.input_dim = OBS; actor.output_dim = 8; actor.head = B200L2F_HEAD_SQUASH_EVAL;
    std::ifstream f(actor_path, std::ios::binary);
    f.read((char*)actor.blob.data(), sizeof(float) * actor.blob.size());
    rlt::utils::assert_exit(device, (size_t)f.gcount() == sizeof(float) * actor.blob.size(), "actor file too short");
    rlt::malloc(device, env, actor);
    rlt::off_policy_runner::Runner<SPEC, N, CAPACITY> runner;
    runner.step_limit = 30;
    rlt::init(device, runner, env, rng);
    rlt::step(device, runner, 100);                               // the rings wrap
    rlt::off_policy_runner::SequentialBatch<SPEC, BATCH> batch1;
    rlt::off_policy_runner::SequentialBatch<SPEC, BATCH, 6, SacSequenceParameters> batch6;
    for(size_t b = 0; b < BATCH; b++){ batch1.rng[b] = 0xAAAAAAAAull + 5000 + b; batch6.rng[b] = 0xAAAAAAAAull + 6000 + b; }
    rlt::gather_batch(device, runner, batch1);
    rlt::gather_batch(device, runner, batch6, 16, 24);            // one group of environments (a teacher's)
    std::ofstream o(out_path, std::ios::binary);
    auto put = [&](const auto& v){ for(auto x : v){ const float y = (float)x; o.write((const char*)&y, sizeof(float)); } };
    put(runner.data); put(runner.position); put(runner.full);
    put(batch1.observations_actions); put(batch1.rewards); put(batch1.terminated);
    put(batch6.observations_actions); put(batch6.rewards); put(batch6.terminated); put(batch6.reset); put(batch6.next_reset); put(batch6.final_step_mask); put(batch6.next_final_step_mask);
    std::printf("rings full %d, batch6 final steps %d\n", (int)runner.full[0], (int)batch6.final_step_mask[0]);
    rlt::free(device, env);
    return 0;
}

int main(int argc, char** argv){
    if(argc >= 4 && std::string(argv[1]) == "sac") return run_sac(argv[2], argv[3]);
    if(argc >= 3 && std::string(argv[1]) == "checkpoint"){        // host only: no GPU needed
        rlt::devices::B200 device;
        rlt::policy::Checkpoint c;
        rlt::load(device, argv[2], c);
        double sum = 0; for(float v : c.blob) sum += v;
        std::printf("arch %d in %d hidden %d out %d floats %zu sum %.9g name %s\n", c.desc.arch, c.desc.input_dim, c.desc.hidden_dim, c.desc.output_dim, c.blob.size(), sum, c.name.c_str());
        return 0;
    }
    if(argc >= 3 && std::string(argv[1]) == "json"){              // host only: parameter row -> the reference's JSON text -> row
        rlt::devices::B200 device;
        float row[B200L2F_PARAMS_DIM], again[B200L2F_PARAMS_DIM];
        std::ifstream f(argv[2], std::ios::binary);
        f.read((char*)row, sizeof(row));
        rlt::utils::assert_exit(device, (size_t)f.gcount() == sizeof(row), "row file too short");
        const std::string text = rlt::json(device, row);
        for(int i = 0; i < B200L2F_PARAMS_DIM; i++) again[i] = -7.0f;
        rlt::from_json(device, text, again);
        std::printf("%zu characters\n", text.size());
        if(argc >= 5){ std::ofstream o(argv[4], std::ios::binary); o.write((const char*)again, sizeof(again)); }
        if(argc >= 4){ std::ofstream o(argv[3]); o << text; }
        return 0;
    }
    if(argc < 4){ std::fprintf(stderr, "usage: ppo_loop raptor|default blobs.f32 out.f32 | ppo_loop checkpoint FILE\n"); return 2; }
    return std::string(argv[1]) == "default" ? run<rlt::l2f::DefaultDRSpecification>(argv[2], argv[3]) : run<rlt::l2f::RaptorDRSpecification>(argv[2], argv[3]);
}
