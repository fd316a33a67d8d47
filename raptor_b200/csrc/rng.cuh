// raptor_b200/csrc/rng.cuh -- per-environment RNG streams.
//
// Contract (DESIGN.md "RNG"): the reference's Generic xorshift64 engine
//   init     rl_tools/random/operations_generic.h:16-18   state = 0xAAAAAAAA + seed
//   next     rl_tools/random/operations_generic.h:26-31   x ^= x<<13; x ^= x>>17; x ^= x<<5
//   uniform  rl_tools/random/operations_generic.h:52-58   state / (T)MAX_INDEX * (hi-lo) + lo      (MAX_INDEX = 2^64-1 -> (float) = 2^64)
//   normal   rl_tools/random/operations_generic.h:59-71   Box-Muller, cosine branch, always two draws
// combined with the CPU device's rule that a zero standard deviation returns the mean WITHOUT touching the
// stream (rl_tools/random/operations_cpu.h:39-41).  One stream per environment, seeded seed + global_env_id
// (precedent: rl_tools/rl/components/on_policy_runner/operations_cpu.h:36-44).  The integer stream is bit-exact
// with the oracle; the float transforms agree to the last ulp of logf/cosf.
#pragma once
#include <cstdint>
#include "fastmath.cuh"

namespace b200l2f {

__host__ __device__ __forceinline__ uint64_t rng_seed_state(uint64_t seed){ return 0xAAAAAAAAull + seed; }

__device__ __forceinline__ void rng_next(uint64_t& s){
    s ^= (s << 13);
    s ^= (s >> 17);
    s ^= (s << 5);
}
// state / (float)MAX_INDEX: the conversion rounds to nearest, the division by 2^64 is exact
__device__ __forceinline__ float rng_unit(uint64_t& s){
    rng_next(s);
    return __ull2float_rn(s) * 5.42101086242752217e-20f;  // 2^-64
}
__device__ __forceinline__ float rng_uniform(uint64_t& s, float lo, float hi){
    float u = rng_unit(s);
    return __fadd_rn(__fmul_rn(u, __fsub_rn(hi, lo)), lo);  // no FMA contraction: bit-exact with the oracle
}
__device__ __forceinline__ float rng_normal_draw(uint64_t& s, float mean, float std){
    float u1 = rng_unit(s);
    float u2 = rng_unit(s);
    // the reference evaluates sqrt(-2.0 * log(u1)) and 2.0 * PI<float> * u2 in double (double literals) and rounds to float; -2 * logf(u1) and
    // 2 * PI<float> are exact in float, so the float evaluation below differs from that only by double rounding (< 1 ulp, rare) and keeps
    // fp64 instructions out of the hot loop
    float x = sqrtf(-2.0f * logf(u1));
    float y = __fmul_rn(6.28318548202514648f, u2);
    float z = __fmul_rn(x, cosf(y));
    return __fadd_rn(__fmul_rn(z, std), mean);
}
// Out-of-line twin: the collection kernels contain ~30 draw sites (observation / action noise, action sampling, reset samplers); the
// inlined logf / cosf bodies would double their size and push them out of the instruction cache (profiles/r01_configs34.md).  The rollout
// kernels have three sites (Langevin target) and keep the inlined draw (a call there costs 4 %).
static __device__ __noinline__ float rng_normal_draw_ool(uint64_t& s, float mean, float std){ return rng_normal_draw(s, mean, std); }
// default-math twin: same two uniforms (the integer stream stays bit-exact), MUFU logarithm / cosine / square root (absolute error ~5e-7)
__device__ __forceinline__ float rng_normal_draw_fast(uint64_t& s, float mean, float std){
    const float u1 = rng_unit(s);
    float u2 = rng_unit(s);
    float l2; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u1));   // u1 >= 2^-64: never denormal, no scaling needed (what __logf adds)
    const float x = sqrt_approx(-1.38629436111989062f * l2);              // -2 ln u1 = -2 ln 2 * log2 u1
    if(u2 > 0.5f) u2 -= 1.0f;                            // cos(2 pi u) with the argument reduced to (-pi, pi], where cos.approx is tight
    const float z = x * __cosf(6.28318548202514648f * u2);
    return fmaf(z, std, mean);
}
// out of line with the stream state passed and returned BY VALUE (registers): a reference parameter makes the caller spill the state to local memory around
// every call (LDL / STL + their latency on the dependent chain of draws)
struct NormalDraw { uint64_t s; float z; };
static __device__ __noinline__ NormalDraw rng_normal_draw_fast_ool_value(uint64_t s, float mean, float std){ NormalDraw r; r.z = rng_normal_draw_fast(s, mean, std); r.s = s; return r; }
__device__ __forceinline__ float rng_normal_draw_fast_ool(uint64_t& s, float mean, float std){ const NormalDraw r = rng_normal_draw_fast_ool_value(s, mean, std); s = r.s; return r.z; }
template <bool OOL, bool FAST = false>
__device__ __forceinline__ float rng_normal_t(uint64_t& s, float mean, float std){
    if(std == 0.0f){ return mean; }
    if constexpr(FAST && OOL) return rng_normal_draw_fast_ool(s, mean, std);
    else if constexpr(FAST) return rng_normal_draw_fast(s, mean, std);
    else if constexpr(OOL) return rng_normal_draw_ool(s, mean, std);
    else return rng_normal_draw(s, mean, std);
}
__device__ __forceinline__ float rng_normal(uint64_t& s, float mean, float std){ return rng_normal_t<false>(s, mean, std); }

}  // namespace b200l2f
