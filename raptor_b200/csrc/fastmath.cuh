// raptor_b200/csrc/fastmath.cuh -- the approximate special functions of the default-math kernels (B200L2F_FLAG_ACCURATE_MATH selects the
// libdevice functions instead).  Each is one MUFU instruction (+ a multiply), ~2 ulp.
#pragma once

namespace b200l2f {

constexpr float LOG2E = 1.4426950408889634f;
// .ftz: a denormal result is 0, which is exact enough next to the 1.0 it is added to in the gate activations
__device__ __forceinline__ float ex2_approx(float x){ float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x){ float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_of_scaled(float t){ return rcp_approx(1.0f + ex2_approx(t)); }              // t = -log2(e) * a
__device__ __forceinline__ float tanh_of_scaled(float t){ return fmaf(-2.0f, rcp_approx(ex2_approx(t) + 1.0f), 1.0f); } // t = 2 log2(e) * a
__device__ __forceinline__ float sqrt_approx(float x){ float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rsqrt_approx(float x){ float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// three-input maximum (max.f32 d, a, b, c: sm_100 FMNMX3, operands take |.| modifiers); NaN operands are ignored like fmaxf
__device__ __forceinline__ float max3(float a, float b, float c){ float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
// acos(1 - u) for u in [0, 1]: sqrt(2 u) Q(u), Q a degree-7 minimax fit (|error| < 2.5e-7 with fp32 Horner); replaces libdevice's acosf (~30
// instructions) in the orientation cost 2 acos(1 - |q_z|) of the default-math kernels (squared/operations_generic.h:24)
__device__ __forceinline__ float acos_1m_fast(float u){
    float q = 8.746944950e-04f;
    q = fmaf(q, u, -1.469472889e-03f); q = fmaf(q, u, 2.444653539e-03f); q = fmaf(q, u, 9.892442031e-04f); q = fmaf(q, u, 5.829527043e-03f);
    q = fmaf(q, u, 1.871711761e-02f); q = fmaf(q, u, 8.333496749e-02f); q = fmaf(q, u, 1.0f);
    return sqrt_approx(u + u) * q;
}
template <bool FAST> __device__ __forceinline__ float sqrt_t(float x){ if constexpr(FAST) return sqrt_approx(x); else return sqrtf(x); }
// clamp: the literal form propagates NaN like the reference's math::clamp; the min/max form is two instructions instead of four
__device__ __forceinline__ float clamp_literal(float x, float lo, float hi){ return x < lo ? lo : (x > hi ? hi : x); }
template <bool FAST> __device__ __forceinline__ float clamp_t(float x, float lo, float hi){ if constexpr(FAST) return fminf(fmaxf(x, lo), hi); else return clamp_literal(x, lo, hi); }

}  // namespace b200l2f
