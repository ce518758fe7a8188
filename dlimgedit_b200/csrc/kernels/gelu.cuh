// gelu.cuh -- exact-erf GELU, x * Phi(x).
//
// gelu_erf (fp32): x / (1 + 2^(-x * P(x^2))).  P is the degree-6 least-squares fit (in x^2) of
// log2(Phi(x) / (1 - Phi(x))) / x, so the expression is the erf GELU itself, not the tanh approximation: max
// |error| against float64 0.5 x (1 + erf(x / sqrt 2)) is 6.0e-7 over [-12, 12].  ~12 instructions, 2 MUFU.
//
// gelu_erf_h2 (packed fp16 pair): 0.5 x (1 + tanh(x * Q(x^2))) with Q the weighted least-squares fit of
// atanh(erf(x / sqrt 2)) / x, i.e. again the erf GELU (NOT the 0.044715 tanh approximation), evaluated with
// HFMA2 and one tanh.approx.f16x2 per pair: 6 fma-pipe + 1 alu + 1 MUFU instructions per TWO values.  GELU sits in
// the epilogue of half the encoder's GEMMs and in the depthwise convolutions (143 M evaluations per image), where
// the fp32 form's ALU/MUFU cost was the bound (profiles/r01a, r01b).  Its result is stored as fp16 anyway: on
// N(0, 1.2) inputs the rms error against float64 is 2.4e-4 versus 1.6e-4 for the correctly rounded fp16 result
// (tools/fit_gelu.py reproduces the fit and both figures).
#pragma once

#include <cuda_fp16.h>

namespace dlimg {

__device__ __forceinline__ float gelu_erf(float x) {
    float const t = x * x;
    float p = 5.393212099136235e-09f;
    p = fmaf(p, t, -3.9339354884759814e-07f);
    p = fmaf(p, t, 1.1591534530452918e-05f);
    p = fmaf(p, t, -0.0001604528952157125f);
    p = fmaf(p, t, -9.176623279927298e-05f);
    p = fmaf(p, t, 0.10483267903327942f);
    p = fmaf(p, t, 2.3022100925445557f);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-x * p));
    return __fdividef(x, 1.0f + e);
}

__device__ __forceinline__ __half2 gelu_erf_h2(__half2 x) {
    // Q(t) on t = min(x^2, 16): beyond |x| = 4 the argument x * Q(16) = 1.3 x already saturates tanh in fp16
    __half2 const t = __hmin2(__hmul2(x, x), __float2half2_rn(16.0f));
    __half2 p = __hfma2(__float2half2_rn(-0.0003587913347867181f), t, __float2half2_rn(0.037050605989113146f));
    p = __hfma2(p, t, __float2half2_rn(0.7974582454000725f));
    __half2 const u = __hmul2(x, p);
    uint32_t th;
    asm("tanh.approx.f16x2 %0, %1;" : "=r"(th) : "r"(*reinterpret_cast<uint32_t const*>(&u)));
    __half2 const hx = __hmul2(x, __float2half2_rn(0.5f));
    return __hfma2(hx, *reinterpret_cast<__half2 const*>(&th), hx);
}

}  // namespace dlimg
