// gelu.cuh -- exact-erf GELU, x * Phi(x), evaluated as x / (1 + 2^(-x * P(x^2))).
//
// P is the degree-6 least-squares fit (in x^2) of log2(Phi(x) / (1 - Phi(x))) / x, so the expression is the
// erf GELU itself, not the tanh approximation: max |error| against float64 0.5 x (1 + erf(x / sqrt 2)) is
// 6.0e-7 over [-12, 12] (7.6e-5 relative where |GELU| > 1e-3), i.e. the accuracy of erff-based fp32 code, for
// ~12 instructions (2 MUFU) instead of ~28.  GELU sits in the epilogue of half the encoder's GEMMs and in the
// depthwise convolutions (143 M evaluations per image), where its ALU cost was the bound (profiles/r01a).
// The leading coefficient is positive, so |x| -> inf saturates correctly to x (x > 0) and -0 (x < 0).
#pragma once

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

}  // namespace dlimg
