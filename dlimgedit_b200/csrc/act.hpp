// act.hpp -- storage type of the encoder's activations and tensor-core operands.
//
// Default is IEEE fp16 (10-bit mantissa): MobileSAM's activations are LayerNorm/BatchNorm bounded, so the
// extra 3 mantissa bits over bf16 cut the accumulated rounding error of the 40-layer encoder ~8x at the same
// tcgen05 (kind::f16) rate; conversions saturate at +-65504 so an outlier can never become inf/NaN.
// -DDLIMG_B200_ACT_BF16 switches every kernel to bf16 storage instead.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace dlimg {

#if defined(DLIMG_B200_ACT_BF16)
using act_t = __nv_bfloat16;
using act2_t = __nv_bfloat162;
constexpr bool kActBf16 = true;
__host__ __device__ inline float act2f(act_t v) { return __bfloat162float(v); }
__host__ __device__ inline act_t f2act(float v) { return __float2bfloat16_rn(v); }
__device__ inline float2 act22f2(act2_t v) { return __bfloat1622float2(v); }
__device__ inline act2_t f22act2(float a, float b) { return __floats2bfloat162_rn(a, b); }
#else
using act_t = __half;
using act2_t = __half2;
constexpr bool kActBf16 = false;
__host__ __device__ inline float act_sat(float v) { return v > 65504.0f ? 65504.0f : (v < -65504.0f ? -65504.0f : v); }
__host__ __device__ inline float act2f(act_t v) { return __half2float(v); }
__host__ __device__ inline act_t f2act(float v) { return __float2half_rn(act_sat(v)); }
__device__ inline float2 act22f2(act2_t v) { return __half22float2(v); }
// one F2FP.SATFINITE.PACK_AB instead of four FMNMX + F2FP: the epilogues and depthwise kernels are issue-bound
__device__ inline act2_t f22act2(float a, float b) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return *reinterpret_cast<act2_t const*>(&r);
}
#endif

}  // namespace dlimg
