// t2i_attention.cu -- token -> image attention core of the SAM two-way transformer, flash-decoding style.
//
// 7 query tokens x 8 heads (16 dims each) attend to 4096 image keys.  Every K / V row is 128 contiguous
// 16-bit values (all heads), so a warp streams whole rows with one coalesced 256-byte load each: lane l owns
// dims 4l..4l+3 (head l/4), the 4 lanes of a head finish the 16-dim dot product with two shuffles, and
// each lane keeps an online-softmax state (max, sum, 4 accumulators) for the 7 tokens.  Keys are split
// over 8 warps x kT2iSplits blocks per prompt; partial (max, sum, acc) triples are merged in shared memory
// and then by a small combine kernel.  One pass over K and V, no score matrix, all loads coalesced.
#include "decoder_kernels.cuh"

#include "../profiler.hpp"

namespace dlimg {
namespace dec {

namespace {

constexpr int kWarps = 8;
constexpr int kChunk = 4;  // keys per warp iteration (8 independent 512-byte row loads in flight)
constexpr int kPartStride = 128 + 16;  // per (split, token): 128 accumulators, 8 maxima, 8 sums

__device__ __forceinline__ float4 load4(act_t const* p) {
    uint2 const u = __ldg(reinterpret_cast<uint2 const*>(p));
    float2 const a = act22f2(*reinterpret_cast<act2_t const*>(&u.x)), b = act22f2(*reinterpret_cast<act2_t const*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
}

__global__ void __launch_bounds__(kWarps * 32) t2i_flash_kernel(float const* __restrict__ q, act_t const* __restrict__ base,
                                                                act_t const* const* __restrict__ ptrs, int64_t prompt_stride,
                                                                int pitch, int v_off, float* __restrict__ part) {
    __shared__ float sm_acc[kWarps][kTokens][128];
    __shared__ float sm_m[kWarps][kTokens][kHeads];
    __shared__ float sm_s[kWarps][kTokens][kHeads];
    int const p = blockIdx.x, split = blockIdx.y;
    int const tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    act_t const* Kp = (ptrs ? ptrs[p] : base + (size_t)p * prompt_stride) + 4 * lane;
    act_t const* Vp = Kp + v_off;

    float qr[kTokens][4];
#pragma unroll
    for (int t = 0; t < kTokens; ++t) {
        float4 const x = *reinterpret_cast<float4 const*>(q + ((size_t)p * kTokens + t) * 128 + 4 * lane);
        qr[t][0] = x.x * 0.25f; qr[t][1] = x.y * 0.25f; qr[t][2] = x.z * 0.25f; qr[t][3] = x.w * 0.25f;  // 1/sqrt(16)
    }
    float m[kTokens], s[kTokens], acc[kTokens][4];
#pragma unroll
    for (int t = 0; t < kTokens; ++t) {
        m[t] = -INFINITY;
        s[t] = 0.f;
        acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;
    }
    int const keys_per_split = kImgTokens / kT2iSplits;
    int const i_end = (split + 1) * keys_per_split;
    for (int i = split * keys_per_split + warp * kChunk; i < i_end; i += kWarps * kChunk) {
        float4 kk[kChunk], vv[kChunk];
#pragma unroll
        for (int c = 0; c < kChunk; ++c) {
            kk[c] = load4(Kp + (size_t)(i + c) * pitch);
            vv[c] = load4(Vp + (size_t)(i + c) * pitch);
        }
#pragma unroll
        for (int t = 0; t < kTokens; ++t) {
            float sc[kChunk];
            float cm = -INFINITY;
#pragma unroll
            for (int c = 0; c < kChunk; ++c) {
                float d = qr[t][0] * kk[c].x;
                d = fmaf(qr[t][1], kk[c].y, d);
                d = fmaf(qr[t][2], kk[c].z, d);
                d = fmaf(qr[t][3], kk[c].w, d);
                d += __shfl_xor_sync(0xffffffffu, d, 1);
                d += __shfl_xor_sync(0xffffffffu, d, 2);
                sc[c] = d;
                cm = fmaxf(cm, d);
            }
            float const mn = fmaxf(m[t], cm);
            float const corr = __expf(m[t] - mn);  // exp(-inf) = 0 on the first chunk
            s[t] *= corr;
            acc[t][0] *= corr; acc[t][1] *= corr; acc[t][2] *= corr; acc[t][3] *= corr;
#pragma unroll
            for (int c = 0; c < kChunk; ++c) {
                float const pc = __expf(sc[c] - mn);
                s[t] += pc;
                acc[t][0] = fmaf(pc, vv[c].x, acc[t][0]);
                acc[t][1] = fmaf(pc, vv[c].y, acc[t][1]);
                acc[t][2] = fmaf(pc, vv[c].z, acc[t][2]);
                acc[t][3] = fmaf(pc, vv[c].w, acc[t][3]);
            }
            m[t] = mn;
        }
    }
#pragma unroll
    for (int t = 0; t < kTokens; ++t) {
        *reinterpret_cast<float4*>(&sm_acc[warp][t][4 * lane]) = make_float4(acc[t][0], acc[t][1], acc[t][2], acc[t][3]);
        if ((lane & 3) == 0) {
            sm_m[warp][t][lane >> 2] = m[t];
            sm_s[warp][t][lane >> 2] = s[t];
        }
    }
    __syncthreads();
    // merge the 8 warps -> one partial per (prompt, split)
    float* dst = part + ((size_t)p * kT2iSplits + split) * kTokens * kPartStride;
    for (int idx = tid; idx < kTokens * 128; idx += kWarps * 32) {
        int const t = idx >> 7, d = idx & 127, h = d >> 4;
        float M = sm_m[0][t][h];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) M = fmaxf(M, sm_m[w][t][h]);
        float A = 0.f, S = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            float const e = __expf(sm_m[w][t][h] - M);
            A = fmaf(sm_acc[w][t][d], e, A);
            S = fmaf(sm_s[w][t][h], e, S);
        }
        dst[t * kPartStride + d] = A;
        if ((d & 15) == 0) {
            dst[t * kPartStride + 128 + h] = M;
            dst[t * kPartStride + 136 + h] = S;
        }
    }
}

__global__ void __launch_bounds__(128) t2i_combine_kernel(float const* __restrict__ part, float* __restrict__ out) {
    int const p = blockIdx.x, t = blockIdx.y, d = threadIdx.x, h = d >> 4;
    float const* src = part + ((size_t)p * kT2iSplits * kTokens + t) * kPartStride;
    size_t const split_stride = (size_t)kTokens * kPartStride;
    float M = -INFINITY;
#pragma unroll
    for (int sp = 0; sp < kT2iSplits; ++sp) M = fmaxf(M, src[sp * split_stride + 128 + h]);
    float A = 0.f, S = 0.f;
#pragma unroll
    for (int sp = 0; sp < kT2iSplits; ++sp) {
        float const e = __expf(src[sp * split_stride + 128 + h] - M);
        A = fmaf(src[sp * split_stride + d], e, A);
        S = fmaf(src[sp * split_stride + 136 + h], e, S);
    }
    out[((size_t)p * kTokens + t) * 128 + d] = A / S;
}

}  // namespace

void token_to_image_attention(cudaStream_t s, float const* q, act_t const* base, act_t const* const* ptrs, int64_t prompt_stride,
                              int pitch, int v_off, int P, float* scratch, float* out) {
    static_assert(kT2iScratchPerPrompt == (size_t)kT2iSplits * kTokens * kPartStride, "scratch layout");
    static_assert(kImgTokens % (kT2iSplits * kWarps * kChunk) == 0, "key split must be even");
    {
        ProfScope prof(s, CAT_DEC_ATTN, 4.0 * P * kTokens * kImgTokens * 128, (double)P * kImgTokens * 128 * 4);
        t2i_flash_kernel<<<dim3(P, kT2iSplits), kWarps * 32, 0, s>>>(q, base, ptrs, prompt_stride, pitch, v_off, scratch);
        KERNEL_CHECK();
    }
    ProfScope prof(s, CAT_DEC_ATTN);
    t2i_combine_kernel<<<dim3(P, kTokens), 128, 0, s>>>(scratch, out);
    KERNEL_CHECK();
}

}  // namespace dec
}  // namespace dlimg
