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

#include <cstdlib>

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

// ---- tensor-core form ----------------------------------------------------------------------------------------------
// The same partials from warp-level MMAs (mma.sync m16n8k16, fp16 operands, fp32 accumulators) -- the CUDA-core kernel
// above spends ~100 instructions per key and lane (profiles/r02a_summary.md: 80-92 us per 64-prompt launch for 134 MB).
// A warp owns 64 consecutive keys of its block's split and walks them in 4 steps of 16 keys, 4 heads at a time:
//   S^T (16 keys x 8 tokens) = K_h (16 x 16 dims, A operand straight from global memory: a fragment register is two
//         neighbouring dims of one key row) * Q_h^T (16 dims x 8 tokens: 7 + one zero row; B operand, resident registers)
//   pass 1 takes the maximum of every (head, token) column over the warp's 64 keys; pass 2 recomputes S^T (the K rows
//         come back from L1), p = 2^(s - max), and accumulates  O (8 tokens x 16 dims) += P (tokens x keys) * V_h:
//         P is S^T transposed -- one movmatrix per 8 x 8 block of the accumulator fragment -- and V's B fragment
//         (two KEYS of one dim per register) is the transposed natural row fragment, again by movmatrix.
// With the final maximum known before pass 2 nothing is ever rescaled.  Scores are kept in log2 units (q carries
// 0.25 * log2 e).  Output: the same per-(prompt, split) partials as above, merged by token_post_t2i.
__device__ __forceinline__ void mma16816_f16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t movm_t(uint32_t x) {
    uint32_t y;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 const v = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t const*>(&v);
}
__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

#if !defined(DLIMG_B200_ACT_BF16)
constexpr int kKeysPerWarp = kImgTokens / (kT2iSplits * kWarps);  // 64
static_assert(kKeysPerWarp == 64, "the tensor-core kernel walks 4 steps of 16 keys per warp");

__global__ void __launch_bounds__(kWarps * 32, 2) t2i_mma_kernel(float const* __restrict__ q, act_t const* __restrict__ base,
                                                              act_t const* const* __restrict__ ptrs, int64_t prompt_stride,
                                                              int pitch, int v_off, float* __restrict__ part) {
    __shared__ float sm_acc[kWarps][kTokens][128];
    __shared__ float sm_m[kWarps][kTokens][kHeads];
    __shared__ float sm_s[kWarps][kTokens][kHeads];
    int const p = blockIdx.x, split = blockIdx.y;
    int const tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int const g = lane >> 2, t = lane & 3;
    act_t const* const Kb = (ptrs ? ptrs[p] : base + (size_t)p * prompt_stride) + (size_t)((split * kWarps + warp) * kKeysPerWarp) * pitch;
    act_t const* const Vb = Kb + v_off;
    float const kQScale = 0.25f * 1.4426950408889634f;  // 1 / sqrt(16) * log2 e
    float const* qrow = q + ((size_t)p * kTokens + g) * 128;  // token g (g == 7: the zero padding row)

#pragma unroll 1
    for (int hg = 0; hg < 2; ++hg) {  // four heads at a time (register budget)
        uint32_t qb[4][2];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            int const c = (hg * 4 + h) * 16 + 2 * t;
            float2 const x0 = g < kTokens ? *reinterpret_cast<float2 const*>(qrow + c) : make_float2(0.f, 0.f);
            float2 const x1 = g < kTokens ? *reinterpret_cast<float2 const*>(qrow + c + 8) : make_float2(0.f, 0.f);
            qb[h][0] = pack_h2(x0.x * kQScale, x0.y * kQScale);
            qb[h][1] = pack_h2(x1.x * kQScale, x1.y * kQScale);
        }
        auto scores = [&](int step, int h, float (&s)[4]) {  // S^T of 16 keys for head hg * 4 + h
            act_t const* r0 = Kb + (size_t)(step * 16 + g) * pitch + (hg * 4 + h) * 16 + 2 * t;
            act_t const* r1 = r0 + (size_t)8 * pitch;
            uint32_t const a0 = __ldg(reinterpret_cast<uint32_t const*>(r0)), a1 = __ldg(reinterpret_cast<uint32_t const*>(r1));
            uint32_t const a2 = __ldg(reinterpret_cast<uint32_t const*>(r0 + 8)), a3 = __ldg(reinterpret_cast<uint32_t const*>(r1 + 8));
            s[0] = s[1] = s[2] = s[3] = 0.f;
            mma16816_f16(s, a0, a1, a2, a3, qb[h][0], qb[h][1]);
        };
        // pass 1: column maxima (tokens 2t, 2t + 1) over the warp's keys
        float mx[4][2];
#pragma unroll
        for (int h = 0; h < 4; ++h) mx[h][0] = mx[h][1] = -INFINITY;
#pragma unroll
        for (int step = 0; step < 4; ++step)
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                float s[4];
                scores(step, h, s);
                mx[h][0] = fmaxf(mx[h][0], fmaxf(s[0], s[2]));
                mx[h][1] = fmaxf(mx[h][1], fmaxf(s[1], s[3]));
            }
#pragma unroll
        for (int h = 0; h < 4; ++h)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                float m = mx[h][e];
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
                mx[h][e] = m;
            }
        // pass 2: probabilities, sums, O += P V
        float o[4][2][4], l[4][2];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            l[h][0] = l[h][1] = 0.f;
#pragma unroll
            for (int j = 0; j < 2; ++j) o[h][j][0] = o[h][j][1] = o[h][j][2] = o[h][j][3] = 0.f;
        }
#pragma unroll
        for (int step = 0; step < 4; ++step)
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                float s[4];
                scores(step, h, s);
                float const p0 = ex2f(s[0] - mx[h][0]), p1 = ex2f(s[1] - mx[h][1]);
                float const p2 = ex2f(s[2] - mx[h][0]), p3 = ex2f(s[3] - mx[h][1]);
                l[h][0] += p0 + p2;
                l[h][1] += p1 + p3;
                uint32_t const pa0 = movm_t(pack_h2(p0, p1));  // P[token g][keys 2t, 2t + 1]
                uint32_t const pa2 = movm_t(pack_h2(p2, p3));  // P[token g][keys 8 + 2t, 9 + 2t]
                act_t const* r0 = Vb + (size_t)(step * 16 + g) * pitch + (hg * 4 + h) * 16 + 2 * t;
                act_t const* r1 = r0 + (size_t)8 * pitch;
                uint32_t const v00 = movm_t(__ldg(reinterpret_cast<uint32_t const*>(r0)));      // keys 0-7,  dims 0-7
                uint32_t const v10 = movm_t(__ldg(reinterpret_cast<uint32_t const*>(r1)));      // keys 8-15, dims 0-7
                uint32_t const v01 = movm_t(__ldg(reinterpret_cast<uint32_t const*>(r0 + 8)));  // keys 0-7,  dims 8-15
                uint32_t const v11 = movm_t(__ldg(reinterpret_cast<uint32_t const*>(r1 + 8)));  // keys 8-15, dims 8-15
                mma16816_f16(o[h][0], pa0, 0u, pa2, 0u, v00, v10);
                mma16816_f16(o[h][1], pa0, 0u, pa2, 0u, v01, v11);
            }
        // partials of this warp: accumulators (token g, dims 8 j + 2t, + 1), maxima / sums (tokens 2t, 2t + 1)
#pragma unroll
        for (int h = 0; h < 4; ++h) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                float s = l[h][e];
                s += __shfl_xor_sync(0xffffffffu, s, 4);
                s += __shfl_xor_sync(0xffffffffu, s, 8);
                s += __shfl_xor_sync(0xffffffffu, s, 16);
                int const tok = 2 * t + e;
                if (g == 0 && tok < kTokens) {
                    sm_m[warp][tok][hg * 4 + h] = mx[h][e] * 0.6931471805599453f;  // back to natural-log units
                    sm_s[warp][tok][hg * 4 + h] = s;
                }
            }
            if (g < kTokens) {
#pragma unroll
                for (int j = 0; j < 2; ++j)
                    *reinterpret_cast<float2*>(&sm_acc[warp][g][(hg * 4 + h) * 16 + 8 * j + 2 * t]) = make_float2(o[h][j][0], o[h][j][1]);
            }
        }
    }
    __syncthreads();
    // merge the 8 warps -> one partial per (prompt, split)
    float* dst = part + ((size_t)p * kT2iSplits + split) * kTokens * kPartStride;
    for (int idx = tid; idx < kTokens * 128; idx += kWarps * 32) {
        int const tk = idx >> 7, d = idx & 127, h = d >> 4;
        float M = sm_m[0][tk][h];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) M = fmaxf(M, sm_m[w][tk][h]);
        float A = 0.f, S = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            float const e = __expf(sm_m[w][tk][h] - M);
            A = fmaf(sm_acc[w][tk][d], e, A);
            S = fmaf(sm_s[w][tk][h], e, S);
        }
        dst[tk * kPartStride + d] = A;
        if ((d & 15) == 0) {
            dst[tk * kPartStride + 128 + h] = M;
            dst[tk * kPartStride + 136 + h] = S;
        }
    }
}
#endif

#if !defined(DLIMG_B200_ACT_BF16)
// Image -> tokens attention on tensor cores: a warp owns 16 consecutive image tokens of one prompt and walks the 8 heads.
//   S (16 image tokens x 8 tokens) = Q_h (16 x 16 dims: A fragments straight from the image stream in global memory)
//                                    * K_h^T (16 dims x 8: B fragments of the prompt's token keys, shared memory)
//   softmax over the 7 tokens of a row = over the four lanes of a quad (the padding column is masked),
//   O (16 x 16 dims) = P (16 x 8, zero-extended to k = 16: the accumulator fragment of S IS the A fragment) * V_h
//                      (B fragments prepared once per block: two TOKENS of one dim per register).
// ~30 instructions per (16 image tokens, head) and lane instead of ~350 per (token, head) and thread on CUDA cores.
__global__ void __launch_bounds__(256) i2t_mma_kernel(act_t const* __restrict__ Q, act_t const* const* __restrict__ Qptrs,
                                                      int64_t q_prompt_stride, int q_pitch, int q_off,
                                                      float const* __restrict__ kt, float const* __restrict__ vt,
                                                      act_t* __restrict__ out) {
    __shared__ uint32_t kb[kHeads][2][32];     // B fragments of K^T per head: [k-half][lane]
    __shared__ uint32_t vb[kHeads][2][32];     // B fragments of V per head: [dim tile][lane] (k = tokens 0..7; 8..15 are zero)
    int const p = blockIdx.y;
    int const tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    float const kScale = 0.25f * 1.4426950408889634f;  // 1 / sqrt(16) * log2 e, carried by K
    for (int i = tid; i < kHeads * 2 * 32; i += 256) {
        int const h = i >> 6, half = (i >> 5) & 1, l = i & 31, gg = l >> 2, tt = l & 3;
        // K^T fragment (k = dim, n = token gg): dims 2tt, 2tt+1 (+8 for the second half) of token gg (token 7 = padding)
        float2 kv = make_float2(0.f, 0.f);
        if (gg < kTokens) kv = *reinterpret_cast<float2 const*>(kt + ((size_t)p * kTokens + gg) * 128 + h * 16 + half * 8 + 2 * tt);
        kb[h][half][l] = pack_h2(kv.x * kScale, kv.y * kScale);
        // V fragment (k = token, n = dim gg of tile `half`): tokens 2tt, 2tt+1
        int const d = h * 16 + half * 8 + gg;
        float const v0 = 2 * tt < kTokens ? vt[((size_t)p * kTokens + 2 * tt) * 128 + d] : 0.f;
        float const v1 = 2 * tt + 1 < kTokens ? vt[((size_t)p * kTokens + 2 * tt + 1) * 128 + d] : 0.f;
        vb[h][half][l] = pack_h2(v0, v1);
    }
    __syncthreads();
    int const tok0 = (blockIdx.x * 8 + warp) * 16;  // 8 warps x 16 image tokens per block
    act_t const* qbase = (Qptrs ? Qptrs[p] : Q + (size_t)p * q_prompt_stride) + q_off;
    act_t const* r0 = qbase + (size_t)(tok0 + g) * q_pitch + 2 * t;
    act_t const* r1 = r0 + (size_t)8 * q_pitch;
    act_t* o0 = out + ((size_t)p * kImgTokens + tok0 + g) * 128 + 2 * t;
    act_t* o1 = o0 + (size_t)8 * 128;
    bool const pad_col = t == 3;  // this lane's second column is token 7: the padding column
#pragma unroll
    for (int h = 0; h < kHeads; ++h) {
        uint32_t const a0 = __ldg(reinterpret_cast<uint32_t const*>(r0 + h * 16)), a1 = __ldg(reinterpret_cast<uint32_t const*>(r1 + h * 16));
        uint32_t const a2 = __ldg(reinterpret_cast<uint32_t const*>(r0 + h * 16 + 8)), a3 = __ldg(reinterpret_cast<uint32_t const*>(r1 + h * 16 + 8));
        float s[4] = {0.f, 0.f, 0.f, 0.f};
        mma16816_f16(s, a0, a1, a2, a3, kb[h][0][lane], kb[h][1][lane]);
        if (pad_col) s[1] = s[3] = -INFINITY;
        float m0 = fmaxf(s[0], s[1]), m1 = fmaxf(s[2], s[3]);
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
        float const p0 = ex2f(s[0] - m0), p1 = ex2f(s[1] - m0), p2 = ex2f(s[2] - m1), p3 = ex2f(s[3] - m1);
        float l0 = p0 + p1, l1 = p2 + p3;
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
        l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        float const i0 = 1.0f / l0, i1 = 1.0f / l1;
        uint32_t const pa0 = pack_h2(p0 * i0, p1 * i0), pa1 = pack_h2(p2 * i1, p3 * i1);  // rows g / g + 8, k = tokens 2t, 2t + 1
        float oa[4] = {0.f, 0.f, 0.f, 0.f}, ob[4] = {0.f, 0.f, 0.f, 0.f};
        mma16816_f16(oa, pa0, pa1, 0u, 0u, vb[h][0][lane], 0u);  // dims 0-7 of the head
        mma16816_f16(ob, pa0, pa1, 0u, 0u, vb[h][1][lane], 0u);  // dims 8-15
        *reinterpret_cast<uint32_t*>(o0 + h * 16) = pack_h2(oa[0], oa[1]);
        *reinterpret_cast<uint32_t*>(o1 + h * 16) = pack_h2(oa[2], oa[3]);
        *reinterpret_cast<uint32_t*>(o0 + h * 16 + 8) = pack_h2(ob[0], ob[1]);
        *reinterpret_cast<uint32_t*>(o1 + h * 16 + 8) = pack_h2(ob[2], ob[3]);
    }
}
#endif

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
#if defined(DLIMG_B200_ACT_BF16)
        t2i_flash_kernel<<<dim3(P, kT2iSplits), kWarps * 32, 0, s>>>(q, base, ptrs, prompt_stride, pitch, v_off, scratch);
#else
        static bool const cuda_core = std::getenv("DLIMG_B200_T2I_SIMT") != nullptr;  // cross-check: the CUDA-core form
        if (cuda_core) t2i_flash_kernel<<<dim3(P, kT2iSplits), kWarps * 32, 0, s>>>(q, base, ptrs, prompt_stride, pitch, v_off, scratch);
        else t2i_mma_kernel<<<dim3(P, kT2iSplits), kWarps * 32, 0, s>>>(q, base, ptrs, prompt_stride, pitch, v_off, scratch);
#endif
        KERNEL_CHECK();
    }
    if (!out) return;  // the partials are merged by token_post_t2i
    ProfScope prof(s, CAT_DEC_ATTN);
    t2i_combine_kernel<<<dim3(P, kTokens), 128, 0, s>>>(scratch, out);
    KERNEL_CHECK();
}

void image_to_token_attention_mma(cudaStream_t s, act_t const* Q, act_t const* const* Qptrs, int64_t q_prompt_stride, int q_pitch,
                                  int q_off, float const* kt, float const* vt, int P, act_t* out) {
#if defined(DLIMG_B200_ACT_BF16)
    image_to_token_attention(s, Q, Qptrs, q_prompt_stride, q_pitch, q_off, kt, vt, P, out);
#else
    static bool const cuda_core = std::getenv("DLIMG_B200_I2T_SIMT") != nullptr;  // cross-check: the CUDA-core form
    if (cuda_core) {
        image_to_token_attention(s, Q, Qptrs, q_prompt_stride, q_pitch, q_off, kt, vt, P, out);
        return;
    }
    ProfScope prof(s, CAT_DEC_ATTN, 4.0 * P * kTokens * kImgTokens * 128, (double)P * kImgTokens * 128 * 4);
    i2t_mma_kernel<<<dim3(kImgTokens / 128, P), 256, 0, s>>>(Q, Qptrs, q_prompt_stride, q_pitch, q_off, kt, vt, out);
    KERNEL_CHECK();
#endif
}

}  // namespace dec
}  // namespace dlimg
