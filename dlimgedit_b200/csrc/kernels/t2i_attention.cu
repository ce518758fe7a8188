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

#if DLIMG_B200_ALT  // CUDA-core form: the bf16 build and the cross-check of t2i_mma_kernel
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
#endif

// ---- tensor-core form ----------------------------------------------------------------------------------------------
// The same partials from warp-level MMAs (mma.sync m16n8k16, fp16 operands, fp32 accumulators) -- the CUDA-core kernel
// above spends ~100 instructions per key and lane (80-92 us per 64-prompt launch for 134 MB).
__device__ __forceinline__ void mma16816_f16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
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
constexpr int kTileKeys = 32;                    // keys per pipeline stage
constexpr int kT2iStages = 3;
constexpr int kTileBytes = kTileKeys * 512;      // [K tile: 32 rows x 256 B | V tile: 32 rows x 256 B]
constexpr int kSplitKeys = kImgTokens / kT2iSplits;
constexpr int kT2iSmem = kT2iStages * kTileBytes;  // 48 KB: four blocks per SM, so 64 prompts x 8 splits are one wave
static_assert(kWarps == kHeads, "one warp per head");
static_assert(kSplitKeys % kTileKeys == 0, "whole tiles per split");

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}

// Second form of the tensor-core kernel (the first read its fragments straight from global memory with 4-byte loads and
// was bound by L1 sector throughput, 76 %): a block owns one (prompt, split) = 512 keys and streams them through a
// 3-stage cp.async ring of 32-key tiles -- whole 256-byte K and V rows, 16-byte chunks XOR-swizzled by the row -- and
// each of the 8 warps owns one HEAD, so no cross-warp merge is needed.  Per tile and warp:
//   S (tokens x 32 keys)  = Q_h (A: 7 tokens + zero rows, resident) * K_h^T (B: ldmatrix of the K rows)  -- 4 mma
//   online softmax per tile (maximum over the quad, one rescale of the accumulators per tile), scores stay in registers
//   O (tokens x 16 dims) += P (A: the accumulator fragment of S, packed to fp16) * V_h (B: ldmatrix.trans) -- 4 mma
// No transposes, no second pass over K.  Scores are kept in log2 units (q carries 0.25 * log2 e).
// Output: per-(prompt, split) partials, merged by token_post_t2i.
__global__ void __launch_bounds__(kWarps * 32, 4) t2i_mma_kernel(float const* __restrict__ q, act_t const* __restrict__ base,
                                                              act_t const* const* __restrict__ ptrs, int64_t prompt_stride,
                                                              int pitch, int v_off, float* __restrict__ part) {
    extern __shared__ __align__(128) uint8_t t2i_smem[];
    int const p = blockIdx.x, split = blockIdx.y;
    int const tid = threadIdx.x, h = tid >> 5, lane = tid & 31;
    int const g = lane >> 2, t = lane & 3;
    act_t const* const Kb = (ptrs ? ptrs[p] : base + (size_t)p * prompt_stride) + (size_t)(split * kSplitKeys) * pitch;
    act_t const* const Vb = Kb + v_off;
    uint32_t const smem_s = (uint32_t)__cvta_generic_to_shared(t2i_smem);

    auto issue = [&](int tile) {  // 32 keys x (16 + 16) chunks of 16 bytes: 4 per thread
        uint32_t const st = smem_s + (tile % kT2iStages) * kTileBytes;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int const c = i * 256 + tid, kv = c >> 9, row = (c >> 4) & 31, ch = c & 15;
            act_t const* src = (kv ? Vb : Kb) + (size_t)(tile * kTileKeys + row) * pitch + ch * 8;
            uint32_t const dst = st + kv * (kTileKeys * 256) + row * 256 + ((ch ^ (row & 7)) << 4);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        }
    };
    issue(0);
    asm volatile("cp.async.commit_group;" ::: "memory");
    issue(1);
    asm volatile("cp.async.commit_group;" ::: "memory");

    float const kQScale = 0.25f * 1.4426950408889634f;  // 1 / sqrt(16) * log2 e
    uint32_t qa0 = 0u, qa2 = 0u;  // A fragment of Q_h: row g = token g (row 7 and rows 8-15 are zero)
    if (g < kTokens) {
        float const* qrow = q + ((size_t)p * kTokens + g) * 128 + h * 16 + 2 * t;
        float2 const x0 = *reinterpret_cast<float2 const*>(qrow), x1 = *reinterpret_cast<float2 const*>(qrow + 8);
        qa0 = pack_h2(x0.x * kQScale, x0.y * kQScale);
        qa2 = pack_h2(x1.x * kQScale, x1.y * kQScale);
    }
    // ldmatrix.x4 lane addressing: matrices (keys 0-7 | 8-15) x (dims 0-7 | 8-15) of this head
    int const lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lchunk = 2 * h + (lane >> 4);
    float m = -INFINITY, l = 0.f;
    float o[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    constexpr int kTiles = kSplitKeys / kTileKeys;
#pragma unroll 1
    for (int tile = 0; tile < kTiles; ++tile) {
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();  // tile is visible; every warp is done with tile - 1, whose stage the next copy reuses
        if (tile + 2 < kTiles) issue(tile + 2);
        asm volatile("cp.async.commit_group;" ::: "memory");
        uint32_t const st = smem_s + (tile % kT2iStages) * kTileBytes;
        float sc[kTileKeys / 16][2][4];
#pragma unroll
        for (int step = 0; step < kTileKeys / 16; ++step) {
            int const row = step * 16 + lrow;
            uint32_t k0, k1, k2, k3;  // (keys 0-7, dims 0-7), (keys 8-15, dims 0-7), (keys 0-7, dims 8-15), (keys 8-15, dims 8-15)
            ldsm_x4(st + row * 256 + ((lchunk ^ (row & 7)) << 4), k0, k1, k2, k3);
#pragma unroll
            for (int j = 0; j < 2; ++j) sc[step][j][0] = sc[step][j][1] = sc[step][j][2] = sc[step][j][3] = 0.f;
            mma16816_f16(sc[step][0], qa0, 0u, qa2, 0u, k0, k2);  // keys 0-7 of the step
            mma16816_f16(sc[step][1], qa0, 0u, qa2, 0u, k1, k3);  // keys 8-15
        }
        float mt = -INFINITY;
#pragma unroll
        for (int step = 0; step < kTileKeys / 16; ++step)
#pragma unroll
            for (int j = 0; j < 2; ++j) mt = fmaxf(mt, fmaxf(sc[step][j][0], sc[step][j][1]));
        mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, 1));
        mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, 2));
        float const m_new = fmaxf(m, mt);
        float const alpha = ex2f(m - m_new);  // first tile: 2^(-inf) = 0
        m = m_new;
        l *= alpha;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            o[j][0] *= alpha;
            o[j][1] *= alpha;
        }
#pragma unroll
        for (int step = 0; step < kTileKeys / 16; ++step) {
            float const p0 = ex2f(sc[step][0][0] - m), p1 = ex2f(sc[step][0][1] - m);
            float const p2 = ex2f(sc[step][1][0] - m), p3 = ex2f(sc[step][1][1] - m);
            l += (p0 + p1) + (p2 + p3);
            uint32_t const pa0 = pack_h2(p0, p1), pa2 = pack_h2(p2, p3);  // P[token g][keys 2t, 2t + 1 | 8 + 2t, 9 + 2t]
            int const row = step * 16 + lrow;
            uint32_t v00, v10, v01, v11;  // B fragments of V: (keys 0-7 | 8-15) x (dims 0-7 | 8-15), two KEYS per register
            ldsm_x4_t(st + kTileKeys * 256 + row * 256 + ((lchunk ^ (row & 7)) << 4), v00, v10, v01, v11);
            mma16816_f16(o[0], pa0, 0u, pa2, 0u, v00, v10);
            mma16816_f16(o[1], pa0, 0u, pa2, 0u, v01, v11);
        }
    }
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    if (g < kTokens) {
        float* dst = part + (((size_t)p * kT2iSplits + split) * kTokens + g) * kPartStride;
#pragma unroll
        for (int j = 0; j < 2; ++j) *reinterpret_cast<float2*>(dst + h * 16 + 8 * j + 2 * t) = make_float2(o[j][0], o[j][1]);
        if (t == 0) {
            dst[128 + h] = m * 0.6931471805599453f;  // back to natural-log units
            dst[136 + h] = l;
        }
    }
}
#endif

#if !defined(DLIMG_B200_ACT_BF16)
// Image -> tokens attention on tensor cores: a warp owns 16 consecutive image tokens of one prompt and walks the 8 heads.
//   S (16 image tokens x 8 tokens) = Q_h (16 x 16 dims: A fragments by ldmatrix from the warp's staged tile)
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
    __shared__ __align__(16) uint8_t tile[8][16 * 256];  // per warp: the Q tile, then the output tile
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
    // The warp's Q tile (16 tokens x 128 dims, 4 KB) goes through shared memory: 16-byte loads of whole rows (fully used
    // sectors; fragment-shaped 4-byte loads kept L1 at 81 % of its sector rate), stored with the 16-byte chunk index
    // XOR-swizzled by the row so that ldmatrix and the fragment-shaped output stores are conflict free.  The outputs of a
    // head overwrite the chunks its Q fragments came from; the tile leaves again as whole rows.
    uint8_t* const tw = tile[warp];
    {
        uint4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int const c = i * 32 + lane, row = c >> 4, ch = c & 15;
            v[i] = __ldg(reinterpret_cast<uint4 const*>(qbase + (size_t)(tok0 + row) * q_pitch) + ch);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int const c = i * 32 + lane, row = c >> 4, ch = c & 15;
            *reinterpret_cast<uint4*>(tw + row * 256 + ((ch ^ (row & 7)) << 4)) = v[i];
        }
    }
    __syncwarp();
    int const lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lhalf = lane >> 4;  // ldmatrix.x4: matrices (rows 0-7 | 8-15) x (dims 0-7 | 8-15)
    uint32_t const tw_s = (uint32_t)__cvta_generic_to_shared(tw);
    bool const pad_col = t == 3;  // this lane's second column is token 7: the padding column
#pragma unroll
    for (int h = 0; h < kHeads; ++h) {
        uint32_t a0, a1, a2, a3;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                     : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3)
                     : "r"(tw_s + lrow * 256 + (((2 * h + lhalf) ^ (lrow & 7)) << 4))
                     : "memory");
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
        uint8_t* const c0 = tw + (((2 * h) ^ g) << 4) + 4 * t;      // dims 0-7: chunk 2h (rows g and g + 8 swizzle alike)
        uint8_t* const c1 = tw + (((2 * h + 1) ^ g) << 4) + 4 * t;  // dims 8-15
        *reinterpret_cast<uint32_t*>(c0 + g * 256) = pack_h2(oa[0], oa[1]);
        *reinterpret_cast<uint32_t*>(c0 + (g + 8) * 256) = pack_h2(oa[2], oa[3]);
        *reinterpret_cast<uint32_t*>(c1 + g * 256) = pack_h2(ob[0], ob[1]);
        *reinterpret_cast<uint32_t*>(c1 + (g + 8) * 256) = pack_h2(ob[2], ob[3]);
    }
    __syncwarp();
    act_t* const obase = out + ((size_t)p * kImgTokens + tok0) * 128;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int const c = i * 32 + lane, row = c >> 4, ch = c & 15;
        reinterpret_cast<uint4*>(obase + (size_t)row * 128)[ch] = *reinterpret_cast<uint4 const*>(tw + row * 256 + ((ch ^ (row & 7)) << 4));
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
#if DLIMG_B200_ALT
        static bool const cuda_core = dev_switch("DLIMG_B200_T2I_SIMT");  // cross-check: the CUDA-core form
        if (cuda_core) t2i_flash_kernel<<<dim3(P, kT2iSplits), kWarps * 32, 0, s>>>(q, base, ptrs, prompt_stride, pitch, v_off, scratch);
        else
#endif
            t2i_mma_kernel<<<dim3(P, kT2iSplits), kWarps * 32, kT2iSmem, s>>>(q, base, ptrs, prompt_stride, pitch, v_off, scratch);
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
#if DLIMG_B200_ALT
    static bool const cuda_core = dev_switch("DLIMG_B200_I2T_SIMT");  // cross-check: the CUDA-core form
    if (cuda_core) {
        image_to_token_attention(s, Q, Qptrs, q_prompt_stride, q_pitch, q_off, kt, vt, P, out);
        return;
    }
#endif
    ProfScope prof(s, CAT_DEC_ATTN, 4.0 * P * kTokens * kImgTokens * 128, (double)P * kImgTokens * 128 * 4);
    i2t_mma_kernel<<<dim3(kImgTokens / 128, P), 256, 0, s>>>(Q, Qptrs, q_prompt_stride, q_pitch, q_off, kt, vt, out);
    KERNEL_CHECK();
#endif
}

}  // namespace dec
}  // namespace dlimg
