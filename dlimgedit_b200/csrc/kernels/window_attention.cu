// window_attention.cu -- TinyViT windowed attention (7x7 or 14x14 windows, head_dim 32) on tensor cores.
//
// Input is the QKV projection of the UN-partitioned token grid, (B*res*res, heads*96) with per-head [q|k|v]; the
// kernel does the window partition itself: a CTA gathers the tokens of one window (for a group of kHG heads) with
// cp.async, positions that fall outside the image (TinyViT zero-pads the grid to a multiple of the window BEFORE the
// in-attention LayerNorm and does NOT mask them) get the constant projection of LN(0) = beta, and the
// result is scattered straight back to token order.  So neither a partitioned copy of the activations nor the QKV
// rows of padding tokens ever exist in HBM.
//
// CTAs are persistent over (window, head group) items with a fixed head group, so the relative-position bias of
// those heads is staged once (7x7 windows: fp16, already in mma accumulator-fragment order and scaled by log2 e; 14x14: fp32
// fragments read through the read-only path as the initial accumulators of Q K^T) and the next item's tokens stream in
// (double buffer) while the current one is computed.  One warp per (head, 16-query tile): S = Q K^T with mma.sync
// m16n8k16 fed by ldmatrix, online softmax over chunks of 32 keys in registers (exp2; lazy running maxima, attend_chunk), P
// re-used in place as the A operand of O += P V (V through ldmatrix.trans, no transposed copy).  These windows are far
// below a 128-row tcgen05 tile and the attention core is 3-19 % of a stage's MACs, so the warp-level
// MMA is the right tool; the kernel is bound by instruction issue, which is what this structure minimises.
#include "encoder_kernels.cuh"

#include "../profiler.hpp"

#include <cmath>
#include <cuda_fp16.h>

namespace dlimg {
namespace enc {

namespace {

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
#if defined(DLIMG_B200_ACT_BF16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
#else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
#endif
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    act2_t const v = f22act2(a, b);
    return *reinterpret_cast<uint32_t const*>(&v);
}
__device__ __forceinline__ uint32_t smem_u32(void const* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int kWS, int kHG>
struct Cfg {
    static constexpr int n = kWS * kWS;          // tokens per window: 49 / 196
    static constexpr int NQ = (n + 15) / 16;     // 16-query tiles: 4 / 13
    static constexpr int NK8 = (n + 7) / 8;      // 8-key blocks: 7 / 25
    static constexpr int NKP = NQ * 16;          // token rows held in shared memory: 64 / 208
    static constexpr int kWarps = kHG * NQ;      // one warp per (head, query tile)
    static constexpr int kThreads = kWarps * 32;
    static constexpr int kRowBytes = kHG * 192 + 16;  // [q|k|v] of kHG heads + 16 B: rows land in distinct bank groups
    static constexpr int kTileBytes = NKP * kRowBytes;
    static constexpr int kBiasBytes = kHG * NQ * NK8 * 32 * 8;
    // 14x14 windows: the 83 KB bias table of a head would cap the SM at ONE 13-warp CTA, and the kernel is latency
    // bound; the table stays in global memory instead (read-only path, L1 / L2 resident) so that two CTAs fit.
    static constexpr bool kBiasInSmem = kBiasBytes <= 48 * 1024;
    static constexpr int kSmemBias = kBiasInSmem ? kBiasBytes : 0;
    static constexpr int kSmemBytes = kSmemBias + 2 * kTileBytes;
    static constexpr int kMinBlocks = !kBiasInSmem ? 2 : (kHG > 2 ? 1 : (kHG == 2 ? 2 : 4));
    static constexpr int kCPT = kHG * 12;        // 16-byte chunks per token
};

// One chunk of kCnt key blocks (8 keys each) of the online softmax for one (head, query tile).
//
// kMode 0: the running row maxima follow every chunk (the accumulators are rescaled each time).
// kMode 1: LAZY maxima.  The reference point m of a row only moves when some score of the chunk exceeds it by more than
//          kLazyGrowth (in log2 units); until then the probabilities are taken relative to the old m and may reach
//          2^kLazyGrowth, which a 16-bit float holds with the same relative precision, and the final division by the row
//          sum (accumulated from the same rounded values) is unchanged.  The test is per lane plus one warp vote, so the
//          common chunk drops the quad reductions, two exponentials and the 18 accumulator multiplications.
// Measured and dropped: (a) every fourth exponential as a degree-3 polynomial on the FMA pipe -- 1.44 ms of attention per step
// against 1.39, the kernel is short of issue slots and L1 data-pipe cycles, not of MUFU throughput; (b) ex2.approx.f16x2:
// ptxas lowers it to two MUFU.EX2.F16 plus a PRMT on sm_100a, no XU cycles saved; (c) fp32 accumulator-initialising bias
// fragments in shared memory for the 7x7 windows as well: 48 instructions fewer per item, but twice the shared-memory
// wavefronts for the table -- 392 us against 374 us per pass for the 128-wide stage; (d) a 16-bit table for the 14x14 windows
// (half the L2 -> L1 traffic of the bias, four conversions more per key block): 1.369 against 1.376 ms, not worth the precision.
constexpr float kLazyGrowth = 8.0f;
template <int kCnt, bool kFirst, bool kBiasInSmem, int kMode>
__device__ __forceinline__ void attend_chunk(int nb0, uint32_t const (&aq)[2][4], uint32_t k_addr, uint32_t v_addr, int row_bytes,
                                             uint32_t bias_addr, uint8_t const* __restrict__ bias_gl, float (&m)[2], float (&l)[4],
                                             float (&o)[4][4], uint32_t ones_b) {
    float const kScale = 0.17677669529663687f * 1.4426950408889634f;  // 32^-0.5 * log2(e)
    float s[kCnt][4];
    float cm0 = -INFINITY, cm1 = -INFINITY;
    if (kBiasInSmem) {
        // bias fragments (fp16, pre-multiplied by log2 e) in shared memory: s = q.k * scale * log2 e + bias
#pragma unroll
        for (int j = 0; j < kCnt; ++j) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4(k_addr + (uint32_t)((nb0 + j) * 8 * row_bytes), b0, b1, b2, b3);
            s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
            mma16816(s[j], aq[0][0], aq[0][1], aq[0][2], aq[0][3], b0, b1);
            mma16816(s[j], aq[1][0], aq[1][1], aq[1][2], aq[1][3], b2, b3);
        }
#pragma unroll
        for (int j = 0; j < kCnt; ++j) {
            uint32_t w0, w1;
            asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(w0), "=r"(w1) : "r"(bias_addr + (uint32_t)((nb0 + j) * 256)));
            float2 const f0 = __half22float2(*reinterpret_cast<__half2 const*>(&w0));
            float2 const f1 = __half22float2(*reinterpret_cast<__half2 const*>(&w1));
            s[j][0] = fmaf(s[j][0], kScale, f0.x);
            s[j][1] = fmaf(s[j][1], kScale, f0.y);
            s[j][2] = fmaf(s[j][2], kScale, f1.x);
            s[j][3] = fmaf(s[j][3], kScale, f1.y);
            cm0 = fmaxf(cm0, fmaxf(s[j][0], s[j][1]));
            cm1 = fmaxf(cm1, fmaxf(s[j][2], s[j][3]));
        }
    } else {
        // bias fragments (fp32, divided by the qk scale) from global memory are the initial accumulators of the MMA:
        // s = q.k + bias / scale, and the scale * log2 e moves into the exponent below -- eight instructions fewer per
        // key block than converting and adding an fp16 bias
#pragma unroll
        for (int j = 0; j < kCnt; ++j) {
            float4 const bw = __ldg(reinterpret_cast<float4 const*>(bias_gl) + (nb0 + j) * 32);
            uint32_t b0, b1, b2, b3;
            ldsm_x4(k_addr + (uint32_t)((nb0 + j) * 8 * row_bytes), b0, b1, b2, b3);
            s[j][0] = bw.x; s[j][1] = bw.y; s[j][2] = bw.z; s[j][3] = bw.w;
            mma16816(s[j], aq[0][0], aq[0][1], aq[0][2], aq[0][3], b0, b1);
            mma16816(s[j], aq[1][0], aq[1][1], aq[1][2], aq[1][3], b2, b3);
            cm0 = fmaxf(cm0, fmaxf(s[j][0], s[j][1]));
            cm1 = fmaxf(cm1, fmaxf(s[j][2], s[j][3]));
        }
    }
    bool moved = true;
    if (!kFirst && (kMode & 1)) {  // (scores of the global-bias path are still unscaled: the growth bound is divided by the scale)
        float const bound = kBiasInSmem ? kLazyGrowth : kLazyGrowth / kScale;
        moved = __any_sync(0xffffffffu, cm0 > m[0] + bound || cm1 > m[1] + bound);
    }
    if (moved) {
        cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 1));
        cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 2));
        cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 1));
        cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 2));
        if (kFirst) {
            m[0] = cm0;
            m[1] = cm1;
        } else {  // every chunk holds at least one unmasked key, so the running maxima are finite from chunk 0 on
            float const n0 = fmaxf(m[0], cm0), n1 = fmaxf(m[1], cm1);
            float const a0 = ex2((m[0] - n0) * (kBiasInSmem ? 1.0f : kScale)), a1 = ex2((m[1] - n1) * (kBiasInSmem ? 1.0f : kScale));
            m[0] = n0;
            m[1] = n1;
            l[0] *= a0;  // (accumulator fragment of the ones column: rows g / g + 8 in elements 0 / 2 of the lanes with t == 0)
            l[2] *= a1;
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                o[d][0] *= a0; o[d][1] *= a0;
                o[d][2] *= a1; o[d][3] *= a1;
            }
        }
    }
    float const e0 = -m[0] * kScale, e1 = -m[1] * kScale;  // (global-bias path: scores are still unscaled)
    uint32_t p[kCnt + 1][2];  // probabilities as packed pairs: [j][0] = row g, [j][1] = row g + 8 (keys 2t, 2t + 1 of block j)
    p[kCnt][0] = p[kCnt][1] = 0u;
#pragma unroll
    for (int j = 0; j < kCnt; ++j) {
        float x0, x1, x2, x3;
        if (kBiasInSmem) {
            x0 = s[j][0] - m[0]; x1 = s[j][1] - m[0];
            x2 = s[j][2] - m[1]; x3 = s[j][3] - m[1];
        } else {
            x0 = fmaf(s[j][0], kScale, e0); x1 = fmaf(s[j][1], kScale, e0);
            x2 = fmaf(s[j][2], kScale, e1); x3 = fmaf(s[j][3], kScale, e1);
        }
        p[j][0] = pack2(ex2(x0), ex2(x1));
        p[j][1] = pack2(ex2(x2), ex2(x3));
    }
    // O += P V, 16 keys per step; an odd trailing key block pairs with zeros
#pragma unroll
    for (int kb = 0; kb < (kCnt + 1) / 2; ++kb) {
        uint32_t const a0 = p[2 * kb][0], a1 = p[2 * kb][1], a2 = p[2 * kb + 1][0], a3 = p[2 * kb + 1][1];
        uint32_t const va = v_addr + (uint32_t)((nb0 + 2 * kb) * 8 * row_bytes);
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(va, b0, b1, b2, b3);  // dims 0-7 (keys 0-7, 8-15), dims 8-15
        mma16816(o[0], a0, a1, a2, a3, b0, b1);
        mma16816(o[1], a0, a1, a2, a3, b2, b3);
        ldsm_x4_t(va + 32u, b0, b1, b2, b3);  // dims 16-23, 24-31
        mma16816(o[2], a0, a1, a2, a3, b0, b1);
        mma16816(o[3], a0, a1, a2, a3, b2, b3);
        // row sums of P from the tensor core as well: a fifth B tile whose column 0 is all ones (the sums of the ROUNDED
        // probabilities, i.e. exactly what the other four products were normalised with; four FADDs per key block less)
        mma16816(l, a0, a1, a2, a3, ones_b, ones_b);
    }
}

template <int kWS, int kHG, int kMode>
__global__ void __launch_bounds__(Cfg<kWS, kHG>::kThreads, Cfg<kWS, kHG>::kMinBlocks)
window_attention_kernel2(act_t const* __restrict__ qkv, int batch, int res, int heads, act_t const* __restrict__ pad_qkv,
                         __half const* __restrict__ bias_frag, act_t* __restrict__ out) {
    using C = Cfg<kWS, kHG>;
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t const bias_s = smem_u32(smem);
    uint32_t const tile0_s = bias_s + C::kSmemBias;  // two token tiles (double buffer) follow the bias table
    int const tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int const n_groups = heads / kHG;
    int const hg = blockIdx.x % n_groups;
    int const nw = (res + kWS - 1) / kWS;
    int const total_windows = batch * nw * nw;
    int const w_first = blockIdx.x / n_groups, w_stride = gridDim.x / n_groups;
    int const ld = heads * 96;

    // bias fragments of this head group (resident for the whole kernel) + zero the tiles once: token rows >= n stay
    // zero for ever (K = 0 is masked by a -inf bias, V = 0 contributes nothing)
    {
        uint4 const* src = reinterpret_cast<uint4 const*>(bias_frag) + (size_t)hg * (C::kBiasBytes / 16);
        if (C::kBiasInSmem)
            for (int i = tid; i < C::kBiasBytes / 16; i += C::kThreads) reinterpret_cast<uint4*>(smem)[i] = __ldg(src + i);
        uint4* t = reinterpret_cast<uint4*>(smem + C::kSmemBias);
        for (int i = tid; i < 2 * C::kTileBytes / 16; i += C::kThreads) t[i] = make_uint4(0, 0, 0, 0);
    }
    pdl_wait();  // the bias table is a constant; qkv / out belong to the neighbouring kernels
    pdl_trigger();
    __syncthreads();

    // Token gather: a thread always moves the same 16-byte piece (`part`) of a token row and walks over tokens with a
    // fixed stride, so the per-copy work is one small constant division and an address add.
    constexpr int kLoadTok = C::kThreads / C::kCPT;  // tokens moved per sweep of the CTA
    int const ld_part = tid % C::kCPT, ld_tok0 = tid / C::kCPT;
    // Window coordinates are stepped from item to item (win -> win + w_stride is (b, wy, wx) plus a constant triple with two
    // carries) and windows that lie inside the image skip the per-token bounds tests: the four run-time divisions and the
    // tests were a quarter (14 x 14 windows) to a half (7 x 7) of the kernel's instructions (profiles/r02_summary.md section 6).
    struct WinPos {
        int b, wy, wx;
    };
    int const nw2 = nw * nw;
    WinPos const wstep{w_stride / nw2, (w_stride % nw2) / nw, w_stride % nw};
    auto advance = [&](WinPos& p) {
        p.wx += wstep.wx;
        if (p.wx >= nw) { p.wx -= nw; ++p.wy; }
        p.wy += wstep.wy;
        if (p.wy >= nw) { p.wy -= nw; ++p.b; }
        p.b += wstep.b;
    };
    auto load_item = [&](WinPos const& p, int buf) {
        int const y0 = p.wy * kWS, x0 = p.wx * kWS;
        act_t const* const src0 = qkv + ((size_t)(p.b * res + y0) * res + x0) * ld + hg * kHG * 96 + ld_part * 8;
        uint32_t const dst0 = tile0_s + (uint32_t)(buf * C::kTileBytes + ld_part * 16);
        bool const inside = y0 + kWS <= res && x0 + kWS <= res;  // block-uniform
        if (ld_tok0 < kLoadTok) {
#pragma unroll
            for (int tok = ld_tok0, i = 0; i < (C::n + kLoadTok - 1) / kLoadTok; ++i, tok += kLoadTok) {
                if (tok < C::n) {
                    int const iy = tok / kWS, ix = tok - iy * kWS;
                    uint32_t const dst = dst0 + (uint32_t)(tok * C::kRowBytes);
                    if (inside || (y0 + iy < res && x0 + ix < res)) {
                        act_t const* src = src0 + (iy * res + ix) * ld;  // (element offsets inside one window fit 32 bits)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                    } else {
                        uint4 const v = __ldg(reinterpret_cast<uint4 const*>(pad_qkv + hg * kHG * 96 + ld_part * 8));
                        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
                    }
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    constexpr int kStoreTok = C::kThreads / (kHG * 4);  // output: kHG * 4 pieces of 16 bytes per token
    int const st_part = tid % (kHG * 4), st_tok0 = tid / (kHG * 4);
    int const hh = warp / C::NQ, qt = warp % C::NQ;
    int const g = lane >> 2, t = lane & 3;
    // per-lane ldmatrix row addresses (relative to the tile)
    uint32_t const q_off = (uint32_t)((qt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * C::kRowBytes + (hh * 96 + ((lane >> 4) & 1) * 8) * 2);
    uint32_t const k_off = (uint32_t)((lane & 7) * C::kRowBytes + (hh * 96 + 32 + (lane >> 3) * 8) * 2);
    uint32_t const v_off = (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * C::kRowBytes + (hh * 96 + 64 + ((lane >> 4) & 1) * 8) * 2);
    uint32_t const bias_addr = bias_s + (uint32_t)((((hh * C::NQ + qt) * C::NK8) * 32 + lane) * 8);
    // (14x14, fp32 fragments: one float4 per lane and key block)
    uint8_t const* const bias_gl = reinterpret_cast<uint8_t const*>(bias_frag) +
                                   ((size_t)hg * (C::kBiasBytes / 8) + ((hh * C::NQ + qt) * C::NK8) * 32 + lane) * 16;

    int it = 0;
    WinPos cur{w_first / nw2, (w_first % nw2) / nw, w_first % nw}, nxt = cur;
    if (w_first < total_windows) load_item(cur, 0);
    for (int win = w_first; win < total_windows; win += w_stride, ++it, cur = nxt) {
        int const buf = it & 1;
        advance(nxt);
        if (win + w_stride < total_windows) {
            load_item(nxt, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();  // this item's tokens are visible to every warp

        uint32_t const tile = tile0_s + (uint32_t)(buf * C::kTileBytes);
        // query tiles that lie entirely below the image (bottom-edge windows) are padding whose output is never stored
        bool const live_tile = (qt * 16) / kWS < res - cur.wy * kWS;
        if (live_tile) {
            uint32_t aq[2][4];
            ldsm_x4(tile + q_off, aq[0][0], aq[0][1], aq[0][2], aq[0][3]);
            ldsm_x4(tile + q_off + 32u, aq[1][0], aq[1][1], aq[1][2], aq[1][3]);
            float m[2] = {-INFINITY, -INFINITY}, l[4] = {0.f, 0.f, 0.f, 0.f};
            uint32_t const ones_b = g == 0 ? (kActBf16 ? 0x3f803f80u : 0x3c003c00u) : 0u;  // B fragment (k = keys 2t, 2t + 1; n = g): column 0 = 1.0
            float o[4][4];
#pragma unroll
            for (int d = 0; d < 4; ++d) o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f;
            if constexpr (C::NK8 <= 8) {
                attend_chunk<C::NK8, true, C::kBiasInSmem, kMode>(0, aq, tile + k_off, tile + v_off, C::kRowBytes, bias_addr, bias_gl, m, l, o, ones_b);
            } else {
                static_assert(C::NK8 <= 8 || C::NK8 == 25, "chunk schedule written for 196-token windows");
                // 25 key blocks in chunks of 4 (+1): with the bias table in global memory the kernel runs two CTAs per SM
                // at 72 registers, and a 4-block chunk (16 score registers) is what fits without spilling
                attend_chunk<4, true, C::kBiasInSmem, kMode>(0, aq, tile + k_off, tile + v_off, C::kRowBytes, bias_addr, bias_gl, m, l, o, ones_b);
#pragma unroll
                for (int c0 = 4; c0 < 24; c0 += 4)
                    attend_chunk<4, false, C::kBiasInSmem, kMode>(c0, aq, tile + k_off, tile + v_off, C::kRowBytes, bias_addr, bias_gl, m, l, o, ones_b);
                attend_chunk<1, false, C::kBiasInSmem, kMode>(24, aq, tile + k_off, tile + v_off, C::kRowBytes, bias_addr, bias_gl, m, l, o, ones_b);
            }
            // the sums sit in the t == 0 lane of every quad
            float const inv0 = 1.0f / __shfl_sync(0xffffffffu, l[0], lane & ~3), inv1 = 1.0f / __shfl_sync(0xffffffffu, l[2], lane & ~3);
            // O (16-bit) goes into this task's own Q slot of the tile; the CTA then writes whole token rows
            uint32_t const o_addr = tile + (uint32_t)((qt * 16 + g) * C::kRowBytes + (hh * 96 + 2 * t) * 2);
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(o_addr + (uint32_t)(d * 16)), "r"(pack2(o[d][0] * inv0, o[d][1] * inv0)) : "memory");
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(o_addr + (uint32_t)(8 * C::kRowBytes + d * 16)), "r"(pack2(o[d][2] * inv1, o[d][3] * inv1)) : "memory");
            }
        }
        __syncthreads();  // all heads' outputs of this window are in the tile

        if (st_tok0 < kStoreTok) {
            int const y0 = cur.wy * kWS, x0 = cur.wx * kWS;
            int const Cout = heads * 32;
            bool const inside = y0 + kWS <= res && x0 + kWS <= res;
            act_t* const dst0 = out + ((size_t)(cur.b * res + y0) * res + x0) * Cout + hg * kHG * 32 + st_part * 8;
            uint32_t const src0 = tile + (uint32_t)(((st_part >> 2) * 96 + (st_part & 3) * 8) * 2);
#pragma unroll
            for (int tok = st_tok0, i = 0; i < (C::n + kStoreTok - 1) / kStoreTok; ++i, tok += kStoreTok) {
                if (tok < C::n) {
                    int const iy = tok / kWS, ix = tok - iy * kWS;
                    if (inside || (y0 + iy < res && x0 + ix < res)) {
                        uint4 v;
                        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(src0 + (uint32_t)(tok * C::kRowBytes)));
                        *reinterpret_cast<uint4*>(dst0 + (iy * res + ix) * Cout) = v;
                    }
                }
            }
        }
        __syncthreads();  // the tile may be overwritten by the load issued two items ahead
    }
}

template <int kWS, int kHG, int kMode>
void launch_attention_mode(cudaStream_t s, act_t const* qkv, int batch, int res, int heads, act_t const* pad_qkv,
                           __half const* bias_frag, act_t* out, int num_sms) {
    using C = Cfg<kWS, kHG>;
    static bool attr_set = false;
    if (!attr_set) {
        CUDA_CHECK(cudaFuncSetAttribute(window_attention_kernel2<kWS, kHG, kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
        attr_set = true;
    }
    int const n_groups = heads / kHG;
    int const nw = (res + kWS - 1) / kWS;
    int const items = batch * nw * nw * n_groups;
    int grid = (num_sms * C::kMinBlocks / n_groups) * n_groups;
    if (grid > items) grid = items;
    launch_pdl(PDL_ATTENTION, window_attention_kernel2<kWS, kHG, kMode>, dim3(grid), dim3(C::kThreads), (size_t)C::kSmemBytes, s, qkv, batch, res, heads,
               pad_qkv, bias_frag, out);
    KERNEL_CHECK();
}

constexpr int kSoftmaxMode = 1;  // attend_chunk's kMode of the release build

int softmax_mode() {
    static int const mode = dev_int("DLIMG_B200_WA_MODE", kSoftmaxMode);  // development A/B of the variants
    return mode;
}

template <int kWS, int kHG>
void launch_attention(cudaStream_t s, act_t const* qkv, int batch, int res, int heads, act_t const* pad_qkv,
                      __half const* bias_frag, act_t* out, int num_sms) {
#if DLIMG_B200_ALT
    int const mode = softmax_mode();
    if (mode == 0) return launch_attention_mode<kWS, kHG, 0>(s, qkv, batch, res, heads, pad_qkv, bias_frag, out, num_sms);
    if (mode == 1) return launch_attention_mode<kWS, kHG, 1>(s, qkv, batch, res, heads, pad_qkv, bias_frag, out, num_sms);
#endif
    launch_attention_mode<kWS, kHG, kSoftmaxMode>(s, qkv, batch, res, heads, pad_qkv, bias_frag, out, num_sms);
}

}  // namespace

// 7 x 7 windows: fp16 fragments pre-multiplied by log2 e (they live in shared memory).  14 x 14 windows: fp32 fragments
// divided by the qk scale (read from global memory straight into the MMA accumulators), two uint16 slots per value.
static bool bias_fragments_fp32(int ws) { return !Cfg<14, 1>::kBiasInSmem && ws == 14; }

size_t attention_bias_fragment_count(int heads, int ws) {
    int const n = ws * ws, nq = (n + 15) / 16, nk8 = (n + 7) / 8;
    return (size_t)heads * nq * nk8 * 32 * 4 * (bias_fragments_fp32(ws) ? 2 : 1);
}

void attention_bias_fragments(float const* dense, int heads, int ws, uint16_t* out) {
    int const n = ws * ws, nq = (n + 15) / 16, nk8 = (n + 7) / 8;
    float const log2e = 1.4426950408889634f;
    bool const f32 = bias_fragments_fp32(ws);
    size_t i = 0;
    for (int h = 0; h < heads; ++h)
        for (int qt = 0; qt < nq; ++qt)
            for (int nb = 0; nb < nk8; ++nb)
                for (int lane = 0; lane < 32; ++lane)
                    for (int e = 0; e < 4; ++e) {
                        int const r = qt * 16 + (lane >> 2) + (e >= 2 ? 8 : 0);
                        int const c = nb * 8 + 2 * (lane & 3) + (e & 1);
                        float v = 0.f;
                        if (c >= n) v = -INFINITY;  // key padding of the tile: masked
                        else if (r < n) v = dense[((size_t)h * n + r) * n + c];
                        if (f32) {
                            if (c < n) v *= 5.656854249492381f;  // / 32^-0.5
                            reinterpret_cast<float*>(out)[i++] = v;
                        } else {
                            if (c < n) v *= log2e;
                            __half const hv = __float2half_rn(v);
                            out[i++] = *reinterpret_cast<uint16_t const*>(&hv);
                        }
                    }
}

void window_attention(cudaStream_t s, act_t const* qkv, int batch, int res, int ws, int heads, act_t const* pad_qkv,
                      uint16_t const* bias_frag, act_t* out, int num_sms) {
    int const nw = (res + ws - 1) / ws, n = ws * ws;
    ProfScope prof(s, CAT_WIN_ATTN, 4.0 * batch * nw * nw * heads * n * n * 32, (double)batch * res * res * heads * 128 * 2);
    __half const* bf = reinterpret_cast<__half const*>(bias_frag);
    // 7x7 windows: ONE head per CTA, four CTAs of four warps per SM.  A CTA meets at two CTA-wide barriers per item (tokens
    // landed / outputs complete), and with all the heads of a window in one 16- or 20-warp CTA the SM idled through every
    // load and store phase: 1.378 ms of attention per step with 4 / 5 heads per CTA, 1.262 with two, 1.246 with one.
#if DLIMG_B200_ALT
    int const hg = dev_int("DLIMG_B200_WA_HG", 1);  // development A/B of the grouping
    if (ws == 7 && hg == 5 && heads % 5 == 0) return launch_attention<7, 5>(s, qkv, batch, res, heads, pad_qkv, bf, out, num_sms);
    if (ws == 7 && hg == 4 && heads % 4 == 0) return launch_attention<7, 4>(s, qkv, batch, res, heads, pad_qkv, bf, out, num_sms);
    if (ws == 7 && hg == 2 && heads % 2 == 0) return launch_attention<7, 2>(s, qkv, batch, res, heads, pad_qkv, bf, out, num_sms);
#endif
    if (ws == 7) launch_attention<7, 1>(s, qkv, batch, res, heads, pad_qkv, bf, out, num_sms);
    else if (ws == 14) launch_attention<14, 1>(s, qkv, batch, res, heads, pad_qkv, bf, out, num_sms);
    else fail("window_attention: unsupported (window, heads) = (" + std::to_string(ws) + ", " + std::to_string(heads) + ")");
}

}  // namespace enc
}  // namespace dlimg
