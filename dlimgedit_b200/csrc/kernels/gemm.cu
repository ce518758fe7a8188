// gemm.cu -- tcgen05 / TMEM / TMA GEMM for sm_100a (see gemm.cuh for the contract).
#include "gemm.cuh"
#include "gelu.cuh"
#include "mask_select.cuh"
#include "tcgen05.cuh"

#include "../profiler.hpp"

#include <cuda_runtime.h>

#include <cstdlib>

#include <map>
#include <mutex>
#include <tuple>

namespace dlimg {
namespace gemm {

namespace {

constexpr int kAStageBytes = kBlockM * kKBytes;     // 16 KiB
constexpr int kEpiWarps = 16;                       // four per TMEM lane quarter; slabs of 16 columns are dealt round-robin
constexpr int kNumThreads = 64 + kEpiWarps * 32;
constexpr int kTmemCols = 512;                      // two accumulator stages of up to 256 fp32 columns
// Output staging for the coalescing epilogue: 128 rows of up to 256 16-bit columns, rows padded by 16 bytes so that
// the 16-byte shared-memory stores of 8 consecutive rows fall into distinct bank groups.
constexpr int kStagePad = 16;


using namespace tc;

// gelu_erf: see gelu.cuh

struct EpiParams {
    float const* bias;
    void const* residual;
    int const* row_map;
    float2 const* ln_stats;
    float2* stats_out;
    int ln_parts;
    float ln_eps;
    int act;
    int out_f32;
    int ldc;
    int res_mod;   // staged residual: residual row = output row % res_mod (0 = output row)
    float const* fuse_a;  // kFuse 1: LayerNorm2d gamma (64); kFuse 2: hypernetwork weights (prompts, 4, 32)
    float const* fuse_b;  // kFuse 1: LayerNorm2d beta (64); kFuse 2, MASKS_BEST: predicted IoUs (prompts, 4)
    int fuse_mode;        // kFuse 2: MaskMode
    int ksplit;           // split-K: M holds ksplit * (rows of A rounded up to 128)
    float* fuse_out;      // kFuse 2: low-resolution mask logits (prompts, 4, 256, 256)
    void const* const* res_table;  // staged residual: base pointer of every group of res_mod output rows (kFuse 3, layer 0)
};

// d += a * b on both lanes of a packed fp32 pair (one FFMA2 on sm_100)
__device__ __forceinline__ void ffma2(float2& d, float2 a, float2 b) {
    uint64_t& dd = reinterpret_cast<uint64_t&>(d);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(reinterpret_cast<uint64_t const&>(a)), "l"(reinterpret_cast<uint64_t const&>(b)));
}

__device__ __forceinline__ void add_bias16(float (&v)[16], float const* bias, int col) {
    float4 const* b4 = reinterpret_cast<float4 const*>(bias + col);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float4 b = __ldg(b4 + i);
        v[4 * i + 0] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
    }
}

// folded LayerNorm with row-centred weights: v = rstd * acc + bias, see Epilogue::ln_stats
__device__ __forceinline__ void ln_bias16(float (&v)[16], float const* bias, int col, float rstd) {
    float4 const* b4 = reinterpret_cast<float4 const*>(bias + col);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float4 const b = __ldg(b4 + i);
        v[4 * i + 0] = fmaf(v[4 * i + 0], rstd, b.x);
        v[4 * i + 1] = fmaf(v[4 * i + 1], rstd, b.y);
        v[4 * i + 2] = fmaf(v[4 * i + 2], rstd, b.z);
        v[4 * i + 3] = fmaf(v[4 * i + 3], rstd, b.w);
    }
}

// The same two with the bias vector staged in shared memory (broadcast reads; the global loads above cost the staged
// epilogues ~10 % of their time in exposed L1 latency, profiles/r01f).
__device__ __forceinline__ void add_bias16_s(float (&v)[16], uint32_t bias_s, int col) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float4 b;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "r"(bias_s + (uint32_t)((col + 4 * i) * 4)));
        v[4 * i + 0] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
    }
}
__device__ __forceinline__ void ln_bias16_s(float (&v)[16], uint32_t bias_s, int col, float rstd) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float4 b;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "r"(bias_s + (uint32_t)((col + 4 * i) * 4)));
        v[4 * i + 0] = fmaf(v[4 * i + 0], rstd, b.x);
        v[4 * i + 1] = fmaf(v[4 * i + 1], rstd, b.y);
        v[4 * i + 2] = fmaf(v[4 * i + 2], rstd, b.z);
        v[4 * i + 3] = fmaf(v[4 * i + 3], rstd, b.w);
    }
}

// activation + conversion of 16 fp32 values to 16-bit storage (two 16-byte vectors)
__device__ __forceinline__ void activate_pack16(float (&v)[16], int act, uint4 (&x)[2]) {
    act2_t* h = reinterpret_cast<act2_t*>(x);
#if !defined(DLIMG_B200_ACT_BF16)
    if (act == ACT_GELU) {
        // fp16 storage: round first, then the packed-half erf GELU (two values per instruction, see gelu.cuh)
#pragma unroll
        for (int j = 0; j < 8; ++j) h[j] = gelu_erf_h2(f22act2(v[2 * j], v[2 * j + 1]));
        return;
    }
#endif
    if (act == ACT_GELU) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = gelu_erf(v[i]);
    } else if (act == ACT_RELU) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.0f);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) h[j] = f22act2(v[2 * j], v[2 * j + 1]);
}

// activation on 16 values already rounded to 16-bit storage
__device__ __forceinline__ void activate_packed16(uint4 (&x)[2], int act) {
    act2_t* h = reinterpret_cast<act2_t*>(x);
    if (act == ACT_NONE) return;
#if !defined(DLIMG_B200_ACT_BF16)
    if (act == ACT_GELU) {
#pragma unroll
        for (int j = 0; j < 8; ++j) h[j] = gelu_erf_h2(h[j]);
        return;
    }
#endif
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float2 f = act22f2(h[j]);
#if defined(DLIMG_B200_ACT_BF16)
        if (act == ACT_GELU) { f.x = gelu_erf(f.x); f.y = gelu_erf(f.y); } else
#endif
        { f.x = fmaxf(f.x, 0.0f); f.y = fmaxf(f.y, 0.0f); }
        h[j] = f22act2(f.x, f.y);
    }
}

// One 16-column slab of one accumulator row, written by its own thread: bias -> residual -> activation -> store.
// (Residual / row-scatter / fp32-output GEMMs; the wide 16-bit outputs take the staged path in the kernel.)
__device__ __forceinline__ void epilogue_store16(uint32_t const (&r)[16], EpiParams const& ep, void* out, int64_t orow,
                                                 int col, float& sum, float& sumsq) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
    if (ep.bias) add_bias16(v, ep.bias, col);
    if (ep.out_f32) {
        float* o = reinterpret_cast<float*>(out) + orow * ep.ldc + col;
        if (ep.residual) {
            float4 const* r4 = reinterpret_cast<float4 const*>(reinterpret_cast<float const*>(ep.residual) + orow * ep.ldc + col);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float4 x = r4[i];
                v[4 * i + 0] += x.x; v[4 * i + 1] += x.y; v[4 * i + 2] += x.z; v[4 * i + 3] += x.w;
            }
        }
        if (ep.act == ACT_GELU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = gelu_erf(v[i]);
        } else if (ep.act == ACT_RELU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.0f);
        }
        float4* o4 = reinterpret_cast<float4*>(o);
#pragma unroll
        for (int i = 0; i < 4; ++i) o4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
        act_t* o = reinterpret_cast<act_t*>(out) + orow * ep.ldc + col;
        if (ep.residual) {
            uint4 const* r4 = reinterpret_cast<uint4 const*>(reinterpret_cast<act_t const*>(ep.residual) + orow * ep.ldc + col);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                uint4 x = r4[i];
                act2_t const* h = reinterpret_cast<act2_t const*>(&x);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float2 f = act22f2(h[j]);
                    v[8 * i + 2 * j] += f.x;
                    v[8 * i + 2 * j + 1] += f.y;
                }
            }
        }
        if (ep.stats_out) {  // LayerNorm statistics of the row this GEMM writes (no activation in this mode)
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                sum += v[i];
                sumsq = fmaf(v[i], v[i], sumsq);
            }
        }
        uint4 x[2];
        activate_pack16(v, ep.act, x);
        uint4* o4 = reinterpret_cast<uint4*>(o);
        o4[0] = x[0];
        o4[1] = x[1];
    }
}

// Byte offset of 16-byte piece `chunk` of row `row` in a warp's staging area (32 rows x cnt * 32 bytes).  The area is
// written row-per-lane (8 consecutive rows, same piece, per quarter-warp) and read piece-per-lane (whole row segments
// across consecutive lanes); both must spread over the eight 16-byte bank groups.  128-byte rows do that with a 16-byte
// pad; 64- and 96-byte rows need an XOR swizzle instead (with padding one of the two phases conflicts 2-way).
__device__ __forceinline__ uint32_t stage_offset(int cnt, int row, int chunk) {
    if (cnt == 2) return (uint32_t)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
    if (cnt == 3) return (uint32_t)(row * 96 + ((chunk ^ ((row >> 2) & 1)) << 4));
    return (uint32_t)(row * (cnt * 32 + kStagePad) + (chunk << 4));
}

// The slabs of one warp for one tile, with a compile-time slab count so that the TMEM double buffer lives in fixed
// registers (no moves) and the loop has no run-time conditions.
struct SlabCtx {
    uint32_t taddr;       // TMEM address of the warp's first slab (lane quarter + accumulator stage + column)
    uint32_t tempty;      // accumulator-empty mbarrier of this stage
    int col0;             // global column of the first slab
    int64_t orow;         // output row of this lane (direct path), -1 = none
    float rstd;           // folded LayerNorm: 1 / sqrt(var + eps) of this lane's row
    int lane;
    uint32_t stage_base;  // shared-memory address of the warp's staging area (kStaged), laid out by stage_offset()
    act_t* out_seg;       // output address of (first row of the warp's 32, first column of its range) (kStaged)
    int64_t out_off;      // the same as an element offset (fp32 outputs)
    int64_t ldc;
    int rows_valid;       // how many of the warp's 32 rows exist (M tail)
    uint32_t bias_s;      // shared-memory copy of the bias vector (kStaged; 0 = read it from global memory)
    int64_t row0;         // global row of the warp's first row
    int rows_total;       // M
};

// kFuse (staged 16-bit GELU epilogues of the mask decoder's upscaling, one warp = one 64- / 32-column group of a row):
//   1  the warp's 64 columns are one LayerNorm2d group (output_upscaling.1 on the (dy, dx) block of the first transposed
//      convolution): statistics from a first pass over the accumulators, then normalise -> GELU -> store as usual.
//   3  (with a staged residual, block_n == N == 256) LayerNorm over the whole 256-wide row of  acc + bias + residual:
//      the pre-norm sums go to the staging area as usual and leave their row sums; the kernel exchanges those between
//      the four warps of the row and ln_copy_out() normalises on the way to global memory (two-way transformer:
//      keys <- LN(keys + out_proj(attention))).
//   2  the warp's 32 columns are the 32 channels of one output pixel (ey, ex) of the second transposed convolution: after
//      GELU they are dotted with the prompt's four hypernetwork vectors and ONLY the four mask logits are written
//      (mask_decoder: masks = hyper_in @ upscaled_embedding) -- the upscaled tensor never reaches memory.
template <int kCnt, bool kStaged, int kAct, bool kLn, bool kRes = false, bool kF32 = false, int kFuse = 0>
__device__ __forceinline__ void epilogue_slabs(SlabCtx const& cx, EpiParams const& ep, void* out, float& sum, float& sumsq) {
    uint32_t r[2][16];
    float ln_mean = 0.f, ln_rstd = 1.f;
    float2 dot2[4] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
    float const* fuse2_hy = nullptr;  // kFuse 2: hypernetwork vectors of this row's prompt, first requested mask
    int fuse2_first = 0, fuse2_count = 4;
    if (kFuse == 2) {
        int64_t const prompt = (cx.row0 + cx.lane) >> 14;  // 16384 blocked pixels per prompt: uniform over the warp
        if (ep.fuse_mode == MASKS_MULTI) {
            fuse2_first = 1;
            fuse2_count = 3;
        } else if (ep.fuse_mode == MASKS_BEST) {
            fuse2_first = best_mask_index(ep.fuse_b + prompt * 4);
            fuse2_count = 1;
        }
        fuse2_hy = ep.fuse_a + prompt * 128 + fuse2_first * 32;
    }
    if (kFuse == 1) {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < kCnt; ++k) {
            tmem_ld16(cx.taddr + (uint32_t)(k * 16), r[0]);
            tmem_ld_wait();
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[0][i]);
            add_bias16_s(v, cx.bias_s, cx.col0 + k * 16);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                s1 += v[i];
                s2 = fmaf(v[i], v[i], s2);
            }
        }
        ln_mean = s1 * (1.0f / (16 * kCnt));
        ln_rstd = rsqrtf(fmaxf(s2 * (1.0f / (16 * kCnt)) - ln_mean * ln_mean, 0.f) + 1e-6f);
    }
    tmem_ld16(cx.taddr, r[0]);
#pragma unroll
    for (int k = 0; k < kCnt; ++k) {
        tmem_ld_wait();  // slab k is in r[k & 1]
        if (k + 1 < kCnt) {
            tmem_ld16(cx.taddr + (uint32_t)((k + 1) * 16), r[(k + 1) & 1]);
        } else {
            tc_fence_before();
            __syncwarp();
            if (cx.lane == 0) mbar_arrive(cx.tempty);  // all of this warp's slabs are out of TMEM
        }
        int const c = cx.col0 + k * 16;
        if (kStaged && kF32) {
            // fp32 output (decoder): the slabs go through the staging area in pairs (128-byte rows, XOR-swizzled; a
            // trailing single slab as 64-byte rows), so that every store instruction writes whole 128-byte lines
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[k & 1][i]);
            add_bias16_s(v, cx.bias_s, c);
            if (ep.act == ACT_GELU) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = gelu_erf(v[i]);
            } else if (ep.act == ACT_RELU) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.0f);
            }
            constexpr int kPairs = (kCnt + 1) / 2;
            int const pair = k >> 1;
            bool const full = 2 * pair + 1 < kCnt;  // this pair holds two slabs (compile-time after unrolling)
            int const cnt = full ? 4 : 2;           // in units of 32-byte half-slabs: reuse the 16-bit layouts
            (void)kPairs;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int const chunk = (k & 1) * 4 + i;
                uint32_t const dst = cx.stage_base + (full ? (uint32_t)(cx.lane * 128 + ((chunk ^ (cx.lane & 7)) << 4))
                                                           : stage_offset(2, cx.lane, chunk));
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "f"(v[4 * i]), "f"(v[4 * i + 1]), "f"(v[4 * i + 2]), "f"(v[4 * i + 3]) : "memory");
            }
            if ((k & 1) == 1 || k == kCnt - 1) {  // pair complete: whole row segments -> global
                __syncwarp();
                int const cpr = 2 * cnt, rows_it = 32 / cpr;  // 16-byte pieces per row (8 / 4), rows per instruction
                int const row0 = cx.lane / cpr, chunk = cx.lane - row0 * cpr;
                float* pdst = reinterpret_cast<float*>(out) + cx.out_off + (int64_t)row0 * cx.ldc + pair * 32 + chunk * 4;
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    int const rr = row0 + it * rows_it;
                    if (it < cpr) {
                        uint32_t const src = cx.stage_base + (full ? (uint32_t)(rr * 128 + ((chunk ^ (rr & 7)) << 4)) : stage_offset(2, rr, chunk));
                        float4 x;
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(src));
                        if (rr < cx.rows_valid) *reinterpret_cast<float4*>(pdst) = x;
                        pdst += (int64_t)rows_it * cx.ldc;
                    }
                }
                __syncwarp();
            }
        } else if (kStaged) {
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[k & 1][i]);
            if (kLn) ln_bias16_s(v, cx.bias_s, c, cx.rstd);  // staged kernels always hold the bias vector (or zeros)
            else add_bias16_s(v, cx.bias_s, c);
            if (kFuse == 1) {
                float4 const* g4 = reinterpret_cast<float4 const*>(ep.fuse_a + (c & 63));
                float4 const* b4 = reinterpret_cast<float4 const*>(ep.fuse_b + (c & 63));
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float4 const g = __ldg(g4 + i), b = __ldg(b4 + i);
                    v[4 * i + 0] = (v[4 * i + 0] - ln_mean) * ln_rstd * g.x + b.x;
                    v[4 * i + 1] = (v[4 * i + 1] - ln_mean) * ln_rstd * g.y + b.y;
                    v[4 * i + 2] = (v[4 * i + 2] - ln_mean) * ln_rstd * g.z + b.z;
                    v[4 * i + 3] = (v[4 * i + 3] - ln_mean) * ln_rstd * g.w + b.w;
                }
            }
            if (kFuse == 2) {
                uint4 x[2];
                activate_pack16(v, kAct, x);  // the same 16-bit values the unfused path stored
                act2_t const* h = reinterpret_cast<act2_t const*>(x);
                float const* hy = fuse2_hy + (c & 31);
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    if (m >= fuse2_count) break;
                    float4 const* h4 = reinterpret_cast<float4 const*>(hy + m * 32);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float4 const w = __ldg(h4 + i);
                        // two fp32 FMAs per instruction: even / odd channels accumulate separately
                        ffma2(dot2[m], act22f2(h[2 * i]), make_float2(w.x, w.y));
                        ffma2(dot2[m], act22f2(h[2 * i + 1]), make_float2(w.z, w.w));
                    }
                }
                continue;
            }
            uint32_t const dst = cx.stage_base + stage_offset(kCnt, cx.lane, 2 * k);
            uint32_t const dst1 = cx.stage_base + stage_offset(kCnt, cx.lane, 2 * k + 1);
            if (kRes && (ep.residual || ep.res_table)) {
                // the residual piece of this tile is already in the staging area (cp.async, whole row segments); this
                // lane adds its row's 16 values in fp32 and puts the rounded sums back in the same place
                uint4 rs[2];
                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(rs[0].x), "=r"(rs[0].y), "=r"(rs[0].z), "=r"(rs[0].w) : "r"(dst));
                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(rs[1].x), "=r"(rs[1].y), "=r"(rs[1].z), "=r"(rs[1].w) : "r"(dst1));
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    act2_t const* h = reinterpret_cast<act2_t const*>(&rs[i]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float2 const f = act22f2(h[j]);
                        v[8 * i + 2 * j] += f.x;
                        v[8 * i + 2 * j + 1] += f.y;
                    }
                }
            }
            if (kRes && (kFuse == 3 || ep.stats_out)) {  // LayerNorm row sums of what this GEMM writes (fp32 values, fixed order)
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    sum += v[i];
                    sumsq = fmaf(v[i], v[i], sumsq);
                }
            }
            uint4 x[2];
            activate_pack16(v, kAct, x);
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(x[0].x), "r"(x[0].y), "r"(x[0].z), "r"(x[0].w) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst1), "r"(x[1].x), "r"(x[1].y), "r"(x[1].z), "r"(x[1].w) : "memory");
        } else if (cx.orow >= 0) {
            epilogue_store16(r[k & 1], ep, out, cx.orow, c, sum, sumsq);
        }
    }
    if (kFuse == 2) {
        // blocked pixel row ((y * 64 + x) * 4 + dy * 2 + dx) of the prompt, column group (ey, ex) -> logit pixel (Y, X)
        int64_t const row = cx.row0 + cx.lane;
        if (row < cx.rows_total) {
            int64_t const prompt = row >> 14;
            int const rr = (int)(row & 16383), pix = rr >> 2, dy = (rr >> 1) & 1, dx = rr & 1, g = (cx.col0 >> 5) & 3;
            int const Y = 4 * (pix >> 6) + 2 * dy + (g >> 1), X = 4 * (pix & 63) + 2 * dx + (g & 1);
            float* o = ep.fuse_out + (prompt * 4 + fuse2_first) * 65536 + Y * 256 + X;
#pragma unroll
            for (int m = 0; m < 4; ++m)
                if (m < fuse2_count) o[(int64_t)m * 65536] = dot2[m].x + dot2[m].y;
        }
        return;
    }
    if (kFuse == 3) {  // the rows are normalised on their way out, once the row sums of all four warps have met
        __syncwarp();
        return;
    }
    if (kStaged && !kF32) {
        // transpose through the warp's staging area: each instruction now covers 32 / cpr whole row segments
        constexpr int kCpr = 2 * kCnt, kRows = 32 / kCpr, kIters = (32 + kRows - 1) / kRows;
        __syncwarp();
        int const row0 = cx.lane / kCpr, chunk = cx.lane - row0 * kCpr;
        if (row0 < kRows) {
            act_t* p = cx.out_seg + (int64_t)row0 * cx.ldc + chunk * 8;
            constexpr int kBatch = 4;  // loads in flight per lane before their stores
#pragma unroll
            for (int it0 = 0; it0 < kIters; it0 += kBatch) {
                uint4 x[kBatch];
#pragma unroll
                for (int j = 0; j < kBatch; ++j) {
                    int const rr = row0 + (it0 + j) * kRows;
                    if (it0 + j < kIters && rr < 32) {
                        uint32_t const src = cx.stage_base + stage_offset(kCnt, rr, chunk);
                        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(x[j].x), "=r"(x[j].y), "=r"(x[j].z), "=r"(x[j].w) : "r"(src));
                    }
                }
#pragma unroll
                for (int j = 0; j < kBatch; ++j) {
                    int const rr = row0 + (it0 + j) * kRows;
                    if (it0 + j < kIters) {
                        if (rr < cx.rows_valid && rr < 32) *reinterpret_cast<uint4*>(p) = x[j];
                        p += (int64_t)kRows * cx.ldc;
                    }
                }
            }
        }
        __syncwarp();  // the staging area is private to this warp
    }
}

// kFuse 3: the warp's 32 x 128-byte piece of pre-norm rows (staging area) -> (x - mean) * rstd * gamma + beta -> global,
// whole row segments per store instruction.  mean / rstd: of this lane's row; the row a lane stores comes by shuffle.
__device__ __forceinline__ void ln_copy_out(SlabCtx const& cx, float mean, float rstd, uint32_t gamma_s, uint32_t beta_s) {
    constexpr int kCnt = 4, kCpr = 2 * kCnt, kRows = 32 / kCpr, kIters = 32 / kRows;  // 8 pieces per row, 4 rows per instruction
    int const row0 = cx.lane / kCpr, chunk = cx.lane - row0 * kCpr;
    int const col = cx.col0 + chunk * 8;
    float g[8], b[8];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(g[4 * i]), "=f"(g[4 * i + 1]), "=f"(g[4 * i + 2]), "=f"(g[4 * i + 3]) : "r"(gamma_s + (uint32_t)((col + 4 * i) * 4)));
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b[4 * i]), "=f"(b[4 * i + 1]), "=f"(b[4 * i + 2]), "=f"(b[4 * i + 3]) : "r"(beta_s + (uint32_t)((col + 4 * i) * 4)));
    }
    act_t* p = cx.out_seg + (int64_t)row0 * cx.ldc + chunk * 8;
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
        int const rr = row0 + it * kRows;
        float const mu = __shfl_sync(0xffffffffu, mean, rr), rs = __shfl_sync(0xffffffffu, rstd, rr);
        uint4 x;
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "r"(cx.stage_base + stage_offset(kCnt, rr, chunk)));
        act2_t* h = reinterpret_cast<act2_t*>(&x);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float2 const f = act22f2(h[j]);
            h[j] = f22act2((f.x - mu) * rs * g[2 * j] + b[2 * j], (f.y - mu) * rs * g[2 * j + 1] + b[2 * j + 1]);
        }
        if (rr < cx.rows_valid) *reinterpret_cast<uint4*>(p) = x;
        p += (int64_t)kRows * cx.ldc;
    }
    __syncwarp();  // the staging area is private to this warp
}

// Implicit-GEMM 3x3 convolution (stride 1, zero padding 1) over a 16-bit NHWC tensor: the A operand of k-block
// (tap, channel block) of a tile of 128 consecutive pixels (128 / W whole image rows) is one 4D TMA box of the input
// shifted by the tap, out-of-image elements zero-filled by TMA -- the im2col matrix never exists.  w == 0: plain GEMM.
struct ConvParams {
    int w = 0, h = 0;  // image width (divides 128) and height (multiple of 128 / w)
    int cblocks = 0;   // channels / 64
};

// Shared-memory plan (all sizes known on the host): [barriers | staging (kStaged) | stages x (A 16 KiB, B block_n*128 B)].
// The stage count is whatever fits (up to kMaxStages): narrow-N GEMMs get deep rings (9 stages at block_n 64), which
// they need because a stage then carries only 24 KiB towards the ~70 KiB per SM that HBM latency x bandwidth asks for.
constexpr int kMaxStages = 10;
constexpr int kBarrierBytes = 256;  // full[10], empty[10], tmem_full[2], tmem_empty[2], tmem slot
constexpr int kSmemLimit = 227 * 1024;

struct SmemPlan {
    int stages, staging_bytes, total_bytes;
};
inline SmemPlan plan_smem(int block_n, bool staged, bool stats = false, int bias_bytes = 0, bool tf32 = false) {
    SmemPlan p;
    // 16 epilogue warps x 32 rows x (64-column share of the tile + pad), see the epilogue
    p.staging_bytes = staged ? (int)round_up64((int64_t)kEpiWarps * 32 * ((((block_n >> 4) + 3) / 4) * 32 + kStagePad), 1024) : 0;
    if (staged && tf32) p.staging_bytes = kEpiWarps * 32 * 128;  // fp32 outputs: pairs of slabs as 128-byte rows
    if (stats) p.staging_bytes += 8192;  // two buffers of 16 warps x 32 lanes x (sum, sum of squares), behind the staging
    p.staging_bytes += bias_bytes;       // the bias vector, behind both
    int const stage_bytes = kAStageBytes + block_n * kKBytes;
    int const fixed = 1024 /*align*/ + 1024 /*barriers, keeps the stages 1024-aligned*/ + p.staging_bytes;
    p.stages = (kSmemLimit - fixed) / stage_bytes;
    if (p.stages > kMaxStages) p.stages = kMaxStages;
    p.total_bytes = fixed + p.stages * stage_bytes;
    return p;
}

// kStaged kernels are additionally specialised on the activation (kAct) and on the folded LayerNorm (kLn), so the
// slab loop carries no run-time branches; the direct kernels read both from EpiParams.
template <int kTF32, bool kStaged, int kAct = ACT_NONE, bool kLn = false, bool kRes = false, int kFuse = 0>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, int M, int N,
               int K, int block_n, int num_stages, int staging_bytes, int bias_bytes, void* out, EpiParams ep, ConvParams conv) {
    extern __shared__ uint8_t smem_raw[];
    uint32_t const smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint32_t const bar_base = smem_base;
    uint32_t const stage_out = smem_base + 1024u;  // epilogue output staging (16-bit rows), kStaged only
    uint32_t const ring_base = stage_out + (uint32_t)staging_bytes;
    uint32_t const b_stage_bytes = (uint32_t)(block_n * kKBytes);
    uint32_t const stage_bytes = kAStageBytes + b_stage_bytes;
    auto a_stage = [&](int s) { return ring_base + (uint32_t)s * stage_bytes; };
    auto b_stage = [&](int s) { return ring_base + (uint32_t)s * stage_bytes + kAStageBytes; };
    // barrier slots (8 bytes each): full[kMaxStages], empty[kMaxStages], tmem_full[2], tmem_empty[2], tmem ptr
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 2 + s); };
    uint32_t const tmem_slot = bar_base + 8u * (2 * kMaxStages + 4);
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    int const warp = threadIdx.x >> 5;
    int const lane = threadIdx.x & 31;

    int const elems_per_kb = kTF32 ? 32 : 64;
    int const elem_bytes = kTF32 ? 4 : 2;
    // split-K (EpiParams::ksplit > 1): M counts the rows of all partial outputs -- m-tile mt is rows (mt % m_tiles_a) * 128 of
    // A and k-blocks [(mt / m_tiles_a) * num_kb, +num_kb); only the producer knows, the rest of the kernel sees a tall problem
    int const num_kb = (K + elems_per_kb - 1) / elems_per_kb / ep.ksplit;
    int const n_tiles = N / block_n;
    int const m_tiles = (M + kBlockM - 1) / kBlockM;
    int const m_tiles_a = m_tiles / ep.ksplit;
    int const total_tiles = m_tiles * n_tiles;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_b) : "memory");
        for (int s = 0; s < num_stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), kEpiWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    uint32_t const bias_s = bias_bytes ? stage_out + (uint32_t)(staging_bytes - bias_bytes) : 0u;
    if (bias_bytes) {  // weights: independent of the previous kernel
        for (int i = threadIdx.x; i < N; i += kNumThreads)
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_s + 4u * (uint32_t)i), "f"(ep.bias ? __ldg(ep.bias + i) : 0.f) : "memory");
        if (kFuse == 3)  // LayerNorm gamma / beta behind the bias vector
            for (int i = threadIdx.x; i < 2 * N; i += kNumThreads)
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_s + 4u * (uint32_t)(N + i)), "f"(__ldg((i < N ? ep.fuse_a : ep.fuse_b - N) + i)) : "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();     // everything above is independent of the previous kernel in the stream
    pdl_trigger();  // the next kernel may run its own prologue as soon as SMs drain
    uint32_t const tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ---------------- TMA producer ----------------
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            uint32_t const tx_bytes = stage_bytes;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int const mt_all = tile / n_tiles, kb0 = (mt_all / m_tiles_a) * num_kb;
                int const m0 = (mt_all % m_tiles_a) * kBlockM;
                int const n0 = (tile % n_tiles) * block_n;
                int img = 0, y0 = 0;
                if (conv.w > 0) {
                    int const rows_pt = kBlockM / conv.w, tiles_per_img = conv.h / rows_pt, mt = tile / n_tiles;
                    img = mt / tiles_per_img;
                    y0 = (mt - img * tiles_per_img) * rows_pt;
                }
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    mbar_expect_tx(full_bar(stage), tx_bytes);
                    if (conv.w > 0) {
                        int const tap = kb / conv.cblocks, cb = kb - tap * conv.cblocks, ky = tap / 3, kx = tap - ky * 3;
                        tma_load_4d(a_stage(stage), &tma_a, full_bar(stage), cb * 64, kx - 1, y0 + ky - 1, img);
                    } else {
                        tma_load_2d(a_stage(stage), &tma_a, full_bar(stage), (kb0 + kb) * elems_per_kb, m0);
                    }
                    tma_load_2d(b_stage(stage), &tma_b, full_bar(stage), (kb0 + kb) * elems_per_kb, n0);
                    if (++stage == num_stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer: the whole warp walks the schedule, one elected lane issues ----------------
        // (under `lane == 0` the compiler wraps every tcgen05.mma in an elect / branch loop and moves its descriptors
        // from vector to uniform registers one by one; converged + elect.sync keeps it all on the uniform datapath)
        {
            // instruction descriptor: D fp32, A/B bf16 or tf32, both K-major, N = block_n, M = 128
            uint32_t const fmt = kTF32 ? 2u : (kActBf16 ? 1u : 0u);  // UMMA F16F32Format: F16 = 0, BF16 = 1, TF32 = 2
            uint32_t const idesc = make_idesc(fmt, kBlockM, block_n);
            uint64_t const adesc0 = make_smem_desc(a_stage(0)), bdesc0 = make_smem_desc(b_stage(0));
            uint64_t const stage_step = (uint64_t)(stage_bytes >> 4);  // descriptor address field: 16-byte units
            int const tail_bytes = ep.ksplit > 1 ? elems_per_kb * elem_bytes : (K - (num_kb - 1) * elems_per_kb) * elem_bytes;
            int const tail_instr = (tail_bytes + 31) >> 5;
            int stage = 0;
            uint32_t phase = 0;
            int local = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
                int const acc = local & 1;
                uint32_t const acc_phase = (uint32_t)(local >> 1) & 1u;
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                tc_fence_after();
                uint32_t const d_tmem = tmem_base + (uint32_t)(acc * kMaxBlockN);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    if (elect_one()) {
                        uint64_t const adesc = adesc0 + (uint64_t)stage * stage_step;
                        uint64_t const bdesc = bdesc0 + (uint64_t)stage * stage_step;
                        // advance 32 bytes of K inside the swizzle atom: +2 in the 16-byte address field
                        if (kb + 1 < num_kb) {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                tc_mma<kTF32>(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
                        } else {
                            for (int k = 0; k < tail_instr; ++k)
                                tc_mma<kTF32>(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
                            tc_commit(tfull_bar(acc));  // accumulator complete -> epilogue
                        }
                        tc_commit(empty_bar(stage));  // frees the smem slot once these MMAs retire
                    }
                    __syncwarp();
                    if (++stage == num_stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else {
        // ---------------- epilogue (16 warps; TMEM lane quarter = warp index mod 4) ----------------
        // The block_n / 16 slabs of 16 columns are dealt to the four warps of a lane quarter as CONTIGUOUS ranges (at
        // most 4 slabs = 64 columns each), so a warp owns a 32-row x (up to) 128-byte piece of the output tile.
        // The slabs are software-pipelined: the tcgen05.ld of slab k+1 is in flight while slab k is biased, activated
        // and written (TMEM reads are ~64 B/clk/SM, ~2000 clk for a 128x256 fp32 tile: they must overlap the
        // arithmetic), and the accumulator stage is handed back to the MMA warp as soon as the last slab has landed
        // in registers.  kStaged (plain 16-bit outputs): the warp transposes its piece through its own staging area
        // in shared memory (row per lane in, 16-byte pieces of whole row segments out), so every global store
        // instruction writes complete 64..128-byte segments and no CTA-level barrier is involved.
        int const quarter = warp & 3;
        int const slab = (warp - 2) >> 2;  // which of the kEpiWarps/4 warps of this lane quarter
        int const nslab = block_n >> 4;
        int const s_cnt = nslab / 4 + (slab < (nslab & 3) ? 1 : 0);           // slabs of this warp
        int const s_first = slab * (nslab / 4) + min(slab, nslab & 3);       // first slab
        uint32_t const pitch = kTF32 ? 128u : (uint32_t)(((nslab + 3) / 4) * 32 + kStagePad);  // bytes per staged row
        uint32_t const my_stage = stage_out + (uint32_t)(quarter * 4 + slab) * 32u * pitch;
        // kRes (staged kernels with a 16-bit residual): the warp's 32 x (s_cnt * 32 B) piece of the residual is copied
        // into its staging area with cp.async as whole row segments -- one tile ahead, right after the staging area has
        // been drained -- so neither the residual reads nor the output writes touch partial 128-byte lines.
        // Tile coordinates are stepped, not divided: tile -> tile + gridDim.x is (mt + g_mt, nt + g_nt) with one carry.  The
        // divisions by the run-time n_tiles (and lane / cpr below) were ~100 of the ~290 instructions a warp spends per tile
        // in the folded-LayerNorm GEMM (profiles/r02_summary.md section 6).
        int const g_mt = (int)gridDim.x / n_tiles, g_nt = (int)gridDim.x - g_mt * n_tiles;
        auto step = [&](int& mt, int& nt) {
            mt += g_mt;
            nt += g_nt;
            if (nt >= n_tiles) {
                nt -= n_tiles;
                ++mt;
            }
        };
        int const res_cpr = 2 * s_cnt, res_rows_it = s_cnt ? 32 / res_cpr : 0;  // 16-byte pieces per row, rows per instruction
        int const res_row0 = s_cnt ? lane / res_cpr : 0, res_chunk = lane - res_row0 * res_cpr;
        auto prefetch_residual = [&](int mt, int nt) {
            if (s_cnt == 0 || (!ep.residual && !ep.res_table)) return;  // (the kernel also serves plain GEMMs that only want row sums)
            int const m0 = mt * kBlockM + quarter * 32;
            int const n0 = nt * block_n + s_first * 16;
            // res_mod: the residual is a (res_mod, N) table shared by every group of res_mod output rows (a multiple of
            // the tile height, so a tile never straddles two groups) -- the decoder's position terms
            int const r0 = ep.res_mod ? m0 % ep.res_mod : m0;
            act_t const* seg = (ep.res_table ? reinterpret_cast<act_t const*>(ep.res_table[m0 / ep.res_mod])
                                             : reinterpret_cast<act_t const*>(ep.residual)) + (int64_t)r0 * ep.ldc + n0;
            int const rows_valid = min(32, M - m0);
            if (res_row0 < res_rows_it) {
                for (int rr = res_row0; rr < rows_valid; rr += res_rows_it) {
                    uint32_t const dst = my_stage + stage_offset(s_cnt, rr, res_chunk);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(seg + (int64_t)rr * ep.ldc + res_chunk * 8) : "memory");
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        int cur_mt = (int)blockIdx.x / n_tiles, cur_nt = (int)blockIdx.x - cur_mt * n_tiles;  // coordinates of `tile` in the loop below
        if (kRes && (int)blockIdx.x < total_tiles) prefetch_residual(cur_mt, cur_nt);
        // folded LayerNorm: (mean, rstd) of this lane's row (ln_parts == 0) or its (sum, sum of squares), the partial
        // sums of the producing kernel added in a fixed order
        // The statistics of the NEXT tile's rows are requested one tile ahead and only touched when that tile starts: the
        // loads return raw (up to kLnRaw partial pairs per row, summed in a fixed order at use).  Summing inside the
        // prefetch made it wait for its own loads -- a quarter of the stall samples of the qkv GEMMs sat on that add.
        constexpr int kLnRaw = 4;
        struct LnRaw {
            float2 v[kLnRaw];
        };
        auto load_ln = [&](int mt) {
            int const row = mt * kBlockM + quarter * 32 + lane;
            LnRaw r;
#pragma unroll
            for (int pp = 0; pp < kLnRaw; ++pp) r.v[pp] = make_float2(0.f, pp == 0 && ep.ln_parts == 0 ? 1.f : 0.f);
            if (row < M) {
                if (ep.ln_parts == 0) {
                    r.v[0] = __ldg(ep.ln_stats + row);
                } else if (ep.ln_parts <= kLnRaw) {
#pragma unroll
                    for (int pp = 0; pp < kLnRaw; ++pp)
                        if (pp < ep.ln_parts) r.v[pp] = __ldg(ep.ln_stats + (int64_t)row * ep.ln_parts + pp);
                } else {  // many partial sums: added here (this path waits for its loads)
                    float sx = 0.f, sq = 0.f;
                    for (int pp = 0; pp < ep.ln_parts; ++pp) {
                        float2 const pv = __ldg(ep.ln_stats + (int64_t)row * ep.ln_parts + pp);
                        sx += pv.x;
                        sq += pv.y;
                    }
                    r.v[0] = make_float2(sx, sq);
                }
            }
            return r;
        };
        auto ln_sum = [&](LnRaw const& r) {  // fixed order: ((p0 + p1) + p2) + p3, zeros for the unused slots
            float sx = r.v[0].x, sq = r.v[0].y;
#pragma unroll
            for (int pp = 1; pp < kLnRaw; ++pp) {
                sx += r.v[pp].x;
                sq += r.v[pp].y;
            }
            return make_float2(sx, sq);
        };
        LnRaw ln_raw;
#pragma unroll
        for (int pp = 0; pp < kLnRaw; ++pp) ln_raw.v[pp] = make_float2(0.f, pp == 0 ? 1.f : 0.f);
        if (kStaged && kLn && (int)blockIdx.x < total_tiles) ln_raw = load_ln(cur_mt);
        int local = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
            int const acc = local & 1;
            uint32_t const acc_phase = (uint32_t)(local >> 1) & 1u;
            int const m0 = cur_mt * kBlockM;
            int const n0 = cur_nt * block_n;
            int const this_nt = cur_nt;
            step(cur_mt, cur_nt);  // from here on: the coordinates of the next tile of this CTA
            bool const has_next = tile + (int)gridDim.x < total_tiles;
            int64_t orow = -1;
            float rstd = 1.f;  // folded LayerNorm: 1/std of this thread's row
            if (!kStaged) {
                int const row = m0 + quarter * 32 + lane;
                if (row < M) orow = ep.row_map ? (int64_t)__ldg(ep.row_map + row) : (int64_t)row;
            } else if (kLn) {
                // the row statistics were fetched one tile ahead (their latency used to sit in front of every tile)
                float2 const ln_cur = ln_sum(ln_raw);
                if (ep.ln_parts == 0) {
                    rstd = ln_cur.y;
                } else {
                    float const inv_k = 1.0f / (float)K, mean = ln_cur.x * inv_k;
                    rstd = rsqrtf(fmaxf(fmaf(-mean, mean, ln_cur.y * inv_k), 0.f) + ep.ln_eps);
                }
                if (has_next) ln_raw = load_ln(cur_mt);
            }
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            if (kRes) {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncwarp();
            }
            SlabCtx cx;
            cx.taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * kMaxBlockN + s_first * 16);
            cx.tempty = tempty_bar(acc);
            cx.col0 = n0 + s_first * 16;
            cx.orow = orow;
            cx.rstd = rstd;
            cx.lane = lane;
            cx.stage_base = my_stage;
            cx.out_off = (int64_t)(m0 + quarter * 32) * ep.ldc + n0 + s_first * 16;
            cx.out_seg = reinterpret_cast<act_t*>(out) + cx.out_off;
            cx.ldc = ep.ldc;
            cx.rows_valid = M - (m0 + quarter * 32);
            cx.bias_s = bias_s;
            cx.row0 = m0 + quarter * 32;
            cx.rows_total = M;
            float row_sum = 0.f, row_sumsq = 0.f;
            if (kFuse == 1) {  // block_n == 256: every warp owns the 64 columns of one LayerNorm2d group
                epilogue_slabs<4, kStaged, kAct, kLn, kRes, kTF32 != 0, 1>(cx, ep, out, row_sum, row_sumsq);
                continue;
            }
            if (kFuse == 2) {  // block_n == 128: every warp owns the 32 channels of one output pixel
                epilogue_slabs<2, kStaged, kAct, kLn, kRes, kTF32 != 0, 2>(cx, ep, out, row_sum, row_sumsq);
                continue;
            }
            if (kFuse == 3) {  // block_n == N == 256: LayerNorm of the whole row, statistics exchanged between its four warps
                epilogue_slabs<4, kStaged, kAct, kLn, kRes, kTF32 != 0, 3>(cx, ep, out, row_sum, row_sumsq);
                uint32_t const red = stage_out + (uint32_t)(staging_bytes - bias_bytes - 8192 + (local & 1) * 4096);
                uint32_t const mine = red + (uint32_t)(((quarter * 4 + slab) * 32 + lane) * 8);
                asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(mine), "f"(row_sum), "f"(row_sumsq) : "memory");
                asm volatile("bar.sync %0, 128;" ::"r"(1 + quarter) : "memory");
                float sx = 0.f, sq = 0.f;
#pragma unroll
                for (int w4 = 0; w4 < 4; ++w4) {
                    float a, b;
                    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(a), "=f"(b) : "r"(red + (uint32_t)(((quarter * 4 + w4) * 32 + lane) * 8)));
                    sx += a;
                    sq += b;
                }
                float const mean = sx * (1.0f / 256.0f);
                float const rstd = rsqrtf(fmaxf(sq * (1.0f / 256.0f) - mean * mean, 0.f) + ep.ln_eps);
                ln_copy_out(cx, mean, rstd, bias_s + 4u * (uint32_t)N, bias_s + 8u * (uint32_t)N);
                if (has_next) prefetch_residual(cur_mt, cur_nt);
                continue;
            }
            switch (s_cnt) {  // warp-uniform
                case 4: epilogue_slabs<4, kStaged, kAct, kLn, kRes, kTF32 != 0>(cx, ep, out, row_sum, row_sumsq); break;
                case 3: epilogue_slabs<3, kStaged, kAct, kLn, kRes, kTF32 != 0>(cx, ep, out, row_sum, row_sumsq); break;
                case 2: epilogue_slabs<2, kStaged, kAct, kLn, kRes, kTF32 != 0>(cx, ep, out, row_sum, row_sumsq); break;
                case 1: epilogue_slabs<1, kStaged, kAct, kLn, kRes, kTF32 != 0>(cx, ep, out, row_sum, row_sumsq); break;
                default:  // narrow tiles (block_n < 64): this warp owns no slab
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty_bar(acc));
                    break;
            }
            if (kRes && has_next) prefetch_residual(cur_mt, cur_nt);
            if ((!kStaged || kRes) && ep.stats_out) {
                // row statistics of this tile: the four warps of a lane quarter hold pieces of the same 32 rows
                uint32_t const red = stage_out + (uint32_t)(staging_bytes - bias_bytes - 8192 + (local & 1) * 4096);
                uint32_t const mine = red + (uint32_t)(((quarter * 4 + slab) * 32 + lane) * 8);
                asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(mine), "f"(row_sum), "f"(row_sumsq) : "memory");
                asm volatile("bar.sync %0, 128;" ::"r"(1 + quarter) : "memory");
                int const row = m0 + quarter * 32 + lane;
                if (slab == 0 && row < M) {
                    float sx = 0.f, sq = 0.f;
#pragma unroll
                    for (int w4 = 0; w4 < 4; ++w4) {
                        float a, b;
                        asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(a), "=f"(b) : "r"(red + (uint32_t)(((quarter * 4 + w4) * 32 + lane) * 8)));
                        sx += a;
                        sq += b;
                    }
                    ep.stats_out[(int64_t)row * n_tiles + this_nt] = make_float2(sx, sq);
                }
                // the other buffer is used by the next tile; a warp can only reach the tile after that once every warp
                // of its quarter (including the reader above) has passed the next tile's barrier
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Fused TinyViT MLP: out = x + fc2(GELU(fc1(LN(x)))) for a 128-row tile, with the 4C-wide hidden activation living only
// in TMEM and shared memory.  LN is folded into fc1 (row-centred weights, 1/std from row sums), fc2's epilogue adds the
// residual and leaves the LayerNorm row sums of the result for the next block's qkv.
//
// The work of one CTA is a stream of chunks q = (tile, h), h = 64 hidden units, NH = 4C / 64 chunks per tile:
//     MMA1(q)   D1[q & 3] (128 x 64 fp32, TMEM) = A(tile) * W1[h*64 .. +64, :]^T       W1 chunk: TMA ring of 2 or 4
//     EPI1(q)   16 warps, one 16-column slab each: TMEM -> rstd * acc + bias -> GELU -> fp16 -> H[q & 1] in shared
//               memory in the 128B-swizzled K-major layout (the A operand of the second GEMM)
//     MMA2(q)   D2 (128 x C fp32, TMEM) += H[q & 1] * W2[:, h*64 .. +64]^T              W2 chunk: TMA ring of 2
//   per tile:   A (128 x C) by TMA, double-buffered;  EPI2: D2 -> + bias + residual -> fp16 -> global, row sums
//
// MMA1 runs kMlpLead chunks ahead of MMA2 (across tile boundaries), so an accumulator of the first GEMM is always
// waiting when the epilogue warps finish a chunk: the round trip "H ready -> MMA2 -> MMA1 -> D1 ready" is off the
// critical path, which is the epilogue warps' instruction issue.  C <= 160 (one N tile for fc2; TMEM: 160 + 4 * 64).
struct MlpParams {
    int M, C;                    // rows, model width (128 or 160)
    float const* b1;             // [4C] folded fc1 bias
    float const* b2;             // [C]
    float2 const* ln_stats;      // [M] partial (sum, sum of squares) of the rows of x (ln_parts == 1)
    float ln_eps;
    act_t* out;                  // [M][C]
    float2* stats_out;           // [M] (sum, sum of squares) of the output rows, or null
};

constexpr int kMlpChunk = 64;
constexpr int kMlpLead = 3;      // MMA1 chunks in flight ahead of MMA2 (D1 ring of kMlpLead + 1 = 4)

struct MlpSmem {
    int kb1, a_bytes, w1_bytes, w2_bytes, w1_ring_log, total;
    // layout: [barriers 1 KiB][row-sum exchange 4 KiB][b1, b2 4 KiB][A x2][W1 x ring][W2 x2][H x2]
};
inline MlpSmem plan_mlp(int C) {
    MlpSmem m;
    m.kb1 = (C + 63) / 64;
    m.a_bytes = m.kb1 * kAStageBytes;
    m.w1_bytes = m.kb1 * kMlpChunk * kKBytes;
    m.w2_bytes = C * kKBytes;
    int const fixed = 1024 + 1024 + 8192 + 2 * (m.a_bytes + m.w2_bytes + kAStageBytes);
    m.w1_ring_log = fixed + 4 * m.w1_bytes <= kSmemLimit ? 2 : 1;
    m.total = fixed + (m.w1_bytes << m.w1_ring_log);
    return m;
}

// fc2 accumulator slabs of one warp: + bias + residual -> 16-bit, row sums.  The residual is the tile's own A operand,
// still in shared memory (128B-swizzled K-major: row r, 16-byte piece j of k-block kb at kb * 16 KiB + r * 128 +
// ((j ^ (r & 7)) << 4)); each lane reads its row's pieces, adds in fp32 and puts the rounded result back in place,
// from where the warp then copies whole row segments to global memory.
template <int kCnt>
__device__ __forceinline__ void mlp_epilogue2(uint32_t taddr, uint32_t d2_empty, int lane, uint32_t bias_s, int col0,
                                              uint32_t a_buf, int trow, bool stats, float& sum, float& sumsq) {
    uint32_t r[2][16];
    tmem_ld16(taddr, r[0]);
#pragma unroll
    for (int k = 0; k < kCnt; ++k) {
        tmem_ld_wait();
        if (k + 1 < kCnt) {
            tmem_ld16(taddr + (uint32_t)((k + 1) * 16), r[(k + 1) & 1]);
        } else {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(d2_empty);
        }
        int const col = col0 + k * 16;
        uint32_t const rowaddr = a_buf + (uint32_t)((col >> 6) * kAStageBytes + trow * 128);
        uint32_t const j0 = (uint32_t)((col & 63) >> 3), sw = (uint32_t)(trow & 7);
        uint32_t const p0 = rowaddr + ((j0 ^ sw) << 4), p1 = rowaddr + (((j0 + 1) ^ sw) << 4);
        uint4 rs[2];
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(rs[0].x), "=r"(rs[0].y), "=r"(rs[0].z), "=r"(rs[0].w) : "r"(p0));
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(rs[1].x), "=r"(rs[1].y), "=r"(rs[1].z), "=r"(rs[1].w) : "r"(p1));
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[k & 1][i]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float4 b;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                         : "r"(bias_s + (uint32_t)((col + 4 * i) * 4)));
            v[4 * i + 0] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            act2_t const* h = reinterpret_cast<act2_t const*>(&rs[i]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float2 const f = act22f2(h[j]);
                v[8 * i + 2 * j] += f.x;
                v[8 * i + 2 * j + 1] += f.y;
            }
        }
        if (stats) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                sum += v[i];
                sumsq = fmaf(v[i], v[i], sumsq);
            }
        }
        uint4 x[2];
        activate_pack16(v, ACT_NONE, x);
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(p0), "r"(x[0].x), "r"(x[0].y), "r"(x[0].z), "r"(x[0].w) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(p1), "r"(x[1].x), "r"(x[1].y), "r"(x[1].z), "r"(x[1].w) : "memory");
    }
}

__global__ void __launch_bounds__(kNumThreads, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tma_x, const __grid_constant__ CUtensorMap tma_w1,
                 const __grid_constant__ CUtensorMap tma_w2, MlpParams p, int kb1, int a_bytes, int w1_bytes, int w2_bytes,
                 int w1_ring_log) {
    extern __shared__ uint8_t smem_raw[];
    uint32_t const smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint32_t const bar_base = smem_base;
    uint32_t const red_base = smem_base + 1024u;
    uint32_t const bias_base = red_base + 4096u;  // b1 (4C floats) then b2 (C floats)
    uint32_t const a_base = red_base + 8192u;
    uint32_t const w1_base = a_base + 2u * (uint32_t)a_bytes;
    uint32_t const w2_base = w1_base + ((uint32_t)w1_bytes << w1_ring_log);
    uint32_t const h_base = w2_base + 2u * (uint32_t)w2_bytes;
    // barrier slots (8 bytes each)
    auto a_full = [&](int s) { return bar_base + 8u * (0 + s); };
    auto a_empty = [&](int s) { return bar_base + 8u * (2 + s); };
    auto w2_full = [&](int s) { return bar_base + 8u * (4 + s); };
    auto w2_empty = [&](int s) { return bar_base + 8u * (6 + s); };
    auto h_full = [&](int s) { return bar_base + 8u * (8 + s); };
    auto h_empty = [&](int s) { return bar_base + 8u * (10 + s); };
    auto w1_full = [&](int s) { return bar_base + 8u * (12 + s); };
    auto w1_empty = [&](int s) { return bar_base + 8u * (16 + s); };
    auto d1_full = [&](int s) { return bar_base + 8u * (20 + s); };
    auto d1_empty = [&](int s) { return bar_base + 8u * (24 + s); };
    uint32_t const d2_full = bar_base + 8u * 28, d2_empty = bar_base + 8u * 29, tmem_slot = bar_base + 8u * 30;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int const C = p.C, NH = (4 * C) / kMlpChunk;
    int const m_tiles = (p.M + kBlockM - 1) / kBlockM;
    int const my_tiles = (int)blockIdx.x < m_tiles ? (m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    uint32_t const total = (uint32_t)(my_tiles * NH);   // chunks this CTA processes
    uint32_t const w1_mask = (1u << w1_ring_log) - 1u;
    int const k_tail_bytes = (C - (kb1 - 1) * 64) * 2;  // valid bytes of the last k-block of the first GEMM

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_w1) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_w2) : "memory");
        for (int s = 0; s < 2; ++s) {
            mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1 + kEpiWarps);
            mbar_init(w2_full(s), 1); mbar_init(w2_empty(s), 1);
            mbar_init(h_full(s), kEpiWarps); mbar_init(h_empty(s), 1);
        }
        for (int s = 0; s < 4; ++s) {
            mbar_init(w1_full(s), 1); mbar_init(w1_empty(s), 1);
            mbar_init(d1_full(s), 1); mbar_init(d1_empty(s), kEpiWarps);
        }
        mbar_init(d2_full, 1);
        mbar_init(d2_empty, kEpiWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 5 * C; i += kNumThreads) {  // both bias vectors: read by every warp for every chunk
        float const b = i < 4 * C ? __ldg(p.b1 + i) : __ldg(p.b2 + (i - 4 * C));
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_base + 4u * (uint32_t)i), "f"(b) : "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();     // barriers, TMEM and the bias vectors above do not depend on the previous kernel in the stream
    pdl_trigger();
    uint32_t const tmem_base = *tmem_slot_ptr;
    uint32_t const tmem_d2 = tmem_base;           // columns [0, C)
    uint32_t const tmem_d1 = tmem_base + 256u;    // four buffers of 64 columns

    if (warp == 0) {
        // ---------------- TMA producer: A and W1 run kMlpLead chunks ahead of W2, like the MMAs that consume them ----
        if (lane == 0 && total > 0) {
            uint32_t q1 = 0;   // next first-GEMM chunk to load
            int h1 = 0, lt1 = 0;
            auto load_first = [&]() {
                if (h1 == 0) {
                    int const as = lt1 & 1;
                    mbar_wait(a_empty(as), ((uint32_t)(lt1 >> 1) & 1u) ^ 1u);
                    mbar_expect_tx(a_full(as), (uint32_t)a_bytes);
                    int const tile = (int)blockIdx.x + lt1 * (int)gridDim.x;
                    for (int kb = 0; kb < kb1; ++kb)
                        tma_load_2d(a_base + as * a_bytes + kb * kAStageBytes, &tma_x, a_full(as), kb * 64, tile * kBlockM);
                }
                int const ws = (int)(q1 & w1_mask);
                mbar_wait(w1_empty(ws), ((q1 >> w1_ring_log) & 1u) ^ 1u);
                mbar_expect_tx(w1_full(ws), (uint32_t)w1_bytes);
                for (int kb = 0; kb < kb1; ++kb)
                    tma_load_2d(w1_base + ws * w1_bytes + kb * (kMlpChunk * kKBytes), &tma_w1, w1_full(ws), kb * 64, h1 * kMlpChunk);
                ++q1;
                if (++h1 == NH) { h1 = 0; ++lt1; }
            };
            for (int j = 0; j < kMlpLead && q1 < total; ++j) load_first();
            int h2 = 0;
            for (uint32_t q = 0; q < total; ++q) {
                int const ws = (int)(q & 1u);
                mbar_wait(w2_empty(ws), ((q >> 1) & 1u) ^ 1u);
                mbar_expect_tx(w2_full(ws), (uint32_t)w2_bytes);
                tma_load_2d(w2_base + ws * w2_bytes, &tma_w2, w2_full(ws), h2 * kMlpChunk, 0);
                if (++h2 == NH) h2 = 0;
                if (q1 < total) load_first();
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer: the whole warp walks the schedule, one elected lane issues ----------------
        if (total > 0) {
            uint32_t const fmt = kActBf16 ? 1u : 0u;
            uint32_t const idesc1 = make_idesc(fmt, kBlockM, kMlpChunk);
            uint32_t const idesc2 = make_idesc(fmt, kBlockM, C);
            uint64_t const a_desc0 = make_smem_desc(a_base), w1_desc0 = make_smem_desc(w1_base);
            uint64_t const h_desc0 = make_smem_desc(h_base), w2_desc0 = make_smem_desc(w2_base);
            // descriptor address fields are in 16-byte units
            uint64_t const a_step = (uint64_t)(a_bytes >> 4), w1_step = (uint64_t)(w1_bytes >> 4), w2_step = (uint64_t)(w2_bytes >> 4);
            int const tail_instr = (k_tail_bytes + 31) >> 5;
            uint32_t q1 = 0;
            int h1 = 0, lt1 = 0;
            auto issue_first = [&]() {
                int const as = lt1 & 1;
                if (h1 == 0) mbar_wait(a_full(as), (uint32_t)(lt1 >> 1) & 1u);
                int const s1 = (int)(q1 & w1_mask), d = (int)(q1 & 3u);
                mbar_wait(w1_full(s1), (q1 >> w1_ring_log) & 1u);
                mbar_wait(d1_empty(d), ((q1 >> 2) & 1u) ^ 1u);
                tc_fence_after();
                if (elect_one()) {
                    uint64_t adesc = a_desc0 + (uint64_t)as * a_step;
                    uint64_t bdesc = w1_desc0 + (uint64_t)s1 * w1_step;
                    uint32_t const dst = tmem_d1 + (uint32_t)(d * kMlpChunk);
                    for (int kb = 0; kb + 1 < kb1; ++kb) {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            tc_mma<0>(dst, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc1, (uint32_t)((kb | k) != 0));
                        adesc += (uint64_t)(kAStageBytes >> 4);
                        bdesc += (uint64_t)((kMlpChunk * kKBytes) >> 4);
                    }
                    for (int k = 0; k < tail_instr; ++k) tc_mma<0>(dst, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc1, 1u);
                    tc_commit(w1_empty(s1));
                    tc_commit(d1_full(d));
                    if (h1 == NH - 1) tc_commit(a_empty(as));
                }
                __syncwarp();
                ++q1;
                if (++h1 == NH) { h1 = 0; ++lt1; }
            };
            for (int j = 0; j < kMlpLead && q1 < total; ++j) issue_first();
            int h2 = 0, lt2 = 0;
            for (uint32_t q = 0; q < total; ++q) {
                int const s2 = (int)(q & 1u);
                uint32_t const par = (q >> 1) & 1u;
                mbar_wait(h_full(s2), par);
                mbar_wait(w2_full(s2), par);
                if (h2 == 0) mbar_wait(d2_empty, ((uint32_t)lt2 & 1u) ^ 1u);
                tc_fence_after();
                if (elect_one()) {
                    uint64_t const adesc = h_desc0 + (uint64_t)s2 * (uint64_t)(kAStageBytes >> 4);
                    uint64_t const bdesc = w2_desc0 + (uint64_t)s2 * w2_step;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc_mma<0>(tmem_d2, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc2, (uint32_t)((h2 | k) != 0));
                    tc_commit(w2_empty(s2));
                    tc_commit(h_empty(s2));
                    if (h2 == NH - 1) tc_commit(d2_full);
                }
                __syncwarp();
                if (++h2 == NH) { h2 = 0; ++lt2; }
                if (q1 < total) issue_first();
            }
        }
    } else {
        // ---------------- epilogue warps ----------------
        int const quarter = warp & 3, slab = (warp - 2) >> 2;
        int const nslab = C >> 4;
        int const s_cnt = nslab / 4 + (slab < (nslab & 3) ? 1 : 0);
        int const s_first = slab * (nslab / 4) + min(slab, nslab & 3);
        int lt = 0;
        uint32_t u = 0;  // chunk counter across tiles
        float2 pv_next = make_float2(0.f, 1.f);  // this lane's row sums, fetched one tile ahead
        if ((int)blockIdx.x * kBlockM + quarter * 32 + lane < p.M) pv_next = __ldg(p.ln_stats + (int)blockIdx.x * kBlockM + quarter * 32 + lane);
        // second GEMM's accumulator of tile number `t` of this CTA: + bias + residual -> 16-bit -> global, row sums.
        // The A buffer of the tile doubles as the residual source and the staging area of the output, so it goes back
        // to the TMA producer from here (a_empty: the MMA warp's commit + one arrival per epilogue warp).
        auto finish_tile = [&](int row, int t) {
            bool const valid = row < p.M;
            uint32_t const a_buf = a_base + (uint32_t)((t & 1) * a_bytes);
            int const trow = quarter * 32 + lane;
            mbar_wait(d2_full, (uint32_t)t & 1u);
            tc_fence_after();
            uint32_t const taddr = tmem_d2 + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(s_first * 16);
            float row_sum = 0.f, row_sumsq = 0.f;
            if (s_cnt == 3) mlp_epilogue2<3>(taddr, d2_empty, lane, bias_base + 16u * (uint32_t)C, s_first * 16, a_buf, trow, p.stats_out != nullptr, row_sum, row_sumsq);
            else mlp_epilogue2<2>(taddr, d2_empty, lane, bias_base + 16u * (uint32_t)C, s_first * 16, a_buf, trow, p.stats_out != nullptr, row_sum, row_sumsq);
            __syncwarp();
            {   // whole row segments of the warp's 32 x (s_cnt * 32 B) piece: shared memory -> global
                int const cpr = 2 * s_cnt, rows_it = 32 / cpr;
                int const row0 = lane / cpr, chunk = lane - row0 * cpr;
                int const gc = s_first * 2 + chunk;  // 16-byte piece within the row
                int const first_row = row - lane;     // global row of the warp's first row
                if (row0 < rows_it) {
                    for (int rr = row0; rr < 32 && first_row + rr < p.M; rr += rows_it) {
                        int const tr = quarter * 32 + rr;
                        uint32_t const src = a_buf + (uint32_t)((gc >> 3) * kAStageBytes + tr * 128) + ((uint32_t)((gc & 7) ^ (tr & 7)) << 4);
                        uint4 x;
                        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "r"(src));
                        *reinterpret_cast<uint4*>(p.out + (int64_t)(first_row + rr) * C + gc * 8) = x;
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic accesses to A before the next TMA write
            __syncwarp();
            if (lane == 0) mbar_arrive(a_empty(t & 1));
            if (p.stats_out) {
                // one exchange buffer is enough: a warp gets to write the next tile's sums only after that tile's second
                // GEMM has finished, which takes an h_full arrival per chunk from the reader below
                uint32_t const mine = red_base + (uint32_t)(((quarter * 4 + slab) * 32 + lane) * 8);
                asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(mine), "f"(row_sum), "f"(row_sumsq) : "memory");
                asm volatile("bar.sync %0, 128;" ::"r"(1 + quarter) : "memory");
                if (slab == 0 && valid) {
                    float sx = 0.f, sq = 0.f;
#pragma unroll
                    for (int w4 = 0; w4 < 4; ++w4) {
                        float a, b;
                        asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(a), "=f"(b) : "r"(red_base + (uint32_t)(((quarter * 4 + w4) * 32 + lane) * 8)));
                        sx += a;
                        sq += b;
                    }
                    p.stats_out[row] = make_float2(sx, sq);
                }
            }
        };
        int prev_row = 0;
        for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++lt) {
            int const row = tile * kBlockM + quarter * 32 + lane;
            float rstd;
            {
                float const inv_k = 1.0f / (float)C, mean = pv_next.x * inv_k;
                rstd = rsqrtf(fmaxf(fmaf(-mean, mean, pv_next.y * inv_k), 0.f) + p.ln_eps);
                int const next_row = row + (int)gridDim.x * kBlockM;
                if (next_row < p.M) pv_next = __ldg(p.ln_stats + next_row);
            }
            for (int h = 0; h < NH; ++h, ++u) {
                int const sb = (int)(u & 1u), d = (int)(u & 3u);
                // The previous tile is finished one chunk late: its last MMA2 is then hidden behind this chunk, and the
                // MMA warp already has kMlpLead first-GEMM chunks of this tile in flight.
                bool const finish_prev = h == 0 && lt > 0;
                mbar_wait(d1_full(d), (u >> 2) & 1u);
                tc_fence_after();
                uint32_t r[16];
                tmem_ld16(tmem_d1 + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(d * kMlpChunk + slab * 16), r);
                float4 bq[4];  // this slab's folded bias (broadcast reads), under the TMEM load
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(bq[i].x), "=f"(bq[i].y), "=f"(bq[i].z), "=f"(bq[i].w)
                                 : "r"(bias_base + (uint32_t)((h * kMlpChunk + slab * 16 + 4 * i) * 4)));
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(d1_empty(d));
                float v[16];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    v[4 * i + 0] = fmaf(__uint_as_float(r[4 * i + 0]), rstd, bq[i].x);
                    v[4 * i + 1] = fmaf(__uint_as_float(r[4 * i + 1]), rstd, bq[i].y);
                    v[4 * i + 2] = fmaf(__uint_as_float(r[4 * i + 2]), rstd, bq[i].z);
                    v[4 * i + 3] = fmaf(__uint_as_float(r[4 * i + 3]), rstd, bq[i].w);
                }
                uint4 x[2];
                activate_pack16(v, ACT_GELU, x);
                mbar_wait(h_empty(sb), ((u >> 1) & 1u) ^ 1u);  // the second GEMM of chunk u - 2 has finished reading this buffer
                int const hrow = quarter * 32 + lane;
                uint32_t const rowaddr = h_base + (uint32_t)(sb * kAStageBytes + hrow * 128);
                uint32_t const c0 = (uint32_t)(slab * 2), sw = (uint32_t)(hrow & 7);
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(rowaddr + ((c0 ^ sw) << 4)), "r"(x[0].x), "r"(x[0].y), "r"(x[0].z), "r"(x[0].w) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(rowaddr + (((c0 + 1) ^ sw) << 4)), "r"(x[1].x), "r"(x[1].y), "r"(x[1].z), "r"(x[1].w) : "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(h_full(sb));
                if (finish_prev) finish_tile(prev_row, lt - 1);
            }
            prev_row = row;
        }
        if (lt > 0) finish_tile(prev_row, lt - 1);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
    }
}

// ---- CUDA-core cross-check / small-M kernel -------------------------------------------------------
template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<act_t>(act_t v) { return act2f(v); }

template <typename T>
__global__ void gemm_simt_kernel(T const* __restrict__ A, int64_t lda, T const* __restrict__ B, int64_t ldb, int M, int N,
                                 int K, void* out, EpiParams ep) {
    // 32x32 output tile per block, 32x8 threads, each thread 4 rows; K tiled by 32 through shared memory
    __shared__ float As[32][33];
    __shared__ float Bs[32][33];
    int const tx = threadIdx.x, ty = threadIdx.y;
    int const m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k0 = 0; k0 < K; k0 += 32) {
        for (int i = ty; i < 32; i += 8) {
            int const m = m0 + i, n = n0 + i, k = k0 + tx;
            As[i][tx] = (m < M && k < K) ? to_f32<T>(A[(int64_t)m * lda + k]) : 0.f;
            Bs[i][tx] = (n < N && k < K) ? to_f32<T>(B[(int64_t)n * ldb + k]) : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
            float const b = Bs[tx][k];
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] += As[ty + 8 * i][k] * b;
        }
        __syncthreads();
    }
    int const n = n0 + tx;
    if (n >= N) return;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int const m = m0 + ty + 8 * i;
        if (m >= M) continue;
        int64_t const orow = ep.row_map ? (int64_t)ep.row_map[m] : (int64_t)m;
        if (orow < 0) continue;
        float v = acc[i];
        if (ep.ln_stats) v *= ep.ln_stats[m].y;  // row-centred weights: no mean term
        if (ep.bias) v += ep.bias[n];
        int64_t const o = orow * ep.ldc + n;
        if (ep.residual)
            v += ep.out_f32 ? reinterpret_cast<float const*>(ep.residual)[o]
                            : act2f(reinterpret_cast<act_t const*>(ep.residual)[o]);
        if (ep.act == ACT_GELU) v = gelu_erf(v);
        else if (ep.act == ACT_RELU) v = fmaxf(v, 0.f);
        if (ep.out_f32) reinterpret_cast<float*>(out)[o] = v;
        else reinterpret_cast<act_t*>(out)[o] = f2act(v);
    }
}

// ---- host side ------------------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, cuuint64_t const*,
                                   cuuint64_t const*, cuuint32_t const*, cuuint32_t const*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        if (!p || qres != cudaDriverEntryPointSuccess) fail("cuTensorMapEncodeTiled is not available in this driver");
        fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

CUtensorMap encode_map(Operand const& op, bool tf32, int box_rows);

// Tensor maps depend only on (pointer, shape, pitch, box, type); the engine's workspaces and weights have stable
// addresses, so encoded maps are memoised instead of calling into the driver twice per GEMM launch.
CUtensorMap make_map(Operand const& op, bool tf32, int box_rows) {
    struct Key {
        void const* ptr;
        int64_t rows, cols, pitch;
        int box_rows, tf32;
        bool operator<(Key const& o) const {
            return std::tie(ptr, rows, cols, pitch, box_rows, tf32) < std::tie(o.ptr, o.rows, o.cols, o.pitch, o.box_rows, o.tf32);
        }
    };
    static std::mutex mutex;
    static std::map<Key, CUtensorMap> cache;
    Key const key{op.ptr, op.rows, op.cols, op.pitch ? op.pitch : op.cols, box_rows, tf32 ? 1 : 0};
    std::lock_guard<std::mutex> lock(mutex);
    auto it = cache.find(key);
    if (it == cache.end()) {
        if (cache.size() > 8192) cache.clear();
        it = cache.emplace(key, encode_map(op, tf32, box_rows)).first;
    }
    return it->second;
}

CUtensorMap encode_map(Operand const& op, bool tf32, int box_rows) {
    CUtensorMap map;
    int const esz = tf32 ? 4 : 2;
    int64_t const pitch = op.pitch ? op.pitch : op.cols;
    DLIMG_ASSERT((reinterpret_cast<uintptr_t>(op.ptr) & 15) == 0);
    DLIMG_ASSERT((pitch * esz) % 16 == 0);
    cuuint64_t dims[2] = {(cuuint64_t)op.cols, (cuuint64_t)op.rows};
    cuuint64_t strides[1] = {(cuuint64_t)(pitch * esz)};
    cuuint32_t box[2] = {(cuuint32_t)(kKBytes / esz), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMapDataType const dtype = tf32 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32
                                           : (kActBf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
    CUresult r = encode_fn()(&map, dtype, 2,
                             const_cast<void*>(op.ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) fail("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return map;
}

EpiParams to_params(Epilogue const& e, int N) {
    EpiParams p;
    p.bias = e.bias;
    p.residual = e.residual;
    p.row_map = e.row_map;
    p.ln_stats = e.ln_stats;
    p.stats_out = e.stats_out;
    p.ln_parts = e.ln_parts;
    p.ln_eps = e.ln_eps;
    p.act = e.act;
    p.out_f32 = e.out_f32;
    p.ldc = e.ldc ? e.ldc : N;
    p.res_mod = e.res_mod;
    p.fuse_a = e.fuse_a;
    p.fuse_b = e.fuse_b;
    p.fuse_mode = e.fuse_mode;
    p.ksplit = e.ksplit > 1 ? e.ksplit : 1;
    p.fuse_out = e.fuse_out;
    p.res_table = e.res_table;
    return p;
}

}  // namespace

CUtensorMap make_tensor_map(Operand const& op, bool tf32, int box_rows) { return make_map(op, tf32, box_rows); }

CUtensorMap make_tensor_map_nhwc(void const* ptr, int batch, int H, int W, int C, int box_c, int box_w, int box_h,
                                 bool swizzle128) {
    CUtensorMap map;
    DLIMG_ASSERT(!swizzle128 || box_c == 64);  // one 128-byte swizzle row per pixel
    DLIMG_ASSERT((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (C * 2) % 16 == 0 && box_c <= 256 && box_w <= 256 && box_h <= 256);
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)batch};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode_fn()(&map, kActBf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                             const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) fail("cuTensorMapEncodeTiled (NHWC) failed with code " + std::to_string((int)r));
    return map;
}

int pick_block_n(int N) {
    for (int bn = kMaxBlockN; bn >= 16; bn -= 16)
        if (N % bn == 0) return bn;
    return 0;
}

namespace {
struct ConvInput {
    int batch, H, W, C;
};
void launch_impl(cudaStream_t stream, bool tf32, Operand const& a, Operand const& b, void* out, Epilogue const& epi,
                 int num_sms, ConvInput const* ci);
}  // namespace

void launch(cudaStream_t stream, bool tf32, Operand const& a, Operand const& b, void* out, Epilogue const& epi,
            int num_sms) {
    launch_impl(stream, tf32, a, b, out, epi, num_sms, nullptr);
}

void launch_conv3x3(cudaStream_t stream, void const* in, int batch, int H, int W, int C, Operand const& b, void* out,
                    Epilogue const& epi, int num_sms) {
    DLIMG_ASSERT(W > 0 && kBlockM % W == 0 && H % (kBlockM / W) == 0 && C % 64 == 0 && b.cols == 9 * (int64_t)C);
    ConvInput const ci{batch, H, W, C};
    launch_impl(stream, false, Operand{in, (int64_t)batch * H * W, 9 * (int64_t)C, 0}, b, out, epi, num_sms, &ci);
}

namespace {
void launch_impl(cudaStream_t stream, bool tf32, Operand const& a, Operand const& b, void* out, Epilogue const& epi,
                 int num_sms, ConvInput const* ci) {
    int const ks = epi.ksplit > 1 ? epi.ksplit : 1;
    int const M = ks > 1 ? (int)(split_rows(a.rows) * ks) : (int)a.rows, N = (int)b.rows, K = (int)a.cols;
    DLIMG_ASSERT(a.cols == b.cols);
    if (ks > 1)  // whole k-blocks per split, plain fp32 outputs: the tall problem the epilogue sees has no other meaning
        DLIMG_ASSERT(!ci && epi.out_f32 && !epi.fuse && !epi.residual && !epi.res_table && !epi.ln_stats && !epi.stats_out && !epi.row_map &&
                     !epi.bias && K % (ks * (tf32 ? 32 : 64)) == 0);
    DLIMG_ASSERT(M > 0 && N > 0 && K > 0);
    int block_n = pick_block_n(N);
    if (block_n == 0) fail("GEMM: N must be a multiple of 16, got " + std::to_string(N));
    EpiParams ep = to_params(epi, N);
    // Small problems (the decoder's token-side Linears, single-image encoder passes): narrower tiles put more SMs to
    // work and shorten the per-CTA k-loop.  The sum over K of an output element does not depend on the tile width, so
    // results are bit-identical; producers of row sums keep their width (the consumer counts the partial sums).
    if (epi.fuse == 3) {
        DLIMG_ASSERT(!tf32 && !ci && ep.act == ACT_NONE && (ep.residual || ep.res_table) && !ep.ln_stats && !ep.stats_out && !ep.row_map && !ep.out_f32);
        DLIMG_ASSERT(N == 256 && ep.bias && ep.fuse_a && ep.fuse_b && (!ep.res_table || ep.res_mod > 0));
    } else if (epi.fuse) {
        DLIMG_ASSERT(!tf32 && !ci && ep.act == ACT_GELU && !ep.residual && !ep.ln_stats && !ep.stats_out && !ep.row_map && !ep.out_f32);
        DLIMG_ASSERT((epi.fuse == 1 && N == 256 && ep.fuse_a && ep.fuse_b) || (epi.fuse == 2 && N == 128 && M % 16384 == 0 && ep.fuse_a && ep.fuse_out && (epi.fuse_mode != MASKS_BEST || ep.fuse_b)));
    } else if (!ep.stats_out) {
        int const tiles_m = ceil_div(M, kBlockM);
        while (block_n >= 128 && block_n % 32 == 0 && N % (block_n / 2) == 0 && 2 * tiles_m * (N / block_n) <= num_sms) block_n /= 2;
    }
    DLIMG_ASSERT(ep.ldc % (ep.out_f32 ? 4 : 8) == 0);
    ConvParams conv;
    if (ci) {
        conv.w = ci->W;
        conv.h = ci->H;
        conv.cblocks = ci->C / 64;
    }
    CUtensorMap ma = ci ? make_tensor_map_nhwc(a.ptr, ci->batch, ci->H, ci->W, ci->C, 64, ci->W, kBlockM / ci->W, true)
                        : make_map(a, tf32, kBlockM);
    CUtensorMap mb = make_map(b, tf32, block_n);
    int const tiles = ceil_div(M, kBlockM) * (N / block_n);
    ProfScope prof(stream, tf32 ? CAT_GEMM_TF32 : CAT_GEMM_BF16, 2.0 * M * N * K / ks,
                   (double)(tf32 ? 4 : 2) * ((double)M * (ci ? ci->C : K) + (double)N * K) +
                       (double)(ep.out_f32 ? 4 : 2) * M * N * (ep.residual ? 2.0 : 1.0));
    int const grid = tiles < num_sms ? tiles : num_sms;
    // plain 16-bit outputs go through the coalescing (staged) epilogue; residual / scatter / fp32 outputs store directly
    static bool const allow_staged = !dev_switch("DLIMG_B200_GEMM_DIRECT");  // A/B switch
    static bool const allow_staged_res = !dev_switch("DLIMG_B200_GEMM_DIRECT_RESIDUAL");  // A/B switch
    bool const res_ok = !(ep.residual || ep.res_table) || (allow_staged_res && ep.act == ACT_NONE && !ep.ln_stats);
    if (ep.res_mod) DLIMG_ASSERT((ep.residual || ep.res_table) && ep.res_mod % kBlockM == 0 && M % ep.res_mod == 0);
    static bool const allow_staged_f32 = !dev_switch("DLIMG_B200_GEMM_DIRECT_F32");  // A/B switch
    bool const staged_f32 = allow_staged && allow_staged_f32 && tf32 && ep.out_f32 && !ep.residual && !ep.row_map && !ep.stats_out &&
                            !ep.ln_stats && block_n >= 64 && N <= 2048;
    bool const staged = staged_f32 ||
                        (allow_staged && !tf32 && res_ok && !ep.row_map && !ep.out_f32 && block_n >= 64 &&
                         (ep.act == ACT_NONE || ep.act == ACT_GELU) && (!ep.stats_out || (ep.act == ACT_NONE && !ep.ln_stats)) && N <= 2048);
    if (ep.res_mod && !(staged && !tf32)) fail("GEMM: a residual table (res_mod) needs the staged 16-bit epilogue");
    if (ep.ln_stats && (!staged || !ep.bias))
        fail("GEMM: the folded LayerNorm needs a plain 16-bit output (staged epilogue) and a bias");
    if (ep.stats_out && (ep.act != ACT_NONE || ep.row_map || ep.out_f32 || block_n < 64))
        fail("GEMM: row statistics are produced by the 16-bit epilogues without activation only");
    int const bias_bytes = staged ? (int)round_up64((int64_t)N * 4 * (epi.fuse == 3 ? 3 : 1), 1024) : 0;  // fuse 3: + gamma, beta
    SmemPlan const sp = plan_smem(block_n, staged, ep.stats_out != nullptr || epi.fuse == 3, bias_bytes, tf32);
    DLIMG_ASSERT(sp.stages >= 2);
    using Kernel = void (*)(CUtensorMap, CUtensorMap, int, int, int, int, int, int, int, void*, EpiParams, ConvParams);
    Kernel kernel;
    if (tf32) kernel = staged ? gemm_tc_kernel<1, true> : gemm_tc_kernel<1, false>;
    else if (!staged) kernel = gemm_tc_kernel<0, false>;
    else if (epi.fuse == 3) kernel = gemm_tc_kernel<0, true, ACT_NONE, false, true, 3>;
    else if (ep.residual || ep.stats_out) kernel = gemm_tc_kernel<0, true, ACT_NONE, false, true>;
    else if (ep.ln_stats) kernel = ep.act == ACT_GELU ? gemm_tc_kernel<0, true, ACT_GELU, true> : gemm_tc_kernel<0, true, ACT_NONE, true>;
    else kernel = ep.act == ACT_GELU ? gemm_tc_kernel<0, true, ACT_GELU, false> : gemm_tc_kernel<0, true, ACT_NONE, false>;
    if (epi.fuse == 3) DLIMG_ASSERT(staged && block_n == 256);
    if (epi.fuse == 1 || epi.fuse == 2) {
        DLIMG_ASSERT(staged);
        kernel = epi.fuse == 1 ? gemm_tc_kernel<0, true, ACT_GELU, false, false, 1> : gemm_tc_kernel<0, true, ACT_GELU, false, false, 2>;
    }
    {
        static std::mutex attr_mutex;
        static std::map<void const*, bool> attr_done;
        std::lock_guard<std::mutex> lock(attr_mutex);
        if (!attr_done[(void const*)kernel]) {
            CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
            attr_done[(void const*)kernel] = true;
        }
    }
    launch_pdl(PDL_GEMM, kernel, dim3(grid), dim3(kNumThreads), (size_t)sp.total_bytes, stream, ma, mb, M, N, K, block_n, sp.stages, sp.staging_bytes, bias_bytes,
               out, ep, conv);
    KERNEL_CHECK();
}
}  // namespace

bool mlp_fused_supported(int C) {
#if defined(DLIMG_B200_ACT_BF16)
    (void)C;
    return false;
#else
    return (C == 128 || C == 160) && plan_mlp(C).total <= kSmemLimit;
#endif
}

void launch_mlp_fused(cudaStream_t stream, void const* x, int64_t rows, int C, void const* w1, float const* b1,
                      float2 const* ln_stats, float ln_eps, void const* w2, float const* b2, void* out, float2* stats_out,
                      int num_sms) {
    DLIMG_ASSERT(mlp_fused_supported(C));
    int const M = (int)rows, H = 4 * C;
    MlpSmem const sp = plan_mlp(C);
    CUtensorMap const mx = make_map(Operand{x, rows, C, C}, false, kBlockM);
    CUtensorMap const m1 = make_map(Operand{w1, H, C, C}, false, kMlpChunk);
    CUtensorMap const m2 = make_map(Operand{w2, C, H, H}, false, C);
    MlpParams p;
    p.M = M;
    p.C = C;
    p.b1 = b1;
    p.b2 = b2;
    p.ln_stats = ln_stats;
    p.ln_eps = ln_eps;
    p.out = static_cast<act_t*>(out);
    p.stats_out = stats_out;
    ProfScope prof(stream, CAT_GEMM_BF16, 4.0 * M * (double)C * H, 2.0 * ((double)M * C * 3 + 2.0 * C * H));
    static std::once_flag once;
    std::call_once(once, [] { CUDA_CHECK(cudaFuncSetAttribute(mlp_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit)); });
    int const tiles = ceil_div(M, kBlockM);
    int const grid = tiles < num_sms ? tiles : num_sms;
    launch_pdl(PDL_GEMM, mlp_fused_kernel, dim3(grid), dim3(kNumThreads), (size_t)sp.total, stream, mx, m1, m2, p, sp.kb1, sp.a_bytes, sp.w1_bytes,
               sp.w2_bytes, sp.w1_ring_log);
    KERNEL_CHECK();
}

void launch_simt(cudaStream_t stream, bool f32_operands, Operand const& a, Operand const& b, void* out,
                 Epilogue const& epi) {
    int const M = (int)a.rows, N = (int)b.rows, K = (int)a.cols;
    DLIMG_ASSERT(a.cols == b.cols && epi.ksplit <= 1);
    EpiParams ep = to_params(epi, N);
    ProfScope prof(stream, CAT_OTHER, 2.0 * M * N * K);
    dim3 grid(ceil_div(N, 32), ceil_div(M, 32)), block(32, 8);
    int64_t const lda = a.pitch ? a.pitch : a.cols, ldb = b.pitch ? b.pitch : b.cols;
    if (f32_operands)
        gemm_simt_kernel<float><<<grid, block, 0, stream>>>((float const*)a.ptr, lda, (float const*)b.ptr, ldb, M, N, K, out, ep);
    else
        gemm_simt_kernel<act_t><<<grid, block, 0, stream>>>((act_t const*)a.ptr, lda,
                                                                     (act_t const*)b.ptr, ldb, M, N, K, out, ep);
    KERNEL_CHECK();
}

}  // namespace gemm
}  // namespace dlimg
