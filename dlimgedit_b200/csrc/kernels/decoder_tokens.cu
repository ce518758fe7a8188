// decoder_tokens.cu -- the token side of the SAM two-way transformer (7 tokens x 256 per prompt) as a few fused kernels.
//
// Per prompt the token side is tiny (7 rows), but as separate launches -- three projections, the attention core, the output
// projection, a LayerNorm, the next projection ... -- it was 23 launches per layer of 5-22 us each, 30 % of a 64-prompt
// pass, all of it launch and dependency latency (profiles/r02a_summary.md).  Here a CLUSTER OF FOUR CTAs owns one prompt
// and keeps its 7 x 256 state in (distributed) shared memory across a whole block of the layer:
//     token_attn_block     [+pe] -> q, k, v projections -> 8-head self-attention -> out projection -> [+residual] LayerNorm
//                          -> query projection of the following tokens->image attention
//     token_post_t2i       merge of the split-key attention partials -> out projection -> +residual -> LayerNorm
//     token_post_mlp       +residual -> LayerNorm -> key / value projections of image->tokens attention (and the query
//                          projection of the final attention)
// CTA c of a cluster computes output features [64c, 64c + 64) of every 256-wide projection (= attention heads 2c, 2c + 1,
// so the attention core needs no exchange) and [32c, 32c + 32) of the 128-wide ones: it reads a QUARTER of each weight
// matrix (a single CTA per prompt had to pull every 256 KB matrix through one SM's L2 port: 60 us per block).  What the
// next step needs from the other three CTAs -- the attention output, the LayerNorm row sums, the normalised rows --
// travels through distributed shared memory (st.shared::cluster) followed by a cluster barrier.
// The 7-row MLP (256 -> 2048 -> 256) stays on the tensor-core GEMM: its 4 MB of weights want to be read once per pass.
// Weights are stored transposed and k-blocked, [K / 4][N][4], so that neighbouring threads read neighbouring output features
// (coalesced 16-byte loads), each value used for the prompt's 7 rows.  fp32 FMA throughout; a prompt's result does not depend on which other
// prompts share the pass.
#include "decoder_kernels.cuh"

#include "../profiler.hpp"

#include <cooperative_groups.h>

#include <map>
#include <mutex>

namespace dlimg {
namespace dec {

namespace {

namespace cg = cooperative_groups;

constexpr int kT = kTokens;      // 7
constexpr int kThreads = 256;
constexpr int kCl = 4;           // CTAs per prompt
constexpr int kS256 = kDim / kCl;  // 64: this CTA's features of a 256-wide projection
constexpr int kS128 = 128 / kCl;   // 32: ... of a 128-wide one

// out_s[r][f] = bias[n0 + f] + sum_k xs[r][k] * Wt[k][n0 + f]   for r < 7, f < NOUT, with all 256 threads: thread
// (part, f) sums its share of k, the 256 / NOUT partial sums meet in shared memory.
// xs: shared [7][K]; Wt: global (K, N) row-major; scratch: shared [256 / NOUT][7][NOUT]; out_s: shared [7][NOUT].
// The projection comes in two halves so that a kernel can request the weight slice of its NEXT projection before it
// reduces the current one (and the first slice before it even loads its inputs): the slice lives in registers only between
// proj_load and proj_run's FMA loop, so one slice is live at a time and the trips to L2 overlap the reductions, the
// attention core and the cluster barriers instead of standing in front of every projection.
template <int NOUT, int K>
struct ProjW {
    static constexpr int kParts = kThreads / NOUT;
    static constexpr int kPer = K / kParts;  // k values per thread: 64 (256 -> 64 wide), 32 (256 -> 32 wide, 128 -> 64 wide)
    static_assert(kPer % 4 == 0 && kPer <= 64, "k slice per thread");
    float4 w[kPer / 4];
};
// Wt is stored as [K / 4][N][4]: the four k values of one output feature are one 16-byte load, so a thread's whole
// weight slice is 1 .. 16 loads, all requested at once (with scalar loads ptxas kept a rolling window of ~20 of the 64 in
// flight: 15 cycles of scoreboard stall per issued instruction)
template <int NOUT, int K>
__device__ __forceinline__ void proj_load(ProjW<NOUT, K>& pw, float const* __restrict__ Wt, int N, int n0) {
    int const f = threadIdx.x % NOUT, part = threadIdx.x / NOUT;
    float4 const* wp = reinterpret_cast<float4 const*>(Wt) + (size_t)(part * ProjW<NOUT, K>::kPer / 4) * N + n0 + f;
#pragma unroll
    for (int k = 0; k < ProjW<NOUT, K>::kPer / 4; ++k) pw.w[k] = __ldg(wp + (size_t)k * N);
}
// partial sums of this thread's k slice -> scratch
template <int NOUT, int K>
__device__ __forceinline__ void proj_fma(ProjW<NOUT, K> const& pw, float const* xs, float* scratch) {
    constexpr int kPer = ProjW<NOUT, K>::kPer;
    int const f = threadIdx.x % NOUT, part = threadIdx.x / NOUT;
    int const k0 = part * kPer;
    float acc[kT];
#pragma unroll
    for (int r = 0; r < kT; ++r) acc[r] = 0.f;
#pragma unroll
    for (int k = 0; k < kPer / 4; ++k) {
#pragma unroll
        for (int r = 0; r < kT; ++r) {
            float4 const x = *reinterpret_cast<float4 const*>(xs + r * K + k0 + 4 * k);
            acc[r] = fmaf(x.x, pw.w[k].x, acc[r]);
            acc[r] = fmaf(x.y, pw.w[k].y, acc[r]);
            acc[r] = fmaf(x.z, pw.w[k].z, acc[r]);
            acc[r] = fmaf(x.w, pw.w[k].w, acc[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < kT; ++r) scratch[(part * kT + r) * NOUT + f] = acc[r];
}
// bias + the partial sums, in a fixed order -> out_s; barriers on both sides
template <int NOUT, int K>
__device__ __forceinline__ void proj_reduce(int n0, float const* __restrict__ bias, float const* scratch, float* out_s) {
    constexpr int kParts = ProjW<NOUT, K>::kParts;
    __syncthreads();
    for (int i = threadIdx.x; i < kT * NOUT; i += kThreads) {
        int const ff = i % NOUT;
        float s = __ldg(bias + n0 + ff);
#pragma unroll
        for (int p = 0; p < kParts; ++p) s += scratch[p * kT * NOUT + i];
        out_s[i] = s;
    }
    __syncthreads();
}
template <int NOUT, int K>
__device__ __forceinline__ void cta_proj(float const* xs, float const* __restrict__ Wt, int N, int n0,
                                         float const* __restrict__ bias, float* scratch, float* out_s) {
    ProjW<NOUT, K> pw;
    proj_load(pw, Wt, N, n0);
    proj_fma(pw, xs, scratch);
    proj_reduce<NOUT, K>(n0, bias, scratch, out_s);
}

// LayerNorm over 256 features that are spread over the cluster: this CTA holds v[r][f] for its 64 features (shared [7][64]).
// Every CTA leaves its per-row (sum, sum of squares) in the `ln_part` buffer of all four; after the cluster barrier each
// CTA has the four partial pairs of every row.  stats (shared [7][2]) <- (mean, rstd).
__device__ __forceinline__ void cluster_row_stats(cg::cluster_group& cluster, float const* v_s, float* ln_part /* [4][7][2] */,
                                                  float* stats) {
    int const rank = (int)cluster.block_rank();
    int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < kT) {  // warp r reduces row r of the local slice
        float const a = v_s[warp * kS256 + lane], b = v_s[warp * kS256 + 32 + lane];
        float s = a + b, q = a * a + b * b;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        if (lane < kCl) {  // lane d delivers to CTA d
            float* dst = cluster.map_shared_rank(ln_part, lane);
            dst[(rank * kT + warp) * 2 + 0] = s;
            dst[(rank * kT + warp) * 2 + 1] = q;
        }
    }
    cluster.sync();
    if (threadIdx.x < kT) {
        float s = 0.f, q = 0.f;
#pragma unroll
        for (int c = 0; c < kCl; ++c) {
            s += ln_part[(c * kT + threadIdx.x) * 2 + 0];
            q += ln_part[(c * kT + threadIdx.x) * 2 + 1];
        }
        float const mean = s * (1.0f / kDim);
        stats[threadIdx.x * 2 + 0] = mean;
        stats[threadIdx.x * 2 + 1] = rsqrtf(fmaxf(q * (1.0f / kDim) - mean * mean, 0.f) + 1e-5f);
    }
    __syncthreads();
}

// copies this CTA's [7][64] slice into columns [64 * rank, +64) of a [7][256] buffer in every CTA of the cluster
__device__ __forceinline__ void cluster_scatter_slice(cg::cluster_group& cluster, float const* slice_s, float* full /* [7][256] */) {
    int const rank = (int)cluster.block_rank();
    for (int i = threadIdx.x; i < kCl * kT * kS256; i += kThreads) {
        int const dst_cta = i / (kT * kS256), j = i % (kT * kS256), r = j / kS256, f = j % kS256;
        cluster.map_shared_rank(full, dst_cta)[r * kDim + rank * kS256 + f] = slice_s[j];
    }
}

struct AttnSmem {
    float xs[kT * kDim], xp[kT * kDim], full[kT * kDim];  // queries, queries (+ pe), gathered rows (attention output / new rows)
    float q[kT * kS256], k[kT * kS256], v[kT * kS256], o[kT * kS256];
    float scratch[kThreads * kT];
    float sc[2 * kT * kT];
    float ln_part[kCl * kT * 2], stats[kT * 2];
};

__global__ void __cluster_dims__(kCl, 1, 1) __launch_bounds__(kThreads) token_attn_block_kernel(TokenAttnBlock p) {
    extern __shared__ __align__(16) uint8_t tok_smem[];
    AttnSmem& sm = *reinterpret_cast<AttnSmem*>(tok_smem);
    cg::cluster_group cluster = cg::this_cluster();
    int const rank = (int)cluster.block_rank();
    int const prompt = blockIdx.x / kCl, n = threadIdx.x;
    int const n0 = rank * kS256;
    float* q_g = p.queries + (size_t)prompt * kT * kDim;
    float const* pe_g = p.pe + (size_t)prompt * kT * kDim;
    ProjW<kS256, kDim> pw;  // one weight slice in flight or in use at a time (see proj_load)
    proj_load(pw, p.wq_t, kDim, n0);
#pragma unroll
    for (int r = 0; r < kT; ++r) {
        float const x = q_g[r * kDim + n], e = pe_g[r * kDim + n];
        sm.xs[r * kDim + n] = x;
        sm.xp[r * kDim + n] = p.with_pe ? x + e : x;
    }
    __syncthreads();
    // q, k from queries (+ pe), v from queries: this CTA's 64 features = heads 2 * rank, 2 * rank + 1
    proj_fma(pw, sm.xp, sm.scratch);
    proj_load(pw, p.wk_t, kDim, n0);
    proj_reduce<kS256, kDim>(n0, p.bq, sm.scratch, sm.q);
    proj_fma(pw, sm.xp, sm.scratch);
    proj_load(pw, p.wv_t, kDim, n0);
    proj_reduce<kS256, kDim>(n0, p.bk, sm.scratch, sm.k);
    proj_fma(pw, sm.xs, sm.scratch);
    proj_load(pw, p.wo_t, kDim, n0);  // in flight during the attention core and the cluster exchange
    proj_reduce<kS256, kDim>(n0, p.bv, sm.scratch, sm.v);
    // scores of the two local heads: 2 x 7 x 7, head_dim 32, scale 1 / sqrt(32)
    if (n < 2 * kT * kT) {
        int const h = n / (kT * kT), t = (n / kT) % kT, u = n % kT;
        float const* qq = sm.q + t * kS256 + h * 32;
        float const* kk = sm.k + u * kS256 + h * 32;
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < 32; ++d) s = fmaf(qq[(d + u) & 31], kk[(d + u) & 31], s);  // rotated start: rows are 64 floats apart
        sm.sc[n] = s * 0.17677669529663687f;
    }
    __syncthreads();
    if (n < 2 * kT) {  // softmax over the 7 keys of (head, query)
        float* row = sm.sc + n * kT;
        float mx = row[0];
#pragma unroll
        for (int u = 1; u < kT; ++u) mx = fmaxf(mx, row[u]);
        float e[kT], sum = 0.f;
#pragma unroll
        for (int u = 0; u < kT; ++u) {
            e[u] = expf(row[u] - mx);
            sum += e[u];
        }
#pragma unroll
        for (int u = 0; u < kT; ++u) row[u] = e[u] / sum;
    }
    __syncthreads();
    if (n < kS256) {  // attention output of feature n (head n / 32)
        int const h = n >> 5;
        float vv[kT];
#pragma unroll
        for (int u = 0; u < kT; ++u) vv[u] = sm.v[u * kS256 + n];
#pragma unroll
        for (int t = 0; t < kT; ++t) {
            float const* w = sm.sc + (h * kT + t) * kT;
            float o = 0.f;
#pragma unroll
            for (int u = 0; u < kT; ++u) o = fmaf(w[u], vv[u], o);
            sm.o[t * kS256 + n] = o;
        }
    }
    __syncthreads();
    cluster_scatter_slice(cluster, sm.o, sm.full);  // every CTA needs all 256 features as the input of the out projection
    cluster.sync();
    proj_fma(pw, sm.full, sm.scratch);
    ProjW<kS128, kDim> pw_next;
    proj_load(pw_next, p.w_next_t, 128, rank * kS128);  // in flight during the LayerNorm and its two cluster barriers
    proj_reduce<kS256, kDim>(n0, p.bo, sm.scratch, sm.o);
    if (p.residual) {
        for (int i = n; i < kT * kS256; i += kThreads) sm.o[i] += sm.xs[(i / kS256) * kDim + n0 + (i % kS256)];
        __syncthreads();
    }
    cluster_row_stats(cluster, sm.o, sm.ln_part, sm.stats);
    for (int i = n; i < kT * kS256; i += kThreads) {
        int const r = i / kS256, c = n0 + (i % kS256);
        float const y = (sm.o[i] - sm.stats[2 * r]) * sm.stats[2 * r + 1] * __ldg(p.gamma + c) + __ldg(p.beta + c);
        q_g[r * kDim + c] = y;
        sm.o[i] = y + pe_g[r * kDim + c];  // input of the following query projection
    }
    __syncthreads();
    // gathered into `xp`, which every CTA stopped reading at its k projection (all are past the barrier behind the attention
    // exchange): `full` is still being read by slower CTAs' out projections, and a buffer of its own saves a cluster barrier
    cluster_scatter_slice(cluster, sm.o, sm.xp);
    cluster.sync();  // the last remote access of the kernel lies before this barrier: CTAs may exit independently afterwards
    // query projection of tokens -> image attention: (queries + pe) W^T + b, 256 -> 128; this CTA's 32 features
    proj_fma(pw_next, sm.xp, sm.scratch);
    proj_reduce<kS128, kDim>(rank * kS128, p.b_next, sm.scratch, sm.q);
    for (int i = n; i < kT * kS128; i += kThreads)
        p.out_next[((size_t)prompt * kT + i / kS128) * 128 + rank * kS128 + (i % kS128)] = sm.q[i];
}

struct PostSmem {
    float in_full[kT * kDim], xp_full[kT * kDim];
    float o[kT * kS256], p32[kT * kS128];
    float scratch[kThreads * kT];
    float ln_part[kCl * kT * 2], stats[kT * 2];
};

// tokens -> image attention, second half: merge the split-key partials (max / sum / accumulators per split), output
// projection 128 -> 256, + residual, LayerNorm.
__global__ void __cluster_dims__(kCl, 1, 1) __launch_bounds__(kThreads) token_post_t2i_kernel(TokenPostT2i p) {
    extern __shared__ __align__(16) uint8_t tok_smem[];
    PostSmem& sm = *reinterpret_cast<PostSmem*>(tok_smem);
    cg::cluster_group cluster = cg::this_cluster();
    int const rank = (int)cluster.block_rank();
    int const prompt = blockIdx.x / kCl, n = threadIdx.x, n0 = rank * kS256;
    constexpr int kPart = 128 + 16;
    ProjW<kS256, 128> pw;
    proj_load(pw, p.wo_t, kDim, n0);  // under the merge of the attention partials
    // every CTA needs all 7 x 128 merged values (they are its GEMM input): each merges the 32 dims of its own two heads --
    // one element per thread, one round of loads -- and delivers them to all four through distributed shared memory
    if (n < kT * 32) {
        int const t = n >> 5, d = rank * 32 + (n & 31), h = d >> 4;
        float const* src = p.partials + ((size_t)prompt * kT2iSplits * kT + t) * kPart;
        size_t const split_stride = (size_t)kT * kPart;
        float mx[kT2iSplits], ac[kT2iSplits], su[kT2iSplits];
#pragma unroll
        for (int sp = 0; sp < kT2iSplits; ++sp) {
            mx[sp] = src[sp * split_stride + 128 + h];
            ac[sp] = src[sp * split_stride + d];
            su[sp] = src[sp * split_stride + 136 + h];
        }
        float M = -INFINITY;
#pragma unroll
        for (int sp = 0; sp < kT2iSplits; ++sp) M = fmaxf(M, mx[sp]);
        float A = 0.f, S = 0.f;
#pragma unroll
        for (int sp = 0; sp < kT2iSplits; ++sp) {
            float const e = __expf(mx[sp] - M);
            A = fmaf(ac[sp], e, A);
            S = fmaf(su[sp], e, S);
        }
        float const o = A / S;
#pragma unroll
        for (int c = 0; c < kCl; ++c) cluster.map_shared_rank(sm.in_full, c)[t * 128 + d] = o;
    }
    cluster.sync();
    float* q_g = p.queries + (size_t)prompt * kT * kDim;
    float res[2];  // this thread's residual values, requested before the projection
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        int const i = n + j * kThreads;
        res[j] = i < kT * kS256 ? q_g[(i / kS256) * kDim + n0 + (i % kS256)] : 0.f;
    }
    proj_fma(pw, sm.in_full, sm.scratch);
    proj_reduce<kS256, 128>(n0, p.bo, sm.scratch, sm.o);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        int const i = n + j * kThreads;
        if (i < kT * kS256) sm.o[i] += res[j];
    }
    __syncthreads();
    cluster_row_stats(cluster, sm.o, sm.ln_part, sm.stats);
    for (int i = n; i < kT * kS256; i += kThreads) {
        int const r = i / kS256, c = n0 + (i % kS256);
        q_g[r * kDim + c] = (sm.o[i] - sm.stats[2 * r]) * sm.stats[2 * r + 1] * __ldg(p.gamma + c) + __ldg(p.beta + c);
    }
    // (no closing cluster barrier: the last remote access is in front of the barrier inside cluster_row_stats)
}

// After the token MLP: queries <- LayerNorm(queries + mlp_out) (mlp_out = bias + the split-K partial sums), then up to three 256 -> 128 projections of the new
// queries (with or without the positional encoding added): keys / values of image -> tokens attention, and after the
// last layer the queries of the final tokens -> image attention.
__global__ void __cluster_dims__(kCl, 1, 1) __launch_bounds__(kThreads) token_post_mlp_kernel(TokenPostMlp p) {
    extern __shared__ __align__(16) uint8_t tok_smem[];
    PostSmem& sm = *reinterpret_cast<PostSmem*>(tok_smem);
    cg::cluster_group cluster = cg::this_cluster();
    int const rank = (int)cluster.block_rank();
    int const prompt = blockIdx.x / kCl, n = threadIdx.x, n0 = rank * kS256;
    float* q_g = p.queries + (size_t)prompt * kT * kDim;
    float const* m_g = p.mlp_out + (size_t)prompt * kT * kDim;
    float const* pe_g = p.pe + (size_t)prompt * kT * kDim;
    ProjW<kS128, kDim> pw;
    proj_load(pw, p.w_t[0], 128, rank * kS128);  // under the residual sum, the LayerNorm and its cluster barriers
    constexpr int kMaxParts = 8;  // split-K partials summed from registers (all loads of a row in flight at once)
    for (int i = n; i < kT * kS256; i += kThreads) {
        int const g = (i / kS256) * kDim + n0 + (i % kS256);
        float part[kMaxParts];
#pragma unroll
        for (int sp = 0; sp < kMaxParts; ++sp) part[sp] = sp < p.mlp_parts ? m_g[(size_t)sp * p.mlp_part_stride + g] : 0.f;
        float v = q_g[g] + __ldg(p.mlp_bias + n0 + (i % kS256));
#pragma unroll
        for (int sp = 0; sp < kMaxParts; ++sp) v += part[sp];  // fixed order; unused slots add zero
        for (int sp = kMaxParts; sp < p.mlp_parts; ++sp) v += m_g[(size_t)sp * p.mlp_part_stride + g];
        sm.o[i] = v;
    }
    __syncthreads();
    cluster_row_stats(cluster, sm.o, sm.ln_part, sm.stats);
    for (int i = n; i < kCl * kT * kS256; i += kThreads) {  // normalised slice -> global + both gathered buffers of every CTA
        int const dst_cta = i / (kT * kS256), j = i % (kT * kS256), r = j / kS256, c = n0 + (j % kS256);
        float const y = (sm.o[j] - sm.stats[2 * r]) * sm.stats[2 * r + 1] * __ldg(p.gamma + c) + __ldg(p.beta + c);
        if (dst_cta == 0) q_g[r * kDim + c] = y;
        cluster.map_shared_rank(sm.in_full, dst_cta)[r * kDim + c] = y;
        cluster.map_shared_rank(sm.xp_full, dst_cta)[r * kDim + c] = y + pe_g[r * kDim + c];
    }
    cluster.sync();
    for (int j = 0; j < p.count; ++j) {
        proj_fma(pw, p.with_pe[j] ? sm.xp_full : sm.in_full, sm.scratch);
        if (j + 1 < p.count) proj_load(pw, p.w_t[j + 1], 128, rank * kS128);
        proj_reduce<kS128, kDim>(rank * kS128, p.b[j], sm.scratch, sm.p32);
        for (int i = n; i < kT * kS128; i += kThreads)
            p.out[j][((size_t)prompt * kT + i / kS128) * 128 + rank * kS128 + (i % kS128)] = sm.p32[i];
        __syncthreads();
    }
    // (no closing cluster barrier: the projections above touch local shared memory only)
}

// IoU head and the four hypernetwork MLPs (256 -> 256 -> 256 -> 4 | 32, ReLU between): a cluster owns one MLP for SEVEN
// PROMPTS (the 7 rows cta_proj works on are prompts here, token m of each), so the MLP's 0.5 MB of weights are pulled once
// per seven prompts, a quarter per CTA, one trip to L2 per layer.  (The earlier kernel -- a CTA per (prompt, MLP), a warp
// per output feature -- walked 32 dependent weight-row loads per layer: 34 us per 64-prompt pass.)
struct Mlp3Smem {
    float x[kT * kDim], y[kT * kDim];
    float o[kT * kS256];
    float scratch[kThreads * kT];
};

__global__ void __cluster_dims__(kCl, 1, 1) __launch_bounds__(kThreads) token_mlp3_kernel(float const* __restrict__ tokens, TokenMlp3 heads,
                                                                                       int P, float* __restrict__ hyper,
                                                                                       float* __restrict__ iou) {
    extern __shared__ __align__(16) uint8_t tok_smem[];
    Mlp3Smem& sm = *reinterpret_cast<Mlp3Smem*>(tok_smem);
    cg::cluster_group cluster = cg::this_cluster();
    int const rank = (int)cluster.block_rank();
    int const groups = (P + kT - 1) / kT;
    int const cl = blockIdx.x / kCl, m = cl / groups, p0 = (cl % groups) * kT;  // m = 0: IoU head (token 0), 1..4: mask token m
    int const n = threadIdx.x, n0 = rank * kS256;
    ProjW<kS256, kDim> pw;
    proj_load(pw, heads.w[m][0], kDim, n0);
    for (int i = n; i < kT * kDim; i += kThreads) {
        int const r = i / kDim;
        sm.x[i] = p0 + r < P ? tokens[((size_t)(p0 + r) * kTokens + m) * kDim + (i % kDim)] : 0.f;
    }
    __syncthreads();
    proj_fma(pw, sm.x, sm.scratch);
    proj_load(pw, heads.w[m][1], kDim, n0);  // under the reduction, the exchange and the cluster barrier
    proj_reduce<kS256, kDim>(n0, heads.b[m][0], sm.scratch, sm.o);
    for (int i = n; i < kT * kS256; i += kThreads) sm.o[i] = fmaxf(sm.o[i], 0.f);
    __syncthreads();
    cluster_scatter_slice(cluster, sm.o, sm.y);
    cluster.sync();
    proj_fma(pw, sm.y, sm.scratch);
    proj_reduce<kS256, kDim>(n0, heads.b[m][1], sm.scratch, sm.o);
    for (int i = n; i < kT * kS256; i += kThreads) sm.o[i] = fmaxf(sm.o[i], 0.f);
    __syncthreads();
    cluster_scatter_slice(cluster, sm.o, sm.x);  // every CTA is past its reads of x: they precede the barrier above
    cluster.sync();
    if (rank != 0) return;  // the narrow last layer is one CTA's work (nothing is written to the others any more)
    if (m == 0) {
        cta_proj<4, kDim>(sm.x, heads.w[0][2], 4, 0, heads.b[0][2], sm.scratch, sm.o);
        for (int i = n; i < kT * 4; i += kThreads)
            if (p0 + i / 4 < P) iou[(size_t)(p0 + i / 4) * 4 + (i % 4)] = sm.o[i];
    } else {
        cta_proj<32, kDim>(sm.x, heads.w[m][2], 32, 0, heads.b[m][2], sm.scratch, sm.o);
        for (int i = n; i < kT * 32; i += kThreads)
            if (p0 + i / 32 < P) hyper[((size_t)(p0 + i / 32) * 4 + (m - 1)) * 32 + (i % 32)] = sm.o[i];
    }
}

template <typename K, typename P> void launch_cluster(K kernel, size_t smem, cudaStream_t s, int prompts, P const& params) {
    static std::mutex mutex;
    static std::map<void const*, bool> done;
    {
        std::lock_guard<std::mutex> lock(mutex);
        if (!done[(void const*)kernel]) {
            CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            done[(void const*)kernel] = true;
        }
    }
    kernel<<<prompts * kCl, kThreads, smem, s>>>(params);  // cluster shape comes from __cluster_dims__
}

}  // namespace

void token_attn_block(cudaStream_t s, TokenAttnBlock const& p, int P) {
    ProfScope prof(s, CAT_DEC_LINEAR, 2.0 * P * kT * kDim * (4.0 * kDim + 128));
    launch_cluster(token_attn_block_kernel, sizeof(AttnSmem), s, P, p);
    KERNEL_CHECK();
}

void token_post_t2i(cudaStream_t s, TokenPostT2i const& p, int P) {
    ProfScope prof(s, CAT_DEC_LINEAR, 2.0 * P * kT * 128 * kDim);
    launch_cluster(token_post_t2i_kernel, sizeof(PostSmem), s, P, p);
    KERNEL_CHECK();
}

void token_post_mlp(cudaStream_t s, TokenPostMlp const& p, int P) {
    DLIMG_ASSERT(p.count >= 1 && p.count <= 3);
    ProfScope prof(s, CAT_DEC_LINEAR, 2.0 * P * kT * kDim * 128 * p.count);
    launch_cluster(token_post_mlp_kernel, sizeof(PostSmem), s, P, p);
    KERNEL_CHECK();
}

void token_mlp3(cudaStream_t s, float const* tokens, int P, TokenMlp3 const& heads, float* hyper, float* iou) {
    ProfScope prof(s, CAT_DEC_LINEAR, 2.0 * P * (4.0 * (2 * kDim * kDim + 32 * kDim) + 2 * kDim * kDim + 4 * kDim));
    static std::once_flag once;
    std::call_once(once, [] { CUDA_CHECK(cudaFuncSetAttribute(token_mlp3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Mlp3Smem))); });
    int const groups = (P + kT - 1) / kT;
    token_mlp3_kernel<<<5 * groups * kCl, kThreads, sizeof(Mlp3Smem), s>>>(tokens, heads, P, hyper, iou);
    KERNEL_CHECK();
}

}  // namespace dec
}  // namespace dlimg
