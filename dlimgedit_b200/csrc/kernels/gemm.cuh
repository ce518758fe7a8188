// gemm.cuh -- the tensor-core workhorse: C[M,N] = epilogue(A[M,K] * B[N,K]^T).
//
// Both operands are K-major (activations: rows = tokens; weights: rows = output features, exactly the
// nn.Linear / 1x1-conv layout).  The kernel is a persistent, warp-specialised sm_100a pipeline:
//   warp 0      TMA producer   (cp.async.bulk.tensor 2D, 128-byte swizzle, mbarrier ring as deep as fits)
//   warp 1      MMA issuer     (tcgen05.mma cta_group::1, M=128, N=block_n, fp32 accumulators in TMEM,
//                               two accumulator stages so the epilogue of tile i overlaps tile i+1)
//   warps 2..17 epilogue       (tcgen05.ld -> bias / residual / activation -> bf16|fp32 global stores,
//                               optional row scatter used for window un-partition)
// Operand element type is bf16 (kind::f16) for the encoder or fp32-as-tf32 (kind::tf32) for the decoder;
// the shared-memory geometry is identical in bytes (128 B of K per row per stage).
#pragma once

#include "../common.hpp"

#include <cuda.h>
#include "../act.hpp"

namespace dlimg {
namespace gemm {

enum Act : int { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2 };

struct Epilogue {
    float const* bias = nullptr;     // [N] fp32, or null
    void const* residual = nullptr;  // same element type / pitch as the output, indexed by OUTPUT row
    int res_mod = 0;                 // != 0 (16-bit staged epilogue only): residual row = output row % res_mod
    int const* row_map = nullptr;    // [M] -> output row (-1 = drop the row), or null for identity
    int act = ACT_NONE;              // applied after bias and residual
    int out_f32 = 0;                 // 0: bf16 output, 1: fp32 output
    int ldc = 0;                     // output row pitch in elements (0 = N)
    // LayerNorm folded into the GEMM (16-bit outputs without residual).  The caller passes the raw, un-normalised rows
    // as A, W'' = W * gamma with every ROW of W'' centred (W''_nk -= mean_k W''_nk) as B and b' = b + W beta as bias:
    //   LN(x) W^T + b = rstd_m * sum_k (x_mk - mean_m) (W gamma)_nk + b'_n = rstd_m * (x W''^T)_mn + b'_n
    // because sum_k x_mk = K mean_m cancels the centring term -- so the epilogue needs only 1/std per row.
    float2 const* ln_stats = nullptr;  // ln_parts == 0: [M] (mean, rstd) per row of A (only rstd is read)
    // ln_parts >= 1: ln_stats is [M][ln_parts] partial (sum x, sum x^2) written by the GEMM that produced A
    // (stats_out below); 1/std = rsqrt(sum x^2 / K - (sum x / K)^2 + ln_eps) is formed in the epilogue.
    int ln_parts = 0;
    float ln_eps = 1e-5f;
    // Direct epilogue without activation: also write, per output row and N tile, (sum v, sum v^2) of the row's values
    // in this tile -> [M][N / block_n] float2.  Summed in a fixed order (no atomics): results are reproducible.
    float2* stats_out = nullptr;
    // Fused tails of the mask decoder's upscaling (16-bit GELU epilogue; see epilogue_slabs in gemm.cu):
    //   fuse = 1, N == 256: LayerNorm2d(64, eps 1e-6) over every 64-column group (gamma = fuse_a, beta = fuse_b) before the GELU
    //   fuse = 2, N == 128, M = prompts * 16384: the GELU'd 32-channel groups are dotted with the prompt's hypernetwork
    //             vectors fuse_a (prompts, 4, 32) and only the mask logits fuse_out (prompts, 4, 256, 256) are written.
    //             fuse_mode (mask_select.cuh MaskMode): all four planes | planes 1..3 | only the plane the predicted IoUs
    //             fuse_b (prompts, 4) select -- the other planes of fuse_out are then left untouched
    //   fuse = 3, N == 256, with a residual: out = LayerNorm(acc + bias + residual) over the whole row (gamma = fuse_a, beta =
    //             fuse_b, eps = ln_eps).  The residual of output rows [g * res_mod, (g + 1) * res_mod) may come from its own
    //             base pointer res_table[g] (device array) instead of `residual`
    int fuse = 0;
    int fuse_mode = 0;
    // Split-K for skinny problems (the decoder's 7-row-per-prompt 2048 -> 256 Linear: 16 tiles with a 64-step k-loop each):
    // split s of ksplit sums k in [s K / ksplit, (s + 1) K / ksplit) into rows [s * Mpad, s * Mpad + M) of `out`, Mpad = M
    // rounded up to 128 (split_rows()); the CONSUMER adds the ksplit partial sums in a fixed order.  fp32 outputs only.
    int ksplit = 1;
    void const* const* res_table = nullptr;
    float const* fuse_a = nullptr;
    float const* fuse_b = nullptr;
    float* fuse_out = nullptr;
};

struct Operand {
    void const* ptr = nullptr;
    int64_t rows = 0;   // M for A, N for B
    int64_t cols = 0;   // K
    int64_t pitch = 0;  // row pitch in elements (0 = cols)
};

constexpr int kBlockM = 128;
constexpr int kMaxBlockN = 256;
constexpr int kKBytes = 128;  // bytes of K per row per stage == swizzle span

// Memoised TMA descriptor of a K-major operand: box = box_rows rows x 128 bytes of K, 128-byte swizzle.
CUtensorMap make_tensor_map(Operand const& op, bool tf32, int box_rows);

// TMA descriptor of a 16-bit NHWC activation tensor for halo-tile loads: box = box_h x box_w pixels x box_c channels
// of one image, no swizzle, out-of-bounds elements (the convolution's zero padding) read as zero.
// swizzle128 (box_c == 64): every pixel is one 128-byte row of a SWIZZLE_128B K-major operand tile (implicit GEMM).
CUtensorMap make_tensor_map_nhwc(void const* ptr, int batch, int H, int W, int C, int box_c, int box_w, int box_h,
                                 bool swizzle128 = false);

// Picks the widest legal tile width (multiple of 16, <= 256) that divides N.
int pick_block_n(int N);

// Encodes the two tensor maps and launches.  tf32 != 0 selects fp32 operands with kind::tf32.
void launch(cudaStream_t stream, bool tf32, Operand const& a, Operand const& b, void* out, Epilogue const& epi,
            int num_sms);

// 3x3 convolution (stride 1, zero padding 1) of a 16-bit NHWC tensor as an implicit GEMM: in (batch, H, W, C), b (N, 9 * C)
// with K index = tap * C + channel (tap = ky * 3 + kx), out (batch * H * W, N).  128 % W == 0, H % (128 / W) == 0,
// C % 64 == 0.  Same epilogues as launch().
void launch_conv3x3(cudaStream_t stream, void const* in, int batch, int H, int W, int C, Operand const& b, void* out,
                    Epilogue const& epi, int num_sms);

// Fused TinyViT MLP for C = 128 / 160: out = x + fc2(GELU(fc1(LN(x)))) with the hidden activation kept in TMEM / shared
// memory.  x (rows, C) 16-bit (also the residual; out may alias x), w1 (4C, C) with the LayerNorm folded in (gamma-scaled,
// row-centred), b1 (4C) folded bias, ln_stats (rows) partial (sum, sum of squares) of x's rows, w2 (C, 4C), b2 (C);
// stats_out (optional, rows): (sum, sum of squares) of the output rows.
inline int64_t split_rows(int64_t M) { return (M + 127) / 128 * 128; }  // rows per split-K partial
bool mlp_fused_supported(int C);
void launch_mlp_fused(cudaStream_t stream, void const* x, int64_t rows, int C, void const* w1, float const* b1,
                      float2 const* ln_stats, float ln_eps, void const* w2, float const* b2, void* out, float2* stats_out,
                      int num_sms);

// Plain CUDA-core GEMM with the same contract; used by tests to cross-check the tensor-core kernel and
// for tiny problems (M < 64) where a 128-row tile would be mostly padding.
void launch_simt(cudaStream_t stream, bool f32_operands, Operand const& a, Operand const& b, void* out,
                 Epilogue const& epi);

}  // namespace gemm
}  // namespace dlimg
