/* dlimg_b200_debug.h -- kernel-level entry points of libdlimgedit.so used ONLY by the parity tests
 * (tests/): they let a test drive one CUDA kernel (or one encoder stage) with its own device buffers and
 * compare against the oracle.  Not part of the supported interface; may change between rounds. */
#ifndef DLIMG_B200_DEBUG_H_
#define DLIMG_B200_DEBUG_H_

#include "dlimg_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dlimg_b200_Debug {
    uint32_t struct_size;
    uint32_t act_is_bf16; /* storage type of encoder activations / 16-bit GEMM operands: 0 = fp16, 1 = bf16 */
    /* C = epilogue(A[M,K] * B[N,K]^T).  Device pointers; elements are bf16 (tf32 == 0) or fp32 (tf32 != 0).
     * simt != 0 runs the CUDA-core cross-check kernel instead of the tcgen05 kernel.  act: 0 none, 1 GELU(erf),
     * 2 ReLU.  row_map (optional): output row per input row, -1 drops the row.  ln_stats / ln_colsum (optional):
     * folded LayerNorm; B must hold gamma-scaled weights with centred rows, then out = rstd * (A B^T) + bias (see
     * csrc/kernels/gemm.cuh): ln_parts == 0 -> per-row (mean, rstd) pairs [M][2]; ln_parts >= 1 -> [M][ln_parts]
     * partial (sum, sum of squares) of the rows of A.  stats_out (optional): [M][N / block_n][2], the same partial
     * sums of the rows this GEMM writes.  simt: 0 = tcgen05 kernel, 1 = CUDA-core cross-check kernel, k >= 2 = tcgen05
     * kernel with split-K over k parts: out is [k][M rounded up to 128][N] fp32 partial sums (no bias). */
    dlimg_Result (*gemm)(void* stream, int tf32, int simt, void const* a, void const* b, int M, int N, int K,
                         float const* bias, void const* residual, int const* row_map, int act, int out_f32, void* out,
                         float const* ln_stats, int ln_parts, float* stats_out);
    /* Encodes `count` device-resident images and copies the activation called `name` (see model.cu) as fp32. */
    dlimg_Result (*encode_tap)(dlimg_Environment, dlimg_ImageView const* dev_views, int count, char const* name,
                               float* dev_out, size_t capacity, size_t* written);
    /* Host-side resampling plan of the resize kernel: returns the tap count; first[out_size],
     * weights[out_size * max_taps]. */
    int (*resize_plan)(int in_size, int out_size, int max_taps, int* first, float* weights);
    void (*srgb_tables)(float* decode256, float* threshold256);
    /* Windowed attention on the un-partitioned token grid: qkv (batch*res*res, heads*96) 16-bit, ws x ws windows,
     * pad_qkv (heads*96) 16-bit = the qkv of a zero-padding position, bias (heads, ws^2, ws^2) fp32 dense ->
     * out (batch*res*res, heads*32) 16-bit. */
    dlimg_Result (*window_attention)(void* stream, void const* qkv, int batch, int res, int ws, int heads,
                                     void const* pad_qkv, float const* bias, void* out);
    /* CUDA-core cross-check of the attention core on partitioned windows: qkv (windows*n, heads*96) ->
     * out (windows*n, heads*32). */
    dlimg_Result (*window_attention_simt)(void* stream, void const* qkv, int windows, int n, int heads,
                                          float const* bias, void* out);
    /* Per-row LayerNorm statistics: in (rows, C) 16-bit -> out (rows, 2) fp32 (mean, rstd). */
    dlimg_Result (*layernorm_stats)(void* stream, void const* in, int rows, int C, float eps, float* out);
    /* Fused TinyViT MLP (C = 128 / 160): out = x + fc2(GELU(fc1(LN(x)))).  w1 (4C, C): gamma-scaled, row-centred fc1
     * weights; b1 (4C): folded bias; ln_sums (rows, 2): (sum, sum of squares) of x's rows; w2 (C, 4C); b2 (C);
     * stats_out (optional, rows x 2): the same sums of the output rows. */
    dlimg_Result (*mlp_fused)(void* stream, void const* x, int rows, int C, void const* w1, float const* b1,
                              float const* ln_sums, void const* w2, float const* b2, void* out, float* stats_out);
    /* 3x3 convolution (stride 1, zero padding 1) as an implicit GEMM: in (batch, H, W, C) 16-bit NHWC, weight (N, 9 * C)
     * 16-bit with K index = (ky * 3 + kx) * C + channel, bias (N) fp32, out (batch * H * W, N) 16-bit.
     * 128 % W == 0, H % (128 / W) == 0, C % 64 == 0, N % 16 == 0. */
    dlimg_Result (*conv3x3)(void* stream, void const* in, int batch, int H, int W, int C, void const* weight,
                            float const* bias, int N, void* out);
    /* TinyViT local_conv: depthwise 3x3 (stride 1, zero padding 1), fp32 accumulation, no activation.  in / out (batch, H,
     * W, C) 16-bit NHWC, weight (9, C) fp32 with tap = ky * 3 + kx, bias (C); stats (batch * H * W, parts, 2) fp32:
     * (sum, sum of squares) of the output pixel's channels (parts = 2 for C = 320 with tma != 0, else 1).
     * tma != 0: the TMA halo-tile kernel (H % 8 == 0, W % 16 == 0, C in {128, 160, 320}); 0: the register-tiled kernel. */
    dlimg_Result (*local_conv)(void* stream, void const* in, int batch, int H, int W, int C, float const* weight,
                               float const* bias, void* out, float* stats, int tma);
    /* Storage behind create_image / load_image / destroy_image (csrc/image_pool.hpp): out[0] page-locked bytes in use,
     * out[1] page-locked bytes cached for re-use, out[2] page-locked allocations made, out[3] re-uses of a cached block,
     * out[4] plain (pageable) allocations. */
    void (*image_pool_stats)(uint64_t* out5);
} dlimg_b200_Debug;

DLIMG_B200_EXPORT dlimg_b200_Debug const* dlimg_b200_debug_init(void);

#ifdef __cplusplus
}
#endif

#endif /* DLIMG_B200_DEBUG_H_ */
