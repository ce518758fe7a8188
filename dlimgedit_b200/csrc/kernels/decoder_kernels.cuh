// decoder_kernels.cuh -- fp32 CUDA-core kernels of the SAM prompt encoder + two-way mask decoder.
// The large image-side projections (4096 rows per prompt) go through the tf32 tcgen05 GEMM (gemm.cuh);
// everything here is token-side (7 rows per prompt), attention cores, LayerNorms and glue.
#pragma once

#include "../common.hpp"

namespace dlimg {
namespace dec {

constexpr int kTokens = 7;          // iou token + 4 mask tokens + 2 prompt points
constexpr int kDim = 256;
constexpr int kImgTokens = 4096;    // 64 x 64
constexpr int kHeads = 8;

struct PromptParams {
    float const* gaussian;       // (2, 128) positional_encoding_gaussian_matrix
    float const* point_embed;    // (4, 256) point_embeddings[0..3]
    float const* not_a_point;    // (256)
    float const* iou_token;      // (256)
    float const* mask_tokens;    // (4, 256)
};

// coords (P, 2, 2) already in 1024-space, labels (P, 2) -> tokens (P, 7, 256).
void prompt_tokens(cudaStream_t s, float const* coords, float const* labels, int P, PromptParams const& pp, float* tokens);

// Dense positional encoding of the 64x64 grid -> (4096, 256) token-major.
void dense_pe(cudaStream_t s, float const* gaussian, float* pos);

// keys0 = emb + no_mask_embed ; kpe0 = keys0 + pos.  rows = batch * 4096 (pos repeats per image).
void embed_prepare(cudaStream_t s, float const* emb, float const* no_mask, float const* pos, int64_t rows, float* keys0,
                   float* kpe0);

// y[r, :] = act( (x[r, :] (+ x2[r, :])) @ W^T + b ).  Row r of x lives at x + r*x_stride (same for x2), of y at
// y + r*y_stride.  W is (N, K) row-major.  relu != 0 applies ReLU.  K <= 2048.
void linear_small(cudaStream_t s, float const* x, int64_t x_stride, float const* x2, int64_t x2_stride, int rows, int K,
                  float const* W, float const* b, int N, bool relu, float* y, int64_t y_stride);

// Token self-attention: q, k, v (P, 7, 256) already projected; 8 heads x 32 -> out (P, 7, 256).
void token_self_attention(cudaStream_t s, float const* q, float const* k, float const* v, int P, float* out);

// Token -> image attention core: q (P, 7, 128); K, V (4096, 128) per prompt at stride kv_stride floats
// (0 = shared by all prompts); 8 heads x 16 -> out (P, 7, 128).
// `scratch` holds P * kT2iSplits * 7 * (128 + 16) floats of partial results (split-key softmax merge).
constexpr int kT2iSplits = 4;
constexpr size_t kT2iScratchPerPrompt = (size_t)kT2iSplits * kTokens * (128 + 16);
void token_to_image_attention(cudaStream_t s, float const* q, float const* K, float const* V, int64_t kv_stride, int P,
                              float* scratch, float* out);
// Older two-pass formulation of the same op, kept as a cross-check for tests.
void token_to_image_attention_twopass(cudaStream_t s, float const* q, float const* K, float const* V, int64_t kv_stride,
                                      int P, float* out);

// Image -> token attention core: Q (4096, 128) per prompt at stride q_stride (0 = shared); kt, vt (P, 7, 128)
// -> out (P, 4096, 128).
void image_to_token_attention(cudaStream_t s, float const* Q, int64_t q_stride, float const* kt, float const* vt, int P,
                              float* out);

// out = LayerNorm_256(x + res) (eps 1e-5); optionally out2 = out + pos.  res row = row % res_mod, pos row =
// row % pos_mod.  res / pos / out2 may be null.  In-place (out == x) is allowed.
void layernorm256(cudaStream_t s, float const* x, float const* res, int64_t res_mod, int64_t rows, float const* gamma,
                  float const* beta, float const* pos, int64_t pos_mod, float* out, float* out2);

// In-place LayerNorm2d over groups of 64 channels (eps 1e-6) followed by exact GELU; rows of 64 floats.
void layernorm64_gelu(cudaStream_t s, float* x, int64_t rows, float const* gamma, float const* beta);

// low[p, m, Y, X] = sum_c hyper[p, m, c] * up2[p, blocked(Y, X), c]  (m = 0..3), where up2 is the blocked
// output of the two transposed convolutions: row ((y*64+x)*4 + dy*2+dx), col (ey*2+ex)*32 + c, with
// Y = 4y + 2dy + ey, X = 4x + 2dx + ex.
void mask_dot(cudaStream_t s, float const* hyper, float const* up2, int P, float* low);

// IoU head (slot 0, on token 0, 256 -> 256 -> 256 -> 4) and the four hypernetwork MLPs (slots 1..4, on mask tokens 1..4,
// 256 -> 256 -> 256 -> 32), ReLU between layers: tokens (P, 7, 256) -> iou (P, 4), hyper (P, 4, 32).
struct TokenMlp3 {
    float const* w[5][3];
    float const* b[5][3];
};
void token_mlp3(cudaStream_t s, float const* tokens, int P, TokenMlp3 const& heads, float* hyper, float* iou);

// Mask selection of the decoder graphs.  iou (P, 4).
//   multi == 0: plane_index[p] = p*4 + argmax(iou + (2 - 2.5) * [1000, 0, 0, 0]); iou_out[p] = iou of that token
//   multi != 0: plane_index[p*3 + i] = p*4 + 1 + i; iou_out[p*3 + i] = iou[p, 1 + i]
void select_masks(cudaStream_t s, float const* iou, int P, int multi, int* plane_index, float* iou_out);

}  // namespace dec
}  // namespace dlimg
