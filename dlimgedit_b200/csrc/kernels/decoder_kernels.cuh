// decoder_kernels.cuh -- CUDA-core kernels of the SAM prompt encoder + two-way mask decoder.
// The image-side stream (4096 rows per prompt) is stored in 16 bits (act_t) and its projections go through the
// tcgen05 GEMM (gemm.cuh); the token side (7 rows per prompt) stays fp32.  Everything here is token-side Linears,
// the attention cores, LayerNorms and glue.
#pragma once

#include "../common.hpp"

#include "../act.hpp"

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

// coords (P, 2, 2) already in 1024-space, labels (P, 2) -> tokens (P, 7, 256), written twice: `tokens` (the token-side
// positional encoding, kept) and `queries` (the transformer's running state).
void prompt_tokens(cudaStream_t s, float const* coords, float const* labels, int P, PromptParams const& pp, float* tokens,
                   float* queries);

// Dense positional encoding of the 64x64 grid -> (4096, 256) token-major.
void dense_pe(cudaStream_t s, float const* gaussian, float* pos);

// y[r, :] = act( (x[r, :] (+ x2[r, :])) @ W^T + b ).  Row r of x lives at x + r*x_stride (same for x2), of y at
// y + r*y_stride.  W is (N, K) row-major.  relu != 0 applies ReLU.  K <= 2048.
void linear_small(cudaStream_t s, float const* x, int64_t x_stride, float const* x2, int64_t x2_stride, int rows, int K,
                  float const* W, float const* b, int N, bool relu, float* y, int64_t y_stride);

// Token -> image attention core: q (P, 7, 128) fp32; K and V rows of the image stream in 16 bits: row i of prompt p at
// base_p + i * pitch (K) and base_p + v_off + i * pitch (V), where base_p = ptrs[p] (per-prompt tables, layer 0: the
// image's prompt-independent projections) or base + p * prompt_stride; 8 heads x 16 -> out (P, 7, 128) fp32.
// `scratch` holds P * kT2iSplits * 7 * (128 + 16) floats of partial results (split-key softmax merge).
constexpr int kT2iSplits = 8;
constexpr size_t kT2iScratchPerPrompt = (size_t)kT2iSplits * kTokens * (128 + 16);
// out == nullptr: leave the partials in `scratch` (token_post_t2i merges them).
void token_to_image_attention(cudaStream_t s, float const* q, act_t const* base, act_t const* const* ptrs, int64_t prompt_stride,
                              int pitch, int v_off, int P, float* scratch, float* out);

// Image -> token attention core: Q rows (16-bit, 128 wide) of prompt p at Qp + q_off + i * q_pitch, Qp = Qptrs[p] or
// Q + p * q_prompt_stride; kt, vt (P, 7, 128) fp32 -> out (P, 4096, 128) 16-bit.
#if DLIMG_B200_ALT  // CUDA-core cross-check form (development builds)
void image_to_token_attention(cudaStream_t s, act_t const* Q, act_t const* const* Qptrs, int64_t q_prompt_stride, int q_pitch,
                              int q_off, float const* kt, float const* vt, int P, act_t* out);
#endif

// The same on tensor cores (mma.sync, t2i_attention.cu); DLIMG_B200_I2T_SIMT selects the CUDA-core form above.
void image_to_token_attention_mma(cudaStream_t s, act_t const* Q, act_t const* const* Qptrs, int64_t q_prompt_stride, int q_pitch,
                                  int q_off, float const* kt, float const* vt, int P, act_t* out);

// fp32 -> 16-bit storage (load-time tables).
void f32_to_act(cudaStream_t s, float const* in, int64_t n, act_t* out);

// IoU head (slot 0, on token 0, 256 -> 256 -> 256 -> 4) and the four hypernetwork MLPs (slots 1..4, on mask tokens 1..4,
// 256 -> 256 -> 256 -> 32), ReLU between layers: tokens (P, 7, 256) -> iou (P, 4), hyper (P, 4, 32).
struct TokenMlp3 {  // weights as [K / 4][N][4] (Linear32T); decoder_tokens.cu
    float const* w[5][3];
    float const* b[5][3];
};
void token_mlp3(cudaStream_t s, float const* tokens, int P, TokenMlp3 const& heads, float* hyper, float* iou);

// ---- fused token-side blocks (decoder_tokens.cu): a 4-CTA cluster per prompt, weights as [K / 4][N][4] ----------------------
// [+pe] -> q, k, v -> 8-head self-attention over the 7 tokens -> out projection -> [+residual] LayerNorm -> queries (in
// place), then out_next (P, 7, 128) = (queries + pe) w_next^T + b_next: the query projection of tokens -> image attention.
struct TokenAttnBlock {
    float* queries;       // (P, 7, 256) in / out
    float const* pe;      // (P, 7, 256) the prompt tokens (positional encoding of the token side)
    float const *wq_t, *bq, *wk_t, *bk, *wv_t, *bv, *wo_t, *bo;  // (256, 256) weights as [K / 4][N][4], (256) biases
    float const *gamma, *beta;
    int with_pe, residual;  // layer 0: q = k = v = queries, output replaces them; later: q = k = queries + pe, residual
    float const *w_next_t, *b_next;  // (256, 128), (128)
    float* out_next;
};
void token_attn_block(cudaStream_t s, TokenAttnBlock const& p, int P);

// Merge of the split-key partials of token_to_image_attention (see `scratch`) -> out projection 128 -> 256 -> + queries ->
// LayerNorm -> queries (in place).
struct TokenPostT2i {
    float const* partials;  // P * kT2iScratchPerPrompt
    float* queries;
    float const *wo_t, *bo;  // (128, 256) transposed, (256)
    float const *gamma, *beta;
};
void token_post_t2i(cudaStream_t s, TokenPostT2i const& p, int P);

// queries <- LayerNorm(queries + mlp_bias + sum of the mlp_out partials); then `count` (<= 3) projections 256 -> 128 of the new queries (+ pe if
// with_pe[j]) -> out[j] (P, 7, 128).
struct TokenPostMlp {
    float* queries;
    float const* mlp_out;        // mlp_parts split-K partial sums of the MLP's second Linear, mlp_part_stride floats apart ...
    int mlp_parts;
    int64_t mlp_part_stride;
    float const* mlp_bias;       // ... and its bias (256)
    float const* pe;
    float const *gamma, *beta;
    int count;
    float const* w_t[3];  // (256, 128) transposed
    float const* b[3];
    int with_pe[3];
    float* out[3];
};
void token_post_mlp(cudaStream_t s, TokenPostMlp const& p, int P);

// Mask selection of the decoder graphs.  iou (P, 4).
//   multi == 0: plane_index[p] = p*4 + argmax(iou + (2 - 2.5) * [1000, 0, 0, 0]); iou_out[p] = iou of that token
//   multi != 0: plane_index[p*3 + i] = p*4 + 1 + i; iou_out[p*3 + i] = iou[p, 1 + i]
void select_masks(cudaStream_t s, float const* iou, int P, int multi, int* plane_index, float* iou_out);

}  // namespace dec
}  // namespace dlimg
