// encoder_kernels.cuh -- non-GEMM kernels of the TinyViT encoder.  Activations are bf16, NHWC
// ("token-major": row = pixel, contiguous channels), so every 1x1 conv / Linear is a plain K-major GEMM.
#pragma once

#include "../common.hpp"

#include "../act.hpp"

#include <cuda.h>

namespace dlimg {
namespace enc {

using dlimg::act_t;

// One input image as the preprocessing kernels see it (device memory).
struct ImageDesc {
    uint8_t const* pixels;  // device pointer, row-major u8
    int stride;             // bytes per row
    int _pad;
};

// Fused: channel-map + (x-mean)/std + zero-pad to 1024^2 (encoder graph preamble) +
// PatchEmbed conv 3x3 s2 p1 (3->32, BN folded) + GELU.  Reads the u8 image once, writes bf16
// (B, 512, 512, 32).  `w`,`h` is the valid (already <= 1024) extent; weight layout [27][32] (tap-major,
// tap = (ky*3+kx)*3+ci), bias [32].
#if DLIMG_B200_ALT  // development / bf16 builds only (common.hpp)
void conv1_preprocess(cudaStream_t s, ImageDesc const* imgs, int batch, int w, int h, int channels,
                      float const* weight, float const* bias, act_t* out);
#endif

// The whole PatchEmbed in one kernel (patch_embed.cu): preprocess + conv1 + GELU + conv2 -> out (B, 256, 256, 64).
// w1_frag: conv1 weights from patch_embed_w1_fragments(); w2_map: TMA descriptor of the conv2 weights as a K-major
// (64, 320) matrix, K = (ky, kx, ci) zero-padded from 288, box 64 rows.  c1_debug (optional): also writes the
// (B, 512, 512, 32) conv1 activation (debug tap).
void patch_embed(cudaStream_t s, ImageDesc const* imgs, int batch, int w, int h, int channels, uint32_t const* w1_frag,
                 float const* b1, CUtensorMap const& w2_map, float const* b2, act_t* out, act_t* c1_debug, int num_sms);
// (27, 32) fp32 conv1 weights [(ky*3+kx)*3+ci][oc] -> 512 packed fp16 pairs in mma B-fragment order.
void patch_embed_w1_fragments(float const* w27x32, uint32_t* out512);

// Second half of an MBConv block in one kernel (mbconv_tail.cu): depthwise 3x3 + GELU on the (B, 256, 256, 256) expanded
// tensor (given as a TMA descriptor from gemm::make_tensor_map_nhwc, box mbconv_tail_unit_channels() ch x 18 x 10), then 1x1 conv 256 -> 64 +
// shortcut + GELU.  w3_map: K-major (64, 256) project weights, box 64 rows.
int mbconv_tail_unit_channels();  // channels of the TMA box of `expanded_map`: 64 (128 in a development A/B run)
void mbconv_tail(cudaStream_t s, CUtensorMap const& expanded_map, int batch, act_t const* dw_w16, float const* dw_b,
                 CUtensorMap const& w3_map, float const* b3, act_t const* shortcut, act_t* out, int num_sms);

// im2col for 3x3 / pad 1 convolutions on NHWC bf16: out[(b,oy,ox)][(ky,kx,c)] (K = 9*C).
#if DLIMG_B200_ALT
void im2col3x3(cudaStream_t s, act_t const* in, int batch, int H, int W, int C, int stride, act_t* out);
#endif

// Depthwise 3x3 / pad 1, NHWC 16-bit, fp32 weights [9][C] + bias [C] (BN folded), optional GELU.  weight16 (optional):
// the same filter in act_t; with fp16 storage the GELU'd convolutions then run in packed-half arithmetic.
void dwconv3x3(cudaStream_t s, act_t const* in, int batch, int H, int W, int C, int stride, float const* weight,
               act_t const* weight16, float const* bias, bool gelu, act_t* out);

// Stride-1 depthwise 3x3 with fp32 accumulation and no activation (TinyViT local_conv) that also writes, per output
// pixel, (sum, sum of squares) over its C channels: the LayerNorm row sums for the GEMM behind it.
void dwconv3x3_stats(cudaStream_t s, act_t const* in, int batch, int H, int W, int C, float const* weight, float const* bias,
                     act_t* out, float2* stats);

// The same operation from TMA-loaded halo tiles (local_conv.cu): persistent CTAs, 8 x 16-pixel tiles, C = 128 / 160 / 320
// (320 as two channel parts: stats then holds local_conv_parts(C) partial sums per pixel, [rows][parts]).  Output
// bit-identical to dwconv3x3_stats.  H % 8 == 0, W % 16 == 0.
bool local_conv_tma_supported(int H, int W, int C);
int local_conv_parts(int C);
void local_conv_tma(cudaStream_t s, act_t const* in, int batch, int H, int W, int C, float const* weight, float const* bias,
                    act_t* out, float2* stats, int num_sms);

// Final LayerNorm2d of the neck over C = 256 channels of (batch, tokens, 256) 16-bit rows, written twice: fp32 NCHW
// (batch, 256, tokens) -- the reference's `image_embeddings` -- and 16-bit token-major (batch, tokens, 256) with
// no_mask (256) added -- the decoder's layer-0 image stream.  tokens % 32 == 0.
void layernorm256_tokens_nchw(cudaStream_t s, act_t const* in, int batch, int tokens, float const* gamma, float const* beta,
                              float eps, float const* no_mask, act_t* out_keys, float* out_nchw);

// Row LayerNorm over C channels.  src_row (optional, length `rows`): gather index into `in`, -1 = the row
// is window padding and the output is LN(0) = beta.  Output bf16 or fp32.
void layernorm_rows(cudaStream_t s, act_t const* in, int rows, int C, int const* src_row, float const* gamma,
                    float const* beta, float eps, void* out, bool out_f32);

// Per-row LayerNorm statistics (mean, 1/sqrt(var + eps)) of (rows, C) 16-bit activations.  The normalisation itself
// is folded into the GEMM that consumes the row (gemm.cuh, Epilogue::ln_stats), so no normalised copy is written.
void layernorm_stats(cudaStream_t s, act_t const* in, int rows, int C, float eps, float2* out);

// Windowed multi-head attention, head_dim 32, on the un-partitioned token grid.  qkv: (batch*res*res, heads*96) with
// per-head [q|k|v]; out: (batch*res*res, heads*32).  Windows are ws x ws over the grid zero-padded to a multiple of
// ws; padded positions use pad_qkv (heads*96), the projection of LN(0).  bias_frag: relative-position bias from
// attention_bias_fragments().
void window_attention(cudaStream_t s, act_t const* qkv, int batch, int res, int ws, int heads, act_t const* pad_qkv,
                      uint16_t const* bias_frag, act_t* out, int num_sms);
// Number of fp16 values of the fragment-ordered bias table.
size_t attention_bias_fragment_count(int heads, int ws);
// dense (heads, n, n) fp32 -> fp16 [head][query tile (16)][key block (8)][lane (32)][4], values scaled by log2(e); for
// lane = 4*g + t the four values are (row g, col 2t), (g, 2t+1), (g+8, 2t), (g+8, 2t+1) of the 16 x 8 block; key columns
// beyond n hold -inf (which masks them in the softmax), query rows beyond n hold 0.
void attention_bias_fragments(float const* dense, int heads, int ws, uint16_t* out);
// CUDA-core reference of the attention core on already partitioned windows, dense (heads, n, n) fp32 bias; qkv
// (windows*n, heads*96) -> out (windows*n, heads*32).  Cross-check in tests.
void window_attention_simt(cudaStream_t s, act_t const* qkv, int windows, int n, int heads, float const* bias, act_t* out);

// (tokens, C) fp32 -> (C, tokens) fp32 per image: the reference's NCHW `image_embeddings` layout.
void tokens_to_nchw(cudaStream_t s, float const* in, int batch, int tokens, int C, float* out);

// bf16 -> fp32 copy (debug taps).
void act_to_f32(cudaStream_t s, act_t const* in, int64_t n, float* out);

}  // namespace enc
}  // namespace dlimg
