// model.hpp -- MobileSAM on the device: weights (folded / re-laid-out at load), the TinyViT encoder
// launch sequence and the batched prompt decoder.  Takes the place of the reference's
// SegmentAnythingModel + three onnxruntime sessions (segmentation.hpp:17-32, session.cpp:57-136).
#pragma once

#include "common.hpp"

#include <cuda.h>
#include "kernels/decoder_kernels.cuh"
#include "kernels/encoder_kernels.cuh"
#include "weights.hpp"

#include "act.hpp"

#include <string>
#include <vector>

namespace dlimg {


template <typename T> class DeviceBuffer {
  public:
    DeviceBuffer() = default;
    explicit DeviceBuffer(size_t n) { allocate(n); }
    DeviceBuffer(DeviceBuffer const&) = delete;
    DeviceBuffer& operator=(DeviceBuffer const&) = delete;
    DeviceBuffer(DeviceBuffer&& o) noexcept : ptr_(o.ptr_), n_(o.n_) { o.ptr_ = nullptr; o.n_ = 0; }
    DeviceBuffer& operator=(DeviceBuffer&& o) noexcept {
        if (this != &o) { release(); ptr_ = o.ptr_; n_ = o.n_; o.ptr_ = nullptr; o.n_ = 0; }
        return *this;
    }
    ~DeviceBuffer() { release(); }

    void allocate(size_t n) {
        release();
        if (n) CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&ptr_), n * sizeof(T)));
        n_ = n;
    }
    void release() {
        if (ptr_) cudaFree(ptr_);
        ptr_ = nullptr;
        n_ = 0;
    }
    void upload(std::vector<T> const& host) {
        allocate(host.size());
        if (!host.empty()) CUDA_CHECK(cudaMemcpy(ptr_, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
    }
    T* get() const { return ptr_; }
    size_t size() const { return n_; }
    explicit operator bool() const { return ptr_ != nullptr; }

  private:
    T* ptr_ = nullptr;
    size_t n_ = 0;
};

// ---- weights ----------------------------------------------------------------------------------
struct Linear16 {  // act_t weight (N, K) row-major + fp32 bias (N) -- encoder GEMM operand B
    DeviceBuffer<act_t> w;
    DeviceBuffer<float> b;
    bool ln_folded = false;  // weights hold W * gamma with centred rows, bias holds b + W beta (LayerNorm folded in)
    int n = 0, k = 0;
};
struct Linear32 {  // fp32 weight (N, K) + bias -- decoder
    DeviceBuffer<float> w;
    DeviceBuffer<float> b;
    int n = 0, k = 0;
};
struct DwConv {  // depthwise 3x3: weight (9, C) fp32 (+ the same in act_t), bias (C)
    DeviceBuffer<float> w, b;
    DeviceBuffer<act_t> w16;
    int c = 0;
};
struct Norm {
    DeviceBuffer<float> g, b;
};

struct MBConvW { Linear16 conv1; DwConv conv2; Linear16 conv3; };
struct MergeW { Linear16 conv1; DwConv conv2; Linear16 conv3; int stride = 2; };
struct BlockW {
    // attn.norm and mlp.norm are folded into qkv and fc1 (gamma into the weights, beta into the bias; the per-row
    // mean / rstd are applied in the GEMM epilogue), see gemm.cuh Epilogue::ln_stats
    Linear16 qkv, proj;
    DeviceBuffer<act_t> qkv_pad;        // (3C): qkv of a zero-padding token = projection of LN(0) = beta
    DeviceBuffer<uint16_t> attn_bias;   // fp16 fragment-ordered relative-position bias (attention_bias_fragments)
    DwConv local_conv;
    Linear16 fc1, fc2;
};
struct StageCfg { int dim, res, depth, heads, ws; };

struct EncoderW {
    DeviceBuffer<float> conv1_w, conv1_b;  // (27, 32), (32)
    Linear16 conv2;                        // (64, 288) K = (ky, kx, ci)
    // fused PatchEmbed kernel: conv1 as mma fragments, conv2 K-padded to 320 with its TMA descriptor
    DeviceBuffer<uint32_t> conv1_frag;
    DeviceBuffer<act_t> conv2_k320;
    CUtensorMap conv2_map;
    MBConvW mb[2];
    MergeW merge[3];
    std::vector<BlockW> blocks[3];
    Linear16 neck1;  // (256, 320)
    Norm neck_ln1;
    Linear16 neck2;  // (256, 2304) K = (ky, kx, ci)
    Norm neck_ln2;
};

constexpr int kMlpSplit = 8;  // k-splits of the decoder token MLP's 2048 -> 256 Linear (gemm.cuh Epilogue::ksplit)

struct Linear32T {  // fp32 weight transposed and k-blocked, [K / 4][N][4], + bias -- the fused token-side kernels (decoder_tokens.cu)
    DeviceBuffer<float> wt;
    DeviceBuffer<float> b;
    int n = 0, k = 0;
};
struct AttnW { Linear32T q, k, v, o; };  // token-side projections (fp32); the image-side ones live in DecoderW below
struct DecLayerW {
    AttnW self_attn;
    Linear32T t2i_q, t2i_o;  // tokens -> image attention: query and output projections (token rows)
    Linear32T i2t_k, i2t_v;  // image -> tokens attention: key / value projections (token rows)
    Linear16 i2t_out;        // (256, 128) output projection of image -> tokens attention (image rows, 16-bit)
    Norm n1, n2, n3, n4;
    Linear32 lin1, lin2;
};
struct DecoderW {
    DeviceBuffer<float> gaussian, point_embed, not_a_point, no_mask, iou_token, mask_tokens;
    DecLayerW layers[2];
    Linear32T final_q, final_o;
    Norm norm_final;
    // Image-side projections as 16-bit GEMM operands.  The projections that see `keys + pos` (attention keys of
    // tokens -> image, queries of image -> tokens) are split as keys W^T + (pos W^T): the second term does not depend on the
    // image or the prompt and is a (4096, N) table added by the GEMM epilogue (Epilogue::res_mod), so `keys + pos`
    // is never materialised.
    Linear16 kvq[2];                // (384, 256): [t2i.k | t2i.v | i2t.q] of layer 0 / 1, concatenated along N
    DeviceBuffer<act_t> pos_kvq[2]; // (4096, 384): [pos Wk^T | 0 | pos Wq^T]
    Linear16 kv_final;              // (256, 256): [k | v] of the final tokens -> image attention
    DeviceBuffer<act_t> pos_kv_final;  // (4096, 256): [pos Wk^T | 0]
    Linear16 up1;  // (256 = (dy,dx,co), 256)
    Norm up_ln;    // LayerNorm2d(64)
    Linear16 up2;  // (128 = (ey,ex,c2), 64)
    Linear32T hyper[4][3];
    Linear32T iou[3];
};

// ---- workspaces -------------------------------------------------------------------------------
struct EncoderWorkspace {
    int max_batch = 0;
    DeviceBuffer<act_t> c1, col;       // development builds / the conv1 debug tap only
    DeviceBuffer<act_t> xa, xb, big[3];
    DeviceBuffer<float2> stats;    // (max_batch * 16384): per-token LayerNorm (mean, rstd) of the current block input
    DeviceBuffer<float2> stats_parts;  // (max_batch * 16384 * 2): partial (sum, sum of squares) written by fc2's epilogue
    explicit EncoderWorkspace(int max_batch);
};

// Per-prompt inputs of one decoder pass, uploaded by the engine in ONE copy: prompt p of the pass reads the
// prompt-independent tensors of ITS image through keys0[p] / kvq0[p], so prompts of different images share a pass.
struct DecoderParams {
    float* coords = nullptr;               // (P, 2, 2) in 1024-space
    float* labels = nullptr;               // (P, 2)
    act_t const** keys0 = nullptr;         // (P) -> (4096, 256): embedding + no_mask_embed of the prompt's image
    act_t const** kvq0 = nullptr;          // (P) -> (4096, 384): its layer-0 [K | V | Q] projections
    static size_t bytes(int P) { return (size_t)P * (6 * sizeof(float) + 2 * sizeof(void*)); }
};

struct DecoderWorkspace {
    int max_prompts = 0;
    DeviceBuffer<uint8_t> param_block;                       // DecoderParams::bytes(max_prompts), laid out per pass by layout()
    DeviceBuffer<float> tok0, queries;                       // (P,7,256)
    DeviceBuffer<float> tmp;                                 // split-K partials of the token MLP output
    DeviceBuffer<float> t128a, t128b, t128c;                 // (P,7,128)
    DeviceBuffer<float> hid;                                 // (P,7,2048)
    DeviceBuffer<float> hyper, iou;                          // (P,4,32), (P,4)
    DeviceBuffer<act_t> keys, big;                           // (P,4096,256) image stream / GEMM output in front of a LayerNorm
    DeviceBuffer<act_t> kvq;                                 // (P,4096,384) [K | V | Q] of layer 1; (P,4096,256) [K | V] final
    DeviceBuffer<act_t> ao;                                  // (P,4096,128) image -> tokens attention output
    DeviceBuffer<float> low;                                 // (P,4,256,256)
    DeviceBuffer<int> plane_index;                           // (P*3)
    DeviceBuffer<float> iou_sel;                             // (P*3)
    DeviceBuffer<float> t2i_scratch;                         // split-key partials of the token->image attention
    explicit DecoderWorkspace(int max_prompts);
    // The parameter block of a pass with P prompts, based at `base` (the device block, or a host staging copy of it):
    // [coords P*4 floats | labels P*2 floats | keys0 P pointers | kvq0 P pointers], DecoderParams::bytes(P) in all.
    static DecoderParams layout(uint8_t* base, int P);
};

struct Tap {  // debug: copy a named activation as fp32 to a device buffer
    char const* name = nullptr;
    float* out = nullptr;
    size_t capacity = 0;  // floats
    size_t written = 0;
};

class SamModel {
  public:
    SamModel(std::string const& weight_path, int num_sms);

    // images: `batch` device descriptors of u8 images with identical (w, h, channels), w, h <= 1024.
    // Outputs per image: emb_nchw_out (batch, 256, 64, 64) fp32 (the reference's `image_embeddings`), keys0_out (batch, 4096,
    // 256) 16-bit (embedding + no_mask_embed, token-major) and kvq0_out (batch, 4096, 384) 16-bit (the decoder's layer-0
    // image-side projections, which depend on the image only).  finish = false stops in front of the final LayerNorm2d
    // (its input is ws.big[0]); neck_finish() then writes the outputs -- the engine captures the trunk into a CUDA graph
    // and finishes eagerly into whichever store the call owns.
    void encode(cudaStream_t s, EncoderWorkspace& ws, enc::ImageDesc const* images, int batch, int w, int h, int channels,
                float* emb_nchw_out, act_t* keys0_out, act_t* kvq0_out, Tap* tap = nullptr, bool finish = true) const;
    void neck_finish(cudaStream_t s, EncoderWorkspace& ws, int batch, float* emb_nchw_out, act_t* keys0_out, act_t* kvq0_out) const;

    // Runs the prompt encoder + mask decoder for P prompts whose parameter block (DecoderWorkspace::layout(ws.param_block, P))
    // has been filled by the engine.  Results: ws.low (P, 4, 256, 256) logits and ws.iou (P, 4).
    // mask_mode (kernels/mask_select.cuh MaskMode): which planes of ws.low the pass writes -- all four, masks 1..3, or only
    // the plane the predicted IoUs select (single-mask calls: a quarter of the hypernetwork products and logit writes)
    void decode(cudaStream_t s, DecoderWorkspace& ws, int P, int mask_mode = 0) const;

    static StageCfg stage(int i);  // i = 1..3

  private:
    void gemm16(cudaStream_t s, act_t const* a, int64_t rows, Linear16 const& l, void* out, int act, act_t const* residual,
                float2 const* ln_stats = nullptr, bool out_f32 = false, int ln_parts = 0, float2* stats_out = nullptr,
                int res_mod = 0) const;
    void gemm32(cudaStream_t s, float const* a, int64_t rows, Linear32 const& l, float* out, int act) const;
    void lin(cudaStream_t s, float const* x, int64_t xs, float const* x2, int rows, Linear32 const& l, bool relu, float* y,
             int64_t ys) const;

    EncoderW enc_;
    DecoderW dec_;
    int num_sms_ = 148;
};

}  // namespace dlimg
