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

struct AttnW { Linear32 q, k, v, o; };
struct DecLayerW {
    AttnW self_attn, t2i, i2t;
    Norm n1, n2, n3, n4;
    Linear32 lin1, lin2;
};
struct DecoderW {
    DeviceBuffer<float> gaussian, point_embed, not_a_point, no_mask, iou_token, mask_tokens;
    DecLayerW layers[2];
    AttnW final_attn;
    Norm norm_final;
    Linear32 up1;  // (256 = (dy,dx,co), 256)
    Norm up_ln;    // LayerNorm2d(64)
    Linear32 up2;  // (128 = (ey,ex,c2), 64)
    Linear32 hyper[4][3];
    Linear32 iou[3];
    DeviceBuffer<float> dense_pe;  // (4096, 256)
};

// ---- workspaces -------------------------------------------------------------------------------
struct EncoderWorkspace {
    int max_batch = 0;
    DeviceBuffer<act_t> c1, col, xa, xb, big[4];
    DeviceBuffer<float2> stats;    // (max_batch * 16384): per-token LayerNorm (mean, rstd) of the current block input
    DeviceBuffer<float2> stats_parts;  // (max_batch * 16384 * 2): partial (sum, sum of squares) written by fc2's epilogue
    explicit EncoderWorkspace(int max_batch);
};

struct DecoderWorkspace {
    int max_prompts = 0;
    DeviceBuffer<float> coords, labels;                      // (P,2,2), (P,2)
    DeviceBuffer<float> tok0, queries, tq, tk, tv, ta, tmp;  // (P,7,256)
    DeviceBuffer<float> t128a, t128b;                        // (P,7,128)
    DeviceBuffer<float> hid;                                 // (P,7,2048)
    DeviceBuffer<float> h1, h2;                              // (P*4,256)
    DeviceBuffer<float> hyper, iou;                          // (P,4,32), (P,4)
    DeviceBuffer<float> keys, kpe, big256;                   // (P,4096,256)
    DeviceBuffer<float> Kp, Vp, Qp, ao;                      // (P,4096,128)
    DeviceBuffer<float> up2;                                 // (P,16384,128)
    DeviceBuffer<float> low;                                 // (P,4,256,256)
    DeviceBuffer<int> plane_index;                           // (P*3)
    DeviceBuffer<float> iou_sel;                             // (P*3)
    DeviceBuffer<float> t2i_scratch;                         // split-key partials of the token->image attention
    explicit DecoderWorkspace(int max_prompts);
};

// Prompt-independent decoder inputs derived once per image embedding.
struct EmbeddingCache {
    DeviceBuffer<float> keys0, kpe0;     // (4096, 256): embedding + no_mask_embed, and + dense PE
    DeviceBuffer<float> K0, V0, Q0i;     // (4096, 128): layer-0 projections that do not depend on the prompt
    bool ready = false;
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
    // emb_out: (batch, 4096, 256) fp32, token-major (row = y*64 + x).
    // emb_nchw_out (optional): the same embedding as (batch, 256, 64, 64) fp32.  finish = false stops in front of the
    // final LayerNorm2d (its input is ws.big[0]); neck_finish() then writes the embedding -- the engine captures the
    // trunk into a CUDA graph and finishes eagerly into whichever store the call owns.
    void encode(cudaStream_t s, EncoderWorkspace& ws, enc::ImageDesc const* images, int batch, int w, int h, int channels,
                float* emb_out, Tap* tap = nullptr, float* emb_nchw_out = nullptr, bool finish = true) const;
    void neck_finish(cudaStream_t s, EncoderWorkspace& ws, int batch, float* emb_out, float* emb_nchw_out, Tap* tap = nullptr) const;

    void prepare_embedding(cudaStream_t s, float const* emb, EmbeddingCache& cache) const;

    // Runs the prompt encoder + mask decoder for P prompts (coords/labels already uploaded into ws).
    // Results: ws.low (P, 4, 256, 256) logits and ws.iou (P, 4).
    void decode(cudaStream_t s, DecoderWorkspace& ws, EmbeddingCache const& cache, int P) const;

    static StageCfg stage(int i);  // i = 1..3

  private:
    void gemm16(cudaStream_t s, act_t const* a, int64_t rows, Linear16 const& l, void* out, int act, act_t const* residual,
                float2 const* ln_stats = nullptr, bool out_f32 = false, int ln_parts = 0, float2* stats_out = nullptr) const;
    void gemm32(cudaStream_t s, float const* a, int64_t rows, Linear32 const& l, float* out, int act) const;
    void lin(cudaStream_t s, float const* x, int64_t xs, float const* x2, int rows, Linear32 const& l, bool relu, float* y,
             int64_t ys) const;
    void attn_tokens(cudaStream_t s, DecoderWorkspace& ws, AttnW const& a, bool with_pe, bool residual, Norm const& n,
                     int P) const;

    EncoderW enc_;
    DecoderW dec_;
    int num_sms_ = 148;
};

}  // namespace dlimg
