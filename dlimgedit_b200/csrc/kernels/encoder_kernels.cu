// encoder_kernels.cu -- see encoder_kernels.cuh.
#include "encoder_kernels.cuh"
#include "gelu.cuh"

#include "../profiler.hpp"

#include <cstdlib>

namespace dlimg {
namespace enc {

namespace {

// gelu_erf: see gelu.cuh

__device__ __forceinline__ void unpack8(uint4 const& v, float (&f)[8]) {
    act2_t const* h = reinterpret_cast<act2_t const*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 t = act22f2(h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ uint4 pack8(float const (&f)[8]) {
    uint4 v;
    act2_t* h = reinterpret_cast<act2_t*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = f22act2(f[2 * i], f[2 * i + 1]);
    return v;
}

#if DLIMG_B200_ALT  // unfused PatchEmbed conv1 and im2col: cross-check forms of patch_embed.cu / the implicit-GEMM neck conv
// ---------------------------------------------------------------------------------------------
struct Conv1Params {
    int w, h, bpp;
    int c0, c1, c2;  // byte offsets of R, G, B in a pixel
};

constexpr int kC1Tile = 16;
constexpr int kC1In = 2 * kC1Tile + 1;  // 33

__global__ void __launch_bounds__(256) conv1_preprocess_kernel(ImageDesc const* __restrict__ imgs, Conv1Params p,
                                                               float const* __restrict__ weight,
                                                               float const* __restrict__ bias, act_t* __restrict__ out) {
    __shared__ float tile[kC1In * kC1In * 3];
    __shared__ __align__(16) float wsm[27 * 32];
    __shared__ float bsm[32];
    int const tid = threadIdx.y * kC1Tile + threadIdx.x;
    int const b = blockIdx.z;
    ImageDesc const img = imgs[b];
    int const oy0 = blockIdx.y * kC1Tile, ox0 = blockIdx.x * kC1Tile;
    int const iy0 = 2 * oy0 - 1, ix0 = 2 * ox0 - 1;
    for (int i = tid; i < 27 * 32; i += 256) wsm[i] = weight[i];
    if (tid < 32) bsm[tid] = bias[tid];
    // mean / std of the encoder graph's preamble
    float const mean[3] = {123.675f, 116.28f, 103.53f};
    float const sd[3] = {58.395f, 57.12f, 57.375f};
    int const coff[3] = {p.c0, p.c1, p.c2};
    for (int i = tid; i < kC1In * kC1In; i += 256) {
        int const r = i / kC1In, c = i % kC1In;
        int const iy = iy0 + r, ix = ix0 + c;
        float v[3] = {0.f, 0.f, 0.f};
        if (iy >= 0 && iy < p.h && ix >= 0 && ix < p.w) {
            uint8_t const* px = img.pixels + (size_t)iy * img.stride + (size_t)ix * p.bpp;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) v[ch] = __fdiv_rn((float)px[coff[ch]] - mean[ch], sd[ch]);
        }
        tile[i * 3 + 0] = v[0];
        tile[i * 3 + 1] = v[1];
        tile[i * 3 + 2] = v[2];
    }
    __syncthreads();
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = bsm[i];
    int const ty = threadIdx.y, tx = threadIdx.x;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            float const* src = tile + ((2 * ty + ky) * kC1In + 2 * tx + kx) * 3;
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) {
                float const v = src[ci];
                float4 const* w4 = reinterpret_cast<float4 const*>(wsm + ((ky * 3 + kx) * 3 + ci) * 32);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float4 const w = w4[q];
                    acc[4 * q + 0] = fmaf(v, w.x, acc[4 * q + 0]);
                    acc[4 * q + 1] = fmaf(v, w.y, acc[4 * q + 1]);
                    acc[4 * q + 2] = fmaf(v, w.z, acc[4 * q + 2]);
                    acc[4 * q + 3] = fmaf(v, w.w, acc[4 * q + 3]);
                }
            }
        }
    }
    int const oy = oy0 + ty, ox = ox0 + tx;
    uint4* o = reinterpret_cast<uint4*>(out + (((size_t)b * 512 + oy) * 512 + ox) * 32);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = gelu_erf(acc[8 * q + i]);
        o[q] = pack8(f);
    }
}

// ---------------------------------------------------------------------------------------------
__global__ void im2col3x3_kernel(act_t const* __restrict__ in, int H, int W, int C8, int stride, int Ho, int Wo,
                                 int64_t total, act_t* __restrict__ out) {
    int64_t const t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    int const c8 = (int)(t % C8);
    int64_t r = t / C8;
    int const tap = (int)(r % 9);
    r /= 9;  // output pixel index (b, oy, ox)
    int const ox = (int)(r % Wo);
    int64_t r2 = r / Wo;
    int const oy = (int)(r2 % Ho);
    int const b = (int)(r2 / Ho);
    int const iy = oy * stride + tap / 3 - 1, ix = ox * stride + tap % 3 - 1;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W)
        v = reinterpret_cast<uint4 const*>(in)[(((int64_t)b * H + iy) * W + ix) * C8 + c8];
    reinterpret_cast<uint4*>(out)[t] = v;  // t == (r*9 + tap)*C8 + c8
}

#endif  // DLIMG_B200_ALT

// ---------------------------------------------------------------------------------------------
// Depthwise 3x3.  Each thread produces kTX horizontally adjacent outputs for 8 channels: the filter taps are
// loaded once into registers and every input column is read once per row and shared by up to three outputs.
// Threads are ordered (channel group fastest) so a warp reads whole contiguous pixels.  The channel count is a
// template parameter so all column offsets are immediates (one 64-bit base per input row), and threads whose
// window lies inside the row take a predicate-free path; grid = (x, output row, image).
template <int kStride, int kTX, int kC8, bool kEdge, typename Acc>
__device__ __forceinline__ void dwconv_rows(act_t const* __restrict__ in, int H, int W, int b, int oy, int c8, int ix0, Acc& acc) {
    constexpr int kCols = (kTX - 1) * kStride + 3;  // input columns feeding kTX outputs
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        int const iy = oy * kStride + ky - 1;
        if (iy < 0 || iy >= H) continue;  // block-uniform
        uint4 const* row = reinterpret_cast<uint4 const*>(in) + (((int64_t)b * H + iy) * W + ix0) * kC8 + c8;
        uint4 v[kCols];
#pragma unroll
        for (int c = 0; c < kCols; ++c) {
            if (kEdge) {
                int const ix = ix0 + c;
                v[c] = (ix >= 0 && ix < W) ? __ldg(row + c * kC8) : make_uint4(0, 0, 0, 0);
            } else {
                v[c] = __ldg(row + c * kC8);
            }
        }
#pragma unroll
        for (int c = 0; c < kCols; ++c) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                // input column c is tap kx of output o when c == o * kStride + kx
                if ((c - kx) % kStride != 0) continue;
                int const o = (c - kx) / kStride;
                if (o < 0 || o >= kTX) continue;
                acc.tap(o, ky * 3 + kx, v[c]);
            }
        }
    }
}

template <int kTX>
struct DwAcc32 {  // fp32 taps and accumulators
    static constexpr int kMinBlocks = 2;
    float w[9][8];
    float acc[kTX][8];
    __device__ __forceinline__ void init(float const* __restrict__ weight, float const* __restrict__ bias, int C, int c8) {
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            float4 const* w4 = reinterpret_cast<float4 const*>(weight + k * C + c8 * 8);
            float4 const w0 = __ldg(w4), w1 = __ldg(w4 + 1);
            w[k][0] = w0.x; w[k][1] = w0.y; w[k][2] = w0.z; w[k][3] = w0.w;
            w[k][4] = w1.x; w[k][5] = w1.y; w[k][6] = w1.z; w[k][7] = w1.w;
        }
        float4 const* b4 = reinterpret_cast<float4 const*>(bias + c8 * 8);
        float4 const b0 = __ldg(b4), b1 = __ldg(b4 + 1);
#pragma unroll
        for (int o = 0; o < kTX; ++o) {
            acc[o][0] = b0.x; acc[o][1] = b0.y; acc[o][2] = b0.z; acc[o][3] = b0.w;
            acc[o][4] = b1.x; acc[o][5] = b1.y; acc[o][6] = b1.z; acc[o][7] = b1.w;
        }
    }
    __device__ __forceinline__ void tap(int o, int k, uint4 const& v) {
        float f[8];
        unpack8(v, f);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[o][i] = fmaf(f[i], w[k][i], acc[o][i]);
    }
    __device__ __forceinline__ uint4 result(int o, bool gelu) {
        if (gelu) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[o][i] = gelu_erf(acc[o][i]);
        }
        return pack8(acc[o]);
    }
};

#if !defined(DLIMG_B200_ACT_BF16)
// Packed-half variant for the GELU'd depthwise convolutions (MBConv conv2, PatchMerging conv2): taps, bias and the
// erf GELU all run as HFMA2 on fp16 pairs, so the inputs need no unpacking and every instruction covers two
// channels.  The fp32 form spent ~48 issue slots per output (profiles/r01b: issue-bound at 62 %, 128 registers).
// Accumulation is fp16: nine products of GELU-bounded activations with BN-folded weights, an error of ~3 ulp on a
// value that is rounded to fp16 immediately afterwards anyway.
template <int kTX>
struct DwAccH2 {
    static constexpr int kMinBlocks = 3;  // 4 blocks (64 registers) measured the same: the spills eat the occupancy gain
    uint4 w[9];
    __half2 acc[kTX][4];
    __device__ __forceinline__ void init(act_t const* __restrict__ weight, float const* __restrict__ bias, int C, int c8) {
#pragma unroll
        for (int k = 0; k < 9; ++k) w[k] = __ldg(reinterpret_cast<uint4 const*>(weight + k * C + c8 * 8));
        float4 const* b4 = reinterpret_cast<float4 const*>(bias + c8 * 8);
        float4 const b0 = __ldg(b4), b1 = __ldg(b4 + 1);
        __half2 const h0 = f22act2(b0.x, b0.y), h1 = f22act2(b0.z, b0.w), h2 = f22act2(b1.x, b1.y), h3 = f22act2(b1.z, b1.w);
#pragma unroll
        for (int o = 0; o < kTX; ++o) { acc[o][0] = h0; acc[o][1] = h1; acc[o][2] = h2; acc[o][3] = h3; }
    }
    __device__ __forceinline__ void tap(int o, int k, uint4 const& v) {
        __half2 const* f = reinterpret_cast<__half2 const*>(&v);
        __half2 const* wk = reinterpret_cast<__half2 const*>(&w[k]);
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[o][i] = __hfma2(f[i], wk[i], acc[o][i]);
    }
    __device__ __forceinline__ uint4 result(int o, bool gelu) {
        uint4 ov;
        __half2* oh = reinterpret_cast<__half2*>(&ov);
#pragma unroll
        for (int i = 0; i < 4; ++i) oh[i] = gelu ? gelu_erf_h2(acc[o][i]) : acc[o][i];
        return ov;
    }
};
#endif

template <int kStride, int kTX, int kC8, typename Acc, typename WeightT>
__global__ void __launch_bounds__(256, Acc::kMinBlocks) dwconv3x3_kernel(act_t const* __restrict__ in, int H, int W, int Ho, int Wo, int xgroups,
                                                        WeightT const* __restrict__ weight, float const* __restrict__ bias,
                                                        int gelu, act_t* __restrict__ out) {
    pdl_wait();
    pdl_trigger();
    int const t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= xgroups * kC8) return;
    int const c8 = t % kC8, xg = t / kC8;
    int const oy = blockIdx.y, b = blockIdx.z;
    int const ox0 = xg * kTX;
    int const ix0 = ox0 * kStride - 1;
    constexpr int kCols = (kTX - 1) * kStride + 3;
    Acc acc;
    acc.init(weight, bias, kC8 * 8, c8);
    if (ix0 >= 0 && ix0 + kCols <= W) dwconv_rows<kStride, kTX, kC8, false>(in, H, W, b, oy, c8, ix0, acc);
    else dwconv_rows<kStride, kTX, kC8, true>(in, H, W, b, oy, c8, ix0, acc);
    uint4* orow = reinterpret_cast<uint4*>(out) + (((int64_t)b * Ho + oy) * Wo + ox0) * kC8 + c8;
#pragma unroll
    for (int o = 0; o < kTX; ++o) {
        if (ox0 + o >= Wo) break;
        orow[o * kC8] = acc.result(o, gelu != 0);
    }
}

// fp32 accumulators with the filter in shared memory ([tap][channel] floats): 72 fewer registers per thread than
// DwAcc32, which is what lets three or four blocks of the latency-bound local_conv kernel share an SM.
template <int kTX>
struct DwAcc32S {
    float const* w;  // shared memory, this thread's 8 channels of tap 0; taps are C floats apart
    int C;
    float acc[kTX][8];
    __device__ __forceinline__ void init(float const* wsm, float const* __restrict__ bias, int C_, int c8) {
        w = wsm + c8 * 8;
        C = C_;
        float4 const* b4 = reinterpret_cast<float4 const*>(bias + c8 * 8);
        float4 const b0 = __ldg(b4), b1 = __ldg(b4 + 1);
#pragma unroll
        for (int o = 0; o < kTX; ++o) {
            acc[o][0] = b0.x; acc[o][1] = b0.y; acc[o][2] = b0.z; acc[o][3] = b0.w;
            acc[o][4] = b1.x; acc[o][5] = b1.y; acc[o][6] = b1.z; acc[o][7] = b1.w;
        }
    }
    __device__ __forceinline__ void tap(int o, int k, uint4 const& v) {
        float f[8];
        unpack8(v, f);
        float4 const w0 = *reinterpret_cast<float4 const*>(w + k * C), w1 = *reinterpret_cast<float4 const*>(w + k * C + 4);
        acc[o][0] = fmaf(f[0], w0.x, acc[o][0]); acc[o][1] = fmaf(f[1], w0.y, acc[o][1]);
        acc[o][2] = fmaf(f[2], w0.z, acc[o][2]); acc[o][3] = fmaf(f[3], w0.w, acc[o][3]);
        acc[o][4] = fmaf(f[4], w1.x, acc[o][4]); acc[o][5] = fmaf(f[5], w1.y, acc[o][5]);
        acc[o][6] = fmaf(f[6], w1.z, acc[o][6]); acc[o][7] = fmaf(f[7], w1.w, acc[o][7]);
    }
    __device__ __forceinline__ uint4 result(int o, bool) { return pack8(acc[o]); }
};

// local_conv of a TinyViT block (stride 1, fp32 accumulation, no activation) that also leaves the LayerNorm row sums
// of its output for the fc1 GEMM / fused MLP behind it: a block holds kG whole groups of 4 pixels x kC8 channel
// octets, every thread writes the (sum, sum of squares) of its 8 channels per pixel to shared memory and kG * 4
// threads add the kC8 partials of one pixel in a fixed order (reproducible, no atomics).  When a group's kC8 threads
// sit inside one warp (kC8 = 16) the exchange needs a warp barrier only.
template <int kC8, int kG, int kMinBlocks>
__global__ void __launch_bounds__(kC8 * kG, kMinBlocks) dwconv3x3_stats_kernel(act_t const* __restrict__ in, int H, int W, int xgroups,
                                                                   float const* __restrict__ weight, float const* __restrict__ bias,
                                                                   act_t* __restrict__ out, float2* __restrict__ stats) {
    constexpr int kTX = 4, kCols = kTX + 2, kC = kC8 * 8;
    constexpr bool kWarpLocal = 32 % kC8 == 0;  // the kC8 threads of a group never straddle a warp
    __shared__ float2 part[kG][kTX][kC8];
    __shared__ __align__(16) float wsm[9 * kC];
    for (int i = threadIdx.x; i < 9 * kC / 4; i += kC8 * kG)
        reinterpret_cast<float4*>(wsm)[i] = __ldg(reinterpret_cast<float4 const*>(weight) + i);
    pdl_wait();  // the filter is a constant; in / out / stats belong to the neighbouring kernels
    pdl_trigger();
    __syncthreads();
    int const c8 = threadIdx.x % kC8, gl = threadIdx.x / kC8;
    int const xg = blockIdx.x * kG + gl;  // xgroups is a multiple of kG (checked on the host)
    int const oy = blockIdx.y, b = blockIdx.z;
    int const ox0 = xg * kTX, ix0 = ox0 - 1;
    DwAcc32S<kTX> acc;
    acc.init(wsm, bias, kC, c8);
    if (ix0 >= 0 && ix0 + kCols <= W) dwconv_rows<1, kTX, kC8, false>(in, H, W, b, oy, c8, ix0, acc);
    else dwconv_rows<1, kTX, kC8, true>(in, H, W, b, oy, c8, ix0, acc);
    int64_t const row0 = ((int64_t)b * H + oy) * W + ox0;
    uint4* orow = reinterpret_cast<uint4*>(out) + row0 * kC8 + c8;
#pragma unroll
    for (int o = 0; o < kTX; ++o) {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            s1 += acc.acc[o][i];
            s2 = fmaf(acc.acc[o][i], acc.acc[o][i], s2);
        }
        part[gl][o][c8] = make_float2(s1, s2);
        orow[o * kC8] = acc.result(o, false);
    }
    if (kWarpLocal) {
        __syncwarp();
        if (c8 < kTX) {  // lanes 0..3 of each group: one pixel each
            float s1 = 0.f, s2 = 0.f;
#pragma unroll 4
            for (int k = 0; k < kC8; ++k) {
                float2 const v = part[gl][c8][k];
                s1 += v.x;
                s2 += v.y;
            }
            stats[row0 + c8] = make_float2(s1, s2);
        }
    } else {
        __syncthreads();
        if (threadIdx.x < kG * kTX) {
            int const g2 = threadIdx.x / kTX, o = threadIdx.x % kTX;
            float s1 = 0.f, s2 = 0.f;
#pragma unroll 4
            for (int k = 0; k < kC8; ++k) {
                float2 const v = part[g2][o][k];
                s1 += v.x;
                s2 += v.y;
            }
            stats[((int64_t)b * H + oy) * W + (blockIdx.x * kG + g2) * kTX + o] = make_float2(s1, s2);
        }
    }
}

// ---------------------------------------------------------------------------------------------
constexpr int kLnMaxPairs = 5;  // C <= 320

__global__ void __launch_bounds__(256) layernorm_rows_kernel(act_t const* __restrict__ in, int rows, int C,
                                                             int const* __restrict__ src_row,
                                                             float const* __restrict__ gamma,
                                                             float const* __restrict__ beta, float eps,
                                                             void* __restrict__ out, int out_f32) {
    pdl_wait();
    pdl_trigger();
    int const row = blockIdx.x * 8 + (threadIdx.x >> 5);
    int const lane = threadIdx.x & 31;
    if (row >= rows) return;
    int const pairs = C >> 1;
    int64_t src = row;
    if (src_row) src = src_row[row];
    float2 x[kLnMaxPairs];
    float sum = 0.f;
    if (src >= 0) {
        act2_t const* p = reinterpret_cast<act2_t const*>(in + src * C);
#pragma unroll
        for (int i = 0; i < kLnMaxPairs; ++i) {
            int const idx = lane + 32 * i;
            x[i] = idx < pairs ? act22f2(p[idx]) : make_float2(0.f, 0.f);
            sum += x[i].x + x[i].y;
        }
    }
    float rstd = 0.f, mean = 0.f;
    if (src >= 0) {  // warp-uniform
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        mean = sum / (float)C;
        float var = 0.f;
#pragma unroll
        for (int i = 0; i < kLnMaxPairs; ++i) {
            int const idx = lane + 32 * i;
            if (idx < pairs) {
                float const a = x[i].x - mean, b = x[i].y - mean;
                var += a * a + b * b;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
        rstd = rsqrtf(var / (float)C + eps);
    }
#pragma unroll
    for (int i = 0; i < kLnMaxPairs; ++i) {
        int const idx = lane + 32 * i;
        if (idx >= pairs) continue;
        float2 const g = reinterpret_cast<float2 const*>(gamma)[idx];
        float2 const bt = reinterpret_cast<float2 const*>(beta)[idx];
        float2 y;
        if (src >= 0) {
            y.x = (x[i].x - mean) * rstd * g.x + bt.x;
            y.y = (x[i].y - mean) * rstd * g.y + bt.y;
        } else {
            y = bt;  // LayerNorm of an all-zero padding token
        }
        if (out_f32) reinterpret_cast<float2*>(reinterpret_cast<float*>(out) + (int64_t)row * C)[idx] = y;
        else reinterpret_cast<act2_t*>(reinterpret_cast<act_t*>(out) + (int64_t)row * C)[idx] = f22act2(y.x, y.y);
    }
}

// ---------------------------------------------------------------------------------------------
// Final LayerNorm2d of the neck (C = 256) writing the embedding in both layouts the engine keeps: token-major fp32 (the
// decoder's input) and NCHW fp32 (what get_embedding returns).  A block normalises 32 consecutive tokens of one image
// (same arithmetic as layernorm_rows_kernel), writes the rows, and transposes through shared memory so that every
// channel's 32 tokens leave as one 128-byte segment.
__global__ void __launch_bounds__(256) layernorm256_tokens_nchw_kernel(act_t const* __restrict__ in, int tokens,
                                                                       float const* __restrict__ gamma, float const* __restrict__ beta,
                                                                       float eps, float const* __restrict__ no_mask,
                                                                       act_t* __restrict__ out_keys, float* __restrict__ out_nchw) {
    constexpr int C = 256;
    __shared__ float tile[C][33];
    int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int const b = blockIdx.y, t0 = blockIdx.x * 32;
    float2 g[4], bt[4], nm[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        g[i] = reinterpret_cast<float2 const*>(gamma)[lane + 32 * i];
        bt[i] = reinterpret_cast<float2 const*>(beta)[lane + 32 * i];
        nm[i] = reinterpret_cast<float2 const*>(no_mask)[lane + 32 * i];
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        int const tl = warp * 4 + r;
        int64_t const row = (int64_t)b * tokens + t0 + tl;
        act2_t const* p = reinterpret_cast<act2_t const*>(in + row * C);
        float2 x[4];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            x[i] = act22f2(p[lane + 32 * i]);
            sum += x[i].x + x[i].y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        float const mean = sum / (float)C;
        float var = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float const a = x[i].x - mean, c = x[i].y - mean;
            var += a * a + c * c;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
        float const rstd = rsqrtf(var / (float)C + eps);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int const idx = lane + 32 * i;
            float2 y;
            y.x = (x[i].x - mean) * rstd * g[i].x + bt[i].x;
            y.y = (x[i].y - mean) * rstd * g[i].y + bt[i].y;
            // the decoder's layer-0 image stream: embedding + no_mask_embed (the dense prompt embedding of the reference's
            // constant empty mask input, segmentation.cpp:43-45), 16-bit token-major
            reinterpret_cast<act2_t*>(out_keys + row * C)[idx] = f22act2(y.x + nm[i].x, y.y + nm[i].y);
            tile[2 * idx][tl] = y.x;
            tile[2 * idx + 1][tl] = y.y;
        }
    }
    __syncthreads();
    float* const dst = out_nchw + (int64_t)b * C * tokens + t0 + lane;
#pragma unroll 8
    for (int c = warp * 32; c < warp * 32 + 32; ++c) dst[(int64_t)c * tokens] = tile[c][lane];
}

// ---------------------------------------------------------------------------------------------
// Row statistics for the LayerNorms folded into GEMM epilogues: one warp per kStatRows rows, 16-byte loads, all
// rows' loads issued before the first reduction, two-pass (mean, then centred variance) in registers.
constexpr int kStatRows = 4;
constexpr int kStatChunks = 2;  // 16-byte chunks per lane per row: C <= 512

__global__ void __launch_bounds__(256) layernorm_stats_kernel(act_t const* __restrict__ in, int rows, int C8, float eps,
                                                              float2* __restrict__ out) {
    pdl_wait();
    pdl_trigger();
    int const warp = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    int const row0 = warp * kStatRows;
    if (row0 >= rows) return;
    uint4 v[kStatRows][kStatChunks];
#pragma unroll
    for (int r = 0; r < kStatRows; ++r)
#pragma unroll
        for (int c = 0; c < kStatChunks; ++c) {
            int const ch = lane + 32 * c;
            v[r][c] = (row0 + r < rows && ch < C8) ? __ldg(reinterpret_cast<uint4 const*>(in) + (int64_t)(row0 + r) * C8 + ch)
                                                   : make_uint4(0, 0, 0, 0);
        }
    float const inv_c = 1.0f / (float)(C8 * 8);
#pragma unroll
    for (int r = 0; r < kStatRows; ++r) {
        float f[kStatChunks][8];
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < kStatChunks; ++c) {
            unpack8(v[r][c], f[c]);
#pragma unroll
            for (int i = 0; i < 8; ++i) sum += f[c][i];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        float const mean = sum * inv_c;
        float var = 0.f;
#pragma unroll
        for (int c = 0; c < kStatChunks; ++c) {
            if (lane + 32 * c < C8) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float const d = f[c][i] - mean;
                    var = fmaf(d, d, var);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
        if (lane == 0 && row0 + r < rows) out[row0 + r] = make_float2(mean, rsqrtf(var * inv_c + eps));
    }
}

// ---------------------------------------------------------------------------------------------
constexpr int kAttnWarps = 8;
constexpr int kAttnMaxJ = 7;  // n <= 224 keys

__global__ void __launch_bounds__(kAttnWarps * 32) window_attention_kernel(act_t const* __restrict__ qkv, int n, int heads,
                                                                           float const* __restrict__ bias,
                                                                           act_t* __restrict__ out) {
    extern __shared__ float sm[];
    float* Ks = sm;                       // [n][33]
    float* Vs = Ks + n * 33;              // [n][32]
    float* Qs = Vs + n * 32;              // [warps][32]
    float* Ps = Qs + kAttnWarps * 32;     // [warps][n]
    int const win = blockIdx.x / heads, h = blockIdx.x % heads;
    int const ld = heads * 96, C = heads * 32;
    int64_t const row0 = (int64_t)win * n;
    int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < n * 32; i += blockDim.x) {
        int const j = i >> 5, d = i & 31;
        act_t const* base = qkv + (row0 + j) * ld + h * 96;
        Ks[j * 33 + d] = act2f(base[32 + d]);
        Vs[j * 32 + d] = act2f(base[64 + d]);
    }
    __syncthreads();
    float const scale = 0.17677669529663687f;  // 32^-0.5
    float const* bias_h = bias + (int64_t)h * n * n;
    for (int i = warp; i < n; i += kAttnWarps) {
        Qs[warp * 32 + lane] = act2f(qkv[(row0 + i) * ld + h * 96 + lane]);
        __syncwarp();
        float s[kAttnMaxJ];
        float mx = -INFINITY;
#pragma unroll
        for (int jj = 0; jj < kAttnMaxJ; ++jj) {
            int const j = jj * 32 + lane;
            s[jj] = -INFINITY;
            if (j < n) {
                float a = 0.f;
#pragma unroll
                for (int d = 0; d < 32; ++d) a = fmaf(Qs[warp * 32 + d], Ks[j * 33 + d], a);
                s[jj] = a * scale + __ldg(bias_h + (int64_t)i * n + j);
            }
            mx = fmaxf(mx, s[jj]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
#pragma unroll
        for (int jj = 0; jj < kAttnMaxJ; ++jj) {
            int const j = jj * 32 + lane;
            float const e = j < n ? __expf(s[jj] - mx) : 0.f;
            s[jj] = e;
            sum += e;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        float const inv = 1.0f / sum;
#pragma unroll
        for (int jj = 0; jj < kAttnMaxJ; ++jj) {
            int const j = jj * 32 + lane;
            if (j < n) Ps[warp * n + j] = s[jj] * inv;
        }
        __syncwarp();
        float o = 0.f;
        for (int j = 0; j < n; ++j) o = fmaf(Ps[warp * n + j], Vs[j * 32 + lane], o);
        out[(row0 + i) * C + h * 32 + lane] = f2act(o);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
__global__ void tokens_to_nchw_kernel(float const* __restrict__ in, int tokens, int C, float* __restrict__ out) {
    pdl_wait();
    pdl_trigger();
    __shared__ float tile[32][33];
    int const b = blockIdx.z;
    int const t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    float const* src = in + (int64_t)b * tokens * C;
    float* dst = out + (int64_t)b * tokens * C;
    for (int i = threadIdx.y; i < 32; i += 8) {
        int const t = t0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (t < tokens && c < C) ? src[(int64_t)t * C + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        int const c = c0 + i, t = t0 + threadIdx.x;
        if (t < tokens && c < C) dst[(int64_t)c * tokens + t] = tile[threadIdx.x][i];
    }
}

__global__ void act_to_f32_kernel(act_t const* __restrict__ in, int64_t n, float* __restrict__ out) {
    int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = act2f(in[i]);
}

}  // namespace

#if DLIMG_B200_ALT
// ---------------------------------------------------------------------------------------------
void conv1_preprocess(cudaStream_t s, ImageDesc const* imgs, int batch, int w, int h, int channels,
                      float const* weight, float const* bias, act_t* out) {
    DLIMG_ASSERT(w >= 1 && h >= 1 && w <= kImageSize && h <= kImageSize);
    int cmap[3];
    channel_map(channels, cmap);
    Conv1Params p{w, h, bytes_per_pixel(channels), cmap[0], cmap[1], cmap[2]};
    ProfScope prof(s, CAT_CONV1, 2.0 * batch * 512 * 512 * 27 * 32,
                   (double)batch * ((double)w * h * p.bpp + 512.0 * 512 * 32 * 2));
    dim3 grid(512 / kC1Tile, 512 / kC1Tile, batch), block(kC1Tile, kC1Tile);
    conv1_preprocess_kernel<<<grid, block, 0, s>>>(imgs, p, weight, bias, out);
    KERNEL_CHECK();
}

void im2col3x3(cudaStream_t s, act_t const* in, int batch, int H, int W, int C, int stride, act_t* out) {
    DLIMG_ASSERT(C % 8 == 0);
    int const Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
    int64_t const total = (int64_t)batch * Ho * Wo * 9 * (C / 8);
    ProfScope prof(s, CAT_IM2COL, 0, (double)batch * ((double)H * W * C * 2 + (double)Ho * Wo * 9 * C * 2));
    im2col3x3_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, s>>>(in, H, W, C / 8, stride, Ho, Wo, total, out);
    KERNEL_CHECK();
}

#endif  // DLIMG_B200_ALT

namespace {
template <int kStride, int kC8>
void launch_dwconv(cudaStream_t s, dim3 grid, act_t const* in, int H, int W, int Ho, int Wo, int xgroups, float const* weight,
                   act_t const* weight16, float const* bias, bool gelu, act_t* out) {
    constexpr int kTX = 4;
#if !defined(DLIMG_B200_ACT_BF16)
    static bool const acc32 = dev_switch("DLIMG_B200_DW_ACC32");  // A/B switch for parity experiments
    if (weight16 && !acc32) {
        launch_pdl(PDL_DWCONV, dwconv3x3_kernel<kStride, kTX, kC8, DwAccH2<kTX>, act_t>, grid, dim3(256), 0, s, in, H, W, Ho, Wo, xgroups, weight16, bias, gelu ? 1 : 0, out);
        return;
    }
#endif
    (void)weight16;
    launch_pdl(PDL_DWCONV, dwconv3x3_kernel<kStride, kTX, kC8, DwAcc32<kTX>, float>, grid, dim3(256), 0, s, in, H, W, Ho, Wo, xgroups, weight, bias, gelu ? 1 : 0, out);
}
}  // namespace

void dwconv3x3(cudaStream_t s, act_t const* in, int batch, int H, int W, int C, int stride, float const* weight,
               act_t const* weight16, float const* bias, bool gelu, act_t* out) {
    int const Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
    DLIMG_ASSERT(stride == 1 || stride == 2);
    int const xgroups = ceil_div(Wo, 4);
    ProfScope prof(s, CAT_DWCONV, 2.0 * batch * Ho * Wo * C * 9, (double)batch * ((double)H * W + (double)Ho * Wo) * C * 2);
    dim3 const grid((unsigned)ceil_div(xgroups * (C / 8), 256), (unsigned)Ho, (unsigned)batch);
#define DLIMG_DW_CASE(S, CC) \
    if (stride == S && C == CC) { launch_dwconv<S, CC / 8>(s, grid, in, H, W, Ho, Wo, xgroups, weight, weight16, bias, gelu, out); KERNEL_CHECK(); return; }
    DLIMG_DW_CASE(1, 256) DLIMG_DW_CASE(2, 128) DLIMG_DW_CASE(2, 160) DLIMG_DW_CASE(1, 320)
    DLIMG_DW_CASE(1, 128) DLIMG_DW_CASE(1, 160)
#undef DLIMG_DW_CASE
    fail("dwconv3x3: unsupported (stride, channels) = (" + std::to_string(stride) + ", " + std::to_string(C) + ")");
}

void dwconv3x3_stats(cudaStream_t s, act_t const* in, int batch, int H, int W, int C, float const* weight, float const* bias,
                     act_t* out, float2* stats) {
    int const xgroups = W / 4;
    DLIMG_ASSERT(W % 4 == 0);
    ProfScope prof(s, CAT_DWCONV, 2.0 * batch * H * W * C * 9, (double)batch * 2.0 * H * W * C * 2);
#define DLIMG_DWS_CASE(CC, G, MB)                                                                                           \
    if (C == CC && xgroups % G == 0) {                                                                                    \
        dim3 const grid((unsigned)(xgroups / G), (unsigned)H, (unsigned)batch);                                           \
        launch_pdl(PDL_LOCAL_CONV, dwconv3x3_stats_kernel<CC / 8, G, MB>, grid, dim3((CC / 8) * G), 0, s, in, H, W, xgroups, weight, bias, out, stats); \
        KERNEL_CHECK();                                                                                                   \
        return;                                                                                                           \
    }
    DLIMG_DWS_CASE(128, 16, 3) DLIMG_DWS_CASE(160, 8, 5) DLIMG_DWS_CASE(320, 4, 5)  // 256 / 160 / 160 threads
#undef DLIMG_DWS_CASE
    fail("dwconv3x3_stats: unsupported (channels, width) = (" + std::to_string(C) + ", " + std::to_string(W) + ")");
}

void layernorm_rows(cudaStream_t s, act_t const* in, int rows, int C, int const* src_row, float const* gamma,
                    float const* beta, float eps, void* out, bool out_f32) {
    DLIMG_ASSERT(C % 2 == 0 && C <= kLnMaxPairs * 64);
    ProfScope prof(s, CAT_LAYERNORM, 0, (double)rows * C * (2 + (out_f32 ? 4 : 2)));
    launch_pdl(PDL_ROWS, layernorm_rows_kernel, dim3(ceil_div(rows, 8)), dim3(256), 0, s, in, rows, C, src_row, gamma, beta, eps, out, out_f32 ? 1 : 0);
    KERNEL_CHECK();
}

void layernorm256_tokens_nchw(cudaStream_t s, act_t const* in, int batch, int tokens, float const* gamma, float const* beta,
                              float eps, float const* no_mask, act_t* out_keys, float* out_nchw) {
    DLIMG_ASSERT(tokens % 32 == 0);
    ProfScope prof(s, CAT_LAYERNORM, 0, (double)batch * tokens * 256 * (2 + 2 + 4));
    layernorm256_tokens_nchw_kernel<<<dim3((unsigned)(tokens / 32), (unsigned)batch), 256, 0, s>>>(in, tokens, gamma, beta, eps, no_mask, out_keys, out_nchw);
    KERNEL_CHECK();
}

void layernorm_stats(cudaStream_t s, act_t const* in, int rows, int C, float eps, float2* out) {
    DLIMG_ASSERT(C % 8 == 0 && C / 8 <= 32 * kStatChunks);
    ProfScope prof(s, CAT_LAYERNORM, 0, (double)rows * (C * 2 + 8));
    launch_pdl(PDL_ROWS, layernorm_stats_kernel, dim3(ceil_div(rows, 8 * kStatRows)), dim3(256), 0, s, in, rows, C / 8, eps, out);
    KERNEL_CHECK();
}

void window_attention_simt(cudaStream_t s, act_t const* qkv, int windows, int n, int heads, float const* bias, act_t* out) {
    DLIMG_ASSERT(n <= kAttnMaxJ * 32);
    ProfScope prof(s, CAT_OTHER);
    size_t const smem = sizeof(float) * ((size_t)n * 33 + (size_t)n * 32 + kAttnWarps * 32 + (size_t)kAttnWarps * n);
    static bool attr_set = false;
    if (!attr_set) {
        CUDA_CHECK(cudaFuncSetAttribute(window_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        attr_set = true;
    }
    window_attention_kernel<<<windows * heads, kAttnWarps * 32, smem, s>>>(qkv, n, heads, bias, out);
    KERNEL_CHECK();
}

void tokens_to_nchw(cudaStream_t s, float const* in, int batch, int tokens, int C, float* out) {
    ProfScope prof(s, CAT_OTHER);
    dim3 grid(ceil_div(tokens, 32), ceil_div(C, 32), batch), block(32, 8);
    launch_pdl(PDL_ROWS, tokens_to_nchw_kernel, grid, block, 0, s, in, tokens, C, out);
    KERNEL_CHECK();
}

void act_to_f32(cudaStream_t s, act_t const* in, int64_t n, float* out) {
    ProfScope prof(s, CAT_OTHER);
    act_to_f32_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, s>>>(in, n, out);
    KERNEL_CHECK();
}

}  // namespace enc
}  // namespace dlimg
