// decoder_kernels.cu -- see decoder_kernels.cuh.
#include "decoder_kernels.cuh"
#include "mask_select.cuh"
#include "gelu.cuh"

#include "../profiler.hpp"

namespace dlimg {
namespace dec {

namespace {

// gelu_erf: see gelu.cuh
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// random-Fourier positional encoding of a point in [0,1]^2 (SAM PositionEmbeddingRandom._pe_encoding)
__device__ __forceinline__ float pe_feature(float cx, float cy, float const* __restrict__ G, int j) {
    int const jj = j & 127;
    float const x = 2.0f * cx - 1.0f, y = 2.0f * cy - 1.0f;
    float v = fmaf(y, G[128 + jj], x * G[jj]);
    v *= 6.283185307179586f;
    return j < 128 ? sinf(v) : cosf(v);
}

__global__ void __launch_bounds__(256) prompt_tokens_kernel(float const* __restrict__ coords, float const* __restrict__ labels,
                                                            PromptParams pp, float* __restrict__ tokens,
                                                            float* __restrict__ queries) {
    int const p = blockIdx.x, j = threadIdx.x;
    float* out = tokens + (size_t)p * kTokens * kDim;
    float* out2 = queries + (size_t)p * kTokens * kDim;  // the transformer's running token state starts as a copy
    out[j] = out2[j] = pp.iou_token[j];
#pragma unroll
    for (int m = 0; m < 4; ++m) out[(1 + m) * kDim + j] = out2[(1 + m) * kDim + j] = pp.mask_tokens[m * kDim + j];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        float const cx = (coords[(p * 2 + i) * 2 + 0] + 0.5f) / 1024.0f;
        float const cy = (coords[(p * 2 + i) * 2 + 1] + 0.5f) / 1024.0f;
        float const lab = labels[p * 2 + i];
        float v = pe_feature(cx, cy, pp.gaussian, j);
        v = v * (lab != -1.0f ? 1.0f : 0.0f);
        v = v + pp.not_a_point[j] * (lab == -1.0f ? 1.0f : 0.0f);
#pragma unroll
        for (int e = 0; e < 4; ++e) v = v + pp.point_embed[e * kDim + j] * (lab == (float)e ? 1.0f : 0.0f);
        out[(5 + i) * kDim + j] = out2[(5 + i) * kDim + j] = v;
    }
}

__global__ void __launch_bounds__(256) dense_pe_kernel(float const* __restrict__ G, float* __restrict__ pos) {
    int const tok = blockIdx.x, j = threadIdx.x;
    int const iy = tok / kEmbedRes, ix = tok % kEmbedRes;
    float const cx = ((float)ix + 0.5f) / (float)kEmbedRes, cy = ((float)iy + 0.5f) / (float)kEmbedRes;
    pos[(size_t)tok * kDim + j] = pe_feature(cx, cy, G, j);
}

// ---------------------------------------------------------------------------------------------
constexpr int kLinRows = 8;   // rows per block
constexpr int kLinWarps = 8;
constexpr int kLinColsPerBlock = 64;

__global__ void __launch_bounds__(kLinWarps * 32) linear_small_kernel(float const* __restrict__ x, int64_t x_stride,
                                                                      float const* __restrict__ x2, int64_t x2_stride,
                                                                      int rows, int K, float const* __restrict__ W,
                                                                      float const* __restrict__ b, int N, int relu,
                                                                      float* __restrict__ y, int64_t y_stride) {
    extern __shared__ float xs[];  // [kLinRows][K]
    int const r0 = blockIdx.x * kLinRows;
    for (int i = threadIdx.x; i < kLinRows * K; i += blockDim.x) {
        int const r = i / K, k = i % K;
        float v = 0.f;
        if (r0 + r < rows) {
            v = x[(int64_t)(r0 + r) * x_stride + k];
            if (x2) v += x2[(int64_t)(r0 + r) * x2_stride + k];
        }
        xs[i] = v;
    }
    __syncthreads();
    int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int const n_end = min(N, (int)(blockIdx.y + 1) * kLinColsPerBlock);
    for (int n = blockIdx.y * kLinColsPerBlock + warp; n < n_end; n += kLinWarps) {
        float acc[kLinRows];
#pragma unroll
        for (int r = 0; r < kLinRows; ++r) acc[r] = 0.f;
        float const* w = W + (int64_t)n * K;
        for (int k = lane; k < K; k += 32) {
            float const wv = __ldg(w + k);
#pragma unroll
            for (int r = 0; r < kLinRows; ++r) acc[r] = fmaf(xs[r * K + k], wv, acc[r]);
        }
#pragma unroll
        for (int r = 0; r < kLinRows; ++r) acc[r] = warp_sum(acc[r]);
        if (lane < kLinRows && r0 + lane < rows) {
            float v = 0.f;
#pragma unroll
            for (int r = 0; r < kLinRows; ++r) v = lane == r ? acc[r] : v;
            if (b) v += b[n];
            if (relu) v = fmaxf(v, 0.f);
            y[(int64_t)(r0 + lane) * y_stride + n] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) token_self_attention_kernel(float const* __restrict__ q, float const* __restrict__ k,
                                                                   float const* __restrict__ v, float* __restrict__ out) {
    int const p = blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
    size_t const base = (size_t)p * kTokens * kDim + h * 32 + lane;
    float kk[kTokens], vv[kTokens];
#pragma unroll
    for (int u = 0; u < kTokens; ++u) {
        kk[u] = k[base + u * kDim];
        vv[u] = v[base + u * kDim];
    }
    float const scale = 0.17677669529663687f;  // 1/sqrt(32)
#pragma unroll
    for (int t = 0; t < kTokens; ++t) {
        float const qv = q[base + t * kDim];
        float s[kTokens];
        float mx = -INFINITY;
#pragma unroll
        for (int u = 0; u < kTokens; ++u) {
            s[u] = warp_sum(qv * kk[u]) * scale;
            mx = fmaxf(mx, s[u]);
        }
        float sum = 0.f;
#pragma unroll
        for (int u = 0; u < kTokens; ++u) {
            s[u] = expf(s[u] - mx);
            sum += s[u];
        }
        float o = 0.f;
#pragma unroll
        for (int u = 0; u < kTokens; ++u) o = fmaf(s[u] / sum, vv[u], o);
        out[base + t * kDim] = o;
    }
}

// ---------------------------------------------------------------------------------------------
// Image -> token attention core, one thread per (image token, head).  Q rows are 16-bit (the [K|V|Q] projection of the
// image stream, row pitch q_pitch elements); the 7 token keys / values of the prompt sit in shared memory as
// [token][d / 4][head] float4.  Output: 16-bit (P, 4096, 128), the A operand of the out-projection GEMM.
__global__ void __launch_bounds__(256) i2t_attention_kernel(act_t const* __restrict__ Q, act_t const* const* __restrict__ Qptrs,
                                                            int64_t q_prompt_stride, int q_pitch, int q_off,
                                                            float const* __restrict__ kt, float const* __restrict__ vt,
                                                            act_t* __restrict__ out) {
    __shared__ float4 ks[kTokens * 4 * 8];
    __shared__ float4 vs[kTokens * 4 * 8];
    int const p = blockIdx.y;
    for (int i = threadIdx.x; i < kTokens * 32; i += blockDim.x) {
        int const t = i >> 5, hh = (i >> 2) & 7, dq = i & 3;  // source order: [t][head][d / 4]
        float4 const* ksrc = reinterpret_cast<float4 const*>(kt + (size_t)p * kTokens * 128);
        float4 const* vsrc = reinterpret_cast<float4 const*>(vt + (size_t)p * kTokens * 128);
        ks[(t * 4 + dq) * 8 + hh] = ksrc[i];
        vs[(t * 4 + dq) * 8 + hh] = vsrc[i];
    }
    __syncthreads();
    // warp = head, lane = image token: the 7 x 16 keys / values of the warp's head are the same shared-memory words for
    // every lane (one broadcast wavefront per load; with lane = head each 16-byte load took four), and every lane reads /
    // writes whole 32-byte sectors of its own token row
    int const i = blockIdx.x * 32 + (threadIdx.x & 31), h = threadIdx.x >> 5;
    act_t const* qbase = (Qptrs ? Qptrs[p] : Q + (size_t)p * q_prompt_stride) + q_off;
    float qv[16];
    {
        uint4 const* q4 = reinterpret_cast<uint4 const*>(qbase + (size_t)i * q_pitch + h * 16);
        uint4 const a = __ldg(q4), b = __ldg(q4 + 1);
        act2_t const* ha = reinterpret_cast<act2_t const*>(&a);
        act2_t const* hb = reinterpret_cast<act2_t const*>(&b);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float2 const fa = act22f2(ha[j]), fb = act22f2(hb[j]);
            qv[2 * j] = fa.x; qv[2 * j + 1] = fa.y;
            qv[8 + 2 * j] = fb.x; qv[8 + 2 * j + 1] = fb.y;
        }
    }
    float s[kTokens];
    float mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < kTokens; ++t) {
        float a = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float4 const k4 = ks[(t * 4 + c) * 8 + h];
            a = fmaf(qv[4 * c + 0], k4.x, a);
            a = fmaf(qv[4 * c + 1], k4.y, a);
            a = fmaf(qv[4 * c + 2], k4.z, a);
            a = fmaf(qv[4 * c + 3], k4.w, a);
        }
        s[t] = a * 0.25f;
        mx = fmaxf(mx, s[t]);
    }
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < kTokens; ++t) {
        s[t] = expf(s[t] - mx);
        sum += s[t];
    }
    float const inv = 1.0f / sum;
    float o[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) o[c] = 0.f;
#pragma unroll
    for (int t = 0; t < kTokens; ++t) {
        float const w = s[t] * inv;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float4 const v4 = vs[(t * 4 + c) * 8 + h];
            o[4 * c + 0] = fmaf(w, v4.x, o[4 * c + 0]);
            o[4 * c + 1] = fmaf(w, v4.y, o[4 * c + 1]);
            o[4 * c + 2] = fmaf(w, v4.z, o[4 * c + 2]);
            o[4 * c + 3] = fmaf(w, v4.w, o[4 * c + 3]);
        }
    }
    uint4 ov[2];
    act2_t* oh = reinterpret_cast<act2_t*>(ov);
#pragma unroll
    for (int j = 0; j < 8; ++j) oh[j] = f22act2(o[2 * j], o[2 * j + 1]);
    uint4* o4 = reinterpret_cast<uint4*>(out + ((size_t)p * kImgTokens + i) * 128 + h * 16);
    o4[0] = ov[0];
    o4[1] = ov[1];
}

// keys <- LayerNorm_256(x + res) on the 16-bit image stream (eps 1e-5, fp32 statistics): one warp per row, 16-byte
// loads.  The residual of prompt p is res_ptrs[p] (layer 0: the image's own prompt-independent keys) or
// res + p * 4096 * 256.  In-place (out == res) is allowed: a row is read and written by the same lane.
__global__ void __launch_bounds__(256) layernorm256_img_kernel(act_t const* __restrict__ x, act_t const* res,
                                                               act_t const* const* __restrict__ res_ptrs, int64_t rows,
                                                               float const* __restrict__ gamma, float const* __restrict__ beta,
                                                               act_t* out) {
    int64_t const row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    int const lane = threadIdx.x & 31;
    if (row >= rows) return;
    int64_t const p = row / kImgTokens, tok = row - p * kImgTokens;
    act_t const* r = (res_ptrs ? res_ptrs[p] : res + p * kImgTokens * kDim) + tok * kDim;
    uint4 const xa = __ldg(reinterpret_cast<uint4 const*>(x + row * kDim) + lane);
    uint4 const ra = *(reinterpret_cast<uint4 const*>(r) + lane);
    act2_t const* xh = reinterpret_cast<act2_t const*>(&xa);
    act2_t const* rh = reinterpret_cast<act2_t const*>(&ra);
    float v[8];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float2 const a = act22f2(xh[j]), b = act22f2(rh[j]);
        v[2 * j] = a.x + b.x;
        v[2 * j + 1] = a.y + b.y;
        sum += v[2 * j] + v[2 * j + 1];
    }
    float const mean = warp_sum(sum) * (1.0f / kDim);
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float const d = v[i] - mean;
        var = fmaf(d, d, var);
    }
    float const rstd = rsqrtf(warp_sum(var) * (1.0f / kDim) + 1e-5f);
    float4 const g0 = __ldg(reinterpret_cast<float4 const*>(gamma) + 2 * lane), g1 = __ldg(reinterpret_cast<float4 const*>(gamma) + 2 * lane + 1);
    float4 const b0 = __ldg(reinterpret_cast<float4 const*>(beta) + 2 * lane), b1 = __ldg(reinterpret_cast<float4 const*>(beta) + 2 * lane + 1);
    float const gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    float const bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    uint4 ov;
    act2_t* oh = reinterpret_cast<act2_t*>(&ov);
#pragma unroll
    for (int j = 0; j < 4; ++j)
        oh[j] = f22act2((v[2 * j] - mean) * rstd * gg[2 * j] + bb[2 * j], (v[2 * j + 1] - mean) * rstd * gg[2 * j + 1] + bb[2 * j + 1]);
    *(reinterpret_cast<uint4*>(out + row * kDim) + lane) = ov;
}

// ---------------------------------------------------------------------------------------------
// `res` and `out` may alias (in-place residual update), so neither is __restrict__.
__global__ void __launch_bounds__(256) layernorm256_kernel(float const* x, float const* res, int64_t res_mod, int64_t rows,
                                                           float const* __restrict__ gamma, float const* __restrict__ beta,
                                                           float const* __restrict__ pos, int64_t pos_mod, float* out,
                                                           float* out2) {
    int64_t const row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    int const lane = threadIdx.x & 31;
    if (row >= rows) return;
    float v[8];
    {
        float4 const a = reinterpret_cast<float4 const*>(x + row * kDim)[lane];
        float4 const b = reinterpret_cast<float4 const*>(x + row * kDim)[32 + lane];
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    if (res) {
        float const* r = res + (row % res_mod) * kDim;
        float4 const a = reinterpret_cast<float4 const*>(r)[lane];
        float4 const b = reinterpret_cast<float4 const*>(r)[32 + lane];
        v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
    }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) sum += v[i];
    float const mean = warp_sum(sum) * (1.0f / kDim);
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float const d = v[i] - mean;
        var = fmaf(d, d, var);
    }
    float const rstd = rsqrtf(warp_sum(var) * (1.0f / kDim) + 1e-5f);
    float y[8];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        float4 const g = reinterpret_cast<float4 const*>(gamma)[half * 32 + lane];
        float4 const bt = reinterpret_cast<float4 const*>(beta)[half * 32 + lane];
        y[4 * half + 0] = (v[4 * half + 0] - mean) * rstd * g.x + bt.x;
        y[4 * half + 1] = (v[4 * half + 1] - mean) * rstd * g.y + bt.y;
        y[4 * half + 2] = (v[4 * half + 2] - mean) * rstd * g.z + bt.z;
        y[4 * half + 3] = (v[4 * half + 3] - mean) * rstd * g.w + bt.w;
        reinterpret_cast<float4*>(out + row * kDim)[half * 32 + lane] =
            make_float4(y[4 * half], y[4 * half + 1], y[4 * half + 2], y[4 * half + 3]);
    }
    if (out2) {
        float const* q = pos + (row % pos_mod) * kDim;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float4 const a = reinterpret_cast<float4 const*>(q)[half * 32 + lane];
            reinterpret_cast<float4*>(out2 + row * kDim)[half * 32 + lane] =
                make_float4(y[4 * half] + a.x, y[4 * half + 1] + a.y, y[4 * half + 2] + a.z, y[4 * half + 3] + a.w);
        }
    }
}

// In-place LayerNorm2d over groups of 64 channels (eps 1e-6) + exact GELU on 16-bit rows of 64: eight lanes per row
// (16 bytes each), four rows per warp.
__global__ void __launch_bounds__(256) layernorm64_gelu_kernel(act_t* __restrict__ x, int64_t rows,
                                                               float const* __restrict__ gamma,
                                                               float const* __restrict__ beta) {
    int64_t const row = ((int64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * 4 + ((threadIdx.x & 31) >> 3);
    int const l8 = threadIdx.x & 7;
    bool const ok = row < rows;
    uint4 xa = make_uint4(0, 0, 0, 0);
    if (ok) xa = *(reinterpret_cast<uint4 const*>(x + row * 64) + l8);
    act2_t const* xh = reinterpret_cast<act2_t const*>(&xa);
    float v[8];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float2 const a = act22f2(xh[j]);
        v[2 * j] = a.x;
        v[2 * j + 1] = a.y;
        sum += a.x + a.y;
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    float const mean = sum * (1.0f / 64.0f);
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float const d = v[i] - mean;
        var = fmaf(d, d, var);
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
    float const rstd = rsqrtf(var * (1.0f / 64.0f) + 1e-6f);
    float4 const g0 = __ldg(reinterpret_cast<float4 const*>(gamma) + 2 * l8), g1 = __ldg(reinterpret_cast<float4 const*>(gamma) + 2 * l8 + 1);
    float4 const b0 = __ldg(reinterpret_cast<float4 const*>(beta) + 2 * l8), b1 = __ldg(reinterpret_cast<float4 const*>(beta) + 2 * l8 + 1);
    float const gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    float const bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    uint4 ov;
    act2_t* oh = reinterpret_cast<act2_t*>(&ov);
#pragma unroll
    for (int j = 0; j < 4; ++j)
        oh[j] = f22act2(gelu_erf((v[2 * j] - mean) * rstd * gg[2 * j] + bb[2 * j]),
                        gelu_erf((v[2 * j + 1] - mean) * rstd * gg[2 * j + 1] + bb[2 * j + 1]));
    if (ok) *(reinterpret_cast<uint4*>(x + row * 64) + l8) = ov;
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mask_dot_kernel(float const* __restrict__ hyper, act_t const* __restrict__ up2,
                                                       float* __restrict__ low) {
    __shared__ float hs[4 * 32];
    int const p = blockIdx.y;
    if (threadIdx.x < 128) hs[threadIdx.x] = hyper[(size_t)p * 128 + threadIdx.x];
    __syncthreads();
    int const pix = blockIdx.x * blockDim.x + threadIdx.x;  // Y*256 + X
    int const Y = pix >> 8, X = pix & 255;
    int const y = Y >> 2, dy = (Y >> 1) & 1, ey = Y & 1;
    int const x = X >> 2, dx = (X >> 1) & 1, ex = X & 1;
    size_t const row = ((size_t)(y * 64 + x) * 4 + dy * 2 + dx);
    uint4 const* u4 = reinterpret_cast<uint4 const*>(up2 + ((size_t)p * 16384 + row) * 128 + (ey * 2 + ex) * 32);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint4 const u = __ldg(u4 + c);
        act2_t const* uh = reinterpret_cast<act2_t const*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float2 const f = act22f2(uh[j]);
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                float const* hm = hs + m * 32 + c * 8 + 2 * j;
                acc[m] = fmaf(f.x, hm[0], acc[m]);
                acc[m] = fmaf(f.y, hm[1], acc[m]);
            }
        }
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) low[((size_t)p * 4 + m) * 65536 + pix] = acc[m];
}

__global__ void f32_to_act_kernel(float const* __restrict__ in, int64_t n, act_t* __restrict__ out) {
    int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = f2act(in[i]);
}

__global__ void select_masks_kernel(float const* __restrict__ iou, int P, int multi, int* __restrict__ plane_index,
                                    float* __restrict__ iou_out) {
    int const p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    float const* s = iou + p * 4;
    if (multi) {
        for (int i = 0; i < 3; ++i) {
            plane_index[p * 3 + i] = p * 4 + 1 + i;
            if (iou_out) iou_out[p * 3 + i] = s[1 + i];
        }
    } else {
        int const bi = best_mask_index(s);  // mask_select.cuh
        plane_index[p] = p * 4 + bi;
        if (iou_out) iou_out[p] = s[bi];
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
void prompt_tokens(cudaStream_t s, float const* coords, float const* labels, int P, PromptParams const& pp, float* tokens,
                   float* queries) {
    ProfScope prof(s, CAT_DEC_MISC);
    prompt_tokens_kernel<<<P, 256, 0, s>>>(coords, labels, pp, tokens, queries);
    KERNEL_CHECK();
}

void dense_pe(cudaStream_t s, float const* gaussian, float* pos) {
    ProfScope prof(s, CAT_DEC_MISC);
    dense_pe_kernel<<<kImgTokens, 256, 0, s>>>(gaussian, pos);
    KERNEL_CHECK();
}

void linear_small(cudaStream_t s, float const* x, int64_t x_stride, float const* x2, int64_t x2_stride, int rows, int K,
                  float const* W, float const* b, int N, bool relu, float* y, int64_t y_stride) {
    ProfScope prof(s, CAT_DEC_LINEAR);
    DLIMG_ASSERT(K <= 2048);
    static bool attr_set = false;
    if (!attr_set) {
        CUDA_CHECK(cudaFuncSetAttribute(linear_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kLinRows * 2048 * 4));
        attr_set = true;
    }
    dim3 grid(ceil_div(rows, kLinRows), ceil_div(N, kLinColsPerBlock));
    linear_small_kernel<<<grid, kLinWarps * 32, sizeof(float) * kLinRows * K, s>>>(x, x_stride, x2, x2_stride, rows, K, W, b,
                                                                                  N, relu ? 1 : 0, y, y_stride);
    KERNEL_CHECK();
}

void token_self_attention(cudaStream_t s, float const* q, float const* k, float const* v, int P, float* out) {
    ProfScope prof(s, CAT_DEC_ATTN);
    token_self_attention_kernel<<<P, 256, 0, s>>>(q, k, v, out);
    KERNEL_CHECK();
}

void image_to_token_attention(cudaStream_t s, act_t const* Q, act_t const* const* Qptrs, int64_t q_prompt_stride, int q_pitch,
                              int q_off, float const* kt, float const* vt, int P, act_t* out) {
    ProfScope prof(s, CAT_DEC_ATTN, 4.0 * P * kTokens * kImgTokens * 128, (double)P * kImgTokens * 128 * 4);
    dim3 grid(kImgTokens * kHeads / 256, P);
    i2t_attention_kernel<<<grid, 256, 0, s>>>(Q, Qptrs, q_prompt_stride, q_pitch, q_off, kt, vt, out);
    KERNEL_CHECK();
}

void layernorm256_img(cudaStream_t s, act_t const* x, act_t const* res, act_t const* const* res_ptrs, int P, float const* gamma,
                      float const* beta, act_t* out) {
    int64_t const rows = (int64_t)P * kImgTokens;
    ProfScope prof(s, CAT_DEC_NORM, 0, (double)rows * kDim * 6);
    layernorm256_img_kernel<<<(unsigned)ceil_div64(rows, 8), 256, 0, s>>>(x, res, res_ptrs, rows, gamma, beta, out);
    KERNEL_CHECK();
}

void layernorm256(cudaStream_t s, float const* x, float const* res, int64_t res_mod, int64_t rows, float const* gamma,
                  float const* beta, float const* pos, int64_t pos_mod, float* out, float* out2) {
    ProfScope prof(s, CAT_DEC_NORM);
    if (res_mod <= 0) res_mod = rows;
    if (pos_mod <= 0) pos_mod = rows;
    layernorm256_kernel<<<(unsigned)ceil_div64(rows, 8), 256, 0, s>>>(x, res, res_mod, rows, gamma, beta, pos, pos_mod, out,
                                                                     out2);
    KERNEL_CHECK();
}

void layernorm64_gelu(cudaStream_t s, act_t* x, int64_t rows, float const* gamma, float const* beta) {
    ProfScope prof(s, CAT_DEC_NORM, 0, (double)rows * 64 * 4);
    layernorm64_gelu_kernel<<<(unsigned)ceil_div64(rows, 32), 256, 0, s>>>(x, rows, gamma, beta);
    KERNEL_CHECK();
}

void mask_dot(cudaStream_t s, float const* hyper, act_t const* up2, int P, float* low) {
    ProfScope prof(s, CAT_DEC_MISC, 2.0 * P * 65536 * 128, (double)P * (16384.0 * 128 * 2 + 4 * 65536 * 4));
    dim3 grid(65536 / 256, P);
    mask_dot_kernel<<<grid, 256, 0, s>>>(hyper, up2, low);
    KERNEL_CHECK();
}

void f32_to_act(cudaStream_t s, float const* in, int64_t n, act_t* out) {
    ProfScope prof(s, CAT_DEC_MISC);
    f32_to_act_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, s>>>(in, n, out);
    KERNEL_CHECK();
}

void select_masks(cudaStream_t s, float const* iou, int P, int multi, int* plane_index, float* iou_out) {
    ProfScope prof(s, CAT_DEC_MISC);
    select_masks_kernel<<<ceil_div(P, 128), 128, 0, s>>>(iou, P, multi, plane_index, iou_out);
    KERNEL_CHECK();
}

}  // namespace dec
}  // namespace dlimg
