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

#if DLIMG_B200_ALT  // CUDA-core form, kept as the cross-check of i2t_mma_kernel (DLIMG_B200_I2T_SIMT)
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
#endif

__global__ void f32_to_act_kernel(float const* __restrict__ in, int64_t n, act_t* __restrict__ out) {
    int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = f2act(in[i]);
}

// eight values per thread: two 16-byte loads, one 16-byte store (n % 8 == 0, both pointers 16-byte aligned)
__global__ void f32_to_act8_kernel(float4 const* __restrict__ in, int64_t n8, uint4* __restrict__ out) {
    int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    float4 const a = __ldg(in + 2 * i), b = __ldg(in + 2 * i + 1);
    act2_t const p0 = f22act2(a.x, a.y), p1 = f22act2(a.z, a.w), p2 = f22act2(b.x, b.y), p3 = f22act2(b.z, b.w);
    uint4 o;
    o.x = *reinterpret_cast<uint32_t const*>(&p0);
    o.y = *reinterpret_cast<uint32_t const*>(&p1);
    o.z = *reinterpret_cast<uint32_t const*>(&p2);
    o.w = *reinterpret_cast<uint32_t const*>(&p3);
    out[i] = o;
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

#if DLIMG_B200_ALT
void image_to_token_attention(cudaStream_t s, act_t const* Q, act_t const* const* Qptrs, int64_t q_prompt_stride, int q_pitch,
                              int q_off, float const* kt, float const* vt, int P, act_t* out) {
    ProfScope prof(s, CAT_DEC_ATTN, 4.0 * P * kTokens * kImgTokens * 128, (double)P * kImgTokens * 128 * 4);
    dim3 grid(kImgTokens * kHeads / 256, P);
    i2t_attention_kernel<<<grid, 256, 0, s>>>(Q, Qptrs, q_prompt_stride, q_pitch, q_off, kt, vt, out);
    KERNEL_CHECK();
}
#endif

void f32_to_act(cudaStream_t s, float const* in, int64_t n, act_t* out) {
    ProfScope prof(s, CAT_DEC_MISC);
    if (n % 8 == 0 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0)
        f32_to_act8_kernel<<<(unsigned)ceil_div64(n / 8, 256), 256, 0, s>>>(reinterpret_cast<float4 const*>(in), n / 8, reinterpret_cast<uint4*>(out));
    else
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
