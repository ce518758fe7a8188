// window_attention.cu -- TinyViT windowed attention (49- or 196-token windows, head_dim 32) on tensor cores.
//
// One CTA per (window, head).  Q, K and V^T of that head sit in shared memory as 16-bit values; each warp
// owns 16-query tiles: S = Q K^T with mma.sync m16n8k16 (fp32 accumulators, whole S row block in registers),
// learned relative-position bias + softmax in registers (quad shuffles), P re-used in place as the A operand
// of O = P V.  These windows are far below a 128-row tcgen05 tile (SURVEY section 7 "hard parts"), and the
// attention core is only ~3-19 % of a stage's MACs, so the legacy warp-level MMA is the right tool here; the
// tcgen05 kernel (gemm.cu) carries the QKV / proj / MLP GEMMs around it.
#include "encoder_kernels.cuh"

#include "../profiler.hpp"

#include <cmath>

namespace dlimg {
namespace enc {

namespace {

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
#if defined(DLIMG_B200_ACT_BF16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
#else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
#endif
}

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    act2_t const v = f22act2(a, b);
    return *reinterpret_cast<uint32_t const*>(&v);
}

constexpr int kWarps = 4;
constexpr int kQKStride = 40;  // 16-bit elements per Q / K row (32 + 8 padding: conflict-free fragment loads)

// kNPad: window tokens rounded up to a multiple of 16 (49 -> 64, 196 -> 208)
template <int kNPad>
__global__ void __launch_bounds__(kWarps * 32) window_attention_mma_kernel(act_t const* __restrict__ qkv, int n, int heads,
                                                                           float const* __restrict__ bias,
                                                                           act_t* __restrict__ out) {
    constexpr int kVStride = kNPad + 8;  // 16-bit elements per V^T row
    extern __shared__ __align__(16) uint8_t smem_raw[];
    act_t* Qs = reinterpret_cast<act_t*>(smem_raw);  // [kNPad][kQKStride]
    act_t* Ks = Qs + kNPad * kQKStride;              // [kNPad][kQKStride]
    act_t* Vt = Ks + kNPad * kQKStride;              // [32][kVStride]

    int const win = blockIdx.x / heads, h = blockIdx.x % heads;
    int const ld = heads * 96, C = heads * 32;
    int64_t const row0 = (int64_t)win * n;
    int const tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int const g = lane >> 2, t = lane & 3;

    // ---- stage Q, K (row-major) and V (transposed) of this head; padded tokens are zero ----
    for (int i = tid; i < kNPad * 12; i += kWarps * 32) {
        int const j = i / 12, part = i % 12;  // 12 x 16-byte chunks per token: 4 of q, 4 of k, 4 of v
        uint4 v = make_uint4(0, 0, 0, 0);
        if (j < n) v = *reinterpret_cast<uint4 const*>(qkv + (row0 + j) * ld + h * 96 + part * 8);
        if (part < 4) {
            *reinterpret_cast<uint4*>(Qs + j * kQKStride + part * 8) = v;
        } else if (part < 8) {
            *reinterpret_cast<uint4*>(Ks + j * kQKStride + (part - 4) * 8) = v;
        } else {
            act_t const* e = reinterpret_cast<act_t const*>(&v);
            int const d0 = (part - 8) * 8;
#pragma unroll
            for (int k = 0; k < 8; ++k) Vt[(d0 + k) * kVStride + j] = e[k];
        }
    }
    __syncthreads();

    float const scale = 0.17677669529663687f;  // 32^-0.5
    // bias is pre-arranged in accumulator-fragment order: [head][query tile][key block][lane] x float4
    float4 const* bias_h = reinterpret_cast<float4 const*>(bias) + (size_t)h * (kNPad / 16) * (kNPad / 8) * 32;
    uint32_t const* Qw = reinterpret_cast<uint32_t const*>(Qs);
    uint32_t const* Kw = reinterpret_cast<uint32_t const*>(Ks);
    uint32_t const* Vw = reinterpret_cast<uint32_t const*>(Vt);

    for (int qt = warp; qt * 16 < n; qt += kWarps) {
        int const r0 = qt * 16 + g, r1 = r0 + 8;  // the two query rows this thread holds
        // Q fragments for the two 16-wide k steps
        uint32_t aq[2][4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            aq[ks][0] = Qw[(r0 * kQKStride + ks * 16 + 2 * t) >> 1];
            aq[ks][1] = Qw[(r1 * kQKStride + ks * 16 + 2 * t) >> 1];
            aq[ks][2] = Qw[(r0 * kQKStride + ks * 16 + 8 + 2 * t) >> 1];
            aq[ks][3] = Qw[(r1 * kQKStride + ks * 16 + 8 + 2 * t) >> 1];
        }
        // ---- S = Q K^T ----
        float s[kNPad / 8][4];
#pragma unroll
        for (int nb = 0; nb < kNPad / 8; ++nb) {
            s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f;
            int const key = nb * 8 + g;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                uint32_t const b0 = Kw[(key * kQKStride + ks * 16 + 2 * t) >> 1];
                uint32_t const b1 = Kw[(key * kQKStride + ks * 16 + 8 + 2 * t) >> 1];
                mma16816(s[nb], aq[ks][0], aq[ks][1], aq[ks][2], aq[ks][3], b0, b1);
            }
        }
        // ---- scale + bias (+ mask: the table holds -inf for padded key columns), row max ----
        float m0 = -INFINITY, m1 = -INFINITY;
        bool const v0 = r0 < n, v1 = r1 < n;
        float4 const* bias_q = bias_h + (size_t)qt * (kNPad / 8) * 32 + lane;
#pragma unroll
        for (int nb = 0; nb < kNPad / 8; ++nb) {
            float4 const bv = __ldg(bias_q + nb * 32);  // one coalesced 16-byte load per accumulator quad
            s[nb][0] = fmaf(s[nb][0], scale, bv.x);
            s[nb][1] = fmaf(s[nb][1], scale, bv.y);
            s[nb][2] = fmaf(s[nb][2], scale, bv.z);
            s[nb][3] = fmaf(s[nb][3], scale, bv.w);
            m0 = fmaxf(m0, fmaxf(s[nb][0], s[nb][1]));
            m1 = fmaxf(m1, fmaxf(s[nb][2], s[nb][3]));
        }
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
        // ---- exp + row sum ----
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int nb = 0; nb < kNPad / 8; ++nb) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                s[nb][e] = __expf(s[nb][e] - m0);
                s[nb][2 + e] = __expf(s[nb][2 + e] - m1);
                l0 += s[nb][e];
                l1 += s[nb][2 + e];
            }
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
        l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        // ---- O = P V (P fragments come straight from the S accumulators) ----
        float o[4][4];
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) o[nb][0] = o[nb][1] = o[nb][2] = o[nb][3] = 0.f;
#pragma unroll
        for (int kb = 0; kb < kNPad / 16; ++kb) {
            uint32_t const a0 = pack2(s[2 * kb][0], s[2 * kb][1]);
            uint32_t const a1 = pack2(s[2 * kb][2], s[2 * kb][3]);
            uint32_t const a2 = pack2(s[2 * kb + 1][0], s[2 * kb + 1][1]);
            uint32_t const a3 = pack2(s[2 * kb + 1][2], s[2 * kb + 1][3]);
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                int const d = nb * 8 + g;
                uint32_t const b0 = Vw[(d * kVStride + kb * 16 + 2 * t) >> 1];
                uint32_t const b1 = Vw[(d * kVStride + kb * 16 + 8 + 2 * t) >> 1];
                mma16816(o[nb], a0, a1, a2, a3, b0, b1);
            }
        }
        float const inv0 = 1.0f / l0, inv1 = 1.0f / l1;
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
            int const d = h * 32 + nb * 8 + 2 * t;
            if (v0) *reinterpret_cast<act2_t*>(out + (row0 + r0) * C + d) = f22act2(o[nb][0] * inv0, o[nb][1] * inv0);
            if (v1) *reinterpret_cast<act2_t*>(out + (row0 + r1) * C + d) = f22act2(o[nb][2] * inv1, o[nb][3] * inv1);
        }
    }
}

template <int kNPad>
void launch_mma(cudaStream_t s, act_t const* qkv, int windows, int n, int heads, float const* bias, act_t* out) {
    size_t const smem = sizeof(act_t) * ((size_t)2 * kNPad * kQKStride + (size_t)32 * (kNPad + 8));
    static bool attr_set = false;
    if (!attr_set) {
        CUDA_CHECK(cudaFuncSetAttribute(window_attention_mma_kernel<kNPad>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    window_attention_mma_kernel<kNPad><<<windows * heads, kWarps * 32, smem, s>>>(qkv, n, heads, bias, out);
    KERNEL_CHECK();
}

}  // namespace

void attention_bias_fragments(float const* dense, int heads, int n, float* out) {
    int const np = window_pad(n);
    size_t i = 0;
    for (int h = 0; h < heads; ++h)
        for (int qt = 0; qt < np / 16; ++qt)
            for (int nb = 0; nb < np / 8; ++nb)
                for (int lane = 0; lane < 32; ++lane)
                    for (int e = 0; e < 4; ++e) {
                        int const r = qt * 16 + (lane >> 2) + (e >= 2 ? 8 : 0);
                        int const c = nb * 8 + 2 * (lane & 3) + (e & 1);
                        float v = 0.f;
                        if (c >= n) v = -INFINITY;
                        else if (r < n) v = dense[((size_t)h * n + r) * n + c];
                        out[i++] = v;
                    }
}

void window_attention(cudaStream_t s, act_t const* qkv, int windows, int n, int heads, float const* bias, act_t* out) {
    ProfScope prof(s, CAT_WIN_ATTN, 4.0 * windows * heads * n * n * 32, (double)windows * n * heads * 128 * 2);
    if (n <= 64) launch_mma<64>(s, qkv, windows, n, heads, bias, out);
    else if (n <= 208) launch_mma<208>(s, qkv, windows, n, heads, bias, out);
    else fail("window_attention: unsupported window size " + std::to_string(n));
}

}  // namespace enc
}  // namespace dlimg
