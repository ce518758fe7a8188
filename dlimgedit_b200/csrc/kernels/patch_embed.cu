// patch_embed.cu -- TinyViT PatchEmbed as ONE kernel: u8 image -> (B, 256, 256, 64) 16-bit activations.
//
//   preprocess (channel map, (x - mean) / std, zero pad to 1024^2)
//   conv 3x3 s2 p1 3 -> 32 (+BN) + GELU          mma.sync m16n8k16 on a tile held in shared memory
//   conv 3x3 s2 p1 32 -> 64 (+BN)                implicit GEMM: im2col rows assembled in shared memory in the
//                                                128-byte-swizzled K-major layout, tcgen05.mma with the accumulator
//                                                in TMEM, weights resident in shared memory (TMA, once per CTA)
//
// The unfused form wrote the 512x512x32 intermediate (134 MB per 8 images), re-read it, wrote a 302 MB im2col
// buffer and read that again (972 MB of traffic, 397 us per 8 images in profiles/r01c); fused, the kernel reads the
// u8 pixels and writes the 64-channel output (100 MB).
//
// A CTA loops over tiles of 8 x 16 output pixels (= the 128 rows of one UMMA tile):
//   P1  35 x 67 input pixels -> normalised fp16 in shared memory
//   P2  17 x 33 conv1 outputs (the halo conv2 needs) -> GELU -> shared memory; positions outside the 512^2 map are
//       conv2's zero padding
//   P3  gather into the A operand: row r = (ty, tx), K = (tap, ci) with tap = ky*3+kx; k-block kb holds taps 2kb, 2kb+1
//   P4  one thread issues the 20 MMAs (K = 320, the last 32 are zero), commit -> mbarrier
//   P5  all warps: TMEM -> + bias -> 16-bit -> global
#include "encoder_kernels.cuh"
#include "gelu.cuh"
#include "tcgen05.cuh"

#include "../profiler.hpp"

namespace dlimg {
namespace enc {

namespace {

using namespace tc;

constexpr int kTH = 8, kTW = 16;               // output tile
constexpr int kC1H = 2 * kTH + 1, kC1W = 2 * kTW + 1;  // 17 x 33 conv1 outputs
constexpr int kInH = 2 * kC1H + 1, kInW = 2 * kC1W + 1;  // 35 x 67 input pixels
constexpr int kInPitch = 68;                   // pixels per tile row (even: keeps (k, k+1) pairs 4-byte aligned)
constexpr int kC1Pitch = 80;                   // bytes per conv1 pixel in shared memory (64 + 16; the data starts at 0 or 16)
constexpr int kThreads = 576;                  // 18 warps: the 36 conv1 mma tiles split evenly
constexpr int kKBlocks = 5;                    // K = 9 taps * 32 channels = 288, padded to 320
constexpr int kABlockBytes = 128 * 128;        // one k-block of A: 128 rows x 128 B
constexpr int kBBlockBytes = 64 * 128;         // one k-block of W2: 64 rows x 128 B
constexpr int kSmemW2 = 0;
constexpr int kSmemA = kSmemW2 + kKBlocks * kBBlockBytes;            // 40960
constexpr int kSmemC1 = kSmemA + kKBlocks * kABlockBytes;            // 122880
constexpr int kSmemIn = kSmemC1 + ((kC1H * kC1W * kC1Pitch + 127) / 128) * 128;
constexpr int kSmemBar = kSmemIn + ((kInH * kInPitch * 3 * 2 + 16 + 127) / 128) * 128;
constexpr int kSmemBytes = kSmemBar + 64 + 1024 /* alignment slack */;

struct PatchParams {
    int w, h, bpp;
    int c0, c1, c2;  // byte offsets of R, G, B in a pixel
    int tiles;       // batch * 32 * 16
    int sel;         // __byte_perm selector picking (c0, c1, c2) out of the raw pixel word
};

__device__ __forceinline__ void mma16816_f16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                             uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// conv1 reduction index k' = ky * 10 + (kx * 3 + ci) (slot 9 of each ky and k' = 30, 31 are padding with zero weights)
// -> offset in halves from the top-left input pixel of the 3x3 window
__device__ __forceinline__ int conv1_k_offset(int k) {
    if (k >= 30) k = 28;  // padding: any valid address, its weight is zero
    return (k / 10) * (kInPitch * 3) + (k % 10);
}

template <bool kDebug>
__global__ void __launch_bounds__(kThreads, 1)
patch_embed_kernel(ImageDesc const* __restrict__ imgs, PatchParams p, uint32_t const* __restrict__ w1_frag,
                   float const* __restrict__ b1, const __grid_constant__ CUtensorMap w2_map, float const* __restrict__ b2,
                   __half* __restrict__ out, __half* __restrict__ c1_debug) {
    extern __shared__ uint8_t smem_raw[];
    uint32_t const base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* const gen = smem_raw + (base - smem_u32(smem_raw));  // generic pointer to the aligned base
    uint32_t const w2_s = base + kSmemW2, a_s = base + kSmemA, c1_s = base + kSmemC1, in_s = base + kSmemIn;
    uint32_t const bar_w = base + kSmemBar, bar_mma = bar_w + 8, tmem_slot = bar_w + 16;
    __half* const in_g = reinterpret_cast<__half*>(gen + kSmemIn);
    int const tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int const g = lane >> 2, t = lane & 3;

    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&w2_map) : "memory");
        mbar_init(bar_w, 1);
        mbar_init(bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(tmem_slot) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // A: zero once (the upper half of k-block 4 stays zero for ever); the input tile's pad column too
    for (int i = tid; i < kKBlocks * kABlockBytes / 16; i += kThreads)
        reinterpret_cast<uint4*>(gen + kSmemA)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < (kSmemBar - kSmemIn) / 16; i += kThreads)
        reinterpret_cast<uint4*>(gen + kSmemIn)[i] = make_uint4(0, 0, 0, 0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();  // first kernel of an encoder pass: its output buffer may still be read by the previous pass
    pdl_trigger();
    uint32_t const tmem = *reinterpret_cast<uint32_t const*>(gen + kSmemBar + 16);
    if (tid == 0) {  // conv2 weights: 5 k-blocks of (64 x 64) halves, resident for the whole kernel
        mbar_expect_tx(bar_w, kKBlocks * kBBlockBytes);
        for (int kb = 0; kb < kKBlocks; ++kb) tma_load_2d(w2_s + kb * kBBlockBytes, &w2_map, bar_w, kb * 64, 0);
    }

    // conv1 weights as mma B fragments, bias of this thread's channels
    uint32_t wb[2][4][2];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
            uint2 const v = __ldg(reinterpret_cast<uint2 const*>(w1_frag) + (ks * 4 + nb) * 32 + lane);
            wb[ks][nb][0] = v.x;
            wb[ks][nb][1] = v.y;
        }
    float2 bias1[4];
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) bias1[nb] = __ldg(reinterpret_cast<float2 const*>(b1 + nb * 8 + 2 * t));
    int koff[2][2];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        koff[ks][0] = conv1_k_offset(ks * 16 + 2 * t);
        koff[ks][1] = conv1_k_offset(ks * 16 + 8 + 2 * t);
    }
    float const mean[3] = {123.675f, 116.28f, 103.53f};
    float const inv_sd[3] = {1.0f / 58.395f, 1.0f / 57.12f, 1.0f / 57.375f};  // the result is rounded to fp16 right away
    uint32_t const idesc = make_idesc(0u, 128, 64);
    // P3 / P5 geometry of this thread
    // A row (output pixel) and 16-byte piece of the 64-byte tap.  A quarter-warp holds 8 consecutive rows and ONE piece:
    // its swizzled STS.128 then hit eight different 16-byte bank groups (rows x 4 pieces per quarter-warp collided 2-way)
    int const a_row = (tid >> 5) * 8 + (tid & 7), a_cc = (tid >> 3) & 3;
    int const a_ty = a_row >> 4, a_tx = a_row & 15;
    int const quarter = warp & 3, col0 = (warp >> 2) * 16;    // epilogue: TMEM lane quarter, 16 of the 64 columns

    // P1 is split in two: the u8 loads of tile i+1 are issued right after tile i's pixels have been consumed (they are
    // in flight during P2..P4 of tile i), the conversion + shared-memory store happens at the top of the next iteration.
    constexpr int kPxIters = (kInH * kInW + kThreads - 1) / kThreads;
    // Raw registers only: nothing may consume a loaded value before the next iteration, or the warp would sit out the
    // DRAM latency here.  Aligned 4-byte pixels (the common case) take one 32-bit load; the channel order is applied at
    // conversion time.  Other layouts pack three byte loads (and do wait here).
    uint32_t pxr[kPxIters];
    uint32_t pxvalid = 0;  // bit it: pixel `it` lies inside the image
    bool const fast_px = p.bpp == 4;
    auto fetch_pixels = [&](int tile) {
        int const b = tile >> 9, tr = tile & 511;
        int const iy0 = 4 * (tr >> 4) * kTH - 3, ix0 = 4 * (tr & 15) * kTW - 3;
        ImageDesc const img = imgs[b];
        bool const aligned = fast_px && ((reinterpret_cast<uintptr_t>(img.pixels) | (uintptr_t)img.stride) & 3u) == 0;
        pxvalid = 0;
#pragma unroll
        for (int it = 0; it < kPxIters; ++it) {
            int const i = tid + it * kThreads;
            int const r = i / kInW, c = i - r * kInW;
            int const iy = iy0 + r, ix = ix0 + c;
            uint32_t v = 0;
            if (i < kInH * kInW && iy >= 0 && iy < p.h && ix >= 0 && ix < p.w) {
                uint8_t const* px = img.pixels + (size_t)iy * img.stride + (size_t)ix * p.bpp;
                pxvalid |= 1u << it;
                if (aligned) v = __ldg(reinterpret_cast<uint32_t const*>(px));
                else {
                    v = (uint32_t)__ldg(px);
                    if (p.bpp >= 3) v |= ((uint32_t)__ldg(px + 1) << 8) | ((uint32_t)__ldg(px + 2) << 16);
                    if (p.bpp == 4) v |= (uint32_t)__ldg(px + 3) << 24;
                }
            }
            pxr[it] = v;
        }
    };
    auto store_pixels = [&]() {
#pragma unroll
        for (int it = 0; it < kPxIters; ++it) {
            int const i = tid + it * kThreads;
            if (i >= kInH * kInW) continue;
            int const r = i / kInW, c = i - r * kInW;
            uint32_t const v = __byte_perm(pxr[it], 0u, (uint32_t)p.sel);  // -> R | G << 8 | B << 16
            float v0 = 0.f, v1 = 0.f, v2 = 0.f;
            if ((pxvalid >> it) & 1u) {  // zero outside the image: padding happens after normalisation
                v0 = ((float)(v & 255u) - mean[0]) * inv_sd[0];
                v1 = ((float)((v >> 8) & 255u) - mean[1]) * inv_sd[1];
                v2 = ((float)((v >> 16) & 255u) - mean[2]) * inv_sd[2];
            }
            __half* d = in_g + (r * kInPitch + c) * 3;
            d[0] = __float2half_rn(v0);
            d[1] = __float2half_rn(v1);
            d[2] = __float2half_rn(v2);
        }
    };
    // P5 of a finished tile: accumulator -> + bias -> 16-bit -> global
    auto epilogue = [&](int tile) {
        if (warp >= 16) return;
        int const b = tile >> 9, tr = tile & 511;
        int const oy0 = (tr >> 4) * kTH, ox0 = (tr & 15) * kTW;
        uint32_t r[16];
        tmem_ld16(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)col0, r);
        tmem_ld_wait();
        int const row = quarter * 32 + lane;
        int const oy = oy0 + (row >> 4), ox = ox0 + (row & 15);
        float4 const* bb = reinterpret_cast<float4 const*>(b2 + col0);
        uint4 o[2];
        __half2* oh = reinterpret_cast<__half2*>(o);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float4 const bv = __ldg(bb + i);
            oh[2 * i] = f22act2(__uint_as_float(r[4 * i]) + bv.x, __uint_as_float(r[4 * i + 1]) + bv.y);
            oh[2 * i + 1] = f22act2(__uint_as_float(r[4 * i + 2]) + bv.z, __uint_as_float(r[4 * i + 3]) + bv.w);
        }
        uint4* dst = reinterpret_cast<uint4*>(out + (((size_t)b * 256 + oy) * 256 + ox) * 64 + col0);
        dst[0] = o[0];
        dst[1] = o[1];
    };

    uint32_t mma_phase = 0;
    bool w2_ready = false;
    int prev_tile = -1;
    if ((int)blockIdx.x < p.tiles) fetch_pixels(blockIdx.x);
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        int const tr = tile & 511;
        int const b = tile >> 9;
        int const oy0 = (tr >> 4) * kTH, ox0 = (tr & 15) * kTW;

        // ---- P1: input pixels (fetched during the previous tile) -> normalised fp16 tile ----
        store_pixels();
        __syncthreads();
        if (tile + (int)gridDim.x < p.tiles) fetch_pixels(tile + gridDim.x);

        // ---- P2: conv1 + GELU on the 17 x 33 halo, 16 positions per mma tile ----
        for (int mt = warp; mt < (kC1H * kC1W + 15) / 16; mt += kThreads / 32) {
            int q[2], qc[2];
            uint32_t pix[2];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                q[hh] = mt * 16 + g + 8 * hh;
                qc[hh] = min(q[hh], kC1H * kC1W - 1);
                int const cy = qc[hh] / kC1W, cx = qc[hh] - cy * kC1W;
                pix[hh] = in_s + (uint32_t)(((2 * cy) * kInPitch + 2 * cx) * 3 * 2);
            }
            float acc[4][4];
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) acc[nb][0] = acc[nb][1] = acc[nb][2] = acc[nb][3] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                uint32_t a[4];
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(a[hh]) : "r"(pix[hh] + (uint32_t)(koff[ks][0] * 2)));
                    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(a[2 + hh]) : "r"(pix[hh] + (uint32_t)(koff[ks][1] * 2)));
                }
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) mma16816_f16(acc[nb], a[0], a[1], a[2], a[3], wb[ks][nb][0], wb[ks][nb][1]);
            }
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                if (q[hh] >= kC1H * kC1W) continue;
                int const cy = q[hh] / kC1W, cx = q[hh] - cy * kC1W;
                int const Y = 2 * oy0 - 1 + cy, X = 2 * ox0 - 1 + cx;  // position in the 512 x 512 conv1 map
                bool const inside = Y >= 0 && X >= 0;                   // the upper bounds cannot be exceeded
                // pixels 8..15, 24..31 of a row sit 16 bytes further into their 80-byte slot: the P3 reads of 8
                // consecutive output pixels (conv1 x stride 2 = 160 bytes) then cover all eight 16-byte bank groups
                uint32_t const dst = c1_s + (uint32_t)(q[hh] * kC1Pitch + ((cx >> 3) & 1) * 16 + 2 * t * 2);
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) {
                    __half2 v = gelu_erf_h2(__floats2half2_rn(acc[nb][2 * hh] + bias1[nb].x, acc[nb][2 * hh + 1] + bias1[nb].y));
                    if (!inside) v = __float2half2_rn(0.f);
                    asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst + (uint32_t)(nb * 16)), "r"(*reinterpret_cast<uint32_t*>(&v)) : "memory");
                    if (kDebug && inside && cy >= 1 && cx >= 1)  // interior of the halo: each position owned by one tile
                        *reinterpret_cast<__half2*>(c1_debug + (((size_t)b * 512 + Y) * 512 + X) * 32 + nb * 8 + 2 * t) = v;
                }
            }
        }

        // ---- P5 of the previous tile: its MMAs had all of P1 + P2 to finish; frees A and TMEM ----
        if (prev_tile >= 0) {
            mbar_wait(bar_mma, mma_phase);
            mma_phase ^= 1u;
            tc_fence_after();
            epilogue(prev_tile);
            tc_fence_before();
        }
        __syncthreads();

        // ---- P3: gather the im2col rows into the swizzled A operand ----
        if (tid < 512) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                int const cy = 2 * a_ty + tap / 3, cx = 2 * a_tx + tap % 3;
                uint4 v;
                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                             : "r"(c1_s + (uint32_t)((cy * kC1W + cx) * kC1Pitch + ((cx >> 3) & 1) * 16 + a_cc * 16)));
                int const j = (tap & 1) * 4 + a_cc;  // 16-byte piece within the 128-byte row of k-block tap / 2
                uint32_t const dst = a_s + (uint32_t)((tap >> 1) * kABlockBytes + a_row * 128 + ((j ^ (a_row & 7)) << 4));
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
        }
        __syncthreads();

        // ---- P4: conv2 as 128 x 64 x 320 on the tensor core (asynchronous; collected in the next iteration) ----
        if (tid < 32) {  // converged warp + one elected lane: the 20 MMAs issue back to back from uniform registers
            if (!w2_ready) mbar_wait(bar_w, 0);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int kb = 0; kb < kKBlocks; ++kb) {
                    uint64_t const adesc = make_smem_desc(a_s + kb * kABlockBytes);
                    uint64_t const bdesc = make_smem_desc(w2_s + kb * kBBlockBytes);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc_mma<0>(tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
                }
                tc_commit(bar_mma);
            }
            __syncwarp();
        }
        w2_ready = true;
        prev_tile = tile;
    }
    if (prev_tile >= 0) {
        mbar_wait(bar_mma, mma_phase);
        tc_fence_after();
        epilogue(prev_tile);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

}  // namespace

void patch_embed_w1_fragments(float const* w27x32, uint32_t* out512) {
    // w27x32: [(ky*3+kx)*3+ci][oc] (BN folded).  Fragment order [ks][nb][lane][2] for mma.m16n8k16 B (col):
    // register i holds B[k][n], B[k+1][n] with k = ks*16 + i*8 + 2t, n = nb*8 + g, k in the padded order k' = ky*10 + kx*3 + ci.
    auto weight = [&](int k, int n) -> float {
        if (k >= 30) return 0.f;
        int const ky = k / 10, j = k % 10;
        if (j >= 9) return 0.f;
        return w27x32[(size_t)(ky * 9 + j) * 32 + n];
    };
    for (int ks = 0; ks < 2; ++ks)
        for (int nb = 0; nb < 4; ++nb)
            for (int lane = 0; lane < 32; ++lane)
                for (int i = 0; i < 2; ++i) {
                    int const g = lane >> 2, t = lane & 3;
                    int const k = ks * 16 + i * 8 + 2 * t, n = nb * 8 + g;
                    __half const lo = __float2half_rn(weight(k, n)), hi = __float2half_rn(weight(k + 1, n));
                    out512[((ks * 4 + nb) * 32 + lane) * 2 + i] =
                        (uint32_t)(*reinterpret_cast<uint16_t const*>(&lo)) | ((uint32_t)(*reinterpret_cast<uint16_t const*>(&hi)) << 16);
                }
}

void patch_embed(cudaStream_t s, ImageDesc const* imgs, int batch, int w, int h, int channels, uint32_t const* w1_frag,
                 float const* b1, CUtensorMap const& w2_map, float const* b2, act_t* out, act_t* c1_debug, int num_sms) {
#if defined(DLIMG_B200_ACT_BF16)
    (void)imgs; (void)batch; (void)w; (void)h; (void)channels; (void)w1_frag; (void)b1; (void)w2_map; (void)b2; (void)out;
    (void)c1_debug; (void)num_sms; (void)s;
    fail("patch_embed: the fused kernel is built for fp16 activations");
#else
    DLIMG_ASSERT(w >= 1 && h >= 1 && w <= kImageSize && h <= kImageSize);
    int cmap[3];
    channel_map(channels, cmap);
    PatchParams p{w, h, bytes_per_pixel(channels), cmap[0], cmap[1], cmap[2], batch * 512,
                  cmap[0] | (cmap[1] << 4) | (cmap[2] << 8) | (4 << 12)};  // byte 3 <- 0
    ProfScope prof(s, CAT_CONV1, 2.0 * batch * (512.0 * 512 * 27 * 32 + 256.0 * 256 * 288 * 64),
                   (double)batch * ((double)w * h * p.bpp + 256.0 * 256 * 64 * 2));
    static bool attr_set = false;
    if (!attr_set) {
        CUDA_CHECK(cudaFuncSetAttribute(patch_embed_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        CUDA_CHECK(cudaFuncSetAttribute(patch_embed_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        attr_set = true;
    }
    int const grid = p.tiles < num_sms ? p.tiles : num_sms;
    if (c1_debug) launch_pdl(PDL_PATCH_EMBED, patch_embed_kernel<true>, dim3(grid), dim3(kThreads), (size_t)kSmemBytes, s, imgs, p, w1_frag, b1, w2_map, b2, out, c1_debug);
    else launch_pdl(PDL_PATCH_EMBED, patch_embed_kernel<false>, dim3(grid), dim3(kThreads), (size_t)kSmemBytes, s, imgs, p, w1_frag, b1, w2_map, b2, out, (act_t*)nullptr);
    KERNEL_CHECK();
#endif
}

}  // namespace enc
}  // namespace dlimg
