// mbconv_tail.cu -- second half of a TinyViT MBConv block as ONE kernel:
//
//   depthwise 3x3 (+BN) + GELU on the 256 expanded channels      packed-fp16 CUDA-core math on a TMA-loaded halo tile
//   1x1 conv 256 -> 64 (+BN) + shortcut + GELU                   tcgen05.mma, A operand (the depthwise output) assembled
//                                                                in shared memory, accumulator in TMEM
//
// Unfused, the depthwise output (537 MB per 16 images) was written to HBM and read back by the project GEMM
// (245 + 143 us per MBConv in profiles/r01d).  Here the kernel reads the expanded tensor and the shortcut and writes
// the 64-channel result.
//
// A CTA loops over tiles of 8 x 16 output pixels (the 128 rows of one UMMA tile); each tile is processed as two
// channel halves (units) so that the 10 x 18-pixel halo of 128 channels (46 KB) can be double-buffered:
//   TMA   4-D box (128 ch, 18, 10, 1) at (half*128, x0-1, y0-1, image); out-of-bounds = the convolution's zero padding
//   DW    thread = (channel octet, 4-pixel run): 18 LDS.128, 144 HFMA2, packed-half erf GELU, 4 STS.128 into the
//         128B-swizzled K-major A operand (k-blocks 2*half, 2*half+1)
//   MMA   one thread: 2 k-blocks x 4 tcgen05.mma (M=128, N=64, K=16); after the second half, commit -> mbarrier
//   EPI   of the PREVIOUS tile (its MMAs had a whole unit to finish; accumulators are double-buffered in TMEM):
//         TMEM -> + bias + shortcut -> GELU -> global
#include "encoder_kernels.cuh"
#include "gelu.cuh"
#include "tcgen05.cuh"

#include "../profiler.hpp"

namespace dlimg {
namespace enc {

namespace {

using namespace tc;

constexpr int kTH = 8, kTW = 16, kHH = kTH + 2, kHW = kTW + 2;  // tile and halo
constexpr int kCexp = 256, kChalf = 128, kCout = 64;
constexpr int kThreads = 512;
constexpr int kInBytes = kHH * kHW * kChalf * 2;        // 46080: one halo tile of one channel half
constexpr int kABlockBytes = 128 * 128;                  // one k-block (64 channels) of A
constexpr int kWBlockBytes = kCout * 128;                // one k-block of the project weights
constexpr int kSmemW = 0;                                // 4 k-blocks: 32 KB
constexpr int kSmemA = kSmemW + 4 * kWBlockBytes;        // 2 k-blocks: 32 KB, rewritten by every unit once its MMAs retired
constexpr int kInBufs = 3;                               // halo tiles in flight: two units (~3 us) of prefetch lead
constexpr int kSmemIn = kSmemA + 2 * kABlockBytes;       // kInBufs buffers of 46080 B
constexpr int kSmemDw = kSmemIn + kInBufs * kInBytes;          // depthwise filter (9, 256) fp16 + bias (256) as packed fp16 pairs
constexpr int kSmemBar = kSmemDw + 9 * kCexp * 2 + kCexp * 2;  // barriers + TMEM slot
constexpr int kSmemBytes = kSmemBar + 128 + 1024;

#if !defined(DLIMG_B200_ACT_BF16)
__global__ void __launch_bounds__(kThreads, 1)
mbconv_tail_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap w3_map,
                   __half const* __restrict__ dw_w, float const* __restrict__ dw_b, float const* __restrict__ b3,
                   __half const* __restrict__ shortcut, __half* __restrict__ out, int tiles) {
    extern __shared__ uint8_t smem_raw[];
    uint32_t const base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* const gen = smem_raw + (base - smem_u32(smem_raw));
    uint32_t const w_s = base + kSmemW, a_s = base + kSmemA, in_s = base + kSmemIn, dw_s = base + kSmemDw, bar = base + kSmemBar;
    // barriers: w3 at 0, in_full[3] at 8..24, a_free at 32 (the unit's MMAs retired), acc_full[2] at 40,48; TMEM slot at 64
    uint32_t const bar_w = bar, bar_in0 = bar + 8, bar_afree = bar + 32, bar_acc0 = bar + 40, tmem_slot = bar + 64;
    int const tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&in_map) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&w3_map) : "memory");
        for (int i = 0; i < 7; ++i) mbar_init(bar + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {   // depthwise filter and bias (as fp16) stay in shared memory for the whole kernel
        __half* dws = reinterpret_cast<__half*>(gen + kSmemDw);
        for (int i = tid; i < 9 * kCexp; i += kThreads) dws[i] = dw_w[i];
        for (int i = tid; i < kCexp; i += kThreads) dws[9 * kCexp + i] = __float2half_rn(dw_b[i]);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(tmem_slot) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();  // barriers, TMEM and the depthwise filter above do not depend on the previous kernel in the stream
    pdl_trigger();
    uint32_t const tmem = *reinterpret_cast<uint32_t const*>(gen + kSmemBar + 64);

    int const my_tiles = tiles > (int)blockIdx.x ? (tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    int const n_units = 2 * my_tiles;
    auto issue_load = [&](int u) {  // unit u = (local tile u / 2, channel half u % 2) -> input buffer u % kInBufs
        int const tile = blockIdx.x + (u >> 1) * gridDim.x, half = u & 1, buf = u % kInBufs;
        int const b = tile >> 9, tr = tile & 511;
        int const oy0 = (tr >> 4) * kTH, ox0 = (tr & 15) * kTW;
        mbar_expect_tx(bar_in0 + 8 * buf, kInBytes);
        tma_load_4d(in_s + buf * kInBytes, &in_map, bar_in0 + 8 * buf, half * kChalf, ox0 - 1, oy0 - 1, b);
    };
    if (tid == 0) {
        mbar_expect_tx(bar_w, 4 * kWBlockBytes);
        for (int kb = 0; kb < 4; ++kb) tma_load_2d(w_s + kb * kWBlockBytes, &w3_map, bar_w, kb * 64, 0);
        for (int u = 0; u < kInBufs && u < n_units; ++u) issue_load(u);
    }

    // depthwise geometry of this thread: channel octet c8 of the half, output row ty, columns 4*xg .. 4*xg+3
    int const c8 = tid & 15, xg = (tid >> 4) & 3, ty = tid >> 6;
    uint32_t const idesc = make_idesc(0u, 128, kCout);
    // epilogue geometry: TMEM lane quarter, 16 of the 64 output channels
    int const quarter = warp & 3, col0 = (warp >> 2) * 16;
    int const erow = quarter * 32 + lane, ety = erow >> 4, etx = erow & 15;

    auto pixel_offset = [&](int tile) {
        int const b = tile >> 9, tr = tile & 511;
        int const oy = (tr >> 4) * kTH + ety, ox = (tr & 15) * kTW + etx;
        return (((size_t)b * 256 + oy) * 256 + ox) * kCout + col0;
    };
    auto epilogue = [&](int tile, int acc, uint4 const& s0, uint4 const& s1) {
        size_t const pix = pixel_offset(tile);
        uint32_t r[16];
        tmem_ld16(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * kCout + col0), r);
        tmem_ld_wait();
        float4 const* bb = reinterpret_cast<float4 const*>(b3 + col0);
        uint4 sc[2] = {s0, s1};
        __half2 const* sh = reinterpret_cast<__half2 const*>(sc);
        uint4 o[2];
        __half2* oh = reinterpret_cast<__half2*>(o);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float4 const bv = __ldg(bb + i);
            float2 const r0 = __half22float2(sh[2 * i]), r1 = __half22float2(sh[2 * i + 1]);
            oh[2 * i] = gelu_erf_h2(f22act2(__uint_as_float(r[4 * i]) + bv.x + r0.x, __uint_as_float(r[4 * i + 1]) + bv.y + r0.y));
            oh[2 * i + 1] = gelu_erf_h2(f22act2(__uint_as_float(r[4 * i + 2]) + bv.z + r1.x, __uint_as_float(r[4 * i + 3]) + bv.w + r1.y));
        }
        uint4* dst = reinterpret_cast<uint4*>(out + pix);
        dst[0] = o[0];
        dst[1] = o[1];
    };

    uint32_t in_phase = 0, afree_phase = 0, acc_phase = 0;  // bit b = parity of barrier b of the pair
    bool w_ready = false;
    uint4 sc0 = make_uint4(0, 0, 0, 0), sc1 = sc0;  // shortcut of the tile whose epilogue is pending
    for (int u = 0; u < n_units; ++u) {
        int const half = u & 1, lt = u >> 1, acc = lt & 1;
        int const tile = blockIdx.x + lt * gridDim.x;
        // ---- depthwise 3x3 + GELU of this channel half: halo tile -> A operand ----
        if (half == 0 && lt > 0) {  // shortcut of the previous tile: in flight during this unit's depthwise work
            uint4 const* sp = reinterpret_cast<uint4 const*>(shortcut + pixel_offset(tile - (int)gridDim.x));
            sc0 = __ldg(sp);
            sc1 = __ldg(sp + 1);
        }
        uint4 wv[9];
        uint32_t const dw_ch = dw_s + (uint32_t)((half * kChalf + c8 * 8) * 2);
#pragma unroll
        for (int k = 0; k < 9; ++k)
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(wv[k].x), "=r"(wv[k].y), "=r"(wv[k].z), "=r"(wv[k].w) : "r"(dw_ch + (uint32_t)(k * kCexp * 2)));
        __half2 accv[4][4];
        {
            uint4 bv;
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(bv.x), "=r"(bv.y), "=r"(bv.z), "=r"(bv.w) : "r"(dw_ch + (uint32_t)(9 * kCexp * 2)));
            __half2 const* bh = reinterpret_cast<__half2 const*>(&bv);
#pragma unroll
            for (int o = 0; o < 4; ++o) { accv[o][0] = bh[0]; accv[o][1] = bh[1]; accv[o][2] = bh[2]; accv[o][3] = bh[3]; }
        }
        int const buf = u % kInBufs;
        mbar_wait(bar_in0 + 8 * buf, (in_phase >> buf) & 1u);
        in_phase ^= 1u << buf;
        uint32_t const tile_in = in_s + buf * kInBytes + (uint32_t)(((ty * kHW + xg * 4) * kChalf + c8 * 8) * 2);
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            uint4 v[6];
#pragma unroll
            for (int c = 0; c < 6; ++c)
                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(v[c].x), "=r"(v[c].y), "=r"(v[c].z), "=r"(v[c].w)
                             : "r"(tile_in + (uint32_t)(((ky * kHW + c) * kChalf) * 2)));
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                __half2 const* f = reinterpret_cast<__half2 const*>(&v[c]);
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    int const o = c - kx;
                    if (o < 0 || o >= 4) continue;
                    __half2 const* w = reinterpret_cast<__half2 const*>(&wv[ky * 3 + kx]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) accv[o][i] = __hfma2(f[i], w[i], accv[o][i]);
                }
            }
        }
        // the MMAs of the previous unit, which read the A blocks, must have retired
        if (u > 0) {
            mbar_wait(bar_afree, afree_phase);
            afree_phase ^= 1u;
        }
        {
            int const kb = c8 >> 3, chunk = c8 & 7;
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                int const row = ty * kTW + xg * 4 + o;
                uint4 ov;
                __half2* oh = reinterpret_cast<__half2*>(&ov);
#pragma unroll
                for (int i = 0; i < 4; ++i) oh[i] = gelu_erf_h2(accv[o][i]);
                uint32_t const dst = a_s + (uint32_t)(kb * kABlockBytes + row * 128 + ((chunk ^ (row & 7)) << 4));
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(ov.x), "r"(ov.y), "r"(ov.z), "r"(ov.w) : "memory");
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();  // A blocks of this half complete; the input buffer has been consumed by every thread

        if (tid < 32) {  // converged warp + one elected lane: TMA refill and the 8 MMAs issue from uniform registers
            if (!w_ready) mbar_wait(bar_w, 0);
            tc_fence_after();
            if (elect_one()) {
                if (u + kInBufs < n_units) issue_load(u + kInBufs);  // refill the buffer just consumed
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    int const kb = 2 * half + j;
                    uint64_t const adesc = make_smem_desc(a_s + j * kABlockBytes);
                    uint64_t const bdesc = make_smem_desc(w_s + kb * kWBlockBytes);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc_mma<0>(tmem + (uint32_t)(acc * kCout), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                                  (uint32_t)((kb | k) != 0));
                }
                tc_commit(bar_afree);
                if (half == 1) tc_commit(bar_acc0 + 8 * acc);
            }
            __syncwarp();
        }
        w_ready = true;

        // ---- epilogue of the previous tile, once per tile (after its second half has been issued a unit ago) ----
        if (half == 0 && lt > 0) {
            int const pacc = (lt - 1) & 1;
            mbar_wait(bar_acc0 + 8 * pacc, (acc_phase >> pacc) & 1u);
            acc_phase ^= 1u << pacc;
            tc_fence_after();
            epilogue(tile - (int)gridDim.x, pacc, sc0, sc1);
            tc_fence_before();
        }
    }
    if (my_tiles > 0) {
        int const lt = my_tiles - 1, pacc = lt & 1;
        uint4 const* sp = reinterpret_cast<uint4 const*>(shortcut + pixel_offset(blockIdx.x + lt * gridDim.x));
        sc0 = __ldg(sp);
        sc1 = __ldg(sp + 1);
        mbar_wait(bar_acc0 + 8 * pacc, (acc_phase >> pacc) & 1u);
        tc_fence_after();
        epilogue(blockIdx.x + lt * gridDim.x, pacc, sc0, sc1);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}
#endif

}  // namespace

void mbconv_tail(cudaStream_t s, CUtensorMap const& expanded_map, int batch, act_t const* dw_w16, float const* dw_b,
                 CUtensorMap const& w3_map, float const* b3, act_t const* shortcut, act_t* out, int num_sms) {
#if defined(DLIMG_B200_ACT_BF16)
    (void)s; (void)expanded_map; (void)batch; (void)dw_w16; (void)dw_b; (void)w3_map; (void)b3; (void)shortcut; (void)out; (void)num_sms;
    fail("mbconv_tail: the fused kernel is built for fp16 activations");
#else
    int const tiles = batch * 512;
    ProfScope prof(s, CAT_DWCONV, 2.0 * batch * 65536.0 * (256 * 9 + 256 * 64), (double)batch * 65536.0 * (256 + 64 + 64) * 2);
    static bool attr_set = false;
    if (!attr_set) {
        CUDA_CHECK(cudaFuncSetAttribute(mbconv_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        attr_set = true;
    }
    int const grid = tiles < num_sms ? tiles : num_sms;
    launch_pdl(PDL_MBCONV, mbconv_tail_kernel, dim3(grid), dim3(kThreads), (size_t)kSmemBytes, s, expanded_map, w3_map, dw_w16, dw_b, b3, shortcut, out, tiles);
    KERNEL_CHECK();
#endif
}

}  // namespace enc
}  // namespace dlimg
