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
// A CTA loops over tiles of 8 x 16 output pixels (the 128 rows of one UMMA tile); each tile is processed in UNITS of kCU
// channels so that the 10 x 18-pixel halo of a unit can be multi-buffered:
//   TMA   4-D box (kCU ch, 18, 10, 1) at (unit * kCU, x0-1, y0-1, image); out-of-bounds = the convolution's zero padding
//   DW    thread = (channel quad, 2 x 4 pixel patch): 24 + 10 LDS.64, 144 HFMA2, 16 packed-half erf GELUs, 8 STS.64 into the
//         128B-swizzled K-major A operand (the unit's k-blocks)
//   MMA   one thread: kCU / 16 tcgen05.mma (M=128, N=64, K=16); after the last unit of a tile, commit -> mbarrier
//   EPI   of the PREVIOUS tile (its MMAs had a whole unit to finish; accumulators are double-buffered in TMEM):
//         TMEM -> + bias + shortcut -> GELU -> global
// Two shapes of the same kernel.  kCU = 128: one 512-thread CTA per SM, three 46 KB halo buffers (rounds 1 / 2a: 631 us per
// MBConv at batch 32, issue 40 %: within a unit every warp is in the same phase -- shared-memory bound depthwise loads, then
// the MUFU / HFMA2 bound GELU, then a CTA-wide barrier -- and nothing overlaps them).  kCU = 64 (default): 256-thread CTAs
// with 100 KB of shared memory, TWO per SM, which drift apart and fill each other's barrier and latency bubbles.
#include "encoder_kernels.cuh"
#include "gelu.cuh"
#include "tcgen05.cuh"

#include "../profiler.hpp"

#include <algorithm>
#include <map>
#include <mutex>

namespace dlimg {
namespace enc {

namespace {

using namespace tc;

constexpr int kTH = 8, kTW = 16, kHH = kTH + 2, kHW = kTW + 2;  // tile and halo
constexpr int kCexp = 256, kCout = 64;
constexpr int kABlockBytes = 128 * 128;                  // one k-block (64 channels) of A
constexpr int kWBlockBytes = kCout * 128;                // one k-block of the project weights

template <int kCU>
struct Tail {
    static constexpr int kUnits = kCexp / kCU;             // units per tile: 2 / 4
    static constexpr int kKB = kCU / 64;                   // k-blocks of A per unit: 2 / 1
    static constexpr int kQuads = kCU / 4;                 // channel quads per unit
    static constexpr int kThreads = 16 * kQuads;           // (quad, 2 x 4-pixel patch) work items per unit: 512 / 256
    static constexpr int kWarps = kThreads / 32;
    static constexpr int kEpiCols = kCout / (kWarps / 4);  // output channels per epilogue warp: 16 / 32
    static constexpr int kInBytes = kHH * kHW * kCU * 2;   // one halo tile of one unit: 46080 / 23040
    static constexpr int kInBufs = kCU == 128 ? 3 : 2;     // halo tiles in flight
    static constexpr int kMinBlocks = kCU == 128 ? 1 : 2;
    static constexpr int kSmemW = 0;                                  // 4 k-blocks: 32 KB
    static constexpr int kSmemA = kSmemW + 4 * kWBlockBytes;          // the unit's k-blocks, rewritten once its MMAs retired
    static constexpr int kSmemIn = kSmemA + kKB * kABlockBytes;
    static constexpr int kSmemDw = kSmemIn + kInBufs * kInBytes;      // depthwise filter (9, 256) fp16 + bias (256)
    static constexpr int kSmemBar = kSmemDw + 9 * kCexp * 2 + kCexp * 2;  // barriers + TMEM slot
    static constexpr int kSmemBytes = kSmemBar + 128 + 1024;
};

#if !defined(DLIMG_B200_ACT_BF16)
template <int kCU>
__global__ void __launch_bounds__(Tail<kCU>::kThreads, Tail<kCU>::kMinBlocks)
mbconv_tail_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap w3_map,
                   __half const* __restrict__ dw_w, float const* __restrict__ dw_b, float const* __restrict__ b3,
                   __half const* __restrict__ shortcut, __half* __restrict__ out, int tiles) {
    using T = Tail<kCU>;
    constexpr int kThreads = T::kThreads, kInBufs = T::kInBufs, kInBytes = T::kInBytes, kUnits = T::kUnits, kKB = T::kKB;
    extern __shared__ uint8_t smem_raw[];
    uint32_t const base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* const gen = smem_raw + (base - smem_u32(smem_raw));
    uint32_t const w_s = base + T::kSmemW, a_s = base + T::kSmemA, in_s = base + T::kSmemIn, dw_s = base + T::kSmemDw, bar = base + T::kSmemBar;
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
        __half* dws = reinterpret_cast<__half*>(gen + T::kSmemDw);
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
    uint32_t const tmem = *reinterpret_cast<uint32_t const*>(gen + T::kSmemBar + 64);

    int const my_tiles = tiles > (int)blockIdx.x ? (tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    int const n_units = kUnits * my_tiles;
    auto issue_load = [&](int u) {  // unit u = (local tile u / kUnits, channel group u % kUnits) -> input buffer u % kInBufs
        int const tile = blockIdx.x + (u / kUnits) * gridDim.x, cu = u % kUnits, buf = u % kInBufs;
        int const b = tile >> 9, tr = tile & 511;
        int const oy0 = (tr >> 4) * kTH, ox0 = (tr & 15) * kTW;
        mbar_expect_tx(bar_in0 + 8 * buf, kInBytes);
        tma_load_4d(in_s + buf * kInBytes, &in_map, bar_in0 + 8 * buf, cu * kCU, ox0 - 1, oy0 - 1, b);
    };
    if (tid == 0) {
        mbar_expect_tx(bar_w, 4 * kWBlockBytes);
        for (int kb = 0; kb < 4; ++kb) tma_load_2d(w_s + kb * kWBlockBytes, &w3_map, bar_w, kb * 64, 0);
        for (int u = 0; u < kInBufs && u < n_units; ++u) issue_load(u);
    }

    // depthwise geometry of this thread: channel quad c4 of the unit, output rows 2*ty2, 2*ty2+1, columns 4*xg .. 4*xg+3
    int const c4 = tid % T::kQuads, xg = (tid / T::kQuads) & 3, ty2 = tid / (4 * T::kQuads);
    uint32_t const idesc = make_idesc(0u, 128, kCout);
    // epilogue geometry: TMEM lane quarter, kEpiCols of the 64 output channels
    int const quarter = warp & 3, col0 = (warp >> 2) * T::kEpiCols;
    int const erow = quarter * 32 + lane, ety = erow >> 4, etx = erow & 15;

    auto pixel_offset = [&](int tile) {
        int const b = tile >> 9, tr = tile & 511;
        int const oy = (tr >> 4) * kTH + ety, ox = (tr & 15) * kTW + etx;
        return (((size_t)b * 256 + oy) * 256 + ox) * kCout + col0;
    };
    constexpr int kSc = T::kEpiCols / 8;  // 16-byte pieces of the shortcut / output row segment per thread: 2 / 4
    auto epilogue = [&](int tile, int acc, uint4 const (&sc)[kSc]) {
        size_t const pix = pixel_offset(tile);
        uint4* dst = reinterpret_cast<uint4*>(out + pix);
#pragma unroll
        for (int part = 0; part < kSc / 2; ++part) {  // 16 channels at a time
            uint32_t r[16];
            tmem_ld16(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * kCout + col0 + part * 16), r);
            tmem_ld_wait();
            float4 const* bb = reinterpret_cast<float4 const*>(b3 + col0 + part * 16);
            __half2 const* sh = reinterpret_cast<__half2 const*>(&sc[2 * part]);
            uint4 o[2];
            __half2* oh = reinterpret_cast<__half2*>(o);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float4 const bv = __ldg(bb + i);
                float2 const r0 = __half22float2(sh[2 * i]), r1 = __half22float2(sh[2 * i + 1]);
                oh[2 * i] = gelu_erf_h2(f22act2(__uint_as_float(r[4 * i]) + bv.x + r0.x, __uint_as_float(r[4 * i + 1]) + bv.y + r0.y));
                oh[2 * i + 1] = gelu_erf_h2(f22act2(__uint_as_float(r[4 * i + 2]) + bv.z + r1.x, __uint_as_float(r[4 * i + 3]) + bv.w + r1.y));
            }
            dst[2 * part] = o[0];
            dst[2 * part + 1] = o[1];
        }
    };
    auto load_shortcut = [&](int tile, uint4 (&sc)[kSc]) {
        uint4 const* sp = reinterpret_cast<uint4 const*>(shortcut + pixel_offset(tile));
#pragma unroll
        for (int i = 0; i < kSc; ++i) sc[i] = __ldg(sp + i);
    };

    uint32_t in_phase = 0, afree_phase = 0, acc_phase = 0;  // bit b = parity of barrier b of the pair
    bool w_ready = false;
    uint4 sc[kSc];  // shortcut of the tile whose epilogue is pending
#pragma unroll
    for (int i = 0; i < kSc; ++i) sc[i] = make_uint4(0, 0, 0, 0);
    for (int u = 0; u < n_units; ++u) {
        int const cu = u % kUnits, lt = u / kUnits, acc = lt & 1;
        int const tile = blockIdx.x + lt * gridDim.x;
        // ---- depthwise 3x3 + GELU of this channel group: halo tile -> A operand ----
        if (cu == 0 && lt > 0) load_shortcut(tile - (int)gridDim.x, sc);  // previous tile's: in flight during this unit's depthwise work
        // A thread owns a 2 x 4 pixel patch of 4 channels: 24 + 10 eight-byte shared-memory loads for 8 pixels (a 1 x 4 run
        // of 8 channels took 18 + 10 sixteen-byte loads for 4: the kernel is bound by shared-memory bandwidth, this form moves
        // 28 % fewer bytes for the same 144 HFMA2 and 16 packed GELUs per thread).
        uint2 wv[9];
        uint32_t const dw_ch = dw_s + (uint32_t)((cu * kCU + c4 * 4) * 2);
#pragma unroll
        for (int k = 0; k < 9; ++k)
            asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(wv[k].x), "=r"(wv[k].y) : "r"(dw_ch + (uint32_t)(k * kCexp * 2)));
        __half2 accv[2][4][2];
        {
            uint2 bv;
            asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(bv.x), "=r"(bv.y) : "r"(dw_ch + (uint32_t)(9 * kCexp * 2)));
            __half2 const* bh = reinterpret_cast<__half2 const*>(&bv);
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int o = 0; o < 4; ++o) { accv[r][o][0] = bh[0]; accv[r][o][1] = bh[1]; }
        }
        int const buf = u % kInBufs;
        mbar_wait(bar_in0 + 8 * buf, (in_phase >> buf) & 1u);
        in_phase ^= 1u << buf;
        uint32_t const tile_in = in_s + buf * kInBytes + (uint32_t)(((2 * ty2 * kHW + xg * 4) * kCU + c4 * 4) * 2);
#pragma unroll
        for (int iy = 0; iy < 4; ++iy) {  // halo rows 2*ty2 + iy feed output rows 2*ty2 (ky = iy) and 2*ty2 + 1 (ky = iy - 1)
            uint2 v[6];
#pragma unroll
            for (int c = 0; c < 6; ++c)
                asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(v[c].x), "=r"(v[c].y) : "r"(tile_in + (uint32_t)(((iy * kHW + c) * kCU) * 2)));
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                int const ky = iy - r;
                if (ky < 0 || ky > 2) continue;
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    __half2 const* f = reinterpret_cast<__half2 const*>(&v[c]);
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        int const o = c - kx;
                        if (o < 0 || o >= 4) continue;
                        __half2 const* w = reinterpret_cast<__half2 const*>(&wv[ky * 3 + kx]);
                        accv[r][o][0] = __hfma2(f[0], w[0], accv[r][o][0]);
                        accv[r][o][1] = __hfma2(f[1], w[1], accv[r][o][1]);
                    }
                }
            }
        }
        // the MMAs of the previous unit, which read the A blocks, must have retired
        if (u > 0) {
            mbar_wait(bar_afree, afree_phase);
            afree_phase ^= 1u;
        }
        {
            int const kb = c4 >> 4, chunk = (c4 & 15) >> 1, half8 = (c4 & 1) * 8;
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    int const row = (2 * ty2 + r) * kTW + xg * 4 + o;
                    __half2 const g0 = gelu_erf_h2(accv[r][o][0]), g1 = gelu_erf_h2(accv[r][o][1]);
                    uint32_t const dst = a_s + (uint32_t)(kb * kABlockBytes + row * 128 + ((chunk ^ (row & 7)) << 4) + half8);
                    asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(dst), "r"(*reinterpret_cast<uint32_t const*>(&g0)),
                                 "r"(*reinterpret_cast<uint32_t const*>(&g1))
                                 : "memory");
                }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();  // A blocks of this unit complete; the input buffer has been consumed by every thread

        if (tid < 32) {  // converged warp + one elected lane: TMA refill and the MMAs issue from uniform registers
            if (!w_ready) mbar_wait(bar_w, 0);
            tc_fence_after();
            if (elect_one()) {
                if (u + kInBufs < n_units) issue_load(u + kInBufs);  // refill the buffer just consumed
#pragma unroll
                for (int j = 0; j < kKB; ++j) {
                    int const kb = kKB * cu + j;
                    uint64_t const adesc = make_smem_desc(a_s + j * kABlockBytes);
                    uint64_t const bdesc = make_smem_desc(w_s + kb * kWBlockBytes);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc_mma<0>(tmem + (uint32_t)(acc * kCout), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                                  (uint32_t)((kb | k) != 0));
                }
                tc_commit(bar_afree);
                if (cu == kUnits - 1) tc_commit(bar_acc0 + 8 * acc);
            }
            __syncwarp();
        }
        w_ready = true;

        // ---- epilogue of the previous tile, once per tile (after its last unit has been issued a unit ago) ----
        if (cu == 0 && lt > 0) {
            int const pacc = (lt - 1) & 1;
            mbar_wait(bar_acc0 + 8 * pacc, (acc_phase >> pacc) & 1u);
            acc_phase ^= 1u << pacc;
            tc_fence_after();
            epilogue(tile - (int)gridDim.x, pacc, sc);
            tc_fence_before();
        }
    }
    if (my_tiles > 0) {
        int const lt = my_tiles - 1, pacc = lt & 1;
        load_shortcut(blockIdx.x + lt * gridDim.x, sc);
        mbar_wait(bar_acc0 + 8 * pacc, (acc_phase >> pacc) & 1u);
        tc_fence_after();
        epilogue(blockIdx.x + lt * gridDim.x, pacc, sc);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}
#endif

}  // namespace

int mbconv_tail_unit_channels() {
    static int const cu = dev_switch("DLIMG_B200_MBCONV_128") ? 128 : 64;  // A/B: the one-CTA-per-SM shape
    return cu;
}

void mbconv_tail(cudaStream_t s, CUtensorMap const& expanded_map, int batch, act_t const* dw_w16, float const* dw_b,
                 CUtensorMap const& w3_map, float const* b3, act_t const* shortcut, act_t* out, int num_sms) {
#if defined(DLIMG_B200_ACT_BF16)
    (void)s; (void)expanded_map; (void)batch; (void)dw_w16; (void)dw_b; (void)w3_map; (void)b3; (void)shortcut; (void)out; (void)num_sms;
    fail("mbconv_tail: the fused kernel is built for fp16 activations");
#else
    int const tiles = batch * 512;
    ProfScope prof(s, CAT_DWCONV, 2.0 * batch * 65536.0 * (256 * 9 + 256 * 64), (double)batch * 65536.0 * (256 + 64 + 64) * 2);
    auto go = [&](auto kernel, int threads, int smem, int ctas_per_sm) {
        static std::mutex mutex;
        static std::map<void const*, bool> done;
        {
            std::lock_guard<std::mutex> lock(mutex);
            if (!done[(void const*)kernel]) {
                CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
                done[(void const*)kernel] = true;
            }
        }
        int const grid = std::min(tiles, ctas_per_sm * num_sms);
        launch_pdl(PDL_MBCONV, kernel, dim3(grid), dim3(threads), (size_t)smem, s, expanded_map, w3_map, dw_w16, dw_b, b3, shortcut, out, tiles);
    };
#if DLIMG_B200_ALT
    if (mbconv_tail_unit_channels() == 128) go(mbconv_tail_kernel<128>, Tail<128>::kThreads, Tail<128>::kSmemBytes, 1);
    else
#endif
        go(mbconv_tail_kernel<64>, Tail<64>::kThreads, Tail<64>::kSmemBytes, 2);
    KERNEL_CHECK();
#endif
}

}  // namespace enc
}  // namespace dlimg
