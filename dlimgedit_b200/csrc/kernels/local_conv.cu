// local_conv.cu -- the depthwise 3x3 `local_conv` of a TinyViT block (stride 1, zero padding 1, BN folded, no activation)
// from TMA-loaded halo tiles, with the LayerNorm row sums of its output for the MLP behind it.
//
// The register-tiled form (dwconv3x3_stats_kernel, encoder_kernels.cu) issues 18 global 16-byte loads per 4 x 8
// outputs and waits for them at 17-24 resident warps per SM: 2.7-3.4x off the HBM time of its 2 x C bytes per pixel
// (profiles/r01f).  Here a persistent CTA walks over units = (8 x 16-pixel tile, channel part of <= 160 channels):
//   TMA   4-D box (channels, 18, 10, 1) at (part * Cu, x0 - 1, y0 - 1, image), three units in flight; out-of-image
//         elements arrive as zeros = the convolution's padding, so the arithmetic has no edge cases
//   DW    thread = (channel octet, run of 4 pixels, tile row): 18 LDS.128 of activations, the fp32 filter from shared
//         memory, fp32 accumulation in exactly the order of the register-tiled kernel (bit-identical output)
//   OUT   16-byte stores (a pixel's octets are adjacent lanes), per-pixel (sum, sum of squares) through a
//         double-buffered shared-memory exchange, added in a fixed order by one thread per pixel
//   SYNC  no CTA-wide barrier: the threads of two tile rows (4 or 5 whole warps) meet at a named barrier for the row-sum
//         exchange, and the last warp to finish reading an input buffer (a shared-memory counter) issues the TMA load
//         that refills it, so the warp groups drift apart and cover each other's shared-memory latency
// C = 320 runs as two channel parts of 160; the consumer adds the two partial row sums (gemm Epilogue::ln_parts = 2).
#include "encoder_kernels.cuh"
#include "gemm.cuh"
#include "tcgen05.cuh"

#include "../profiler.hpp"

namespace dlimg {
namespace enc {

namespace {

using namespace tc;

constexpr int kTH = 8, kTW = 16, kHH = kTH + 2, kHW = kTW + 2;  // tile and halo
constexpr int kInBufs = 3;

template <int kC8>
struct LcCfg {
    static constexpr int kCu = kC8 * 8;                       // channels per unit
    static constexpr int kThreads = kC8 * 4 * kTH;            // (octet, 4-pixel run, row)
    static constexpr int kGroup = kC8 * 8;                    // threads of two tile rows: 4 or 5 whole warps
    static constexpr int kMaxRegs = (65536 / kThreads) / 8 * 8;  // one CTA per SM: give the compiler the whole file
    static constexpr int kInBytes = kHH * kHW * kCu * 2;      // one halo tile
    static constexpr int kPartPitch = kC8 + 1;                // float2 per pixel row of the exchange (+1: conflict-free reads)
    static constexpr int kPartBytes = kTH * kTW * kPartPitch * 8;  // (sum, sum of squares) per (pixel, octet)
    static constexpr int kSmemIn = 0;
    static constexpr int kSmemPart = kSmemIn + kInBufs * kInBytes;
    static constexpr int kSmemBar = kSmemPart + 2 * kPartBytes;
    static constexpr int kSmemW = kSmemBar + 64;              // filter [9][C] fp32 + bias [C], C known at run time
};

// d = a * b + c on both lanes of a packed fp32 pair: one FFMA2, per-lane result identical to fmaf
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(reinterpret_cast<uint64_t&>(d))
        : "l"(reinterpret_cast<uint64_t const&>(a)), "l"(reinterpret_cast<uint64_t const&>(b)), "l"(reinterpret_cast<uint64_t const&>(c)));
    return d;
}

template <int kC8>
__global__ void __launch_bounds__(LcCfg<kC8>::kThreads, 1) __maxnreg__(LcCfg<kC8>::kMaxRegs)
local_conv_kernel(const __grid_constant__ CUtensorMap in_map, float const* __restrict__ weight, float const* __restrict__ bias,
                  act_t* __restrict__ out, float2* __restrict__ stats, int H, int W, int C, int units) {
    using L = LcCfg<kC8>;
    extern __shared__ uint8_t smem_raw[];
    uint32_t const base = (smem_u32(smem_raw) + 127u) & ~127u;
    uint8_t* const gen = smem_raw + (base - smem_u32(smem_raw));
    uint32_t const in_s = base + L::kSmemIn, bar = base + L::kSmemBar;
    float2* const part = reinterpret_cast<float2*>(gen + L::kSmemPart);
    float* const wsm = reinterpret_cast<float*>(gen + L::kSmemW);  // [9][C], then bias [C]
    int const tid = threadIdx.x;
    int const parts = C / L::kCu, tiles_x = W / kTW, tiles_per_img = tiles_x * (H / kTH);

    static_assert(L::kGroup % 32 == 0, "row-pair groups must be whole warps");
    // barriers: full[kInBufs] at 0; drained-warp counters [kInBufs] (int) at 32
    int* const drained = reinterpret_cast<int*>(gen + L::kSmemBar + 32);
    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&in_map) : "memory");
        for (int i = 0; i < kInBufs; ++i) {
            mbar_init(bar + 8 * i, 1);
            drained[i] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // filter and bias in shared memory as [tap][half of the octet][octet][4]: the 16 / 20 lanes of a pixel read
    // consecutive 16-byte pieces (the natural [tap][channel] order makes every such read a 2-way bank conflict)
    for (int i = tid; i < 10 * C; i += L::kThreads) {
        int const k = i / C, ch = i - k * C;
        wsm[((k * 2 + ((ch >> 2) & 1)) * (C >> 3) + (ch >> 3)) * 4 + (ch & 3)] = k < 9 ? __ldg(weight + i) : __ldg(bias + ch);
    }
    __syncthreads();
    pdl_wait();
    pdl_trigger();

    int const my_units = units > (int)blockIdx.x ? (units - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    // unit -> (image, tile origin, channel part); the parts of a tile are adjacent units
    auto locate = [&](int u, int& b, int& oy0, int& ox0, int& pt) {
        int const unit = (int)blockIdx.x + u * (int)gridDim.x;
        int const tile = unit / parts;
        pt = unit - tile * parts;
        b = tile / tiles_per_img;
        int const tr = tile - b * tiles_per_img;
        oy0 = (tr / tiles_x) * kTH;
        ox0 = (tr % tiles_x) * kTW;
    };

    auto issue_load = [&](int u) {
        int b, oy0, ox0, pt;
        locate(u, b, oy0, ox0, pt);
        int const buf = u % kInBufs;
        mbar_expect_tx(bar + 8 * buf, (uint32_t)L::kInBytes);
        tma_load_4d(in_s + buf * L::kInBytes, &in_map, bar + 8 * buf, pt * L::kCu, ox0 - 1, oy0 - 1, b);
    };
    if (tid == 0)
        for (int u = 0; u < kInBufs && u < my_units; ++u) issue_load(u);

    int const c8 = tid % kC8, xg = (tid / kC8) & 3, ty = tid / (kC8 * 4);
    int const grp = tid / L::kGroup, gtid = tid - grp * L::kGroup;  // row pair and index within it
    // unit coordinates of the main loop are stepped (unit += gridDim.x is a constant (image, tile row, tile column, part)
    // quadruple with carries): every thread needs them, and locate() is four run-time divisions
    int const tiles_y = tiles_per_img / tiles_x;
    int sb, sty, stx, spt;
    {
        int const g = (int)gridDim.x, gt = g / parts;
        spt = g - gt * parts;
        sb = gt / tiles_per_img;
        int const r = gt - sb * tiles_per_img;
        sty = r / tiles_x;
        stx = r - sty * tiles_x;
    }
    int cb, cty, ctx, cpt;
    {
        int oy0, ox0;
        locate(0, cb, oy0, ox0, cpt);
        cty = oy0 / kTH;
        ctx = ox0 / kTW;
    }
    for (int u = 0; u < my_units; ++u) {
        int const b = cb, oy0 = cty * kTH, ox0 = ctx * kTW, pt = cpt;
        cpt += spt;
        if (cpt >= parts) { cpt -= parts; ++ctx; }
        ctx += stx;
        if (ctx >= tiles_x) { ctx -= tiles_x; ++cty; }
        cty += sty;
        if (cty >= tiles_y) { cty -= tiles_y; ++cb; }
        cb += sb;
        int const buf = u % kInBufs;
        float const* const w = wsm + (pt * kC8 + c8) * 4;  // tap k: halves at w + k * C and w + k * C + C / 2
        float2 acc2[4][4];  // [pixel][channel pair]: fp32 accumulation on packed pairs (FFMA2: half the FMA instructions)
        {
            float4 const b0 = *reinterpret_cast<float4 const*>(w + 9 * C), b1 = *reinterpret_cast<float4 const*>(w + 9 * C + (C >> 1));
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                acc2[o][0] = make_float2(b0.x, b0.y); acc2[o][1] = make_float2(b0.z, b0.w);
                acc2[o][2] = make_float2(b1.x, b1.y); acc2[o][3] = make_float2(b1.z, b1.w);
            }
        }
        mbar_wait(bar + 8 * buf, ((uint32_t)(u / kInBufs)) & 1u);
        uint32_t const tile_in = in_s + buf * L::kInBytes + (uint32_t)(((ty * kHW + xg * 4) * L::kCu + c8 * 8) * 2);
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            uint4 v[6];
#pragma unroll
            for (int c = 0; c < 6; ++c)
                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(v[c].x), "=r"(v[c].y), "=r"(v[c].z), "=r"(v[c].w)
                             : "r"(tile_in + (uint32_t)(((ky * kHW + c) * L::kCu) * 2)));
            if (ky == 2) {  // this warp's last reads of the buffer have been issued: the last warp to get here refills it
                asm volatile("" ::: "memory");  // (compiler: keep the counter update behind the loads above)
                __syncwarp();
                if ((tid & 31) == 0 && atomicAdd(&drained[buf], 1) == L::kThreads / 32 - 1) {
                    drained[buf] = 0;
                    if (u + kInBufs < my_units) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        issue_load(u + kInBufs);
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                act2_t const* h = reinterpret_cast<act2_t const*>(&v[c]);
                float2 const f0 = act22f2(h[0]), f1 = act22f2(h[1]), f2 = act22f2(h[2]), f3 = act22f2(h[3]);
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    int const o = c - kx;
                    if (o < 0 || o >= 4) continue;
                    float const* wk = w + (ky * 3 + kx) * C;
                    float4 const w0 = *reinterpret_cast<float4 const*>(wk), w1 = *reinterpret_cast<float4 const*>(wk + (C >> 1));
                    acc2[o][0] = fma2(f0, make_float2(w0.x, w0.y), acc2[o][0]);
                    acc2[o][1] = fma2(f1, make_float2(w0.z, w0.w), acc2[o][1]);
                    acc2[o][2] = fma2(f2, make_float2(w1.x, w1.y), acc2[o][2]);
                    acc2[o][3] = fma2(f3, make_float2(w1.z, w1.w), acc2[o][3]);
                }
            }
        }
        // output + per-(pixel, octet) partial row sums
        int const oy = oy0 + ty, ox = ox0 + xg * 4;
        int64_t const row0 = ((int64_t)b * H + oy) * W + ox;
        uint4* const orow = reinterpret_cast<uint4*>(out + row0 * C + pt * L::kCu) + c8;
        float2* const mypart = part + (u & 1) * (kTH * kTW * L::kPartPitch) + ((ty * kTW + xg * 4) * L::kPartPitch + c8);
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            float const* acc_o = reinterpret_cast<float const*>(acc2[o]);
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                s1 += acc_o[i];
                s2 = fmaf(acc_o[i], acc_o[i], s2);
            }
            mypart[o * L::kPartPitch] = make_float2(s1, s2);
            uint4 ov;
            act2_t* oh = reinterpret_cast<act2_t*>(&ov);
#pragma unroll
            for (int i = 0; i < 4; ++i) oh[i] = f22act2(acc2[o][i].x, acc2[o][i].y);
            orow[(size_t)o * (C / 8)] = ov;
        }
        // the partial sums of this row pair's 32 pixels are visible to its first warp
        asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(L::kGroup) : "memory");
        if (gtid < 2 * kTW) {  // one thread per pixel of the row pair: fixed-order sum over the unit's octets
            int const pix = grp * 2 * kTW + gtid;
            float2 const* p = part + (u & 1) * (kTH * kTW * L::kPartPitch) + pix * L::kPartPitch;
            float s1 = 0.f, s2 = 0.f;
#pragma unroll 4
            for (int k = 0; k < kC8; ++k) {
                float2 const v = p[k];
                s1 += v.x;
                s2 += v.y;
            }
            int const py = pix / kTW, px = pix - py * kTW;
            stats[(((int64_t)b * H + oy0 + py) * W + ox0 + px) * parts + pt] = make_float2(s1, s2);
        }
        // (the other exchange buffer is written by the next unit; this one again only after the next unit's barrier,
        //  which the readers above reach after they are done with it)
    }
}

template <int kC8>
void launch_local_conv(cudaStream_t s, act_t const* in, int batch, int H, int W, int C, float const* weight, float const* bias,
                       act_t* out, float2* stats, int num_sms) {
    using L = LcCfg<kC8>;
    int const smem = L::kSmemW + 10 * C * 4 + 128;
    DLIMG_ASSERT(smem <= 227 * 1024);
    static bool attr_set = false;
    if (!attr_set) {
        CUDA_CHECK(cudaFuncSetAttribute(local_conv_kernel<kC8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    CUtensorMap const map = gemm::make_tensor_map_nhwc(in, batch, H, W, C, L::kCu, kHW, kHH);
    int const units = batch * (H / kTH) * (W / kTW) * (C / L::kCu);
    int const grid = units < num_sms ? units : num_sms;
    launch_pdl(PDL_LOCAL_CONV, local_conv_kernel<kC8>, dim3(grid), dim3(L::kThreads), (size_t)smem, s, map, weight, bias, out, stats, H, W, C, units);
    KERNEL_CHECK();
}

}  // namespace

int local_conv_parts(int C) { return C == 320 ? 2 : 1; }

bool local_conv_tma_supported(int H, int W, int C) {
    return H % kTH == 0 && W % kTW == 0 && (C == 128 || C == 160 || C == 320);
}

void local_conv_tma(cudaStream_t s, act_t const* in, int batch, int H, int W, int C, float const* weight, float const* bias,
                    act_t* out, float2* stats, int num_sms) {
    DLIMG_ASSERT(local_conv_tma_supported(H, W, C));
    ProfScope prof(s, CAT_DWCONV, 2.0 * batch * H * W * C * 9, (double)batch * 2.0 * H * W * C * 2);
    if (C == 128) launch_local_conv<16>(s, in, batch, H, W, C, weight, bias, out, stats, num_sms);
    else launch_local_conv<20>(s, in, batch, H, W, C, weight, bias, out, stats, num_sms);  // C = 160, or 320 as two parts
}

}  // namespace enc
}  // namespace dlimg
