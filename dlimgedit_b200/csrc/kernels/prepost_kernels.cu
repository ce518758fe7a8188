// prepost_kernels.cu -- see prepost_kernels.cuh.
#include "prepost_kernels.cuh"

#include "../profiler.hpp"

#include <cstdlib>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <limits>
#include <map>
#include <mutex>

namespace dlimg {
namespace prepost {

// ---------------------------------------------------------------------------------------------
// Host: resampling plans and sRGB tables
// ---------------------------------------------------------------------------------------------
namespace {

float mitchell(float x) {
    x = std::fabs(x);
    if (x < 1.0f) return (16 + x * x * (21 * x - 36)) / 18;
    if (x < 2.0f) return (32 + x * (-60 + x * (36 - 7 * x))) / 18;
    return 0.0f;
}

float catmull_rom(float x) {
    x = std::fabs(x);
    if (x < 1.0f) return 1 - x * x * (2.5f - 1.5f * x);
    if (x < 2.0f) return 2 - x * (4 + x * (0.5f * x - 2.5f));
    return 0.0f;
}

struct Span {
    int lo = 0, hi = -1;
    std::vector<float> w;
};

void enlarge_plan(AxisPlan& plan, float scale) {
    // one tap list per OUTPUT pixel; kernel support is 2 input pixels either side
    std::vector<Span> spans((size_t)plan.out_size);
    float const radius = 2.0f * scale;
    int taps = 0;
    for (int o = 0; o < plan.out_size; ++o) {
        float const oc = (float)o + 0.5f;
        float const centre = oc / scale;
        int lo = (int)std::floor((oc - radius) / scale + 0.5f);
        int const hi = (int)std::floor((oc + radius) / scale - 0.5f);
        std::vector<float> w;
        float total = 0;
        for (int i = lo; i <= hi; ++i) {
            float const c = catmull_rom(centre - ((float)i + 0.5f));
            if (w.empty() && c == 0.0f) { ++lo; continue; }  // leading zero tap is dropped
            w.push_back(c);
            total += c;
        }
        float const norm = 1 / total;
        for (float& c : w) c *= norm;
        while (!w.empty() && w.back() == 0.0f) w.pop_back();
        spans[(size_t)o].lo = lo;
        spans[(size_t)o].hi = lo + (int)w.size() - 1;
        spans[(size_t)o].w = std::move(w);
        taps = std::max(taps, (int)spans[(size_t)o].w.size());
    }
    plan.taps = taps;
    plan.first.resize((size_t)plan.out_size);
    plan.weights.assign((size_t)plan.out_size * taps, 0.0f);
    for (int o = 0; o < plan.out_size; ++o) {
        plan.first[(size_t)o] = spans[(size_t)o].lo;
        std::copy(spans[(size_t)o].w.begin(), spans[(size_t)o].w.end(), plan.weights.begin() + (size_t)o * taps);
    }
}

void shrink_plan(AxisPlan& plan, float scale) {
    // scatter form: every INPUT pixel (including clamped margin pixels) lists the outputs it feeds
    int const margin = (int)std::ceil(2.0f * 2 / scale) / 2;
    int const n_src = plan.in_size + 2 * margin;
    float const radius = 2.0f / scale;
    std::vector<Span> spans((size_t)n_src);
    for (int j = 0; j < n_src; ++j) {
        float const ic = (float)(j - margin) + 0.5f;
        float const centre = ic * scale;
        Span& sp = spans[(size_t)j];
        sp.lo = (int)std::floor((ic - radius) * scale + 0.5f);
        int const hi = (int)std::floor((ic + radius) * scale - 0.5f);
        for (int o = sp.lo; o <= hi; ++o) sp.w.push_back(mitchell(((float)o + 0.5f) - centre) * scale);
        while (!sp.w.empty() && sp.w.back() == 0.0f) sp.w.pop_back();
        sp.hi = sp.lo + (int)sp.w.size() - 1;
    }
    // gather form + per-output normalisation (sum in ascending source order)
    plan.first.assign((size_t)plan.out_size, INT_MAX);
    std::vector<int> last((size_t)plan.out_size, INT_MIN);
    for (int j = 0; j < n_src; ++j)
        for (int o = std::max(spans[(size_t)j].lo, 0); o <= std::min(spans[(size_t)j].hi, plan.out_size - 1); ++o) {
            plan.first[(size_t)o] = std::min(plan.first[(size_t)o], j - margin);
            last[(size_t)o] = std::max(last[(size_t)o], j - margin);
        }
    int taps = 0;
    for (int o = 0; o < plan.out_size; ++o) taps = std::max(taps, last[(size_t)o] - plan.first[(size_t)o] + 1);
    plan.taps = taps;
    plan.weights.assign((size_t)plan.out_size * taps, 0.0f);
    for (int o = 0; o < plan.out_size; ++o) {
        float total = 0;
        for (int n = plan.first[(size_t)o]; n <= last[(size_t)o]; ++n) {
            Span const& sp = spans[(size_t)(n + margin)];
            if (o >= sp.lo && o <= sp.hi) total += sp.w[(size_t)(o - sp.lo)];
        }
        float const norm = 1 / total;
        for (int n = plan.first[(size_t)o]; n <= last[(size_t)o]; ++n) {
            Span const& sp = spans[(size_t)(n + margin)];
            if (o >= sp.lo && o <= sp.hi)
                plan.weights[(size_t)o * taps + (size_t)(n - plan.first[(size_t)o])] = sp.w[(size_t)(o - sp.lo)] * norm;
        }
    }
}

double srgb_encode_exact(float x) {
    double const v = x <= 0.0031308 ? 12.92 * (double)x : 1.055 * std::pow((double)x, 1.0 / 2.4) - 0.055;
    return std::floor(v * 255.0 + 0.5);
}

}  // namespace

AxisPlan make_axis_plan(int in_size, int out_size) {
    DLIMG_ASSERT(in_size > 0 && out_size > 0);
    AxisPlan plan;
    plan.in_size = in_size;
    plan.out_size = out_size;
    float const scale = (float)out_size / in_size;
    if (scale > 1) enlarge_plan(plan, scale);
    else shrink_plan(plan, scale);
    return plan;
}

SrgbTables const& srgb_tables() {
    static SrgbTables const tables = [] {
        SrgbTables t;
        for (int i = 0; i < 256; ++i) {
            double const c = i / 255.0;
            double const lin = c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4);
            t.decode[i] = (float)(std::floor(lin * 1e6 + 0.5) / 1e6);
        }
        // encode_threshold[i]: smallest positive float whose correctly rounded sRGB8 code is >= i.
        // The encoder is monotonic, so a bisection over float bit patterns in (0, 1) finds it exactly.
        t.encode_threshold[0] = -std::numeric_limits<float>::infinity();
        for (int i = 1; i < 256; ++i) {
            uint32_t lo = 0x00000001u, hi = 0x3f800000u;  // smallest denormal .. 1.0f
            while (lo < hi) {
                uint32_t const mid = lo + (hi - lo) / 2;
                float f;
                std::memcpy(&f, &mid, 4);
                if (srgb_encode_exact(f) >= (double)i) hi = mid;
                else lo = mid + 1;
            }
            std::memcpy(&t.encode_threshold[i], &lo, 4);
        }
        return t;
    }();
    return tables;
}

LongestSide resize_longest_side(int w, int h, int max_side) {
    LongestSide r;
    r.orig_w = w;
    r.orig_h = h;
    r.scale = float(max_side) / float(std::max(w, h));
    r.needs_resize = r.scale != 1;
    r.w = r.needs_resize ? scale_coord(w, r.scale) : w;
    r.h = r.needs_resize ? scale_coord(h, r.scale) : h;
    return r;
}

// ---------------------------------------------------------------------------------------------
// Device kernels
// ---------------------------------------------------------------------------------------------
namespace {

// Correctly rounded linear -> sRGB8: code = number of thresholds <= acc (ascending, thr[0] = -inf); eight
// branch-free steps of a binary search over the shared-memory table.
__device__ __forceinline__ int srgb_encode(float acc, float const* __restrict__ thr) {
    int lo = 0;
#pragma unroll
    for (int step = 128; step > 0; step >>= 1) lo += acc >= thr[lo + step] ? step : 0;
    return lo;
}

// ---- resize: ONE kernel, separable through shared memory (reference image.cpp:37-51) ----------------------------
// A block produces a tile of tr x tc output pixels.  The input rows the tile needs are decoded (u8 -> linear float,
// table look-up, edge pixels clamped) `ch` rows at a time into shared memory; thread (row group, j) owns H-pass column
// j = (output column, channel) with its filter taps in registers, filters two input rows at a time (two independent
// add chains) and leaves the horizontally filtered rows in shared memory; the vertical pass + linear->sRGB8 encode then
// reads four neighbouring columns per thread (LDS.128) and writes four bytes at once.  No intermediate ever reaches HBM
// (the two-pass form wrote in_h x out_w x bpp floats and read them back).  Multiplications and additions stay separate
// (__fmul_rn / __fadd_rn), in ascending tap order from 0, exactly like the reference filter -- results are bit-identical
// to oracle/c/prepost_ref.c.  Tall, narrow tiles (32 x 40 at 4K) keep the halo the two filters re-read at 9 % / 7 %.
// The kernel is bound by FP32 issue, not by HBM: 4K -> 1024 needs 15 + 15 taps of 2 instructions per output element on
// top of the decode, ~6e8 thread instructions against 27 MB of traffic.
struct ResizeTile {
    int tr = 0, tc = 0;          // output rows / columns per block; tc * bpp <= 256 and a multiple of 4
    int nr_max = 0, nc_max = 0;  // most input rows / columns any tile needs (exact, from the axis plans)
    int ch = 8;                  // input rows decoded per pass (one per warp)
    int raw_pitch = 0;           // bytes per row of the raw (cp.async) staging buffers
    int smem_bytes = 0;
};

template <int BPP, int KT>
__global__ void __launch_bounds__(256) resize_tile_kernel(uint8_t const* __restrict__ in, int in_w, int in_h, int stride,
                                                          ResizeDeviceTables t, ResizeTile g, uint8_t* __restrict__ out,
                                                          int out_w, int out_h) {
    extern __shared__ __align__(16) float rs_smem[];
    int const row_elems = g.tc * BPP;            // H-pass outputs per input row
    int const px_per_row = g.nc_max + KT;        // staged pixels per input row (tail zero-filled: taps >= htaps read it)
    int const in_elems = (px_per_row * BPP + 32 + 3) & ~3;  // floats per staged row (+32: the streaming path's alignment slack)
    float* const s_dec = rs_smem;
    float* const s_thr = s_dec + 256;
    float* const s_vw = s_thr + 256;
    float* const s_hb = s_vw + ((g.tr * t.vtaps + 3) & ~3);
    float* const s_in = s_hb + g.nr_max * row_elems;
    int const tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int const ox0 = blockIdx.x * g.tc, oy0 = blockIdx.y * g.tr;
    int const tcv = min(g.tc, out_w - ox0), trv = min(g.tr, out_h - oy0);
    s_dec[tid] = __ldg(t.decode + tid);
    s_thr[tid] = __ldg(t.encode_threshold + tid);
    for (int i = tid; i < trv * t.vtaps; i += 256) s_vw[i] = __ldg(t.vweights + (size_t)oy0 * t.vtaps + i);
    int const r_lo = __ldg(t.vfirst + oy0), r_hi = __ldg(t.vfirst + oy0 + trv - 1) + t.vtaps - 1;
    int const nr = r_hi - r_lo + 1;
    int const c_lo = __ldg(t.hfirst + ox0), c_hi = __ldg(t.hfirst + ox0 + tcv - 1) + t.htaps - 1;
    int const nc = c_hi - c_lo + 1;
    // H-pass geometry: 256 / row_elems row groups share the rows of a chunk
    int const groups = min(256 / row_elems, g.ch), rows_per_group = g.ch / groups;
    int const grp = tid / row_elems, j = tid - grp * row_elems;
    bool const hactive = grp < groups && j < tcv * BPP;
    float w[KT];
    int foff = 0;
    {
        int const ox = ox0 + j / BPP, c = j % BPP;
#pragma unroll
        for (int k = 0; k < KT; ++k) w[k] = (hactive && k < t.htaps) ? __ldg(t.hweights + (size_t)ox * t.htaps + k) : 0.0f;
        if (hactive) foff = (__ldg(t.hfirst + ox) - c_lo) * BPP + c;
    }
    bool const interior = c_lo >= 0 && c_hi < in_w;  // no horizontal clamping in this tile: staged bytes are contiguous
    int const row_bytes = nc * BPP;
    // Interior tiles of 16-byte aligned images stream their input rows with cp.async, one chunk (8 rows, one per warp) ahead
    // of the chunk being filtered: whole 16-byte granules from the aligned address below the tile's first byte into a raw
    // byte buffer (double-buffered), decoded from there four bytes per shared-memory load.  The trip to L2 / HBM no longer
    // sits in front of every chunk (byte loads straight from global memory left the kernel at 30 % issue utilisation, half
    // of its stalls on them), and the copy holds no registers.  Reading up to 15 bytes past the tile's last byte stays
    // inside the image unless the tile touches the last row: those tiles, clamped ones and unaligned images decode with
    // plain loads.
    int const byte_lo = c_lo * BPP, a_lo = byte_lo & ~15, d0 = byte_lo - a_lo;
    int const raw_len = (d0 + row_bytes + 15) & ~15;
    bool const use_async = interior && g.ch == 8 && ((reinterpret_cast<uintptr_t>(in) | (uintptr_t)stride) & 15) == 0 &&
                           (a_lo + raw_len <= in_w * BPP || r_hi < in_h - 1);
    uint8_t* const s_raw = reinterpret_cast<uint8_t*>(s_in + g.ch * in_elems);  // [2][ch][raw_pitch]
    auto issue = [&](int rc, int buf) {
        if (warp < min(g.ch, nr - rc)) {
            int const y = min(max(r_lo + rc + warp, 0), in_h - 1);
            uint8_t const* src = in + (size_t)y * stride + a_lo;
            uint32_t const dst = (uint32_t)__cvta_generic_to_shared(s_raw + (size_t)(buf * g.ch + warp) * g.raw_pitch);
            for (int o = lane * 16; o < raw_len; o += 512)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + o), "l"(src + o) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (use_async) issue(0, 0);
    // the zero tail behind every staged row (taps >= htaps read it) is written once: decoding never touches it.  The
    // streaming path decodes whole raw rows -- staged element e sits at index d0 + e, the raw bytes around the tile's own are
    // image bytes too (finite values under zero weights) -- so every shared-memory store is an aligned 16-byte one.
    int const e_shift = use_async ? d0 : 0;
    for (int rr = warp; rr < g.ch; rr += 8)
        for (int e = (use_async ? raw_len : row_bytes) + lane; e < in_elems; e += 32) s_in[rr * in_elems + e] = 0.0f;
    __syncthreads();
    int chunk = 0;
    for (int rc = 0; rc < nr; rc += g.ch, ++chunk) {
        int const chv = min(g.ch, nr - rc);
        if (use_async) {
            bool const more = rc + g.ch < nr;
            if (more) issue(rc + g.ch, (chunk + 1) & 1);
            if (more) asm volatile("cp.async.wait_group 1;" ::: "memory");
            else asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();  // (a warp decodes the row it copied itself, but the H-pass below must be done with s_in)
            if (warp < chv) {
                uint32_t const* raw = reinterpret_cast<uint32_t const*>(s_raw + (size_t)((chunk & 1) * g.ch + warp) * g.raw_pitch);
                float4* dst_row = reinterpret_cast<float4*>(s_in + warp * in_elems);
                for (int wi = lane; wi < (raw_len >> 2); wi += 32) {
                    uint32_t const v = raw[wi];
                    dst_row[wi] = make_float4(s_dec[v & 255u], s_dec[(v >> 8) & 255u], s_dec[(v >> 16) & 255u], s_dec[v >> 24]);
                }
            }
        } else {
            // decode: warp = staged row, lanes walk its bytes (32 consecutive bytes per load instruction)
            for (int rr = warp; rr < chv; rr += 8) {
                int const y = min(max(r_lo + rc + rr, 0), in_h - 1);
                uint8_t const* src_row = in + (size_t)y * stride;
                float* dst_row = s_in + rr * in_elems;
                for (int e = lane; e < row_bytes; e += 32) {
                    int const px = e / BPP, c = e - px * BPP;
                    int const x = min(max(c_lo + px, 0), in_w - 1);
                    dst_row[e] = s_dec[__ldg(src_row + (size_t)x * BPP + c)];
                }
            }
        }
        __syncthreads();
        if (hactive) {
            int const r_end = min((grp + 1) * rows_per_group, chv);
            int rr = grp * rows_per_group;
            for (; rr + 3 < r_end; rr += 4) {  // four rows at once: four independent add chains hide the shared-memory latency
                float const* row0 = s_in + rr * in_elems + foff + e_shift;
                float const* row1 = row0 + in_elems;
                float const* row2 = row1 + in_elems;
                float const* row3 = row2 + in_elems;
                float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll
                for (int k = 0; k < KT; ++k) {
                    a0 = __fadd_rn(a0, __fmul_rn(row0[k * BPP], w[k]));
                    a1 = __fadd_rn(a1, __fmul_rn(row1[k * BPP], w[k]));
                    a2 = __fadd_rn(a2, __fmul_rn(row2[k * BPP], w[k]));
                    a3 = __fadd_rn(a3, __fmul_rn(row3[k * BPP], w[k]));
                }
                float* hb = s_hb + (rc + rr) * row_elems + j;
                hb[0] = a0;
                hb[row_elems] = a1;
                hb[2 * row_elems] = a2;
                hb[3 * row_elems] = a3;
            }
            for (; rr + 1 < r_end; rr += 2) {  // two rows at once
                float const* row0 = s_in + rr * in_elems + foff + e_shift;
                float const* row1 = row0 + in_elems;
                float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
                for (int k = 0; k < KT; ++k) {
                    a0 = __fadd_rn(a0, __fmul_rn(row0[k * BPP], w[k]));
                    a1 = __fadd_rn(a1, __fmul_rn(row1[k * BPP], w[k]));
                }
                s_hb[(rc + rr) * row_elems + j] = a0;
                s_hb[(rc + rr + 1) * row_elems + j] = a1;
            }
            if (rr < r_end) {
                float const* row0 = s_in + rr * in_elems + foff + e_shift;
                float a0 = 0.0f;
#pragma unroll
                for (int k = 0; k < KT; ++k) a0 = __fadd_rn(a0, __fmul_rn(row0[k * BPP], w[k]));
                s_hb[(rc + rr) * row_elems + j] = a0;
            }
        }
        if (!use_async) __syncthreads();  // (the streaming path meets at the barrier in front of its next decode)
    }
    __syncthreads();
    // vertical pass + encode: four neighbouring columns per thread
    int const quads = row_elems >> 2, valid_elems = tcv * BPP;
    for (int i = tid; i < trv * quads; i += 256) {
        int const oyl = i / quads, jj = (i - oyl * quads) * 4;
        if (jj >= valid_elems) continue;
        int const rbase = __ldg(t.vfirst + oy0 + oyl) - r_lo;
        float const* vw = s_vw + oyl * t.vtaps;
        float const* col = s_hb + rbase * row_elems + jj;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k < t.vtaps; ++k) {
            float4 const v = *reinterpret_cast<float4 const*>(col + k * row_elems);
            float const wk = vw[k];
            acc.x = __fadd_rn(acc.x, __fmul_rn(v.x, wk));
            acc.y = __fadd_rn(acc.y, __fmul_rn(v.y, wk));
            acc.z = __fadd_rn(acc.z, __fmul_rn(v.z, wk));
            acc.w = __fadd_rn(acc.w, __fmul_rn(v.w, wk));
        }
        uint32_t const b0 = (uint32_t)srgb_encode(acc.x, s_thr), b1 = (uint32_t)srgb_encode(acc.y, s_thr);
        uint32_t const b2 = (uint32_t)srgb_encode(acc.z, s_thr), b3 = (uint32_t)srgb_encode(acc.w, s_thr);
        uint8_t* dst = out + ((size_t)(oy0 + oyl) * out_w + ox0) * BPP + jj;
        if (jj + 4 <= valid_elems && (reinterpret_cast<uintptr_t>(dst) & 3) == 0) {
            *reinterpret_cast<uint32_t*>(dst) = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
        } else {
            uint32_t const b[4] = {b0, b1, b2, b3};
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (jj + e < valid_elems) dst[e] = (uint8_t)b[e];
        }
    }
}

// Two-pass fallback for extreme reductions (more than 32 taps per axis, or tiles that do not fit shared memory).
// Horizontal pass: (in_h, in_w, bpp) u8 -> (in_h, out_w, bpp) linear float.
__global__ void resize_h_kernel(uint8_t const* __restrict__ in, int in_w, int in_h, int stride, int bpp, int out_w,
                                float const* __restrict__ decode, int const* __restrict__ first,
                                float const* __restrict__ weights, int taps, float* __restrict__ out) {
    __shared__ float dec[256];
    dec[threadIdx.x] = decode[threadIdx.x];  // blockDim.x == 256
    __syncthreads();
    int64_t const t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t const total = (int64_t)in_h * out_w * bpp;
    if (t >= total) return;
    int const c = (int)(t % bpp);
    int64_t r = t / bpp;
    int const ox = (int)(r % out_w);
    int const y = (int)(r / out_w);
    uint8_t const* row = in + (size_t)y * stride;
    int const f = first[ox];
    float const* w = weights + (size_t)ox * taps;
    float acc = 0.0f;
    for (int k = 0; k < taps; ++k) {
        int const x = min(max(f + k, 0), in_w - 1);
        // separate multiply and add: the reference filter is not contracted into FMAs
        acc = __fadd_rn(acc, __fmul_rn(dec[row[x * bpp + c]], w[k]));
    }
    out[t] = acc;
}

// Vertical pass + linear->sRGB8 encode: (in_h, out_w*bpp) float -> (out_h, out_w*bpp) u8.
__global__ void resize_v_kernel(float const* __restrict__ in, int in_h, int row_elems, int out_h,
                                float const* __restrict__ thresholds, int const* __restrict__ first,
                                float const* __restrict__ weights, int taps, uint8_t* __restrict__ out) {
    __shared__ float thr[256];
    thr[threadIdx.x] = thresholds[threadIdx.x];
    __syncthreads();
    int64_t const t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t const total = (int64_t)out_h * row_elems;
    if (t >= total) return;
    int const x = (int)(t % row_elems);
    int const oy = (int)(t / row_elems);
    int const f = first[oy];
    float const* w = weights + (size_t)oy * taps;
    float acc = 0.0f;
    for (int k = 0; k < taps; ++k) {
        int const y = min(max(f + k, 0), in_h - 1);
        acc = __fadd_rn(acc, __fmul_rn(in[(size_t)y * row_elems + x], w[k]));
    }
    out[t] = (uint8_t)srgb_encode(acc, thr);
}

// ---- create_image_tensor (reference segmentation.cpp:81-106): four pixels per thread, 48 contiguous output bytes ----
template <int BPP>
__global__ void __launch_bounds__(256) image_tensor_kernel(uint8_t const* __restrict__ in, int w, int h, int stride, int c0,
                                                           int c1, int c2, float* __restrict__ out) {
    int const groups = (w + 3) >> 2;
    int64_t const t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)groups * h) return;
    int const y = (int)(t / groups), x0 = (int)(t - (int64_t)y * groups) * 4;
    int const n = min(4, w - x0);
    uint8_t const* px = in + (size_t)y * stride + (size_t)x0 * BPP;
    float v[12];
    if (BPP == 4 && n == 4 && (reinterpret_cast<uintptr_t>(px) & 15) == 0) {
        uint4 const q = __ldg(reinterpret_cast<uint4 const*>(px));
        uint32_t const p[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[3 * i + 0] = (float)((p[i] >> (8 * c0)) & 255u);
            v[3 * i + 1] = (float)((p[i] >> (8 * c1)) & 255u);
            v[3 * i + 2] = (float)((p[i] >> (8 * c2)) & 255u);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint8_t const* p = px + i * BPP;
            bool const ok = i < n;
            v[3 * i + 0] = ok ? (float)__ldg(p + c0) : 0.f;
            v[3 * i + 1] = ok ? (float)__ldg(p + c1) : 0.f;
            v[3 * i + 2] = ok ? (float)__ldg(p + c2) : 0.f;
        }
    }
    float* o = out + ((size_t)y * w + x0) * 3;
    if (n == 4 && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
        float4* o4 = reinterpret_cast<float4*>(o);
        o4[0] = make_float4(v[0], v[1], v[2], v[3]);
        o4[1] = make_float4(v[4], v[5], v[6], v[7]);
        o4[2] = make_float4(v[8], v[9], v[10], v[11]);
    } else {
        for (int i = 0; i < 3 * n; ++i) o[i] = v[i];
    }
}

struct Lerp {
    int i0, i1;
    float l0, l1;
};
// torch/ONNX "half pixel" bilinear source coordinate (align_corners = false)
__device__ __forceinline__ Lerp lerp_coord(int dst, float scale, int in_size) {
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    src = src < 0.f ? 0.f : src;
    Lerp r;
    r.i0 = min((int)src, in_size - 1);
    r.i1 = r.i0 + (r.i0 < in_size - 1 ? 1 : 0);
    r.l1 = src - (float)r.i0;
    r.l0 = 1.0f - r.l1;
    return r;
}

__device__ __forceinline__ float sample_1024(float const* __restrict__ low, int yy, int xx) {
    Lerp const ly = lerp_coord(yy, 0.25f, kLowRes), lx = lerp_coord(xx, 0.25f, kLowRes);
    float const* r0 = low + ly.i0 * kLowRes;
    float const* r1 = low + ly.i1 * kLowRes;
    float const top = lx.l0 * __ldg(r0 + lx.i0) + lx.l1 * __ldg(r0 + lx.i1);
    float const bot = lx.l0 * __ldg(r1 + lx.i0) + lx.l1 * __ldg(r1 + lx.i1);
    return ly.l0 * top + ly.l1 * bot;
}

// Per-pixel form (no shared memory): the fallback for outputs wider than 32768 pixels.
__global__ void __launch_bounds__(256) mask_post_kernel(float const* __restrict__ low_res, int64_t plane_stride,
                                                        int const* __restrict__ plane_index, int rw, int rh, int w,
                                                        int h, float sx, float sy,
                                                        uint8_t* const* __restrict__ out_planes,
                                                        uint8_t* __restrict__ out_contig) {
    int const plane = blockIdx.z;
    int const y = blockIdx.y;
    int const x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (x0 >= w) return;
    float const* low = low_res + (plane_index ? plane_index[plane] : plane) * plane_stride;
    uint8_t* dst = out_planes ? out_planes[plane] : out_contig + (size_t)plane * w * h;
    dst += (size_t)y * w + x0;
    Lerp const ly = lerp_coord(y, sy, rh);
    uint8_t m[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int const x = x0 + i;
        m[i] = 0;
        if (x < w) {
            Lerp const lx = lerp_coord(x, sx, rw);
            float const top = lx.l0 * sample_1024(low, ly.i0, lx.i0) + lx.l1 * sample_1024(low, ly.i0, lx.i1);
            float const bot = lx.l0 * sample_1024(low, ly.i1, lx.i0) + lx.l1 * sample_1024(low, ly.i1, lx.i1);
            float const v = ly.l0 * top + ly.l1 * bot;
            m[i] = v > 0.f ? 255 : 0;
        }
    }
    if (x0 + 3 < w && ((reinterpret_cast<uintptr_t>(dst) & 3) == 0)) {
        *reinterpret_cast<uint32_t*>(dst) = (uint32_t)m[0] | ((uint32_t)m[1] << 8) | ((uint32_t)m[2] << 16) | ((uint32_t)m[3] << 24);
    } else {
        for (int i = 0; i < 4 && x0 + i < w; ++i) dst[i] = m[i];
    }
}

// ---- mask post-processing: 256 -> 1024 bilinear, crop, -> (h, w) bilinear, > 0 -> 0 / 255 -----------------------------
// (decoder-graph post-processing, SURVEY A.5, + write_mask_image, reference segmentation.cpp:108-116.)
// Same operations in the same order as the per-pixel form above.  The previous block-of-rows kernel issued one byte
// store per lane (32 B per warp instruction) and reached 0.055 of the HBM rate; here every lane produces 4 or 8
// neighbouring output bytes per row, keeps the horizontally interpolated source rows it needs in registers across the
// output rows of its strip, and writes whole 32- / 64-bit words (128 / 256 contiguous bytes per warp instruction).

// Four mask bytes (v > 0 ? 255 : 0) from four values that were computed with weights scaled by kSignScale = 2^100: the
// scaling is exact (power of two) and makes every positive result a NORMAL float, whose bit pattern read as a signed
// integer is >= 2^23; zero stays 0 and negative values are negative integers, so an unsigned-saturating 8-bit pack of the
// raw bits is the threshold -- two instructions per four pixels instead of 4 FSETP + 4 SEL + 3 PRMT.
constexpr float kSignScale = 1.2676506002282294e30f;  // 2^100
__device__ __forceinline__ uint32_t pack_sign4(float const* v) {
    uint32_t hi, d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(__float_as_int(v[3])), "r"(__float_as_int(v[2])), "r"(0));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(__float_as_int(v[1])), "r"(__float_as_int(v[0])), "r"(hi));
    return d;
}

// Packed fp32 pairs (one instruction for two lanes of arithmetic; per-lane rounding identical to the scalar forms the
// compiler emitted for  l0 * a + l1 * b:  t = l1 * b  rounded,  result = fma(l0, a, t)).
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    float2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<uint64_t&>(d)) : "l"(reinterpret_cast<uint64_t const&>(a)), "l"(reinterpret_cast<uint64_t const&>(b)));
    return d;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(reinterpret_cast<uint64_t&>(d))
        : "l"(reinterpret_cast<uint64_t const&>(a)), "l"(reinterpret_cast<uint64_t const&>(b)), "l"(reinterpret_cast<uint64_t const&>(c)));
    return d;
}

// Identity case: the resized extent equals the output extent (long side == 1024), so the second bilinear has weights
// 1 / 0 and output pixel (y, x) is the thresholded 256 -> 1024 interpolation itself.  A thread owns 8 neighbouring
// columns (two low-resolution cells) of an 8-row strip.  The half-pixel interpolation is periodic -- output column
// 4j + b blends low-resolution columns (j-1, j) with weights (0.375, 0.625), (0.125, 0.875) for b = 0, 1 and (j, j+1) with
// (0.875, 0.125), (0.625, 0.375) for b = 2, 3, exactly the values lerp_coord() produces, and likewise for rows -- so the
// weights are constants: 16 loads per thread, all issued up front, no coordinate arithmetic.  The borders stay on the same
// path (a separate border path made every warp that holds column 0 or 1016 run both: half of all warps): the clamped
// source coordinate of the last two columns / rows is reproduced by clamping the LOAD index (0.875 a + 0.125 a, like the
// generic form), that of the first two (weights exactly 1 and 0) by switching the two weights of those lanes.
constexpr int kIdRows = 8;
__global__ void __launch_bounds__(256) mask_post_identity_kernel(float const* __restrict__ low_res, int64_t plane_stride,
                                                                 int const* __restrict__ plane_index, int w, int h,
                                                                 uint8_t* const* __restrict__ out_planes,
                                                                 uint8_t* __restrict__ out_contig) {
    int const plane = blockIdx.y;
    int const x0 = (threadIdx.x & 127) * 8;
    int const ys = (blockIdx.x * 2 + (threadIdx.x >> 7)) * kIdRows;
    if (x0 >= w || ys >= h) return;
    float const* __restrict__ low = low_res + (plane_index ? plane_index[plane] : plane) * plane_stride;
    uint8_t* const plane_out = out_planes ? out_planes[plane] : out_contig + (size_t)plane * w * h;
    uint8_t* dst = plane_out + (size_t)ys * w + x0;
    bool const word_ok = x0 + 8 <= w && (reinterpret_cast<uintptr_t>(dst) & 7) == 0 && (w & 7) == 0;
    int const rows_valid = min(kIdRows, h - ys);
    int const j = x0 >> 2, i = ys >> 2;  // low-resolution cell of column x0 / row ys
    float a[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        float const* row = low + min(max(i - 1 + r, 0), kLowRes - 1) * kLowRes;
#pragma unroll
        for (int c = 0; c < 4; ++c) a[r][c] = __ldg(row + min(max(j - 1 + c, 0), kLowRes - 1));
    }
    // columns 0, 1 and rows 0, 1: source coordinate clamped to 0 -> weights (1, 0) on (low[0], low[1]) = (0, 1) on the
    // (clamped low[-1] = low[0], low[0]) pair loaded here
    bool const left = x0 == 0, top = ys == 0;
    float const xa0 = left ? 0.f : 0.375f, xb0 = left ? 1.f : 0.625f, xa1 = left ? 0.f : 0.125f, xb1 = left ? 1.f : 0.875f;
    float const ya0 = top ? 0.f : 0.375f, yb0 = top ? 1.f : 0.625f, ya1 = top ? 0.f : 0.125f, yb1 = top ? 1.f : 0.875f;
    // two neighbouring columns per instruction (mul2 / fma2): o = wa * q[j] + wb * q[j + 1] with per-column weights
    float2 const wa01 = make_float2(xa0, xa1), wb01 = make_float2(xb0, xb1);
    float2 const wa23 = make_float2(0.875f, 0.625f), wb23 = make_float2(0.125f, 0.375f);
    float2 const wa45 = make_float2(0.375f, 0.125f), wb45 = make_float2(0.625f, 0.875f);
    auto hrow = [&](float const (&q)[4], float2 (&o)[4]) {
        float2 const q0 = make_float2(q[0], q[0]), q1 = make_float2(q[1], q[1]), q2 = make_float2(q[2], q[2]), q3 = make_float2(q[3], q[3]);
        o[0] = fma2(wa01, q0, mul2(wb01, q1));
        o[1] = fma2(wa23, q1, mul2(wb23, q2));
        o[2] = fma2(wa45, q1, mul2(wb45, q2));
        o[3] = fma2(wa23, q2, mul2(wb23, q3));
    };
    float2 ha[4], hb[4];
    auto emit = [&](int t, float l0, float l1) {
        if (t >= rows_valid) return;
        float2 const l0p = make_float2(l0 * kSignScale, l0 * kSignScale), l1p = make_float2(l1 * kSignScale, l1 * kSignScale);  // exact scaling: pack_sign4
        float2 v2[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) v2[e] = fma2(l0p, ha[e], mul2(l1p, hb[e]));
        float const* v = reinterpret_cast<float const*>(v2);
        uint8_t* d = dst + (size_t)t * w;
        if (word_ok) {
            *reinterpret_cast<uint2*>(d) = make_uint2(pack_sign4(v), pack_sign4(v + 4));
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e)
                if (x0 + e < w) d[e] = v[e] > 0.f ? 255 : 0;
        }
    };
    // rows ys .. ys+7 blend low-resolution rows (i-1, i) for t = 0, 1; (i, i+1) for t = 2..5; (i+1, i+2) for t = 6, 7
    hrow(a[0], ha);
    hrow(a[1], hb);
    emit(0, ya0, yb0);
    emit(1, ya1, yb1);
#pragma unroll
    for (int e = 0; e < 4; ++e) ha[e] = hb[e];
    hrow(a[2], hb);
    emit(2, 0.875f, 0.125f);
    emit(3, 0.625f, 0.375f);
    emit(4, 0.375f, 0.625f);
    emit(5, 0.125f, 0.875f);
#pragma unroll
    for (int e = 0; e < 4; ++e) ha[e] = hb[e];
    hrow(a[3], hb);
    emit(6, 0.875f, 0.125f);
    emit(7, 0.625f, 0.375f);
}

// General case.  A block produces g.rows (32 .. 1, by shared-memory need) complete output rows of one mask; every
// intermediate is computed once per block and kept in shared memory:
//   1. the few low-resolution rows the block needs                                     s_low[n_low][256]
//   2. their horizontal interpolation at the 1024 grid                                 s_h[n_low][1024]
//   3. the vertical interpolation = the 1024-grid rows the block needs                 s_grid[n_slots][1024 + pad]
//      (column rw repeats column rw - 1, so the right neighbour of a sample is always at +4 bytes: no second index)
//   4. output pixels: a thread owns PX neighbouring columns (4 / 8 / 16 by output width: wide outputs amortise the per-row
//      bookkeeping over more pixels and store 16 bytes at once); the second bilinear's horizontal half for the two grid
//      rows of an output row is carried in registers from row to row (when enlarging, several output rows share the
//      pair), and each output row's (grid row pair, weight) comes from a small table built once per block instead of
//      being recomputed by every thread (at PX = 4 the coordinate arithmetic was a third of the 15 instructions per pixel).
struct PostGeom {
    int rw, rh, w, h;
    float sx, sy;
    int n_slots, n_low;  // upper bounds of the 1024-grid rows / low-resolution rows a block needs (host)
    int rows;            // output rows per block
};
constexpr int kGridPitch = kImageSize + 4;  // floats per s_grid row (one repeated column + alignment)

template <int PX>
__global__ void __launch_bounds__(256, PX >= 16 ? 2 : 3) mask_post_tile_kernel(float const* __restrict__ low_res, int64_t plane_stride,
                                                                               int const* __restrict__ plane_index, PostGeom g,
                                                                               uint8_t* const* __restrict__ out_planes,
                                                                               uint8_t* __restrict__ out_contig) {
    extern __shared__ __align__(16) uint8_t post_smem[];
    float* const s_low = reinterpret_cast<float*>(post_smem);
    float* const s_h = s_low + g.n_low * kLowRes;
    float* const s_grid = s_h + g.n_low * kImageSize;
    int4* const s_row = reinterpret_cast<int4*>(s_grid + g.n_slots * kGridPitch);  // per output row: slot a, slot b, weight of b
    int const tid = threadIdx.x;
    int const plane = blockIdx.y;
    int const y0 = blockIdx.x * g.rows;
    int const rows_valid = min(g.rows, g.h - y0);
    float const* low = low_res + (plane_index ? plane_index[plane] : plane) * plane_stride;
    uint8_t* const plane_out = out_planes ? out_planes[plane] : out_contig + (size_t)plane * g.w * g.h;

    // rows of the 1024 grid this block needs, and the low-resolution rows behind them (both contiguous ranges)
    int const g_base = lerp_coord(y0, g.sy, g.rh).i0;
    int const g_last = lerp_coord(y0 + rows_valid - 1, g.sy, g.rh).i1;
    int const l_base = lerp_coord(g_base, 0.25f, kLowRes).i0;
    int const l_last = lerp_coord(g_last, 0.25f, kLowRes).i1;
    int const n_low = l_last - l_base + 1, n_slots = g_last - g_base + 1;
    if (n_low > g.n_low || n_slots > g.n_slots) __trap();  // the host bounds are exact upper bounds

    {   // 1. low-resolution rows (16-byte loads; planes are 256 KiB apart)
        float4 const* src = reinterpret_cast<float4 const*>(low + (size_t)l_base * kLowRes);
        float4* dst = reinterpret_cast<float4*>(s_low);
        for (int i = tid; i < n_low * (kLowRes / 4); i += 256) dst[i] = __ldg(src + i);
    }
    if (tid < rows_valid) {
        Lerp const q = lerp_coord(y0 + tid, g.sy, g.rh);
        s_row[tid] = make_int4(q.i0 - g_base, q.i1 - g_base, __float_as_int(q.l1), 0);
    }
    __syncthreads();
    // 2. horizontal interpolation of those rows at grid columns 0 .. rw-1: a thread owns four neighbouring columns (their
    // sample coordinates are computed once, the results leave as one 16-byte store per low-resolution row)
    if (4 * tid < g.rw) {
        Lerp lx[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) lx[e] = lerp_coord(min(4 * tid + e, kImageSize - 1), 0.25f, kLowRes);
        for (int j = 0; j < n_low; ++j) {
            float const* r = s_low + j * kLowRes;
            float4 o;
            o.x = lx[0].l0 * r[lx[0].i0] + lx[0].l1 * r[lx[0].i1];
            o.y = lx[1].l0 * r[lx[1].i0] + lx[1].l1 * r[lx[1].i1];
            o.z = lx[2].l0 * r[lx[2].i0] + lx[2].l1 * r[lx[2].i1];
            o.w = lx[3].l0 * r[lx[3].i0] + lx[3].l1 * r[lx[3].i1];
            *reinterpret_cast<float4*>(s_h + j * kImageSize + 4 * tid) = o;
        }
    }
    __syncthreads();
    // 3. the 1024-grid rows, four columns per thread (columns past rw - 1 are computed but never read, except the copy
    // of column rw - 1 written behind it)
    if (4 * tid < g.rw) {
        for (int slot = 0; slot < n_slots; ++slot) {
            Lerp const q = lerp_coord(g_base + slot, 0.25f, kLowRes);
            float4 const a = *reinterpret_cast<float4 const*>(s_h + (q.i0 - l_base) * kImageSize + 4 * tid);
            float4 const b = *reinterpret_cast<float4 const*>(s_h + (q.i1 - l_base) * kImageSize + 4 * tid);
            float4 o;
            o.x = q.l0 * a.x + q.l1 * b.x;
            o.y = q.l0 * a.y + q.l1 * b.y;
            o.z = q.l0 * a.z + q.l1 * b.z;
            o.w = q.l0 * a.w + q.l1 * b.w;
            float* dst = s_grid + slot * kGridPitch + 4 * tid;
            *reinterpret_cast<float4*>(dst) = o;
            int const last = g.rw - 1 - 4 * tid;  // position of column rw - 1 among this thread's four
            if (last >= 0 && last < 4) dst[last + 1] = last == 0 ? o.x : last == 1 ? o.y : last == 2 ? o.z : o.w;
        }
    }
    __syncthreads();
    // 4. second bilinear + threshold.  The row loop is NOT unrolled over rows (the unrolled form was 3600 instructions, a
    // third of them branches and predicate logic: profiles/r02a_summary.md); the "row pair changed" branches are uniform.
    bool const word_ok = (g.w & 3) == 0 && (reinterpret_cast<uintptr_t>(plane_out) & 3) == 0;
    bool const vec_ok = PX == 16 && (g.w & 15) == 0 && (reinterpret_cast<uintptr_t>(plane_out) & 15) == 0;
    uint32_t const grid_s = (uint32_t)__cvta_generic_to_shared(s_grid);
    for (int x0 = PX * tid; x0 < g.w; x0 += 256 * PX) {
        uint32_t off[PX];  // byte offset of the left sample inside a grid row
        float l1[PX];
#pragma unroll
        for (int e = 0; e < PX; ++e) {
            Lerp const lx = lerp_coord(min(x0 + e, g.w - 1), g.sx, g.rw);
            off[e] = grid_s + 4u * (uint32_t)lx.i0;
            l1[e] = lx.l1;
        }
        auto hrow = [&](int slot, float2 (&o)[PX / 2]) {
            uint32_t const row_off = (uint32_t)slot * (kGridPitch * 4);
#pragma unroll
            for (int e = 0; e < PX; e += 2) {
                float2 r0, r1;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r0.x) : "r"(off[e] + row_off));
                asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(r1.x) : "r"(off[e] + row_off));
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r0.y) : "r"(off[e + 1] + row_off));
                asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(r1.y) : "r"(off[e + 1] + row_off));
                float2 const w1 = make_float2(l1[e], l1[e + 1]), w0 = make_float2(1.0f - l1[e], 1.0f - l1[e + 1]);
                o[e / 2] = fma2(w0, r0, mul2(w1, r1));
            }
        };
        int ca = -1, cb = -1;
        float2 ha[PX / 2], hb[PX / 2];
        uint8_t* dst = plane_out + (size_t)y0 * g.w + x0;
        bool const full = x0 + PX <= g.w;
#pragma unroll 1
        for (int t = 0; t < rows_valid; ++t, dst += g.w) {
            int4 const q = s_row[t];
            int const a = q.x, b = q.y;
            float const q1u = __int_as_float(q.z);
            float const q0 = (1.0f - q1u) * kSignScale, q1 = q1u * kSignScale;  // exact: see pack_sign4
            if (a != ca) {
                if (a == cb) {
#pragma unroll
                    for (int e = 0; e < PX / 2; ++e) ha[e] = hb[e];
                } else {
                    hrow(a, ha);
                }
                ca = a;
            }
            if (b != cb) {
                if (b == ca) {
#pragma unroll
                    for (int e = 0; e < PX / 2; ++e) hb[e] = ha[e];
                } else {
                    hrow(b, hb);
                }
                cb = b;
            }
            float2 v2[PX / 2];
            float2 const q0p = make_float2(q0, q0), q1p = make_float2(q1, q1);
#pragma unroll
            for (int e = 0; e < PX / 2; ++e) v2[e] = fma2(q0p, ha[e], mul2(q1p, hb[e]));
            float const* v = reinterpret_cast<float const*>(v2);
            if (full && vec_ok) {
                *reinterpret_cast<uint4*>(dst) = make_uint4(pack_sign4(v), pack_sign4(v + 4), pack_sign4(v + (PX > 8 ? 8 : 0)),
                                                            pack_sign4(v + (PX > 8 ? 12 : 0)));
            } else if (full && word_ok) {
#pragma unroll
                for (int e = 0; e < PX; e += 4) *reinterpret_cast<uint32_t*>(dst + e) = pack_sign4(v + e);
            } else {
#pragma unroll
                for (int e = 0; e < PX; ++e)  // (fully unrolled: a dynamic index would put v[] in local memory)
                    if (x0 + e < g.w) dst[e] = v[e] > 0.f ? 255 : 0;
            }
        }
    }
}

// write_mask_image (reference segmentation.cpp:108-116): four outputs per thread.
__global__ void __launch_bounds__(256) threshold_kernel(float const* __restrict__ logits, int tw, int w, int h,
                                                        uint8_t* __restrict__ out) {
    int64_t const t0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    int64_t const total = (int64_t)w * h;
    if (t0 >= total) return;
    uint32_t m[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t const t = t0 + i;
        m[i] = 0;
        if (t < total) {
            int const y = (int)(t / w), x = (int)(t - (int64_t)y * w);
            m[i] = __ldg(logits + (size_t)y * tw + x) > 0 ? 255u : 0u;
        }
    }
    if (t0 + 4 <= total && (reinterpret_cast<uintptr_t>(out + t0) & 3) == 0) {
        *reinterpret_cast<uint32_t*>(out + t0) = m[0] | (m[1] << 8) | (m[2] << 16) | (m[3] << 24);
    } else {
        for (int i = 0; i < 4 && t0 + i < total; ++i) out[t0 + i] = (uint8_t)m[i];
    }
}

template <typename K> void set_smem_limit(K kernel, int bytes) {
    static std::mutex mutex;
    static std::map<void const*, int> done;  // per kernel: largest limit requested so far (per process; all devices alike)
    std::lock_guard<std::mutex> lock(mutex);
    int& cur = done[(void const*)kernel];
    if (bytes > cur) {
        CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        cur = bytes;
    }
}

constexpr int kPostSmemBudget = 72 * 1024;  // three blocks per SM

// shared-memory need of a kRows-row block; fills the exact upper bounds of the rows it may touch
int post_plan(PostGeom& g, int rows) {
    // span of i1(last) - i0(first) + 1 over `n` consecutive destinations at source step `scale`: <= scale*(n-1) + 3
    auto span = [](float scale, int n, int limit) { return std::min(limit, (int)std::ceil((double)scale * (n - 1)) + 3); };
    g.n_slots = span(g.sy, rows, g.rh);
    g.n_low = span(0.25f, g.n_slots, kLowRes);
    return g.n_low * (kLowRes + kImageSize) * 4 + g.n_slots * kGridPitch * 4 + rows * 16;
}

void launch_mask_post(cudaStream_t s, float const* low_res, int64_t plane_stride, int const* plane_index, int count, int rw,
                      int rh, int w, int h, uint8_t* const* out_planes, uint8_t* out_contig) {
    DLIMG_ASSERT(count > 0 && w > 0 && h > 0 && rw > 0 && rh > 0 && rw <= kImageSize && rh <= kImageSize);
    DLIMG_ASSERT(count <= 65535);
    // torch: scale = float(input_size) / output_size
    PostGeom g;
    g.rw = rw; g.rh = rh; g.w = w; g.h = h;
    g.sx = (float)rw / (float)w;
    g.sy = (float)rh / (float)h;
    ProfScope prof(s, CAT_MASK_POST, 0, (double)count * (65536.0 * 4 + (double)w * h));
    if (rw == w && rh == h) {  // long side == 1024: the second resize is the identity
        dim3 grid(ceil_div(h, 2 * kIdRows), count);
        mask_post_identity_kernel<<<grid, 256, 0, s>>>(low_res, plane_stride, plane_index, w, h, out_planes, out_contig);
        KERNEL_CHECK();
        return;
    }
    int rows = 0, smem = 0;
    // wide outputs run the 16-column kernel, whose registers allow two blocks per SM anyway: give them taller blocks
    int const budget = w > 2048 ? 100 * 1024 : kPostSmemBudget;
    for (int r : {32, 16, 8, 4, 2, 1}) {
        smem = post_plan(g, r);
        if (smem <= budget) { rows = r; break; }
    }
    if (rows == 0) {  // cannot happen for rh <= 1024 (one row needs 3 + 3 source rows); kept as a guard
        DLIMG_ASSERT(h <= 65535);
        dim3 block(256), grid(ceil_div(ceil_div(w, 4), 256), h, count);
        mask_post_kernel<<<grid, block, 0, s>>>(low_res, plane_stride, plane_index, rw, rh, w, h, g.sx, g.sy, out_planes, out_contig);
        KERNEL_CHECK();
        return;
    }
    g.rows = rows;
    auto go = [&](auto kernel) {
        set_smem_limit(kernel, smem);
        kernel<<<dim3(ceil_div(h, rows), count), 256, smem, s>>>(low_res, plane_stride, plane_index, g, out_planes, out_contig);
    };
    if (w > 2048) go(mask_post_tile_kernel<16>);      // 256 threads x 16 columns cover a 4K row in one pass
    else if (w > 1024) go(mask_post_tile_kernel<8>);
    else go(mask_post_tile_kernel<4>);
    KERNEL_CHECK();
}

// Tile shape for the single-kernel resize: the largest candidate whose shared-memory need fits the budget.
ResizeTile plan_resize_tile(ResizeDeviceTables const& t, int bpp, int kt, int out_w, int out_h) {
    static int const kRowsCand[] = {32, 16, 8, 4, 2, 1};
    // tile height vs blocks per SM: 64 KB = 16-row tiles at 4K -> 1024, three blocks per SM (100 KB: 156 us, 48 KB: 114 / 129 us
    // for RGB / BGRA against 122 / 109 us, profiles/r02_summary.md)
    static int64_t const budget = (int64_t)std::min(std::max(dev_int("DLIMG_B200_RESIZE_SMEM_KB", 64), 16), 200) * 1024;
    // tall, narrow tiles: two row groups of tc * bpp H-pass columns fill the 256 threads (120 / 128 columns each)
    int const tc_full = bpp == 3 ? 40 : 128 / bpp;
    ResizeTile best;
    for (int tc = tc_full; tc >= 8 && best.tr == 0; tc /= 2) {
        if ((tc * bpp) % 4 != 0) continue;
        for (int tr : kRowsCand) {
            ResizeTile g;
            g.tr = tr;
            g.tc = tc;
            for (int oy0 = 0; oy0 < out_h; oy0 += tr)
                g.nr_max = std::max(g.nr_max, t.vfirst_host[std::min(oy0 + tr, out_h) - 1] + t.vtaps - t.vfirst_host[oy0]);
            for (int ox0 = 0; ox0 < out_w; ox0 += tc)
                g.nc_max = std::max(g.nc_max, t.hfirst_host[std::min(ox0 + tc, out_w) - 1] + t.htaps - t.hfirst_host[ox0]);
            g.ch = 8;
            int64_t const floats = 512 + ((tr * t.vtaps + 3) & ~3) + (int64_t)g.nr_max * tc * bpp + (int64_t)g.ch * (((g.nc_max + kt) * bpp + 32 + 3) & ~3);
            g.raw_pitch = (g.nc_max * bpp + 15 + 15) & ~15;
            int64_t const bytes = floats * 4 + 2 * (int64_t)g.ch * g.raw_pitch;
            if (bytes <= budget) {
                g.smem_bytes = (int)bytes;
                best = g;
                break;
            }
        }
    }
    return best;
}

template <int BPP>
bool launch_resize_tile(cudaStream_t s, uint8_t const* in, int in_w, int in_h, int stride, ResizeDeviceTables const& t,
                        uint8_t* out, int out_w, int out_h) {
    int const kt = t.htaps <= 4 ? 4 : t.htaps <= 8 ? 8 : t.htaps <= 16 ? 16 : t.htaps <= 32 ? 32 : 0;
    if (kt == 0 || !t.hfirst_host || !t.vfirst_host) return false;
    ResizeTile const g = plan_resize_tile(t, BPP, kt, out_w, out_h);
    if (g.tr == 0) return false;
    dim3 grid(ceil_div(out_w, g.tc), ceil_div(out_h, g.tr));
    auto go = [&](auto kernel) {
        set_smem_limit(kernel, g.smem_bytes);
        kernel<<<grid, 256, g.smem_bytes, s>>>(in, in_w, in_h, stride, t, g, out, out_w, out_h);
    };
    switch (kt) {
        case 4: go(resize_tile_kernel<BPP, 4>); break;
        case 8: go(resize_tile_kernel<BPP, 8>); break;
        case 16: go(resize_tile_kernel<BPP, 16>); break;
        default: go(resize_tile_kernel<BPP, 32>); break;
    }
    KERNEL_CHECK();
    return true;
}

}  // namespace

size_t resize_scratch_floats(ResizeDeviceTables const& t, int in_h, int bpp, int out_w, int out_h) {
    int const kt = t.htaps <= 4 ? 4 : t.htaps <= 8 ? 8 : t.htaps <= 16 ? 16 : t.htaps <= 32 ? 32 : 0;
    bool const tiled = kt != 0 && t.hfirst_host && t.vfirst_host && (bpp == 1 || bpp == 3 || bpp == 4) &&
                       plan_resize_tile(t, bpp, kt, out_w, out_h).tr != 0;
    return tiled ? 0 : (size_t)in_h * out_w * bpp;  // only the two-pass fallback needs an intermediate
}

void resize_srgb(cudaStream_t s, uint8_t const* in, int in_w, int in_h, int stride, int bpp, ResizeDeviceTables const& t,
                 float* scratch, uint8_t* out, int out_w, int out_h) {
    ProfScope prof(s, CAT_RESIZE, 0, (double)in_h * in_w * bpp + (double)out_h * out_w * bpp);
    bool done = false;
    if (bpp == 1) done = launch_resize_tile<1>(s, in, in_w, in_h, stride, t, out, out_w, out_h);
    else if (bpp == 3) done = launch_resize_tile<3>(s, in, in_w, in_h, stride, t, out, out_w, out_h);
    else if (bpp == 4) done = launch_resize_tile<4>(s, in, in_w, in_h, stride, t, out, out_w, out_h);
    if (done) return;
    DLIMG_ASSERT(scratch != nullptr);
    int64_t const n1 = (int64_t)in_h * out_w * bpp;
    resize_h_kernel<<<(unsigned)ceil_div64(n1, 256), 256, 0, s>>>(in, in_w, in_h, stride, bpp, out_w, t.decode, t.hfirst,
                                                                 t.hweights, t.htaps, scratch);
    KERNEL_CHECK();
    int64_t const n2 = (int64_t)out_h * out_w * bpp;
    resize_v_kernel<<<(unsigned)ceil_div64(n2, 256), 256, 0, s>>>(scratch, in_h, out_w * bpp, out_h, t.encode_threshold,
                                                                 t.vfirst, t.vweights, t.vtaps, out);
    KERNEL_CHECK();
}

void image_tensor(cudaStream_t s, uint8_t const* in, int w, int h, int stride, int channels, float* out) {
    int cmap[3];
    channel_map(channels, cmap);
    int64_t const n = (int64_t)((w + 3) / 4) * h;
    ProfScope prof(s, CAT_IMAGE_TENSOR, 0, (double)h * w * bytes_per_pixel(channels) + (double)w * h * 12);
    unsigned const blocks = (unsigned)ceil_div64(n, 256);
    switch (bytes_per_pixel(channels)) {
        case 1: image_tensor_kernel<1><<<blocks, 256, 0, s>>>(in, w, h, stride, cmap[0], cmap[1], cmap[2], out); break;
        case 3: image_tensor_kernel<3><<<blocks, 256, 0, s>>>(in, w, h, stride, cmap[0], cmap[1], cmap[2], out); break;
        default: image_tensor_kernel<4><<<blocks, 256, 0, s>>>(in, w, h, stride, cmap[0], cmap[1], cmap[2], out); break;
    }
    KERNEL_CHECK();
}

void mask_postprocess(cudaStream_t s, float const* low_res, int64_t plane_stride, int const* plane_index, int count, int rw,
                      int rh, int w, int h, uint8_t* const* out_planes) {
    launch_mask_post(s, low_res, plane_stride, plane_index, count, rw, rh, w, h, out_planes, nullptr);
}

void mask_postprocess_contiguous(cudaStream_t s, float const* low_res, int64_t plane_stride, int const* plane_index,
                                 int count, int rw, int rh, int w, int h, uint8_t* out) {
    launch_mask_post(s, low_res, plane_stride, plane_index, count, rw, rh, w, h, nullptr, out);
}

void threshold_mask(cudaStream_t s, float const* logits, int th, int tw, int w, int h, uint8_t* out) {
    DLIMG_ASSERT(w <= tw && h <= th);
    ProfScope prof(s, CAT_MASK_POST, 0, (double)w * h * 5);
    int64_t const n = ((int64_t)w * h + 3) / 4;
    threshold_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, s>>>(logits, tw, w, h, out);
    KERNEL_CHECK();
}

}  // namespace prepost
}  // namespace dlimg
