// prepost_kernels.cu -- see prepost_kernels.cuh.
#include "prepost_kernels.cuh"

#include "../profiler.hpp"

#include <cstdlib>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <limits>

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
    // code = number of thresholds <= acc, thresholds ascending (thr[0] = -inf)
    int lo = 0, hi = 255;
    while (lo < hi) {
        int const mid = (lo + hi + 1) >> 1;
        if (acc >= thr[mid]) lo = mid;
        else hi = mid - 1;
    }
    out[t] = (uint8_t)lo;
}

__global__ void image_tensor_kernel(uint8_t const* __restrict__ in, int w, int h, int stride, int bpp, int c0, int c1,
                                    int c2, float* __restrict__ out) {
    int64_t const t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)w * h) return;
    int const x = (int)(t % w), y = (int)(t / w);
    uint8_t const* px = in + (size_t)y * stride + (size_t)x * bpp;
    float* o = out + t * 3;
    o[0] = (float)px[c0];
    o[1] = (float)px[c1];
    o[2] = (float)px[c2];
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

// 4 output pixels per thread along x, written as one 32-bit store when aligned.
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

// Same arithmetic, organised per block of kPostRows output rows: the 1024-grid samples those rows need (two grid
// rows each, de-duplicated) are computed once into shared memory, then every output pixel combines four of them.
// The per-pixel form above evaluates sample_1024 four times per output pixel; here it is ~1.25 (1024^2 output) to
// ~0.3 (4K output) times, with bit-identical results (same operations in the same order).
constexpr int kPostRows = 4;
constexpr int kPostSlots = 2 * kPostRows;      // 1024-grid rows a block may need
constexpr int kPostLowRows = 2 * kPostSlots;   // low-resolution rows those may need

__global__ void __launch_bounds__(256) mask_post_rows_kernel(float const* __restrict__ low_res, int64_t plane_stride,
                                                             int const* __restrict__ plane_index, int rw, int rh, int w,
                                                             int h, float sx, float sy,
                                                             uint8_t* const* __restrict__ out_planes,
                                                             uint8_t* __restrict__ out_contig) {
    extern __shared__ __align__(16) float post_smem[];
    float (*grid_rows)[kImageSize] = reinterpret_cast<float (*)[kImageSize]>(post_smem);
    float (*low_rows)[kLowRes] = reinterpret_cast<float (*)[kLowRes]>(post_smem + kPostSlots * kImageSize);
    int const plane = blockIdx.y;
    int const y0 = blockIdx.x * kPostRows;
    float const* low = low_res + (plane_index ? plane_index[plane] : plane) * plane_stride;
    // Every thread derives the (tiny) row tables itself -- no serial set-up phase.  The 1024-grid rows needed by the
    // block's output rows are consecutive when the vertical scale is <= 1.75 (slot = row - first row); otherwise each
    // output row gets its own two slots.  Likewise for the low-resolution rows behind them.
    Lerp ly[kPostRows];
#pragma unroll
    for (int t = 0; t < kPostRows; ++t) ly[t] = lerp_coord(min(y0 + t, h - 1), sy, rh);
    int const g_base = ly[0].i0, g_span = ly[kPostRows - 1].i1 - g_base + 1;
    bool const g_dense = g_span <= kPostSlots;
    int const n_slots = g_dense ? g_span : kPostSlots;
    auto grid_row_of = [&](int slot) {
        if (g_dense) return g_base + slot;
        int r = 0;
#pragma unroll
        for (int t = 0; t < kPostRows; ++t) {
            if (slot == 2 * t) r = ly[t].i0;
            if (slot == 2 * t + 1) r = ly[t].i1;
        }
        return r;
    };
    int const l_base = lerp_coord(grid_row_of(0), 0.25f, kLowRes).i0;
    int const l_span = lerp_coord(grid_row_of(n_slots - 1), 0.25f, kLowRes).i1 - l_base + 1;
    bool const l_dense = g_dense && l_span <= kPostLowRows;
    int const n_low = l_dense ? l_span : 2 * n_slots;
    for (int i = threadIdx.x; i < n_low * kLowRes; i += blockDim.x) {
        int const j = i >> 8;
        int row;
        if (l_dense) {
            row = l_base + j;
        } else {
            Lerp const q = lerp_coord(grid_row_of(j >> 1), 0.25f, kLowRes);
            row = (j & 1) ? q.i1 : q.i0;
        }
        low_rows[j][i & 255] = __ldg(low + row * kLowRes + (i & 255));
    }
    __syncthreads();
    // grid_rows[slot][xx] = sample_1024(low, grid row of slot, xx), taps from shared memory; the horizontal
    // coordinates are computed once per column and reused for every slot
    for (int xx = threadIdx.x; xx < rw; xx += blockDim.x) {
        Lerp const lx = lerp_coord(xx, 0.25f, kLowRes);
        for (int slot = 0; slot < n_slots; ++slot) {
            Lerp const q = lerp_coord(grid_row_of(slot), 0.25f, kLowRes);
            float const* r0 = low_rows[l_dense ? q.i0 - l_base : 2 * slot];
            float const* r1 = low_rows[l_dense ? q.i1 - l_base : 2 * slot + 1];
            float const top = lx.l0 * r0[lx.i0] + lx.l1 * r0[lx.i1];
            float const bot = lx.l0 * r1[lx.i0] + lx.l1 * r1[lx.i1];
            grid_rows[slot][xx] = q.l0 * top + q.l1 * bot;
        }
    }
    __syncthreads();
    uint8_t* const plane_out = out_planes ? out_planes[plane] : out_contig + (size_t)plane * w * h;
    // one pixel per lane and iteration: neighbouring lanes read neighbouring grid samples (no bank conflicts) and
    // their byte stores coalesce into whole sectors; the horizontal coordinate is shared by the block's rows
    for (int x = threadIdx.x; x < w; x += blockDim.x) {
        Lerp const lx = lerp_coord(x, sx, rw);
#pragma unroll
        for (int t = 0; t < kPostRows; ++t) {
            int const y = y0 + t;
            if (y < h) {
                float const* r0 = grid_rows[g_dense ? ly[t].i0 - g_base : 2 * t];
                float const* r1 = grid_rows[g_dense ? ly[t].i1 - g_base : 2 * t + 1];
                float const top = lx.l0 * r0[lx.i0] + lx.l1 * r0[lx.i1];
                float const bot = lx.l0 * r1[lx.i0] + lx.l1 * r1[lx.i1];
                float const v = ly[t].l0 * top + ly[t].l1 * bot;
                plane_out[(size_t)y * w + x] = v > 0.f ? 255 : 0;
            }
        }
    }
}

__global__ void threshold_kernel(float const* __restrict__ logits, int tw, int w, int h, uint8_t* __restrict__ out) {
    int64_t const t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)w * h) return;
    int const x = (int)(t % w), y = (int)(t / w);
    out[t] = logits[(size_t)y * tw + x] > 0 ? 255 : 0;
}

void launch_mask_post(cudaStream_t s, float const* low_res, int64_t plane_stride, int const* plane_index, int count, int rw,
                      int rh, int w, int h, uint8_t* const* out_planes, uint8_t* out_contig) {
    DLIMG_ASSERT(count > 0 && w > 0 && h > 0 && rw > 0 && rh > 0 && rw <= kImageSize && rh <= kImageSize);
    DLIMG_ASSERT(h <= 65535);
    // torch: scale = float(input_size) / output_size
    float const sx = (float)rw / (float)w, sy = (float)rh / (float)h;
    ProfScope prof(s, CAT_MASK_POST, 0, (double)count * (65536.0 * 4 + (double)w * h));
    static bool const per_pixel = std::getenv("DLIMG_B200_MASK_POST_PIXEL") != nullptr;  // A/B: the per-pixel form
    if (per_pixel) {
        dim3 block(256), grid(ceil_div(ceil_div(w, 4), 256), h, count);
        mask_post_kernel<<<grid, block, 0, s>>>(low_res, plane_stride, plane_index, rw, rh, w, h, sx, sy, out_planes, out_contig);
    } else {
        dim3 grid(ceil_div(h, kPostRows), count);
        constexpr int kPostSmem = (kPostSlots * kImageSize + kPostLowRows * kLowRes) * (int)sizeof(float);
        static bool attr_set = false;
        if (!attr_set) {
            CUDA_CHECK(cudaFuncSetAttribute(mask_post_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPostSmem));
            attr_set = true;
        }
        mask_post_rows_kernel<<<grid, 256, kPostSmem, s>>>(low_res, plane_stride, plane_index, rw, rh, w, h, sx, sy, out_planes, out_contig);
    }
    KERNEL_CHECK();
}

}  // namespace

void resize_srgb(cudaStream_t s, uint8_t const* in, int in_w, int in_h, int stride, int bpp, ResizeDeviceTables const& t,
                 float* scratch, uint8_t* out, int out_w, int out_h) {
    int64_t const n1 = (int64_t)in_h * out_w * bpp;
    ProfScope prof(s, CAT_RESIZE, 0, (double)in_h * stride + (double)out_h * out_w * bpp);
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
    int64_t const n = (int64_t)w * h;
    ProfScope prof(s, CAT_IMAGE_TENSOR, 0, (double)h * stride + (double)n * 12);
    image_tensor_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, s>>>(in, w, h, stride, bytes_per_pixel(channels), cmap[0],
                                                                    cmap[1], cmap[2], out);
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
    int64_t const n = (int64_t)w * h;
    threshold_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, s>>>(logits, tw, w, h, out);
    KERNEL_CHECK();
}

}  // namespace prepost
}  // namespace dlimg
