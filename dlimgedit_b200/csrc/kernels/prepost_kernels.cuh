// prepost_kernels.cuh -- byte/pixel kernels either side of the network: resize (reference
// image.cpp:37-51), channel map (segmentation.cpp:81-106), mask upsample + threshold
// (decoder-graph post-processing, SURVEY A.5, + segmentation.cpp:108-116).
#pragma once

#include "../common.hpp"

#include <vector>

namespace dlimg {
namespace prepost {

// Per-axis resampling plan: output pixel o reads taps first[o] .. first[o]+taps-1 (clamped to the
// image) with weights w[o*taps + k].  Built on the host once per (in, out) size pair.
struct AxisPlan {
    int in_size = 0, out_size = 0, taps = 0;
    std::vector<int> first;
    std::vector<float> weights;
};

// stb_image_resize v0.97 semantics of the reference call: Catmull-Rom when
// enlarging, Mitchell-Netravali evaluated in output space when shrinking, clamped edges, weights
// normalised per output pixel.
AxisPlan make_axis_plan(int in_size, int out_size);

// 256-entry sRGB->linear table (values rounded to 6 decimals as in stb's source table) followed by the
// 255 linear thresholds of the correctly rounded linear->sRGB8 encoder.
struct SrgbTables {
    float decode[256];
    float encode_threshold[256];  // [i] = smallest float that encodes to >= i (i = 1..255); [0] = -inf
};
SrgbTables const& srgb_tables();

struct ResizeDeviceTables {
    float const* decode;            // [256]
    float const* encode_threshold;  // [256]
    int const* hfirst; float const* hweights; int htaps;
    int const* vfirst; float const* vweights; int vtaps;
    int const* hfirst_host = nullptr;  // host copies of the first-tap tables: the tile planner reads them
    int const* vfirst_host = nullptr;
};

// Floats of scratch resize_srgb needs: 0 when the single-kernel tile form applies (<= 32 horizontal taps and a tile
// that fits shared memory), in_h*out_w*bpp for the two-pass fallback.
size_t resize_scratch_floats(ResizeDeviceTables const& t, int in_h, int bpp, int out_w, int out_h);

// u8 (in_h, in_w, bpp) with byte stride -> packed u8 (out_h, out_w, bpp).
void resize_srgb(cudaStream_t s, uint8_t const* in, int in_w, int in_h, int stride, int bpp, ResizeDeviceTables const& t,
                 float* scratch, uint8_t* out, int out_w, int out_h);

// u8 strided pixels -> float32 (h, w, 3) with the reference channel map; values 0..255.
void image_tensor(cudaStream_t s, uint8_t const* in, int w, int h, int stride, int channels, float* out);

// `count` low-res logit planes (256x256 fp32, plane i at low_res + i*plane_stride) -> `count` packed u8
// masks (h, w): bilinear 256->1024 (align_corners=false), crop to the resized extent (rh, rw), bilinear to
// (h, w), > 0 -> 255.  out_planes[i] gives the destination of mask i (device pointers, device array).
// plane_index (device, optional): source plane of output i (mask selection happens on the device).
void mask_postprocess(cudaStream_t s, float const* low_res, int64_t plane_stride, int const* plane_index, int count, int rw,
                      int rh, int w, int h, uint8_t* const* out_planes);
// Same, contiguous output (count, h, w).
void mask_postprocess_contiguous(cudaStream_t s, float const* low_res, int64_t plane_stride, int const* plane_index,
                                 int count, int rw, int rh, int w, int h, uint8_t* out);

// write_mask_image: logits plane (th, tw) fp32 -> u8 (h, w), strictly `> 0`.
void threshold_mask(cudaStream_t s, float const* logits, int th, int tw, int w, int h, uint8_t* out);

// ResizeLongestSide (segmentation.cpp:60-74): float32 scale, round-half-up of dim*scale.
struct LongestSide {
    int orig_w = 0, orig_h = 0;
    int w = 0, h = 0;  // resized extent
    float scale = 1.0f;
    bool needs_resize = false;
};
LongestSide resize_longest_side(int w, int h, int max_side);
inline int scale_coord(int c, float scale) { return int(c * scale + 0.5f); }

}  // namespace prepost
}  // namespace dlimg
