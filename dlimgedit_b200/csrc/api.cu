// api.cu -- the C ABI: dlimg_init() (13-slot table, layout of reference dlimgedit.h:44-68, filled like
// dlimgedit.cpp:102-117), dlimg_b200_ext_init() (additive batch / device extension) and
// dlimg_b200_debug_init() (kernel-level entry points used by the parity tests only).
#include "../../include/dlimg_b200.h"
#include "../../include/dlimg_b200_debug.h"
#include "engine.hpp"
#include "image_io.hpp"
#include "image_pool.hpp"
#include "kernels/gemm.cuh"
#include "profiler.hpp"

#include <cstring>

namespace dlimg {
namespace {

thread_local std::string last_error_;  // the reference keeps a racy process-global (dlimgedit.cpp:12)

dlimg_Result set_last_error(char const* what) {
    last_error_ = what;
    return dlimg_error;
}

// reference dlimgedit.cpp:31-40: no exception crosses the boundary
template <typename F> dlimg_Result try_(F const& f) {
    try {
        f();
    } catch (std::exception const& e) {
        return set_last_error(e.what());
    } catch (...) {
        return set_last_error("Unknown error");
    }
    return dlimg_success;
}

EnvironmentImpl& to_impl(dlimg_Environment h) { return *reinterpret_cast<EnvironmentImpl*>(h); }
SegmentationImpl& to_impl(dlimg_Segmentation h) { return *reinterpret_cast<SegmentationImpl*>(h); }

// ---- part 1 -----------------------------------------------------------------------------------
int is_backend_supported(dlimg_Backend backend) { return EnvironmentImpl::is_supported(backend) ? 1 : 0; }

dlimg_Result create_environment(dlimg_Environment* handle, dlimg_Options const* options) {
    return try_([=] {
        DLIMG_ASSERT(handle && options);
        *handle = reinterpret_cast<dlimg_Environment>(new EnvironmentImpl(*options));
    });
}

void destroy_environment(dlimg_Environment handle) { delete reinterpret_cast<EnvironmentImpl*>(handle); }

dlimg_Result process_image_for_segmentation(dlimg_Segmentation* handle, dlimg_ImageView const* img, dlimg_Environment env) {
    return try_([=] {
        DLIMG_ASSERT(handle && img && env);
        auto seg = new SegmentationImpl(to_impl(env));
        *handle = reinterpret_cast<dlimg_Segmentation>(seg);  // assigned first: the caller owns it even on failure
        seg->process(*img);
    });
}

dlimg_Result get_segmentation_mask(dlimg_Segmentation handle, int const* point, int const* region, uint8_t** result_masks,
                                   float* result_accuracy) {
    return try_([=] {
        DLIMG_ASSERT(handle && result_masks);
        to_impl(handle).compute_mask(point, region, result_masks, result_accuracy);
    });
}

void get_segmentation_extent(dlimg_Segmentation handle, int* out_extent) {
    out_extent[0] = to_impl(handle).width();
    out_extent[1] = to_impl(handle).height();
}

void destroy_segmentation(dlimg_Segmentation handle) { delete reinterpret_cast<SegmentationImpl*>(handle); }

dlimg_Result segment_objects(dlimg_ImageView const*, uint8_t*, dlimg_Environment) {
    return set_last_error("segment_objects (BiRefNet) is not supported by the B200 segmentation engine");
}

dlimg_Result load_image_api(char const* filepath, int* out_extent, int* out_channels, uint8_t** out_pixels) {
    return try_([=] { *out_pixels = load_image(filepath, out_extent, out_channels); });
}

dlimg_Result save_image_api(dlimg_ImageView const* img, char const* filepath) {
    return try_([=] { save_image(*img, filepath); });
}

// Pixels handed out by this library (create_image, load_image) all come from image_alloc, so destroy_image's
// image_free is always matched (the reference mixes malloc and delete[]).  While a GPU environment is
// alive these buffers are page-locked and recycled by size (image_pool.hpp).
uint8_t* create_image(int w, int h, int channels) {
    if (w <= 0 || h <= 0 || channels <= 0) return nullptr;
    try {
        return image_alloc((size_t)w * h * channels);
    } catch (std::exception const&) {
        return nullptr;
    }
}
void destroy_image(uint8_t const* pixels) { image_free(pixels); }

char const* last_error() { return last_error_.c_str(); }

dlimg_Api api_;

// ---- part 2 -----------------------------------------------------------------------------------
dlimg_Result ext_set_stream(dlimg_Environment env, void* stream) {
    return try_([=] { to_impl(env).set_stream(static_cast<cudaStream_t>(stream)); });
}

dlimg_Result ext_synchronize(dlimg_Environment env) {
    return try_([=] {
        to_impl(env).synchronize();
    });
}

dlimg_Result ext_get_stats(dlimg_Environment env, dlimg_b200_Stats* out) {
    return try_([=] {
        DLIMG_ASSERT(env && out);
        to_impl(env).stats(*out);
    });
}

dlimg_Result ext_process_batch(dlimg_Environment env, dlimg_ImageView const* views, int count, int on_device,
                               dlimg_Segmentation* out) {
    return try_([=] {
        DLIMG_ASSERT(env && views && out && count >= 0);
        std::vector<SegmentationImpl*> segs((size_t)count);
        for (int i = 0; i < count; ++i) {
            segs[(size_t)i] = new SegmentationImpl(to_impl(env));
            out[i] = reinterpret_cast<dlimg_Segmentation>(segs[(size_t)i]);
        }
        to_impl(env).process_batch(views, count, on_device != 0, segs.data());
    });
}

dlimg_Result ext_compute_masks_batch(dlimg_Environment env, dlimg_Segmentation const* segs, dlimg_b200_Prompt const* prompts,
                                     int count, int multi, uint8_t* const* masks_out, float* ious_out, int on_device) {
    return try_([=] {
        DLIMG_ASSERT(env && segs && prompts && masks_out && count >= 0);
        int const n = multi ? 3 : 1;
        std::vector<SegmentationImpl*> impl((size_t)count);
        std::vector<uint8_t*> planes((size_t)count * n);
        for (int i = 0; i < count; ++i) {
            impl[(size_t)i] = reinterpret_cast<SegmentationImpl*>(segs[i]);
            DLIMG_ASSERT(impl[(size_t)i] && masks_out[i]);
            size_t const plane = (size_t)impl[(size_t)i]->width() * impl[(size_t)i]->height();
            for (int m = 0; m < n; ++m) planes[(size_t)i * n + m] = masks_out[i] + plane * m;
        }
        to_impl(env).compute_masks_batch(impl.data(), prompts, count, multi != 0, planes.data(), ious_out, on_device);
    });
}

dlimg_Result ext_get_embedding(dlimg_Segmentation seg, float* out_host) {
    return try_([=] { to_impl(seg).embedding_nchw(out_host); });
}

dlimg_Result ext_get_embedding_async(dlimg_Segmentation seg, float* out_host) {
    return try_([=] { to_impl(seg).embedding_nchw_async(out_host); });
}

dlimg_Result ext_get_embedding_f16_async(dlimg_Segmentation seg, uint16_t* out_host) {
    return try_([=] { to_impl(seg).embedding_nchw_f16_async(out_host); });
}

dlimg_Result ext_get_low_res_logits(dlimg_Segmentation seg, dlimg_b200_Prompt const* prompt, float* logits, float* iou) {
    return try_([=] { to_impl(seg).environment().low_res_logits(to_impl(seg), *prompt, logits, iou); });
}

dlimg_Result ext_resize_longest_side(dlimg_Environment env, dlimg_ImageView const* v, int max_side, uint8_t* out, int* extent) {
    return try_([=] { to_impl(env).resize_longest_side(*v, max_side, out, extent); });
}
dlimg_Result ext_image_tensor(dlimg_Environment env, dlimg_ImageView const* v, float* out) {
    return try_([=] { to_impl(env).image_tensor(*v, out); });
}
dlimg_Result ext_mask_postprocess(dlimg_Environment env, float const* low, int count, int w, int h, uint8_t* out) {
    return try_([=] { to_impl(env).mask_postprocess(low, count, w, h, out); });
}
dlimg_Result ext_threshold_mask(dlimg_Environment env, float const* logits, int th, int tw, int w, int h, uint8_t* out) {
    return try_([=] { to_impl(env).threshold_mask(logits, th, tw, w, h, out); });
}

dlimg_Result ext_profile_enable(dlimg_Environment env, int on) {
    return try_([=] { to_impl(env).profile_enable(on != 0); });
}
dlimg_Result ext_profile_read(dlimg_Environment env, dlimg_b200_ProfileEntry* out, int capacity, int* count) {
    return try_([=] {
        auto const totals = to_impl(env).profile_collect();
        int n = 0;
        for (int c = 0; c < CAT_COUNT && n < capacity; ++c) {
            if (!totals[(size_t)c].launches) continue;
            dlimg_b200_ProfileEntry& e = out[n++];
            std::memset(&e, 0, sizeof(e));
            std::strncpy(e.name, kernel_cat_name(c), sizeof(e.name) - 1);
            e.launches = totals[(size_t)c].launches;
            e.ms = totals[(size_t)c].ms;
            e.flops = totals[(size_t)c].flops;
            e.bytes = totals[(size_t)c].bytes;
        }
        *count = n;
    });
}

dlimg_b200_Ext ext_;

// ---- debug table (tests only; declared in include/dlimg_b200_debug.h) -----------------------------
using DebugApi = dlimg_b200_Debug;

dlimg_Result dbg_gemm(void* stream, int tf32, int simt, void const* a, void const* b, int M, int N, int K, float const* bias,
                      void const* residual, int const* row_map, int act, int out_f32, void* out, float const* ln_stats,
                      int ln_parts, float* stats_out) {
    return try_([=] {
        gemm::Operand A{a, M, K, K}, B{b, N, K, K};
        gemm::Epilogue e;
        e.bias = bias;
        e.residual = residual;
        e.row_map = row_map;
        e.act = act;
        e.out_f32 = out_f32;
        e.ldc = N;
        e.ln_stats = reinterpret_cast<float2 const*>(ln_stats);
        e.ln_parts = ln_parts;
        e.stats_out = reinterpret_cast<float2*>(stats_out);
        int dev = 0;
        CUDA_CHECK(cudaGetDevice(&dev));
        cudaDeviceProp prop;
        CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
        if (simt >= 2) e.ksplit = simt;  // split-K on the tensor-core kernel: out holds `simt` partial results
        if (simt == 1) gemm::launch_simt(static_cast<cudaStream_t>(stream), tf32 != 0, A, B, out, e);
        else gemm::launch(static_cast<cudaStream_t>(stream), tf32 != 0, A, B, out, e, prop.multiProcessorCount);
    });
}

dlimg_Result dbg_encode_tap(dlimg_Environment env, dlimg_ImageView const* views, int count, char const* name, float* out,
                            size_t capacity, size_t* written) {
    return try_([=] { *written = to_impl(env).encode_tap(views, count, name, out, capacity); });
}

int dbg_resize_plan(int in_size, int out_size, int max_taps, int* first, float* weights) {
    try {
        prepost::AxisPlan const p = prepost::make_axis_plan(in_size, out_size);
        if (p.taps > max_taps) return -1;
        for (int o = 0; o < out_size; ++o) {
            first[o] = p.first[(size_t)o];
            for (int k = 0; k < max_taps; ++k) weights[(size_t)o * max_taps + k] = k < p.taps ? p.weights[(size_t)o * p.taps + k] : 0.f;
        }
        return p.taps;
    } catch (...) {
        return -2;
    }
}

void dbg_srgb_tables(float* decode256, float* threshold256) {
    auto const& t = prepost::srgb_tables();
    std::memcpy(decode256, t.decode, sizeof(t.decode));
    std::memcpy(threshold256, t.encode_threshold, sizeof(t.encode_threshold));
}

dlimg_Result dbg_window_attention(void* stream, void const* qkv, int batch, int res, int ws, int heads, void const* pad_qkv,
                                  float const* bias, void* out) {
    return try_([=] {
        auto s = static_cast<cudaStream_t>(stream);
        int const n = ws * ws;
        // re-order the dense (heads, n, n) table for the tensor-core kernel
        std::vector<float> dense((size_t)heads * n * n);
        std::vector<uint16_t> frag(enc::attention_bias_fragment_count(heads, ws));
        CUDA_CHECK(cudaMemcpy(dense.data(), bias, dense.size() * sizeof(float), cudaMemcpyDeviceToHost));
        enc::attention_bias_fragments(dense.data(), heads, ws, frag.data());
        DeviceBuffer<uint16_t> dfrag;
        dfrag.upload(frag);
        int dev = 0;
        CUDA_CHECK(cudaGetDevice(&dev));
        cudaDeviceProp prop;
        CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
        enc::window_attention(s, static_cast<act_t const*>(qkv), batch, res, ws, heads, static_cast<act_t const*>(pad_qkv),
                              dfrag.get(), static_cast<act_t*>(out), prop.multiProcessorCount);
        CUDA_CHECK(cudaStreamSynchronize(s));
    });
}

dlimg_Result dbg_window_attention_simt(void* stream, void const* qkv, int windows, int n, int heads, float const* bias, void* out) {
    return try_([=] {
        enc::window_attention_simt(static_cast<cudaStream_t>(stream), static_cast<act_t const*>(qkv), windows, n, heads, bias,
                                   static_cast<act_t*>(out));
    });
}

dlimg_Result dbg_local_conv(void* stream, void const* in, int batch, int H, int W, int C, float const* weight, float const* bias,
                            void* out, float* stats, int tma) {
    return try_([=] {
        int dev = 0;
        CUDA_CHECK(cudaGetDevice(&dev));
        cudaDeviceProp prop;
        CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
        auto const s = static_cast<cudaStream_t>(stream);
        if (tma) {
            if (!enc::local_conv_tma_supported(H, W, C)) fail("local_conv: unsupported shape for the TMA kernel");
            enc::local_conv_tma(s, static_cast<act_t const*>(in), batch, H, W, C, weight, bias, static_cast<act_t*>(out),
                                reinterpret_cast<float2*>(stats), prop.multiProcessorCount);
        } else {
            enc::dwconv3x3_stats(s, static_cast<act_t const*>(in), batch, H, W, C, weight, bias, static_cast<act_t*>(out),
                                 reinterpret_cast<float2*>(stats));
        }
    });
}

void dbg_image_pool_stats(uint64_t* out5) {
    ImagePoolStats const st = image_pool_stats();
    out5[0] = st.pinned_in_use;
    out5[1] = st.pinned_cached;
    out5[2] = st.pinned_allocs;
    out5[3] = st.reuses;
    out5[4] = st.plain_allocs;
}

dlimg_Result dbg_conv3x3(void* stream, void const* in, int batch, int H, int W, int C, void const* weight, float const* bias, int N,
                         void* out) {
    return try_([=] {
        int dev = 0;
        CUDA_CHECK(cudaGetDevice(&dev));
        cudaDeviceProp prop;
        CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
        gemm::Epilogue e;
        e.bias = bias;
        e.ldc = N;
        gemm::launch_conv3x3(static_cast<cudaStream_t>(stream), in, batch, H, W, C, gemm::Operand{weight, N, 9 * (int64_t)C, 9 * (int64_t)C},
                             out, e, prop.multiProcessorCount);
    });
}

dlimg_Result dbg_mlp_fused(void* stream, void const* x, int rows, int C, void const* w1, float const* b1, float const* ln_sums,
                           void const* w2, float const* b2, void* out, float* stats_out) {
    return try_([=] {
        int dev = 0;
        CUDA_CHECK(cudaGetDevice(&dev));
        cudaDeviceProp prop;
        CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
        if (!gemm::mlp_fused_supported(C)) fail("mlp_fused: unsupported width " + std::to_string(C));
        gemm::launch_mlp_fused(static_cast<cudaStream_t>(stream), x, rows, C, w1, b1, reinterpret_cast<float2 const*>(ln_sums), 1e-5f, w2,
                               b2, out, reinterpret_cast<float2*>(stats_out), prop.multiProcessorCount);
    });
}

dlimg_Result dbg_layernorm_stats(void* stream, void const* in, int rows, int C, float eps, float* out) {
    return try_([=] {
        enc::layernorm_stats(static_cast<cudaStream_t>(stream), static_cast<act_t const*>(in), rows, C, eps, reinterpret_cast<float2*>(out));
    });
}

DebugApi debug_;

}  // namespace

}  // namespace dlimg

extern "C" {

DLIMG_B200_EXPORT dlimg_Api const* dlimg_init(void) {
    using namespace dlimg;
    api_.is_backend_supported = is_backend_supported;
    api_.create_environment = create_environment;
    api_.destroy_environment = destroy_environment;
    api_.process_image_for_segmentation = process_image_for_segmentation;
    api_.get_segmentation_mask = get_segmentation_mask;
    api_.get_segmentation_extent = get_segmentation_extent;
    api_.destroy_segmentation = destroy_segmentation;
    api_.segment_objects = segment_objects;
    api_.load_image = load_image_api;
    api_.save_image = save_image_api;
    api_.create_image = create_image;
    api_.destroy_image = destroy_image;
    api_.last_error = last_error;
    return &api_;
}

DLIMG_B200_EXPORT dlimg_b200_Ext const* dlimg_b200_ext_init(void) {
    using namespace dlimg;
    ext_.struct_size = sizeof(dlimg_b200_Ext);
    ext_.abi_version = 3;
    ext_.set_stream = ext_set_stream;
    ext_.synchronize = ext_synchronize;
    ext_.get_stats = ext_get_stats;
    ext_.process_batch = ext_process_batch;
    ext_.compute_masks_batch = ext_compute_masks_batch;
    ext_.get_embedding = ext_get_embedding;
    ext_.get_low_res_logits = ext_get_low_res_logits;
    ext_.resize_longest_side = ext_resize_longest_side;
    ext_.image_tensor = ext_image_tensor;
    ext_.mask_postprocess = ext_mask_postprocess;
    ext_.threshold_mask = ext_threshold_mask;
    ext_.profile_enable = ext_profile_enable;
    ext_.profile_read = ext_profile_read;
    ext_.get_embedding_async = ext_get_embedding_async;
    ext_.get_embedding_f16_async = ext_get_embedding_f16_async;
    return &ext_;
}

// Test-only entry points (kernel-level parity checks); not part of the supported interface.
DLIMG_B200_EXPORT dlimg_b200_Debug const* dlimg_b200_debug_init(void) {
    using namespace dlimg;
    debug_.struct_size = sizeof(DebugApi);
    debug_.act_is_bf16 = kActBf16 ? 1 : 0;
    debug_.gemm = dbg_gemm;
    debug_.encode_tap = dbg_encode_tap;
    debug_.resize_plan = dbg_resize_plan;
    debug_.srgb_tables = dbg_srgb_tables;
    debug_.window_attention = dbg_window_attention;
    debug_.window_attention_simt = dbg_window_attention_simt;
    debug_.layernorm_stats = dbg_layernorm_stats;
    debug_.mlp_fused = dbg_mlp_fused;
    debug_.conv3x3 = dbg_conv3x3;
    debug_.local_conv = dbg_local_conv;
    debug_.image_pool_stats = dbg_image_pool_stats;
    return &debug_;
}

}  // extern "C"
