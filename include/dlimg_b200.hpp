// dlimg_b200.hpp -- C++14 wrappers over the additive extension table (dlimg_b200.h part 2), written to sit next to
// the reference's header-only façade: same conventions as src/include/dlimgedit/detail/handle.hpp:27-34 (a lazily
// resolved function table) and dlimgedit.impl.hpp:7-11 / 70-103 (errors become dlimg::Exception, RAII handles).
//
//     #include <dlimgedit/dlimgedit.hpp>   // the reference's own, unchanged
//     #include <dlimg_b200.hpp>
//
//     auto segs = dlimg::b200::process_batch(env, views.data(), (int)views.size());          // one encoder pass
//     dlimg::b200::compute_masks_batch(env, owners.data(), prompts.data(), n, false, masks.data(), ious.data());
//
// With DLIMGEDIT_LOAD_DYNAMIC, resolve "dlimg_b200_ext_init" next to "dlimg_init" and pass its result to
// dlimg::b200::initialize().
#pragma once

#include <dlimgedit/dlimgedit.hpp>

#include "dlimg_b200.h"

#include <vector>

namespace dlimg {
namespace b200 {

namespace detail {
template <typename T> struct Global {
    static dlimg_b200_Ext const* ext_;
};
template <typename T> dlimg_b200_Ext const* Global<T>::ext_{};

// Segmentation's handle slot is protected (handle.hpp:66); a derived type may fill it, then it is moved into a plain one
struct Adopt : Segmentation {
    Adopt() : Segmentation(nullptr) {}
    dlimg_Segmentation_*& slot() noexcept { return emplace(); }
};
}  // namespace detail

inline void initialize(dlimg_b200_Ext const* ext) { detail::Global<void>::ext_ = ext; }

inline dlimg_b200_Ext const& ext() {
#ifndef DLIMGEDIT_LOAD_DYNAMIC
    if (!detail::Global<void>::ext_) detail::Global<void>::ext_ = dlimg_b200_ext_init();
#endif
    if (!detail::Global<void>::ext_) throw Exception("dlimg_b200 extension table is not initialised");
    return *detail::Global<void>::ext_;
}

inline void throw_on_error(dlimg_Result result) {
    if (result == dlimg_error) throw Exception(api().last_error());
}

// A prompt of the batched decoder: a point, or the box of the largest object to segment (dlimgedit.hpp:119-134).
struct Prompt : dlimg_b200_Prompt {
    Prompt(Point p) : dlimg_b200_Prompt{0, p.x, p.y, 0, 0} {}
    Prompt(Region r) : dlimg_b200_Prompt{1, r.top_left.x, r.top_left.y, r.bottom_right.x, r.bottom_right.y} {}
};

enum class Placement : int {
    host = 0,        // masks / scores in host memory, complete when the call returns (like Segmentation::compute_mask)
    device = 1,      // device pointers; asynchronous on the environment's stream
    host_async = 2,  // host (ideally page-locked) memory, complete after synchronize()
};

inline void set_stream(Environment const& env, void* cuda_stream) { throw_on_error(ext().set_stream(env.handle(), cuda_stream)); }
inline void synchronize(Environment const& env) { throw_on_error(ext().synchronize(env.handle())); }
inline dlimg_b200_Stats stats(Environment const& env) {
    dlimg_b200_Stats s{};
    throw_on_error(ext().get_stats(env.handle(), &s));
    return s;
}

// Encodes `count` images of one extent / channel order in one pass (Segmentation::process for a batch).  `views[i].pixels`
// are host pointers, or device pointers with pixels_on_device.  The handles behave like any other Segmentation.
inline std::vector<Segmentation> process_batch(Environment const& env, ImageView const* views, int count, bool pixels_on_device = false) {
    static_assert(sizeof(ImageView) == sizeof(dlimg_ImageView), "ImageView is layout-compatible with dlimg_ImageView (impl.hpp:26-28)");
    std::vector<dlimg_Segmentation> raw((size_t)count, nullptr);
    dlimg_Result const r = ext().process_batch(env.handle(), reinterpret_cast<dlimg_ImageView const*>(views), count,
                                               pixels_on_device ? 1 : 0, raw.data());
    std::vector<Segmentation> out;
    out.reserve((size_t)count);
    for (int i = 0; i < count; ++i) {  // owned even when the call failed (dlimgedit.cpp:55-57)
        detail::Adopt a;
        a.slot() = raw[(size_t)i];
        out.emplace_back(std::move(a));
    }
    throw_on_error(r);
    return out;
}

// Answers `count` prompts; prompt i refers to *segs[i] (handles may repeat, and may belong to different images).
// masks_out[i]: n * W_i * H_i bytes (n = multi ? 3 : 1), ious_out: count * n floats or null.
inline void compute_masks_batch(Environment const& env, Segmentation const* const* segs, Prompt const* prompts, int count, bool multi,
                                uint8_t* const* masks_out, float* ious_out, Placement where = Placement::host) {
    std::vector<dlimg_Segmentation> raw((size_t)count);
    for (int i = 0; i < count; ++i) raw[(size_t)i] = segs[i]->handle();
    throw_on_error(ext().compute_masks_batch(env.handle(), raw.data(), prompts, count, multi ? 1 : 0, masks_out, ious_out, (int)where));
}

// The (1, 256, 64, 64) float32 image embedding, NCHW like the reference's `image_embeddings` tensor (segmentation.cpp:124).
inline void get_embedding(Segmentation const& seg, float* out_host) { throw_on_error(ext().get_embedding(seg.handle(), out_host)); }
inline void get_embedding_async(Segmentation const& seg, float* out_host) {
    throw_on_error(ext().get_embedding_async(seg.handle(), out_host));
}
// half-precision download (256*64*64 uint16 values): complete after synchronize()
inline void get_embedding_f16_async(Segmentation const& seg, uint16_t* out_host) {
    throw_on_error(ext().get_embedding_f16_async(seg.handle(), out_host));
}

}  // namespace b200
}  // namespace dlimg
