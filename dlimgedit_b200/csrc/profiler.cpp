// profiler.cpp -- see profiler.hpp.
#include "profiler.hpp"

namespace dlimg {

char const* kernel_cat_name(int cat) {
    static char const* names[CAT_COUNT] = {"gemm_tcgen05_f16", "gemm_tcgen05_tf32", "conv1_preprocess", "im2col3x3",
                                           "dwconv3x3", "layernorm_rows", "window_attention", "resize_srgb",
                                           "image_tensor", "mask_postprocess", "dec_linear_small", "dec_attention",
                                           "dec_layernorm", "dec_misc", "other"};
    return (cat >= 0 && cat < CAT_COUNT) ? names[cat] : "?";
}

namespace {
thread_local Profiler* tl_profiler = nullptr;
}

Profiler& Profiler::get() {
    if (tl_profiler) return *tl_profiler;
    static Profiler p;
    return p;
}

Profiler* Profiler::bind(Profiler* p) {
    Profiler* const prev = tl_profiler;
    tl_profiler = p;
    return prev;
}

Profiler::~Profiler() {
    for (auto& r : recs_) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    for (auto e : pool_) cudaEventDestroy(e);
}

void Profiler::enable(bool on) {
    std::lock_guard<std::mutex> lock(mutex_);
    enabled_ = on;
}

cudaEvent_t Profiler::take_event() {
    if (!pool_.empty()) {
        cudaEvent_t e = pool_.back();
        pool_.pop_back();
        return e;
    }
    cudaEvent_t e;
    CUDA_CHECK(cudaEventCreate(&e));
    return e;
}

void Profiler::begin(cudaStream_t s, int cat, double flops, double bytes) {
    std::lock_guard<std::mutex> lock(mutex_);
    Rec r{cat, take_event(), take_event(), flops, bytes};
    CUDA_CHECK(cudaEventRecord(r.a, s));
    recs_.push_back(r);
}

void Profiler::end(cudaStream_t s) {
    std::lock_guard<std::mutex> lock(mutex_);
    if (!recs_.empty()) cudaEventRecord(recs_.back().b, s);
}

std::vector<Profiler::Total> Profiler::collect() {
    std::lock_guard<std::mutex> lock(mutex_);
    std::vector<Total> totals((size_t)CAT_COUNT);
    for (auto& r : recs_) {
        CUDA_CHECK(cudaEventSynchronize(r.b));
        float ms = 0;
        CUDA_CHECK(cudaEventElapsedTime(&ms, r.a, r.b));
        Total& t = totals[(size_t)r.cat];
        t.launches += 1;
        t.ms += ms;
        t.flops += r.flops;
        t.bytes += r.bytes;
        pool_.push_back(r.a);
        pool_.push_back(r.b);
    }
    recs_.clear();
    return totals;
}

}  // namespace dlimg
