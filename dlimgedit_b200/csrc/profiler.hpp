// profiler.hpp -- optional per-kernel-category CUDA-event timing (off by default; bench.py switches it on
// for a separate pass to attribute step time to kernels and to compute the roofline of the dominant one).
#pragma once

#include "common.hpp"

#include <mutex>
#include <vector>

namespace dlimg {

enum KernelCat : int {
    CAT_GEMM_BF16 = 0, CAT_GEMM_TF32, CAT_CONV1, CAT_IM2COL, CAT_DWCONV, CAT_LAYERNORM, CAT_WIN_ATTN, CAT_RESIZE,
    CAT_IMAGE_TENSOR, CAT_MASK_POST, CAT_DEC_LINEAR, CAT_DEC_ATTN, CAT_DEC_NORM, CAT_DEC_MISC, CAT_OTHER, CAT_COUNT
};

char const* kernel_cat_name(int cat);

class Profiler {
  public:
    // The profiler of the environment bound to the calling thread (EnvironmentImpl::Scope), else a process-wide one.
    static Profiler& get();
    static Profiler* bind(Profiler* p);  // returns the previous binding of this thread
    ~Profiler();
    bool enabled() const { return enabled_; }
    void enable(bool on);
    void begin(cudaStream_t s, int cat, double flops, double bytes);
    void end(cudaStream_t s);
    struct Total { uint64_t launches = 0; double ms = 0, flops = 0, bytes = 0; };
    // Synchronises the recorded events and returns (and clears) the per-category totals.
    std::vector<Total> collect();

  private:
    struct Rec { int cat; cudaEvent_t a, b; double flops, bytes; };
    bool enabled_ = false;
    std::mutex mutex_;
    std::vector<Rec> recs_;
    std::vector<cudaEvent_t> pool_;
    cudaEvent_t take_event();
};

// RAII scope placed at the top of every kernel-launch wrapper.
struct ProfScope {
    cudaStream_t s;
    bool on;
    ProfScope(cudaStream_t stream, int cat, double flops = 0, double bytes = 0) : s(stream), on(Profiler::get().enabled()) {
        if (on) Profiler::get().begin(s, cat, flops, bytes);
    }
    ~ProfScope() {
        if (on) Profiler::get().end(s);
    }
};

}  // namespace dlimg
