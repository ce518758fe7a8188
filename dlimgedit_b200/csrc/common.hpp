// common.hpp -- error handling, small utilities shared by host and device translation units.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <utility>

namespace dlimg {

// Mirrors dlimg::Exception of the reference façade (dlimgedit.hpp:184-191): everything thrown inside
// the library derives from std::exception and is mapped to dlimg_error at the C boundary.
class Error : public std::runtime_error {
  public:
    explicit Error(std::string const& msg) : std::runtime_error(msg) {}
};

[[noreturn]] inline void fail(std::string const& msg) { throw Error(msg); }

// Development switches.  The library keeps a few alternative code paths to cross-check a fused kernel against its unfused
// form or to A/B a design (DLIMG_B200_UNFUSED_PATCH, DLIMG_B200_T2I_SIMT, ...).  In a release build they are compiled out:
// dev_switch() is a constant false, so the branches fold away, and the kernels that only they reach sit under
// `#if DLIMG_B200_ALT`.  `python -m dlimgedit_b200._build --dev` (-DDLIMG_B200_DEV) and the bf16 build (which runs the
// unfused forms) compile them in.
#if defined(DLIMG_B200_DEV) || defined(DLIMG_B200_ACT_BF16)
#define DLIMG_B200_ALT 1
inline bool dev_switch(char const* name) { return std::getenv(name) != nullptr; }
inline int dev_int(char const* name, int def) {
    char const* v = std::getenv(name);
    return v ? std::atoi(v) : def;
}
#else
#define DLIMG_B200_ALT 0
constexpr bool dev_switch(char const*) { return false; }
constexpr int dev_int(char const*, int def) { return def; }
#endif

// Same contract as the reference's ASSERT (assert.hpp:17-28): report on stderr and throw.
#define DLIMG_ASSERT(cond)                                                                           \
    do {                                                                                             \
        if (!(cond)) {                                                                               \
            std::fprintf(stderr, "Assertion failed at %s:%d: %s\n", __FILE__, __LINE__, #cond);      \
            throw ::dlimg::Error(std::string("Assertion failed: ") + #cond);                         \
        }                                                                                            \
    } while (0)

#define CUDA_CHECK(expr)                                                                             \
    do {                                                                                             \
        cudaError_t err__ = (expr);                                                                  \
        if (err__ != cudaSuccess) {                                                                  \
            throw ::dlimg::Error(std::string("CUDA error ") + cudaGetErrorName(err__) + " (" +       \
                                 cudaGetErrorString(err__) + ") at " + __FILE__ + ":" +              \
                                 std::to_string(__LINE__) + ": " #expr);                             \
        }                                                                                            \
    } while (0)

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t round_up64(int64_t a, int64_t b) { return ceil_div64(a, b) * b; }

// Channels enum values carried in dlimg_ImageView::channels (reference dlimgedit.hpp:29).
enum : int { CH_MASK = 1, CH_RGB = 3, CH_RGBA = 4, CH_BGRA = 5, CH_ARGB = 6 };
inline int bytes_per_pixel(int channels) { return channels > 4 ? 4 : channels; } // impl.hpp:15
inline bool valid_channels(int c) { return c == 1 || c == 3 || c == 4 || c == 5 || c == 6; }

// Source byte offsets of R, G, B inside a pixel (reference segmentation.cpp:82-95).
inline void channel_map(int channels, int cmap[3]) {
    cmap[0] = 0; cmap[1] = 1; cmap[2] = 2;
    if (channels == CH_MASK) { cmap[0] = cmap[1] = cmap[2] = 0; }
    else if (channels == CH_BGRA) { cmap[0] = 2; cmap[1] = 1; cmap[2] = 0; }
    else if (channels == CH_ARGB) { cmap[0] = 1; cmap[1] = 2; cmap[2] = 3; }
}

constexpr int kImageSize = 1024;   // reference segmentation.cpp:17
constexpr int kEmbedDim = 256;     // image embedding channels
constexpr int kEmbedRes = 64;      // image embedding spatial size
constexpr int kLowRes = 256;       // low-resolution mask size

} // namespace dlimg

#include <atomic>
namespace dlimg {
// Kernels launched and bytes copied by this library, per environment (reported through dlimg_b200_Ext::get_stats).
// The engine binds the environment a call belongs to to the calling thread (EnvironmentImpl::Scope); launches outside
// any environment (model load, the test-only debug table) go to a process-wide fallback.
struct EnvCounters {
    std::atomic<uint64_t> kernel_launches{0}, h2d_bytes{0}, d2h_bytes{0};
};
extern EnvCounters g_unbound_counters;
extern thread_local EnvCounters* tl_counters;
inline EnvCounters& counters() { return tl_counters ? *tl_counters : g_unbound_counters; }
inline void count_launch(uint64_t n = 1) { counters().kernel_launches.fetch_add(n, std::memory_order_relaxed); }
// Launch-configuration errors surface immediately; execution errors at the next synchronising call.
#define KERNEL_CHECK()                                                                               \
    do {                                                                                             \
        ::dlimg::count_launch();                                                                     \
        CUDA_CHECK(cudaGetLastError());                                                              \
    } while (0)

// Programmatic dependent launch: a kernel launched through launch_pdl() may start (prologue: barrier / TMEM setup,
// constant staging) while the previous kernel in the stream is still draining; it must execute pdl_wait() before it
// touches anything the previous kernel reads or writes, and calls pdl_trigger() to let its own successor do the same.
// Works in stream capture (the graph gets programmatic edges).  DLIMG_B200_PDL_MASK (bit per kernel family below)
// selects which kernels are launched that way (A/B switch; 0 = none).
enum PdlFamily : int { PDL_GEMM = 0, PDL_ATTENTION, PDL_LOCAL_CONV, PDL_MBCONV, PDL_PATCH_EMBED, PDL_DWCONV, PDL_ROWS };
bool pdl_enabled(int family);
template <typename... KArgs, typename... Args>
inline void launch_pdl(int family, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled(family) ? 1 : 0;
    CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...));
}
#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
} // namespace dlimg
