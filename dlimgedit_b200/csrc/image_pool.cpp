// image_pool.cpp -- see image_pool.hpp.
#include "image_pool.hpp"

#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <unordered_map>

namespace dlimg {

namespace {

constexpr size_t kMinPinned = (size_t)64 << 10;    // smaller buffers are not worth a page-locked block
constexpr size_t kCacheCap = (size_t)512 << 20;    // released blocks kept for re-use
constexpr size_t kInUseCap = (size_t)8 << 30;      // beyond this much page-locked memory in use, fall back to new[]

struct Pool {
    std::mutex mutex;
    int environments = 0;
    int device = 0;
    std::unordered_map<void const*, size_t> in_use;  // page-locked blocks handed out -> rounded size
    std::multimap<size_t, void*> cached;             // released blocks by rounded size
    ImagePoolStats stats;
};

Pool& pool() {
    static Pool* const p = new Pool;  // never destroyed: destroy_image may run during static destruction of the host program
    return *p;
}

// Block size of a request: whole 64 KiB units and always at least one byte more than asked for, so that two blocks that
// happen to be neighbours in the address space never look like ONE contiguous run of images to the engine (it merges
// the copies of buffers that follow each other exactly, and a copy must not span two page-locked allocations).
size_t rounded(size_t bytes) { return (bytes / kMinPinned + 1) * kMinPinned; }

// cudaHostAlloc / cudaFreeHost initialise the CALLING thread's current device; the caller of create_image may never have
// touched CUDA, so the calls run with the environment's device current and leave the thread as they found it.
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != device) cudaSetDevice(device);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

void release_cached(Pool& p) {
    if (p.cached.empty()) return;
    DeviceGuard guard(p.device);
    for (auto& kv : p.cached) cudaFreeHost(kv.second);
    p.cached.clear();
    p.stats.pinned_cached = 0;
}

}  // namespace

uint8_t* image_alloc(size_t bytes) {
    Pool& p = pool();
    if (bytes >= kMinPinned) {
        std::lock_guard<std::mutex> lock(p.mutex);
        size_t const size = rounded(bytes);
        if (p.environments > 0 && p.stats.pinned_in_use + size <= kInUseCap) {
            void* block = nullptr;
            auto it = p.cached.find(size);
            if (it != p.cached.end()) {
                block = it->second;
                p.cached.erase(it);
                p.stats.pinned_cached -= size;
                ++p.stats.reuses;
            } else {
                DeviceGuard guard(p.device);
                if (cudaHostAlloc(&block, size, cudaHostAllocPortable) != cudaSuccess) {
                    cudaGetLastError();  // (not sticky) -- the buffer below is pageable, the copies still work
                    block = nullptr;
                } else {
                    ++p.stats.pinned_allocs;
                }
            }
            if (block) {
                p.in_use.emplace(block, size);
                p.stats.pinned_in_use += size;
                return static_cast<uint8_t*>(block);
            }
        }
    }
    {
        std::lock_guard<std::mutex> lock(p.mutex);
        ++p.stats.plain_allocs;
    }
    return new uint8_t[bytes];
}

void image_free(uint8_t const* pixels) {
    if (!pixels) return;
    Pool& p = pool();
    {
        std::lock_guard<std::mutex> lock(p.mutex);
        auto it = p.in_use.find(pixels);
        if (it != p.in_use.end()) {
            size_t const size = it->second;
            p.in_use.erase(it);
            p.stats.pinned_in_use -= size;
            void* block = const_cast<uint8_t*>(pixels);
            if (p.environments > 0 && p.stats.pinned_cached + size <= kCacheCap) {
                p.cached.emplace(size, block);
                p.stats.pinned_cached += size;
            } else {
                DeviceGuard guard(p.device);
                cudaFreeHost(block);
            }
            return;
        }
    }
    delete[] pixels;
}

void image_pool_attach(int device) {
    Pool& p = pool();
    std::lock_guard<std::mutex> lock(p.mutex);
    if (p.environments++ == 0) p.device = device;
}

void image_pool_detach() {
    Pool& p = pool();
    std::lock_guard<std::mutex> lock(p.mutex);
    if (--p.environments == 0) release_cached(p);
}

ImagePoolStats image_pool_stats() {
    Pool& p = pool();
    std::lock_guard<std::mutex> lock(p.mutex);
    return p.stats;
}

}  // namespace dlimg
