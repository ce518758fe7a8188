// image_pool.hpp -- storage behind the façade's pixel buffers (dlimg_Api::create_image / load_image / destroy_image).
//
// The reference's C++ wrapper takes every Image it owns from these slots (dlimg::Image::Image -> create_image,
// Image::load -> load_image; dlimgedit.impl.hpp:44-66) and hands exactly those buffers to process / compute_mask.  While a
// GPU environment is alive they are page-locked, so the engine's cudaMemcpyAsync calls run at PCIe speed and truly
// asynchronously instead of through the driver's pageable staging path; released blocks are cached by size because
// pinning memory costs milliseconds.  Without an environment (or when pinning fails) the buffers are plain new[] memory.
#pragma once

#include <cstddef>
#include <cstdint>

namespace dlimg {

uint8_t* image_alloc(size_t bytes);
void image_free(uint8_t const* pixels);

// A GPU environment came up on `device` / went away.  The last one to go releases the cached blocks.
void image_pool_attach(int device);
void image_pool_detach();

struct ImagePoolStats {
    size_t pinned_in_use = 0, pinned_cached = 0;  // bytes
    size_t pinned_allocs = 0, reuses = 0, plain_allocs = 0;
};
ImagePoolStats image_pool_stats();

}  // namespace dlimg
