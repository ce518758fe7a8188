// image_io.hpp -- the load_image / save_image slots of the C ABI (reference image.cpp:11-35, there backed
// by stb).  Off the hot path; CPU code.  Loading covers binary PGM/PPM (P5/P6) and PNG (8-bit
// grey / RGB / RGBA, non-interlaced); saving writes PNG for mask / rgb / rgba like the reference.
#pragma once

#include "../../include/dlimg_b200.h"
#include "common.hpp"

namespace dlimg {

// Returns pixels from image_alloc (image_pool.hpp; released by dlimg_Api::destroy_image -> image_free).
uint8_t* load_image(char const* filepath, int* out_extent, int* out_channels);
void save_image(dlimg_ImageView const& img, char const* filepath);

} // namespace dlimg
