// image_io.hpp -- the load_image / save_image slots of the C ABI (reference image.cpp:11-35, there backed
// by stb).  Off the hot path; CPU code.  Loading covers PNG (every colour type / bit depth, palettes, tRNS, Adam7), JPEG
// (sequential and progressive Huffman), BMP, TGA and binary PGM/PPM (P5/P6) -- each restated from stb_image's published
// behaviour so that pixels come out as the reference sees them; saving writes PNG for mask / rgb / rgba like the
// reference.  Malformed input is an exception, never undefined behaviour: tools/fuzz/run.py feeds the readers mutated
// files under AddressSanitizer + UBSan.
#pragma once

#include "../../include/dlimg_b200.h"
#include "common.hpp"

namespace dlimg {

// Returns pixels from image_alloc (image_pool.hpp; released by dlimg_Api::destroy_image -> image_free).
uint8_t* load_image(char const* filepath, int* out_extent, int* out_channels);
void save_image(dlimg_ImageView const& img, char const* filepath);

} // namespace dlimg
