// weights.hpp -- reader for the engine's flat weight container ("DLIMGB2").
//
// The container stores the MobileSAM checkpoint's state-dict tensors under their original names
// (image_encoder.*, prompt_encoder.*, mask_decoder.*) as raw little-endian fp32,
// so a real `mobile_sam.pt` converts 1:1 (tools/convert_checkpoint.py).  All folding (BatchNorm into
// convolutions), re-layout and bf16 conversion happens at load time in model.cu.
//
// Layout: char magic[8] = "DLIMGB2\0"; u32 version (1); u32 count; then `count` records
//   { u16 name_len; char name[name_len]; u8 ndim; u32 dims[ndim]; u64 offset; u64 numel }
// followed by the fp32 payload; `offset` is in bytes from the start of the payload.
#pragma once

#include "common.hpp"

#include <map>
#include <string>
#include <vector>

namespace dlimg {

struct HostTensor {
    std::vector<int64_t> shape;
    std::vector<float> data;
    int64_t numel() const { return (int64_t)data.size(); }
    int64_t dim(int i) const { return shape.at((size_t)i); }
};

class WeightFile {
  public:
    static WeightFile load(std::string const& path);
    bool has(std::string const& name) const { return tensors_.count(name) != 0; }
    HostTensor const& get(std::string const& name) const;
    // get + shape check
    HostTensor const& get(std::string const& name, std::vector<int64_t> const& shape) const;
    size_t size() const { return tensors_.size(); }

  private:
    std::map<std::string, HostTensor> tensors_;
    std::string path_;
};

} // namespace dlimg
