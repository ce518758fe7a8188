// weights.cpp -- see weights.hpp.
#include "weights.hpp"

#include <cstdint>
#include <cstring>
#include <fstream>

namespace dlimg {

namespace {
template <typename T> T read_pod(std::ifstream& f) {
    T v{};
    f.read(reinterpret_cast<char*>(&v), sizeof(T));
    if (!f) fail("Unexpected end of weight file");
    return v;
}
} // namespace

WeightFile WeightFile::load(std::string const& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) fail("Could not open model file " + path);
    char magic[8];
    f.read(magic, 8);
    if (!f || std::memcmp(magic, "DLIMGB2\0", 8) != 0) fail("Not a DLIMGB2 weight container: " + path);
    uint32_t const version = read_pod<uint32_t>(f);
    if (version != 1) fail("Unsupported weight container version " + std::to_string(version));
    uint32_t const count = read_pod<uint32_t>(f);
    // every size in the header is checked against the size of the file before anything is allocated for it
    f.seekg(0, std::ios::end);
    uint64_t const file_size = (uint64_t)f.tellg();
    f.seekg(16, std::ios::beg);
    constexpr uint64_t kMinRecord = 2 + 1 + 8 + 8;  // name length, rank, offset, element count
    if ((uint64_t)count * kMinRecord > file_size) fail("Corrupt weight container: tensor count exceeds the file: " + path);
    struct Rec { std::string name; std::vector<int64_t> shape; uint64_t offset, numel; };
    std::vector<Rec> recs(count);
    for (auto& r : recs) {
        uint16_t const len = read_pod<uint16_t>(f);
        if ((uint64_t)f.tellg() + len > file_size) fail("Corrupt weight container: truncated tensor name");
        r.name.resize(len);
        f.read(&r.name[0], len);
        if (!f) fail("Unexpected end of weight file");
        uint8_t const ndim = read_pod<uint8_t>(f);
        if (ndim > 8) fail("Corrupt weight container: rank " + std::to_string(ndim) + " of " + r.name);
        r.shape.resize(ndim);
        for (auto& d : r.shape) d = read_pod<uint32_t>(f);
        r.offset = read_pod<uint64_t>(f);
        r.numel = read_pod<uint64_t>(f);
        uint64_t n = 1;
        bool overflow = false;
        for (auto d : r.shape) {
            if (d != 0 && n > UINT64_MAX / (uint64_t)d) overflow = true;
            n *= (uint64_t)d;
        }
        if (overflow || n != r.numel) fail("Corrupt weight container: shape/numel mismatch for " + r.name);
        if (r.numel > file_size / sizeof(float) || r.offset > file_size || r.offset + r.numel * sizeof(float) > file_size)
            fail("Corrupt weight container: payload of " + r.name + " lies outside the file");
    }
    std::streamoff const payload = f.tellg();
    WeightFile wf;
    wf.path_ = path;
    for (auto const& r : recs) {
        HostTensor t;
        t.shape = r.shape;
        t.data.resize(r.numel);
        if ((uint64_t)payload + r.offset + r.numel * sizeof(float) > file_size) fail("Corrupt weight container: payload truncated at " + r.name);
        f.seekg(payload + (std::streamoff)r.offset);
        f.read(reinterpret_cast<char*>(t.data.data()), (std::streamsize)(r.numel * sizeof(float)));
        if (!f) fail("Corrupt weight container: payload truncated at " + r.name);
        wf.tensors_.emplace(r.name, std::move(t));
    }
    return wf;
}

HostTensor const& WeightFile::get(std::string const& name) const {
    auto it = tensors_.find(name);
    if (it == tensors_.end()) fail("Tensor '" + name + "' not found in " + path_);
    return it->second;
}

HostTensor const& WeightFile::get(std::string const& name, std::vector<int64_t> const& shape) const {
    HostTensor const& t = get(name);
    if (t.shape != shape) {
        std::string s = "Tensor '" + name + "' has shape (";
        for (auto d : t.shape) s += std::to_string(d) + ",";
        s += "), expected (";
        for (auto d : shape) s += std::to_string(d) + ",";
        fail(s + ") in " + path_);
    }
    return t;
}

} // namespace dlimg
