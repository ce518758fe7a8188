// image_io.cpp -- see image_io.hpp.  Minimal PNG / PNM codec written for this library (no stb).
#include "image_io.hpp"

#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

namespace dlimg {
namespace {

std::vector<uint8_t> read_file(char const* path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) fail(std::string("Failed to load image ") + path + ": cannot open file");
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

uint32_t be32(uint8_t const* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
void put_be32(std::vector<uint8_t>& v, uint32_t x) {
    v.push_back((uint8_t)(x >> 24)); v.push_back((uint8_t)(x >> 16)); v.push_back((uint8_t)(x >> 8)); v.push_back((uint8_t)x);
}

uint32_t crc32(uint8_t const* data, size_t n, uint32_t crc = 0) {
    static uint32_t table[256];
    static bool ready = false;
    if (!ready) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        ready = true;
    }
    crc = ~crc;
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ data[i]) & 0xFF] ^ (crc >> 8);
    return ~crc;
}

uint32_t adler32(uint8_t const* data, size_t n) {
    uint32_t a = 1, b = 0;
    for (size_t i = 0; i < n; ++i) {
        a = (a + data[i]) % 65521;
        b = (b + a) % 65521;
    }
    return (b << 16) | a;
}

// ---- inflate (RFC 1951) ---------------------------------------------------------------------------
struct BitReader {
    uint8_t const* p;
    size_t n, pos = 0;
    uint32_t buf = 0;
    int cnt = 0;
    uint32_t bits(int k) {
        while (cnt < k) {
            if (pos >= n) fail("PNG: truncated deflate stream");
            buf |= (uint32_t)p[pos++] << cnt;
            cnt += 8;
        }
        uint32_t const v = buf & ((k == 32) ? 0xFFFFFFFFu : ((1u << k) - 1));
        buf >>= k;
        cnt -= k;
        return v;
    }
    void align() { buf = 0; cnt = 0; }
};

struct Huffman {
    uint16_t count[16] = {0};
    uint16_t symbol[288] = {0};
    void build(uint8_t const* lengths, int n) {
        std::memset(count, 0, sizeof(count));
        for (int i = 0; i < n; ++i) count[lengths[i]]++;
        count[0] = 0;
        uint16_t offs[16];
        offs[1] = 0;
        for (int i = 1; i < 15; ++i) offs[i + 1] = offs[i] + count[i];
        for (int i = 0; i < n; ++i)
            if (lengths[i]) symbol[offs[lengths[i]]++] = (uint16_t)i;
    }
    int decode(BitReader& br) const {
        int code = 0, first = 0, index = 0;
        for (int len = 1; len <= 15; ++len) {
            code |= (int)br.bits(1);
            int const c = count[len];
            if (code - c < first) return symbol[index + (code - first)];
            index += c;
            first += c;
            first <<= 1;
            code <<= 1;
        }
        fail("PNG: invalid Huffman code");
    }
};

std::vector<uint8_t> inflate(uint8_t const* data, size_t n) {
    static uint16_t const lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static uint16_t const lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static uint16_t const dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static uint16_t const dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    static uint8_t const order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    if (n < 2) fail("PNG: zlib stream too short");
    BitReader br{data + 2, n - 2};
    std::vector<uint8_t> out;
    bool last = false;
    while (!last) {
        last = br.bits(1) != 0;
        uint32_t const type = br.bits(2);
        if (type == 0) {
            br.align();
            if (br.pos + 4 > br.n) fail("PNG: truncated stored block");
            uint32_t const len = br.p[br.pos] | (br.p[br.pos + 1] << 8);
            br.pos += 4;
            if (br.pos + len > br.n) fail("PNG: truncated stored block");
            out.insert(out.end(), br.p + br.pos, br.p + br.pos + len);
            br.pos += len;
        } else if (type == 1 || type == 2) {
            Huffman lit, dist;
            uint8_t lengths[320];
            if (type == 1) {
                int i = 0;
                for (; i < 144; ++i) lengths[i] = 8;
                for (; i < 256; ++i) lengths[i] = 9;
                for (; i < 280; ++i) lengths[i] = 7;
                for (; i < 288; ++i) lengths[i] = 8;
                lit.build(lengths, 288);
                for (i = 0; i < 30; ++i) lengths[i] = 5;
                dist.build(lengths, 30);
            } else {
                int const nlen = (int)br.bits(5) + 257, ndist = (int)br.bits(5) + 1, ncode = (int)br.bits(4) + 4;
                uint8_t cl[19] = {0};
                for (int i = 0; i < ncode; ++i) cl[order[i]] = (uint8_t)br.bits(3);
                Huffman clh;
                clh.build(cl, 19);
                int i = 0;
                while (i < nlen + ndist) {
                    int const sym = clh.decode(br);
                    if (sym < 16) lengths[i++] = (uint8_t)sym;
                    else {
                        int rep = 0;
                        uint8_t val = 0;
                        if (sym == 16) {
                            if (i == 0) fail("PNG: invalid code length repeat");
                            val = lengths[i - 1];
                            rep = 3 + (int)br.bits(2);
                        } else if (sym == 17) rep = 3 + (int)br.bits(3);
                        else rep = 11 + (int)br.bits(7);
                        if (i + rep > nlen + ndist) fail("PNG: code length overflow");
                        while (rep--) lengths[i++] = val;
                    }
                }
                lit.build(lengths, nlen);
                dist.build(lengths + nlen, ndist);
            }
            for (;;) {
                int sym = lit.decode(br);
                if (sym < 256) out.push_back((uint8_t)sym);
                else if (sym == 256) break;
                else {
                    sym -= 257;
                    if (sym >= 29) fail("PNG: invalid length symbol");
                    int const len = lbase[sym] + (int)br.bits(lext[sym]);
                    int const ds = dist.decode(br);
                    if (ds >= 30) fail("PNG: invalid distance symbol");
                    size_t const d = dbase[ds] + br.bits(dext[ds]);
                    if (d > out.size()) fail("PNG: distance too far back");
                    size_t const start = out.size() - d;
                    for (int k = 0; k < len; ++k) out.push_back(out[start + (size_t)k]);
                }
            }
        } else {
            fail("PNG: invalid deflate block type");
        }
    }
    return out;
}

int paeth(int a, int b, int c) {
    int const p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

uint8_t* load_png(std::vector<uint8_t> const& file, char const* path, int* extent, int* channels) {
    size_t pos = 8;
    uint32_t w = 0, h = 0;
    int depth = 0, color = 0, interlace = 0;
    std::vector<uint8_t> idat;
    while (pos + 12 <= file.size()) {
        uint32_t const len = be32(&file[pos]);
        char const* type = reinterpret_cast<char const*>(&file[pos + 4]);
        if (pos + 12 + len > file.size()) fail(std::string("Failed to load image ") + path + ": truncated PNG");
        uint8_t const* d = &file[pos + 8];
        if (!std::memcmp(type, "IHDR", 4)) {
            w = be32(d); h = be32(d + 4); depth = d[8]; color = d[9]; interlace = d[12];
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), d, d + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + len;
    }
    int const ch = color == 0 ? 1 : color == 2 ? 3 : color == 6 ? 4 : 0;
    if (!w || !h || depth != 8 || !ch || interlace)
        fail(std::string("Failed to load image ") + path + ": only 8-bit grey/RGB/RGBA non-interlaced PNG is supported");
    std::vector<uint8_t> raw = inflate(idat.data(), idat.size());
    size_t const row = (size_t)w * ch;
    if (raw.size() < (row + 1) * h) fail(std::string("Failed to load image ") + path + ": PNG data too short");
    uint8_t* px = new uint8_t[row * h];
    for (uint32_t y = 0; y < h; ++y) {
        uint8_t const* src = &raw[(row + 1) * y];
        uint8_t* dst = px + row * y;
        uint8_t const* up = y ? dst - row : nullptr;
        int const ft = src[0];
        ++src;
        for (size_t x = 0; x < row; ++x) {
            int const a = x >= (size_t)ch ? dst[x - ch] : 0, b = up ? up[x] : 0, c = (up && x >= (size_t)ch) ? up[x - ch] : 0;
            int v = src[x];
            switch (ft) {
            case 0: break;
            case 1: v += a; break;
            case 2: v += b; break;
            case 3: v += (a + b) >> 1; break;
            case 4: v += paeth(a, b, c); break;
            default: delete[] px; fail(std::string("Failed to load image ") + path + ": bad PNG filter");
            }
            dst[x] = (uint8_t)v;
        }
    }
    extent[0] = (int)w;
    extent[1] = (int)h;
    *channels = ch;
    return px;
}

uint8_t* load_pnm(std::vector<uint8_t> const& file, char const* path, int* extent, int* channels) {
    size_t pos = 2;
    auto next_int = [&]() -> int {
        for (;;) {
            while (pos < file.size() && std::isspace(file[pos])) ++pos;
            if (pos < file.size() && file[pos] == '#') { while (pos < file.size() && file[pos] != '\n') ++pos; continue; }
            break;
        }
        int v = 0;
        bool any = false;
        while (pos < file.size() && file[pos] >= '0' && file[pos] <= '9') { v = v * 10 + (file[pos++] - '0'); any = true; }
        if (!any) fail(std::string("Failed to load image ") + path + ": malformed PNM header");
        return v;
    };
    int const ch = file[1] == '5' ? 1 : 3;
    int const w = next_int(), h = next_int(), maxv = next_int();
    ++pos;  // single whitespace after maxval
    if (maxv != 255 || w <= 0 || h <= 0 || pos + (size_t)w * h * ch > file.size())
        fail(std::string("Failed to load image ") + path + ": unsupported PNM variant");
    uint8_t* px = new uint8_t[(size_t)w * h * ch];
    std::memcpy(px, &file[pos], (size_t)w * h * ch);
    extent[0] = w;
    extent[1] = h;
    *channels = ch;
    return px;
}

void chunk(std::vector<uint8_t>& out, char const* type, std::vector<uint8_t> const& data) {
    put_be32(out, (uint32_t)data.size());
    size_t const start = out.size();
    out.insert(out.end(), type, type + 4);
    out.insert(out.end(), data.begin(), data.end());
    put_be32(out, crc32(&out[start], out.size() - start));
}

}  // namespace

uint8_t* load_image(char const* filepath, int* out_extent, int* out_channels) {
    std::vector<uint8_t> const file = read_file(filepath);
    static uint8_t const png_sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (file.size() > 8 && !std::memcmp(file.data(), png_sig, 8)) return load_png(file, filepath, out_extent, out_channels);
    if (file.size() > 2 && file[0] == 'P' && (file[1] == '5' || file[1] == '6')) return load_pnm(file, filepath, out_extent, out_channels);
    fail(std::string("Failed to load image ") + filepath + ": unsupported format (PNG and binary PGM/PPM are supported)");
}

// reference image.cpp:25-35: mask / rgb / rgba only, rows written packed (the reference also ignores stride)
void save_image(dlimg_ImageView const& img, char const* filepath) {
    if (!(img.channels == CH_MASK || img.channels == CH_RGB || img.channels == CH_RGBA))
        fail("Unsupported channel order [" + std::to_string(img.channels) + "]");
    int const comp = bytes_per_pixel(img.channels);
    size_t const row = (size_t)img.width * comp;
    std::vector<uint8_t> raw;
    raw.reserve((row + 1) * img.height);
    for (int y = 0; y < img.height; ++y) {
        raw.push_back(0);  // filter type none
        uint8_t const* src = img.pixels + (size_t)y * row;
        raw.insert(raw.end(), src, src + row);
    }
    std::vector<uint8_t> z;
    z.push_back(0x78);
    z.push_back(0x01);
    size_t pos = 0;
    do {  // stored deflate blocks (no compression: the mask path is not size sensitive)
        size_t const n = std::min<size_t>(65535, raw.size() - pos);
        z.push_back(pos + n == raw.size() ? 1 : 0);
        z.push_back((uint8_t)(n & 0xFF)); z.push_back((uint8_t)(n >> 8));
        z.push_back((uint8_t)(~n & 0xFF)); z.push_back((uint8_t)((~n >> 8) & 0xFF));
        z.insert(z.end(), raw.begin() + (long)pos, raw.begin() + (long)(pos + n));
        pos += n;
    } while (pos < raw.size());
    put_be32(z, adler32(raw.data(), raw.size()));

    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    std::vector<uint8_t> ihdr;
    put_be32(ihdr, (uint32_t)img.width);
    put_be32(ihdr, (uint32_t)img.height);
    ihdr.push_back(8);
    ihdr.push_back(comp == 1 ? 0 : comp == 3 ? 2 : 6);
    ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    chunk(out, "IHDR", ihdr);
    chunk(out, "IDAT", z);
    chunk(out, "IEND", {});
    FILE* f = std::fopen(filepath, "wb");
    if (!f) fail(std::string("Failed to save image ") + filepath);
    size_t const written = std::fwrite(out.data(), 1, out.size(), f);
    std::fclose(f);
    if (written != out.size()) fail(std::string("Failed to save image ") + filepath);
}

}  // namespace dlimg
