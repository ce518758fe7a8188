// image_io.cpp -- see image_io.hpp.  Minimal PNG / baseline-JPEG / PNM codec written for this library (no stb).
#include "image_io.hpp"
#include "image_pool.hpp"

#include <algorithm>
#include <cstdint>
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

constexpr uint32_t kMaxImageDim = 1u << 24;          // like stb_image's STBI_MAX_DIMENSIONS
constexpr uint64_t kMaxImageBytes = (uint64_t)1 << 31;  // decoded pixels of one image

// `limit`: the caller knows how many bytes the stream may expand to; anything beyond is a malformed (or hostile) file
std::vector<uint8_t> inflate(uint8_t const* data, size_t n, size_t limit) {
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
            if (out.size() + len > limit) fail("PNG: deflate stream larger than the image it encodes");
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
                if (sym < 256) {
                    if (out.size() >= limit) fail("PNG: deflate stream larger than the image it encodes");
                    out.push_back((uint8_t)sym);
                }
                else if (sym == 256) break;
                else {
                    sym -= 257;
                    if (sym >= 29) fail("PNG: invalid length symbol");
                    int const len = lbase[sym] + (int)br.bits(lext[sym]);
                    int const ds = dist.decode(br);
                    if (ds >= 30) fail("PNG: invalid distance symbol");
                    size_t const d = dbase[ds] + br.bits(dext[ds]);
                    if (d > out.size()) fail("PNG: distance too far back");
                    if (out.size() + (size_t)len > limit) fail("PNG: deflate stream larger than the image it encodes");
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

// PNG as stb_image's 8-bit loader sees it (the reference's stbi_load, image.cpp:11-23): every colour type and bit depth,
// palettes, tRNS transparency (an extra alpha channel), Adam7 interlacing; 16-bit samples keep their high byte, 1 / 2 / 4-bit
// grey is scaled to 0..255.  Channels out: grey 1 (2 with alpha or tRNS), RGB / palette 3 (4 with alpha or tRNS).
uint8_t* load_png(std::vector<uint8_t> const& file, char const* path, int* extent, int* channels) {
    auto bad = [&](char const* why) { fail(std::string("Failed to load image ") + path + ": " + why); };
    size_t pos = 8;
    uint32_t w = 0, h = 0;
    int depth = 0, color = 0, interlace = 0;
    bool have_ihdr = false, have_trns = false;
    uint8_t palette[256][4];
    int pal_len = 0;
    uint16_t key[3] = {0, 0, 0};  // tRNS colour key of grey / RGB images (file precision)
    std::vector<uint8_t> idat;
    while (pos + 12 <= file.size()) {
        uint32_t const len = be32(&file[pos]);
        char const* type = reinterpret_cast<char const*>(&file[pos + 4]);
        if (len > file.size() || pos + 12 + len > file.size()) bad("truncated PNG");
        uint8_t const* d = &file[pos + 8];
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len != 13 || have_ihdr) bad("malformed PNG header");
            have_ihdr = true;
            w = be32(d); h = be32(d + 4); depth = d[8]; color = d[9]; interlace = d[12];
            if (d[10] != 0 || d[11] != 0 || interlace > 1) bad("malformed PNG header");
        } else if (!std::memcmp(type, "PLTE", 4)) {
            if (!have_ihdr || len > 256 * 3 || len % 3) bad("bad PNG palette");
            pal_len = (int)(len / 3);
            for (int i = 0; i < pal_len; ++i) {
                palette[i][0] = d[3 * i]; palette[i][1] = d[3 * i + 1]; palette[i][2] = d[3 * i + 2]; palette[i][3] = 255;
            }
        } else if (!std::memcmp(type, "tRNS", 4)) {
            if (!have_ihdr || !idat.empty()) bad("misplaced tRNS chunk");
            if (color == 3) {
                if (pal_len == 0 || len > (uint32_t)pal_len) bad("bad tRNS chunk");
                for (uint32_t i = 0; i < len; ++i) palette[i][3] = d[i];
            } else if (color == 0 || color == 2) {
                uint32_t const n = color == 0 ? 1 : 3;
                if (len != 2 * n) bad("bad tRNS chunk");
                for (uint32_t i = 0; i < n; ++i) key[i] = (uint16_t)((d[2 * i] << 8) | d[2 * i + 1]);
            } else {
                bad("tRNS chunk in an image with an alpha channel");
            }
            have_trns = true;
        } else if (!std::memcmp(type, "IDAT", 4)) {
            if (!have_ihdr) bad("malformed PNG");
            idat.insert(idat.end(), d, d + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + (size_t)len;
    }
    int const samples = color == 0 ? 1 : color == 2 ? 3 : color == 3 ? 1 : color == 4 ? 2 : color == 6 ? 4 : 0;
    bool const depth_ok = color == 0 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)
                          : color == 3 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8)
                                       : (depth == 8 || depth == 16);
    if (!have_ihdr || !w || !h || !samples || !depth_ok) bad("unsupported PNG colour type / bit depth");
    if (color == 3 && pal_len == 0) bad("PNG palette missing");
    int const out_ch = color == 3 ? (have_trns ? 4 : 3) : samples + ((have_trns && (color == 0 || color == 2)) ? 1 : 0);
    // sizes in 64 bits, bounded before anything is allocated (a 2^31 x 2^31 header must not wrap to a small buffer)
    if (w > kMaxImageDim || h > kMaxImageDim || (uint64_t)w * h * (uint64_t)std::max(out_ch, samples * (depth / 8 ? depth / 8 : 1)) > kMaxImageBytes)
        bad("image too large");
    int const px_bits = depth * samples;
    int const fbpp = std::max(1, px_bits / 8);  // distance of the "left" byte for the filters
    // the seven Adam7 passes (or the one pass of a plain image): origin, spacing, extent
    static int const xorig[7] = {0, 4, 0, 2, 0, 1, 0}, yorig[7] = {0, 0, 4, 0, 2, 0, 1};
    static int const xspc[7] = {8, 8, 4, 4, 2, 2, 1}, yspc[7] = {8, 8, 8, 4, 4, 2, 2};
    struct Pass { uint32_t x0, y0, dx, dy, pw, ph; size_t row; };
    std::vector<Pass> passes;
    size_t expect = 0;
    if (!interlace) {
        passes.push_back(Pass{0, 0, 1, 1, w, h, ((size_t)w * px_bits + 7) / 8});
    } else {
        for (int p = 0; p < 7; ++p) {
            uint32_t const pw = (w + (uint32_t)xspc[p] - 1 - (uint32_t)xorig[p]) / (uint32_t)xspc[p];
            uint32_t const ph = (h + (uint32_t)yspc[p] - 1 - (uint32_t)yorig[p]) / (uint32_t)yspc[p];
            if ((uint32_t)xorig[p] >= w || (uint32_t)yorig[p] >= h || !pw || !ph) continue;
            passes.push_back(Pass{(uint32_t)xorig[p], (uint32_t)yorig[p], (uint32_t)xspc[p], (uint32_t)yspc[p], pw, ph, ((size_t)pw * px_bits + 7) / 8});
        }
    }
    for (Pass const& p : passes) expect += (p.row + 1) * (size_t)p.ph;
    std::vector<uint8_t> raw = inflate(idat.data(), idat.size(), expect);
    if (raw.size() < expect) bad("PNG data too short");

    uint8_t* px = image_alloc((size_t)w * h * out_ch);
    static int const depth_scale[9] = {0, 0xff, 0x55, 0, 0x11, 0, 0, 0, 0x01};
    std::vector<uint8_t> prev, cur, line;  // unfiltered scanlines of the pass; one scanline as 8-bit samples
    size_t at = 0;
    try {
        for (Pass const& p : passes) {
            prev.assign(p.row, 0);
            cur.assign(p.row, 0);
            line.assign((size_t)p.pw * samples * (depth == 16 ? 2 : 1), 0);
            for (uint32_t y = 0; y < p.ph; ++y) {
                int const ft = raw[at++];
                uint8_t const* src = &raw[at];
                at += p.row;
                for (size_t x = 0; x < p.row; ++x) {
                    int const a = x >= (size_t)fbpp ? cur[x - fbpp] : 0, b = prev[x], c = x >= (size_t)fbpp ? prev[x - fbpp] : 0;
                    int v = src[x];
                    switch (ft) {
                    case 0: break;
                    case 1: v += a; break;
                    case 2: v += b; break;
                    case 3: v += (a + b) >> 1; break;
                    case 4: v += paeth(a, b, c); break;
                    default: bad("bad PNG filter");
                    }
                    cur[x] = (uint8_t)v;
                }
                // scanline -> samples (16-bit ones stay two bytes until the colour key has been compared)
                if (depth >= 8) {
                    std::memcpy(line.data(), cur.data(), line.size());
                } else {
                    int const mask = (1 << depth) - 1, scale = color == 0 ? depth_scale[depth] : 1;
                    for (size_t i = 0; i < (size_t)p.pw; ++i) {
                        size_t const bit = i * (size_t)depth;
                        int const v = (cur[bit >> 3] >> (8 - depth - (int)(bit & 7))) & mask;
                        line[i] = (uint8_t)(v * scale);
                    }
                }
                uint8_t* dst_row = px + ((size_t)(p.y0 + y * p.dy) * w) * out_ch;
                for (uint32_t i = 0; i < p.pw; ++i) {
                    uint8_t* o = dst_row + (size_t)(p.x0 + i * p.dx) * out_ch;
                    if (color == 3) {
                        int const idx = line[i];
                        if (idx >= pal_len) bad("PNG palette index out of range");
                        o[0] = palette[idx][0]; o[1] = palette[idx][1]; o[2] = palette[idx][2];
                        if (out_ch == 4) o[3] = palette[idx][3];
                    } else if (depth == 16) {
                        uint8_t const* s16 = &line[(size_t)i * samples * 2];
                        bool is_key = have_trns;
                        for (int k = 0; k < samples; ++k) {
                            o[k] = s16[2 * k];  // the high byte
                            if (have_trns && (uint16_t)((s16[2 * k] << 8) | s16[2 * k + 1]) != key[k]) is_key = false;
                        }
                        if (out_ch > samples) o[samples] = is_key ? 0 : 255;
                    } else {
                        uint8_t const* s8 = &line[(size_t)i * samples];
                        bool is_key = have_trns;
                        for (int k = 0; k < samples; ++k) {
                            o[k] = s8[k];
                            if (have_trns && s8[k] != (uint8_t)((key[k] & 255) * (color == 0 ? depth_scale[depth] : 1))) is_key = false;
                        }
                        if (out_ch > samples) o[samples] = is_key ? 0 : 255;
                    }
                }
                prev.swap(cur);
            }
        }
    } catch (...) {
        image_free(px);
        throw;
    }
    extent[0] = (int)w;
    extent[1] = (int)h;
    *channels = out_ch;
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
        while (pos < file.size() && file[pos] >= '0' && file[pos] <= '9') {
            if (v > (int)kMaxImageDim) fail(std::string("Failed to load image ") + path + ": malformed PNM header");
            v = v * 10 + (file[pos++] - '0');
            any = true;
        }
        if (!any) fail(std::string("Failed to load image ") + path + ": malformed PNM header");
        return v;
    };
    int const ch = file[1] == '5' ? 1 : 3;
    int const w = next_int(), h = next_int(), maxv = next_int();
    ++pos;  // single whitespace after maxval
    if (maxv != 255 || w <= 0 || h <= 0 || (uint32_t)w > kMaxImageDim || (uint32_t)h > kMaxImageDim ||
        (uint64_t)w * h * ch > kMaxImageBytes || pos + (size_t)w * h * ch > file.size())
        fail(std::string("Failed to load image ") + path + ": unsupported PNM variant");
    uint8_t* px = image_alloc((size_t)w * h * ch);
    std::memcpy(px, &file[pos], (size_t)w * h * ch);
    extent[0] = w;
    extent[1] = h;
    *channels = ch;
    return px;
}

// ---- baseline JPEG (ITU T.81 sequential DCT, 8-bit, Huffman) ------------------------------------------------------
// The reference's load_image is stbi_load (image.cpp:11-23), which reads JPEG; test/input/truck.jpg is the one real
// fixture of the reference.  Restated from the published stb_image algorithm so that pixels come out the way the
// reference sees them: integer IDCT with 12-bit constants, "hv_2" triangle-filter chroma upsampling, 20-bit fixed-point
// YCbCr -> RGB.  Sequential (SOF0 / SOF1) and progressive (SOF2) Huffman files, interleaved or one scan per component, with
// restart intervals; lossless, hierarchical and arithmetic-coded files are refused.
struct JpegHuff {
    uint8_t size[257];
    uint16_t code[256];
    uint8_t values[256];
    int maxcode[18];
    int delta[17];
    int n = 0;
    void build(uint8_t const* counts, uint8_t const* vals, int nvals) {
        int k = 0;
        for (int i = 0; i < 16; ++i)
            for (int j = 0; j < counts[i]; ++j) size[k++] = (uint8_t)(i + 1);
        size[k] = 0;
        n = k;
        if (n != nvals || n > 256) fail("JPEG: bad Huffman table");
        std::memcpy(values, vals, (size_t)n);
        int c = 0;
        k = 0;
        for (int j = 1; j <= 16; ++j) {
            delta[j] = k - c;
            if (size[k] == j) {
                while (size[k] == j) code[k++] = (uint16_t)c++;
                if (c - 1 >= (1 << j)) fail("JPEG: bad code lengths");
            }
            maxcode[j] = c << (16 - j);
            c <<= 1;
        }
        maxcode[17] = 0x7fffffff;
    }
};

struct JpegBits {
    uint8_t const* p;
    size_t n, pos;
    uint32_t buf = 0;
    int cnt = 0;
    bool hit_marker = false;
    void fill() {
        while (cnt <= 24) {
            int b = 0;
            if (!hit_marker && pos < n) {
                b = p[pos++];
                if (b == 0xFF) {
                    int const c = pos < n ? p[pos] : 0xD9;
                    if (c == 0) ++pos;           // stuffed zero
                    else { hit_marker = true; --pos; b = 0; }  // a marker: feed zeros from here on
                }
            }
            buf |= (uint32_t)b << (24 - cnt);
            cnt += 8;
        }
    }
    int decode(JpegHuff const& h) {
        if (cnt < 16) fill();
        uint32_t const top = buf >> 16;
        int len = 1;
        while (len <= 16 && (int)top >= h.maxcode[len]) ++len;
        if (len > 16) fail("JPEG: bad Huffman code");
        int const idx = (int)((buf >> (32 - len)) & ((1u << len) - 1)) + h.delta[len];
        if (idx < 0 || idx >= h.n) fail("JPEG: bad Huffman code");
        buf <<= len;
        cnt -= len;
        return h.values[idx];
    }
    int receive_extend(int nbits) {  // T.81 F.2.2.1: nbits magnitude bits, sign-extended
        if (nbits == 0) return 0;
        if (cnt < nbits) fill();
        int const v = (int)(buf >> (32 - nbits));
        buf <<= nbits;
        cnt -= nbits;
        return v < (1 << (nbits - 1)) ? v - (1 << nbits) + 1 : v;
    }
    void reset() { buf = 0; cnt = 0; hit_marker = false; }
};

inline uint8_t jclamp(int x) { return (uint8_t)((unsigned)x > 255 ? (x < 0 ? 0 : 255) : x); }
inline uint8_t jclamp(int64_t x) { return (uint8_t)(x < 0 ? 0 : (x > 255 ? 255 : x)); }

#define DLIMG_F2F(x) ((int64_t)(((x) * 4096 + 0.5)))
#define DLIMG_IDCT_1D(s0, s1, s2, s3, s4, s5, s6, s7)        \
    int64_t t0, t1, t2, t3, p1, p2, p3, p4, p5, x0, x1, x2, x3; \
    p2 = s2;                                                  \
    p3 = s6;                                                  \
    p1 = (p2 + p3) * DLIMG_F2F(0.5411961f);                   \
    t2 = p1 + p3 * DLIMG_F2F(-1.847759065f);                  \
    t3 = p1 + p2 * DLIMG_F2F(0.765366865f);                   \
    p2 = s0;                                                  \
    p3 = s4;                                                  \
    t0 = (p2 + p3) * 4096;                                    \
    t1 = (p2 - p3) * 4096;                                    \
    x0 = t0 + t3;                                             \
    x3 = t0 - t3;                                             \
    x1 = t1 + t2;                                             \
    x2 = t1 - t2;                                             \
    t0 = s7;                                                  \
    t1 = s5;                                                  \
    t2 = s3;                                                  \
    t3 = s1;                                                  \
    p3 = t0 + t2;                                             \
    p4 = t1 + t3;                                             \
    p1 = t0 + t3;                                             \
    p2 = t1 + t2;                                             \
    p5 = (p3 + p4) * DLIMG_F2F(1.175875602f);                 \
    t0 = t0 * DLIMG_F2F(0.298631336f);                        \
    t1 = t1 * DLIMG_F2F(2.053119869f);                        \
    t2 = t2 * DLIMG_F2F(3.072711026f);                        \
    t3 = t3 * DLIMG_F2F(1.501321110f);                        \
    p1 = p5 + p1 * DLIMG_F2F(-0.899976223f);                  \
    p2 = p5 + p2 * DLIMG_F2F(-2.562915447f);                  \
    p3 = p3 * DLIMG_F2F(-1.961570560f);                       \
    p4 = p4 * DLIMG_F2F(-0.390180644f);                       \
    t3 += p1 + p4;                                            \
    t2 += p2 + p3;                                            \
    t1 += p2 + p4;                                            \
    t0 += p1 + p3;

// (64-bit intermediates: the arithmetic is stb_image's, but a corrupt stream can carry coefficients whose products leave 32
// bits, which is undefined behaviour for the int the published code uses; results are identical whenever that one is defined)
void jpeg_idct_block(uint8_t* out, int out_stride, short const data[64]) {
    int64_t val[64], *v = val;
    short const* d = data;
    for (int i = 0; i < 8; ++i, ++d, ++v) {  // columns
        if (d[8] == 0 && d[16] == 0 && d[24] == 0 && d[32] == 0 && d[40] == 0 && d[48] == 0 && d[56] == 0) {
            int64_t const dcterm = (int64_t)d[0] * 4;
            v[0] = v[8] = v[16] = v[24] = v[32] = v[40] = v[48] = v[56] = dcterm;
        } else {
            DLIMG_IDCT_1D(d[0], d[8], d[16], d[24], d[32], d[40], d[48], d[56])
            x0 += 512; x1 += 512; x2 += 512; x3 += 512;  // constants scaled by 1 << 12; keep 2 extra bits
            v[0] = (x0 + t3) >> 10;
            v[56] = (x0 - t3) >> 10;
            v[8] = (x1 + t2) >> 10;
            v[48] = (x1 - t2) >> 10;
            v[16] = (x2 + t1) >> 10;
            v[40] = (x2 - t1) >> 10;
            v[24] = (x3 + t0) >> 10;
            v[32] = (x3 - t0) >> 10;
        }
    }
    v = val;
    uint8_t* o = out;
    for (int i = 0; i < 8; ++i, v += 8, o += out_stride) {  // rows: remove 1 << 17, round, + 128
        DLIMG_IDCT_1D(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7])
        x0 += 65536 + (128 << 17);
        x1 += 65536 + (128 << 17);
        x2 += 65536 + (128 << 17);
        x3 += 65536 + (128 << 17);
        o[0] = jclamp((x0 + t3) >> 17);
        o[7] = jclamp((x0 - t3) >> 17);
        o[1] = jclamp((x1 + t2) >> 17);
        o[6] = jclamp((x1 - t2) >> 17);
        o[2] = jclamp((x2 + t1) >> 17);
        o[5] = jclamp((x2 - t1) >> 17);
        o[3] = jclamp((x3 + t0) >> 17);
        o[4] = jclamp((x3 - t0) >> 17);
    }
}
#undef DLIMG_IDCT_1D
#undef DLIMG_F2F

struct JpegComp {
    int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
    int w2 = 0, h2 = 0;  // padded plane extent
    int dc_pred = 0;
    std::vector<short> coef;    // DCT coefficients, 64 per block, blocks in raster order over the padded plane
    std::vector<uint8_t> data;  // samples
};

// one row of chroma at full horizontal resolution from the two nearest source rows (3/4 near + 1/4 far vertically,
// then 3/4 - 1/4 horizontally), stb_image's stbi__resample_row_hv_2
void resample_hv2(uint8_t* out, uint8_t const* in_near, uint8_t const* in_far, int w) {
    if (w == 1) {
        out[0] = out[1] = (uint8_t)((3 * in_near[0] + in_far[0] + 2) >> 2);
        return;
    }
    int t1 = 3 * in_near[0] + in_far[0];
    out[0] = (uint8_t)((t1 + 2) >> 2);
    for (int i = 1; i < w; ++i) {
        int const t0 = t1;
        t1 = 3 * in_near[i] + in_far[i];
        out[i * 2 - 1] = (uint8_t)((3 * t0 + t1 + 8) >> 4);
        out[i * 2] = (uint8_t)((3 * t1 + t0 + 8) >> 4);
    }
    out[w * 2 - 1] = (uint8_t)((t1 + 2) >> 2);
}
void resample_h2(uint8_t* out, uint8_t const* in, int w) {  // stbi__resample_row_h_2
    if (w == 1) {
        out[0] = out[1] = in[0];
        return;
    }
    out[0] = in[0];
    out[1] = (uint8_t)((in[0] * 3 + in[1] + 2) >> 2);
    int i;
    for (i = 1; i < w - 1; ++i) {
        int const n = 3 * in[i] + 2;
        out[i * 2 + 0] = (uint8_t)((n + in[i - 1]) >> 2);
        out[i * 2 + 1] = (uint8_t)((n + in[i + 1]) >> 2);
    }
    out[i * 2 + 0] = (uint8_t)((in[w - 2] * 3 + in[w - 1] + 2) >> 2);
    out[i * 2 + 1] = in[w - 1];
}

uint8_t* load_jpeg(std::vector<uint8_t> const& file, char const* path, int* extent, int* channels) {
    static uint8_t const zigzag[64 + 15] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                            41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                            30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63,
                                            // (a corrupt run may step past 63: those land on the last coefficient)
                                            63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};
    auto bad = [&](char const* why) { fail(std::string("Failed to load image ") + path + ": " + why); };
    uint16_t quant[4][64] = {};
    JpegHuff dc[4], ac[4];
    bool have_dc[4] = {}, have_ac[4] = {};
    std::vector<JpegComp> comps;
    int W = 0, H = 0, restart = 0, hmax = 1, vmax = 1;
    bool adobe_transform_rgb = false, progressive = false, any_scan = false;
    size_t pos = 2;
    auto u16 = [&](size_t at) {
        if (at + 2 > file.size()) bad("truncated JPEG");
        return (int)((file[at] << 8) | file[at + 1]);
    };
    int mcux = 0, mcuy = 0;

    // One scan (T.81 B.2.3 / G.1): entropy-coded data from file[pos] into the coefficient planes.  Sequential files put whole
    // (dequantised) blocks there; progressive ones build the coefficients up over several scans -- the first DC scan, DC
    // refinement bits, AC bands and their refinements (spectral selection ss..se, successive approximation ah / al) --
    // and are dequantised when the image is complete.  Scans of ONE component walk that component's own blocks in raster
    // order (not the padded MCU grid).
    auto decode_scan = [&](std::vector<JpegComp*> const& sc, int ss, int se, int ah, int al) {
        JpegBits br{file.data(), file.size(), pos};
        int eob_run = 0;
        for (JpegComp* c : sc) c->dc_pred = 0;
        auto getbit = [&]() {
            if (br.cnt < 1) br.fill();
            int const b = (int)(br.buf >> 31);
            br.buf <<= 1;
            --br.cnt;
            return b;
        };
        auto getbits = [&](int n) {
            if (n == 0) return 0;
            if (br.cnt < n) br.fill();
            int const v = (int)(br.buf >> (32 - n));
            br.buf <<= n;
            br.cnt -= n;
            return v;
        };
        auto sequential_block = [&](JpegComp& c, short* blk) {
            int const t = br.decode(dc[c.td]);
            if (t > 15) bad("bad DC code");
            c.dc_pred += br.receive_extend(t);
            if (c.dc_pred < -32768 || c.dc_pred > 32767) bad("bad DC delta");  // (a valid predictor has 12 bits; keeps the sums in range)
            blk[0] = (short)(c.dc_pred * quant[c.tq][0]);
            for (int k = 1; k < 64;) {
                int const rs = br.decode(ac[c.ta]);
                int const s = rs & 15, r = rs >> 4;
                if (s == 0) {
                    if (rs != 0xF0) break;  // end of block
                    k += 16;
                } else {
                    k += r;
                    if (k > 63) bad("bad AC run");
                    int const z = zigzag[k++];
                    blk[z] = (short)(br.receive_extend(s) * quant[c.tq][z]);
                }
            }
        };
        auto progressive_dc = [&](JpegComp& c, short* blk) {
            if (ah == 0) {  // first pass: the difference, scaled by the point transform
                int const t = br.decode(dc[c.td]);
                if (t > 15) bad("bad DC code");
                c.dc_pred += br.receive_extend(t);
                if (c.dc_pred < -32768 || c.dc_pred > 32767) bad("bad DC delta");
                blk[0] = (short)(c.dc_pred * (1 << al));
            } else if (getbit()) {  // refinement: one more bit of precision
                blk[0] = (short)(blk[0] + (1 << al));
            }
        };
        auto progressive_ac = [&](JpegComp& c, short* blk) {
            JpegHuff const& h = ac[c.ta];
            if (ah == 0) {  // first pass over the band ss..se
                if (eob_run) {
                    --eob_run;
                    return;
                }
                int k = ss;
                do {
                    int const rs = br.decode(h);
                    int const s = rs & 15, r = rs >> 4;
                    if (s == 0) {
                        if (r < 15) {  // end of band for this and the next eob_run blocks
                            eob_run = 1 << r;
                            if (r) eob_run += getbits(r);
                            --eob_run;
                            break;
                        }
                        k += 16;
                    } else {
                        k += r;
                        int const z = zigzag[k++];
                        blk[z] = (short)(br.receive_extend(s) * (1 << al));
                    }
                } while (k <= se);
                return;
            }
            // refinement pass: a correction bit for every coefficient that is already non-zero, new +-1 coefficients in between
            short const bit = (short)(1 << al);
            auto refine = [&](short* p) {
                if (getbit() && (*p & bit) == 0) *p = (short)(*p > 0 ? *p + bit : *p - bit);
            };
            if (eob_run) {
                --eob_run;
                for (int k = ss; k <= se; ++k) {
                    short* p = &blk[zigzag[k]];
                    if (*p != 0) refine(p);
                }
                return;
            }
            int k = ss;
            do {
                int const rs = br.decode(h);
                int s = rs & 15, r = rs >> 4;
                if (s == 0) {
                    if (r < 15) {
                        eob_run = (1 << r) - 1;
                        if (r) eob_run += getbits(r);
                        r = 64;  // run to the end of the band
                    }  // else: sixteen zero coefficients (only zeros count)
                } else {
                    if (s != 1) bad("bad refinement code");
                    s = getbit() ? bit : -bit;
                }
                while (k <= se) {
                    short* p = &blk[zigzag[k++]];
                    if (*p != 0) {
                        refine(p);
                    } else {
                        if (r == 0) {
                            *p = (short)s;
                            break;
                        }
                        --r;
                    }
                }
            } while (k <= se);
        };
        auto block = [&](JpegComp& c, int bx, int by) {
            short* blk = &c.coef[((size_t)by * (size_t)(c.w2 / 8) + (size_t)bx) * 64];
            if (!progressive) sequential_block(c, blk);
            else if (ss == 0) progressive_dc(c, blk);
            else progressive_ac(c, blk);
        };
        int todo = restart ? restart : 0x7fffffff;
        auto after_unit = [&]() {
            if (--todo > 0) return;
            // restart interval: byte-align, expect RSTn, reset the predictors and the end-of-band run
            br.reset();
            size_t q = br.pos;
            while (q + 1 < file.size() && !(file[q] == 0xFF && file[q + 1] >= 0xD0 && file[q + 1] <= 0xD7)) {
                if (file[q] == 0xFF && file[q + 1] != 0 && file[q + 1] != 0xFF) break;  // another marker: the scan is over
                ++q;
            }
            if (q + 1 < file.size() && file[q + 1] >= 0xD0 && file[q + 1] <= 0xD7) br.pos = q + 2;
            for (JpegComp* c : sc) c->dc_pred = 0;
            eob_run = 0;
            todo = restart;
        };
        if (sc.size() == 1) {
            JpegComp& c = *sc[0];
            int const bw = ((W * c.h + hmax - 1) / hmax + 7) >> 3, bh = ((H * c.v + vmax - 1) / vmax + 7) >> 3;
            for (int by = 0; by < bh; ++by)
                for (int bx = 0; bx < bw; ++bx) {
                    block(c, bx, by);
                    after_unit();
                }
        } else {
            for (int my = 0; my < mcuy; ++my)
                for (int mx = 0; mx < mcux; ++mx) {
                    for (JpegComp* c : sc)
                        for (int by = 0; by < c->v; ++by)
                            for (int bx = 0; bx < c->h; ++bx) block(*c, mx * c->h + bx, my * c->v + by);
                    after_unit();
                }
        }
        pos = br.pos;
    };

    bool done = false;
    while (!done) {
        // next marker: 0xFF, any number of fill 0xFF, a non-zero code (0xFF 0x00 is a stuffed data byte left over from a scan)
        int marker = 0;
        while (marker == 0) {
            while (pos < file.size() && file[pos] != 0xFF) ++pos;
            while (pos < file.size() && file[pos] == 0xFF) ++pos;
            if (pos >= file.size()) {
                if (!any_scan) bad("no scan in JPEG");
                marker = 0xD9;  // the end-of-image marker is missing: decode what was read
                break;
            }
            marker = file[pos++];
        }
        if (marker == 0xD8 || marker == 0x01 || (marker >= 0xD0 && marker <= 0xD7)) continue;
        if (marker == 0xD9) {
            if (!any_scan) bad("no scan in JPEG");
            break;
        }
        int const len = u16(pos);
        if (len < 2 || pos + (size_t)len > file.size()) bad("truncated JPEG segment");
        uint8_t const* d = &file[pos + 2];
        int const n = len - 2;
        pos += (size_t)len;
        switch (marker) {
        case 0xDB:  // quantisation tables
            for (int i = 0; i < n;) {
                int const pq = d[i] >> 4, tq = d[i] & 15;
                if (tq > 3 || pq > 1 || i + 1 + 64 * (pq + 1) > n) bad("bad DQT");
                for (int k = 0; k < 64; ++k) quant[tq][zigzag[k]] = pq ? (uint16_t)((d[i + 1 + 2 * k] << 8) | d[i + 2 + 2 * k]) : d[i + 1 + k];
                i += 1 + 64 * (pq + 1);
            }
            break;
        case 0xC4:  // Huffman tables (progressive files redefine them between scans)
            for (int i = 0; i < n;) {
                if (i + 17 > n) bad("bad DHT");
                int const tc = d[i] >> 4, th = d[i] & 15;
                int total = 0;
                for (int k = 0; k < 16; ++k) total += d[i + 1 + k];
                if (tc > 1 || th > 3 || total > 256 || i + 17 + total > n) bad("bad DHT");
                (tc ? ac : dc)[th].build(d + i + 1, d + i + 17, total);
                (tc ? have_ac : have_dc)[th] = true;
                i += 17 + total;
            }
            break;
        case 0xC0: case 0xC1: case 0xC2: {  // baseline / extended sequential / progressive, Huffman
            if (!comps.empty()) bad("more than one frame header");
            progressive = marker == 0xC2;
            if (n < 6 || d[0] != 8) bad("only 8-bit JPEG is supported");
            H = (d[1] << 8) | d[2];
            W = (d[3] << 8) | d[4];
            int const nc = d[5];
            if (W <= 0 || H <= 0 || (nc != 1 && nc != 3) || n < 6 + 3 * nc) bad("unsupported JPEG frame");
            if ((uint32_t)W > kMaxImageDim || (uint32_t)H > kMaxImageDim || (uint64_t)W * H * 3 > kMaxImageBytes) bad("image too large");
            comps.resize((size_t)nc);
            for (int i = 0; i < nc; ++i) {
                comps[(size_t)i].id = d[6 + 3 * i];
                comps[(size_t)i].h = d[7 + 3 * i] >> 4;
                comps[(size_t)i].v = d[7 + 3 * i] & 15;
                comps[(size_t)i].tq = d[8 + 3 * i];
                if (comps[(size_t)i].h < 1 || comps[(size_t)i].h > 4 || comps[(size_t)i].v < 1 || comps[(size_t)i].v > 4 || comps[(size_t)i].tq > 3)
                    bad("bad JPEG frame header");
                hmax = std::max(hmax, comps[(size_t)i].h);
                vmax = std::max(vmax, comps[(size_t)i].v);
            }
            // coefficient planes, each padded to whole MCUs
            mcux = (W + 8 * hmax - 1) / (8 * hmax);
            mcuy = (H + 8 * vmax - 1) / (8 * vmax);
            for (auto& c : comps) {
                c.w2 = mcux * c.h * 8;
                c.h2 = mcuy * c.v * 8;
                c.coef.assign((size_t)c.w2 * c.h2, 0);
            }
            break;
        }
        case 0xC3: case 0xC5: case 0xC6: case 0xC7: case 0xC9: case 0xCA: case 0xCB: case 0xCD: case 0xCE: case 0xCF:
            bad("only Huffman-coded sequential and progressive JPEG is supported (no lossless, hierarchical or arithmetic coding)");
            break;
        case 0xDD: restart = n >= 2 ? ((d[0] << 8) | d[1]) : 0; break;
        case 0xEE: if (n >= 12 && !std::memcmp(d, "Adobe", 5)) adobe_transform_rgb = d[11] == 0; break;
        case 0xDA: {
            if (comps.empty()) bad("scan before frame header");
            int const ns = n >= 1 ? d[0] : 0;
            if (ns < 1 || ns > (int)comps.size() || n < 1 + 2 * ns + 3) bad("bad scan header");
            std::vector<JpegComp*> sc;
            for (int i = 0; i < ns; ++i) {
                int const cid = d[1 + 2 * i];
                JpegComp* c = nullptr;
                for (auto& cc : comps)
                    if (cc.id == cid) c = &cc;
                if (!c || std::find(sc.begin(), sc.end(), c) != sc.end()) bad("bad scan component");
                c->td = d[2 + 2 * i] >> 4;
                c->ta = d[2 + 2 * i] & 15;
                if (c->td > 3 || c->ta > 3) bad("bad Huffman table index");
                sc.push_back(c);
            }
            int ss = d[1 + 2 * ns], se = d[2 + 2 * ns];
            int const ah = d[3 + 2 * ns] >> 4, al = d[3 + 2 * ns] & 15;
            if (progressive) {
                if (ss > 63 || se > 63 || ss > se || ah > 13 || al > 13) bad("bad progressive scan parameters");
                if (ss == 0 && se != 0) bad("bad progressive scan parameters");   // DC and AC never share a scan
                if (ss > 0 && ns != 1) bad("bad progressive scan parameters");    // AC scans hold one component
            } else {
                if (ss != 0 || ah != 0 || al != 0) bad("bad sequential scan parameters");
                se = 63;
            }
            for (JpegComp* c : sc) {
                bool const need_dc = !progressive || (ss == 0 && ah == 0), need_ac = !progressive || ss > 0;
                if ((need_dc && !have_dc[c->td]) || (need_ac && !have_ac[c->ta])) bad("missing Huffman table");
            }
            decode_scan(sc, ss, se, ah, al);
            any_scan = true;
            break;
        }
        default: break;  // APPn, COM, ...
        }
    }
    // coefficients -> samples
    for (auto& c : comps) {
        c.data.assign((size_t)c.w2 * c.h2, 0);
        int const bw = c.w2 / 8, bh = c.h2 / 8;
        for (int by = 0; by < bh; ++by)
            for (int bx = 0; bx < bw; ++bx) {
                short* blk = &c.coef[((size_t)by * bw + bx) * 64];
                if (progressive)
                    for (int k = 0; k < 64; ++k) blk[k] = (short)(blk[k] * quant[c.tq][k]);
                jpeg_idct_block(&c.data[(size_t)(by * 8) * c.w2 + (size_t)bx * 8], c.w2, blk);
            }
        std::vector<short>().swap(c.coef);
    }
    // planes -> interleaved pixels
    int const nc = (int)comps.size();
    uint8_t* px = image_alloc((size_t)W * H * (nc == 1 ? 1 : 3));
    if (nc == 1) {
        for (int y = 0; y < H; ++y) std::memcpy(px + (size_t)y * W, &comps[0].data[(size_t)y * comps[0].w2], (size_t)W);
        extent[0] = W; extent[1] = H; *channels = 1;
        return px;
    }
    struct Up { int hs, vs, ystep, ypos, w_lores; uint8_t const *line0, *line1; std::vector<uint8_t> buf; };
    Up up[3];
    for (int k = 0; k < 3; ++k) {
        JpegComp const& c = comps[(size_t)k];
        up[k].hs = hmax / c.h;
        up[k].vs = vmax / c.v;
        if (!((up[k].hs == 1 || up[k].hs == 2) && (up[k].vs == 1 || up[k].vs == 2))) {
            image_free(px);
            bad("unsupported chroma subsampling");
        }
        up[k].ystep = up[k].vs >> 1;
        up[k].ypos = 0;
        up[k].w_lores = (W + up[k].hs - 1) / up[k].hs;
        up[k].line0 = up[k].line1 = c.data.data();
        up[k].buf.assign((size_t)W + 3, 0);
    }
    int const plane_rows[3] = {(H + up[0].vs - 1) / up[0].vs, (H + up[1].vs - 1) / up[1].vs, (H + up[2].vs - 1) / up[2].vs};
    for (int y = 0; y < H; ++y) {
        uint8_t const* row[3];
        for (int k = 0; k < 3; ++k) {
            Up& u = up[k];
            bool const y_bot = u.ystep >= (u.vs >> 1);
            uint8_t const* in_near = y_bot ? u.line1 : u.line0;
            uint8_t const* in_far = y_bot ? u.line0 : u.line1;
            if (u.hs == 1 && u.vs == 1) {
                row[k] = in_near;
            } else if (u.hs == 2 && u.vs == 2) {
                resample_hv2(u.buf.data(), in_near, in_far, u.w_lores);
                row[k] = u.buf.data();
            } else if (u.hs == 2) {
                resample_h2(u.buf.data(), in_near, u.w_lores);
                row[k] = u.buf.data();
            } else {  // vs == 2 only: stbi__resample_row_v_2
                for (int i = 0; i < u.w_lores; ++i) u.buf[(size_t)i] = (uint8_t)((3 * in_near[i] + in_far[i] + 2) >> 2);
                row[k] = u.buf.data();
            }
            if (++u.ystep >= u.vs) {
                u.ystep = 0;
                u.line0 = u.line1;
                if (++u.ypos < plane_rows[k]) u.line1 += comps[(size_t)k].w2;
            }
        }
        uint8_t* out = px + (size_t)y * W * 3;
        if (adobe_transform_rgb) {
            for (int x = 0; x < W; ++x) { out[3 * x] = row[0][x]; out[3 * x + 1] = row[1][x]; out[3 * x + 2] = row[2][x]; }
        } else {
            auto f2fix = [](float v) { return ((int)(v * 4096.0f + 0.5f)) << 8; };
            for (int x = 0; x < W; ++x) {
                int const y_fixed = (row[0][x] << 20) + (1 << 19);
                int const cr = row[2][x] - 128, cb = row[1][x] - 128;
                int r = y_fixed + cr * f2fix(1.40200f);
                int g = y_fixed + (cr * -f2fix(0.71414f)) + ((cb * -f2fix(0.34414f)) & (int)0xffff0000);
                int b = y_fixed + cb * f2fix(1.77200f);
                out[3 * x] = jclamp(r >> 20);
                out[3 * x + 1] = jclamp(g >> 20);
                out[3 * x + 2] = jclamp(b >> 20);
            }
        }
    }
    extent[0] = W; extent[1] = H; *channels = 3;
    return px;
}

// ---- BMP and TGA (the reference's load_image documents "PNG, JPEG, BMP, TGA", dlimgedit.hpp:59) -----------------
// Restated from the published stb_image behaviour the reference sees: BMP with 1 / 4 / 8-bit palettes, 16-bit (5-5-5 or
// bit fields), 24-bit and 32-bit pixels (bit fields or plain, where an all-zero alpha channel reads as opaque), bottom-up
// or top-down, no RLE; TGA of types 1 / 2 / 3 and their run-length forms 9 / 10 / 11 with 8, 15 / 16, 24 or 32 bits per
// pixel and either vertical origin.  Output channels: 3, or 4 when the file carries alpha, or 1 for 8-bit grey TGA.
int high_bit(uint32_t z) {
    int n = 0;
    if (z == 0) return -1;
    if (z >= 0x10000) { n += 16; z >>= 16; }
    if (z >= 0x00100) { n += 8; z >>= 8; }
    if (z >= 0x00010) { n += 4; z >>= 4; }
    if (z >= 0x00004) { n += 2; z >>= 2; }
    if (z >= 0x00002) { n += 1; }
    return n;
}
int bit_count(uint32_t a) {
    int n = 0;
    for (; a; a &= a - 1) ++n;
    return n;
}
// a field of `bits` bits at `shift` -> 8 bits by replicating the bit pattern
int shift_signed(uint32_t v, int shift, int bits) {
    static uint32_t const mul_table[9] = {0, 0xff, 0x55, 0x49, 0x11, 0x21, 0x41, 0x81, 0x01};
    static uint32_t const shift_table[9] = {0, 0, 0, 1, 0, 2, 4, 6, 0};
    if (shift < 0) v <<= -shift;
    else v >>= shift;
    v >>= (8 - bits);
    return (int)((v * mul_table[bits]) >> shift_table[bits]);
}

uint8_t* load_bmp(std::vector<uint8_t> const& file, char const* path, int* extent, int* channels) {
    auto bad = [&](char const* why) { fail(std::string("Failed to load image ") + path + ": " + why); };
    auto le16 = [&](size_t at) -> uint32_t {
        if (at + 2 > file.size()) bad("truncated BMP");
        return (uint32_t)file[at] | ((uint32_t)file[at + 1] << 8);
    };
    auto le32 = [&](size_t at) -> uint32_t { return le16(at) | (le16(at + 2) << 16); };
    uint32_t const offset = le32(10), hsz = le32(14);
    if (hsz != 12 && hsz != 40 && hsz != 56 && hsz != 108 && hsz != 124) bad("unknown BMP header");
    int w, h, bpp;
    uint32_t compress = 0, mr = 0, mg = 0, mb = 0, ma = 0;
    bool all_a_check = false;
    if (hsz == 12) {
        w = (int)le16(18);
        h = (int)le16(20);
        if (le16(22) != 1) bad("bad BMP");
        bpp = (int)le16(24);
    } else {
        w = (int)le32(18);
        h = (int)le32(22);
        if (le16(26) != 1) bad("bad BMP");
        bpp = (int)le16(28);
        compress = le32(30);
        if (compress == 1 || compress == 2) bad("run-length encoded BMP is not supported");
        if (compress >= 4) bad("BMP with embedded JPEG / PNG is not supported");
        if (compress == 3 && bpp != 16 && bpp != 32) bad("bad BMP");
        if (hsz == 40 || hsz == 56) {
            if (bpp == 16 || bpp == 32) {
                if (compress == 0) {
                    if (bpp == 32) { mr = 0xffu << 16; mg = 0xffu << 8; mb = 0xffu; ma = 0xffu << 24; all_a_check = true; }
                    else { mr = 31u << 10; mg = 31u << 5; mb = 31u; }
                } else {  // bit fields behind the 40-byte header (56-byte headers hold them inside)
                    mr = le32(54); mg = le32(58); mb = le32(62);
                    if (hsz == 56) ma = le32(66);
                    if (mr == mg && mg == mb) bad("bad BMP");
                }
            }
        } else {  // V4 / V5 headers always carry the masks
            mr = le32(54); mg = le32(58); mb = le32(62); ma = le32(66);
            if (compress != 3 && (bpp == 16 || bpp == 32)) {
                if (bpp == 32) { mr = 0xffu << 16; mg = 0xffu << 8; mb = 0xffu; ma = 0xffu << 24; all_a_check = true; }
                else { mr = 31u << 10; mg = 31u << 5; mb = 31u; ma = 0; }
            }
        }
    }
    bool const flip = h > 0;  // positive height: rows are stored bottom-up
    if (h < 0) h = -h;
    if (h == INT32_MIN || w <= 0 || h <= 0 || (uint32_t)w > kMaxImageDim || (uint32_t)h > kMaxImageDim) bad("bad BMP extent");
    if (bpp != 1 && bpp != 4 && bpp != 8 && bpp != 16 && bpp != 24 && bpp != 32) bad("unsupported BMP bit depth");
    int const out_ch = ma ? 4 : 3;
    if ((uint64_t)w * h * out_ch > kMaxImageBytes) bad("image too large");
    int psize = 0;
    if (bpp < 16) {
        if (hsz == 12) psize = ((int)offset - 14 - 12) / 3;
        else psize = ((int)offset - 14 - (int)hsz) >> 2;
        if (psize <= 0 || psize > 256) bad("bad BMP palette");
    }
    size_t const row_in = (((size_t)w * bpp + 31) / 32) * 4;
    if ((size_t)offset + row_in * (size_t)h > file.size()) bad("truncated BMP");
    uint8_t* px = image_alloc((size_t)w * h * out_ch);
    if (bpp < 16) {
        uint8_t pal[256][3] = {};
        size_t pp = 14 + hsz;
        for (int i = 0; i < psize; ++i) {
            if (pp + 3 > file.size()) { image_free(px); bad("truncated BMP"); }
            pal[i][2] = file[pp]; pal[i][1] = file[pp + 1]; pal[i][0] = file[pp + 2];
            pp += hsz == 12 ? 3 : 4;
        }
        for (int y = 0; y < h; ++y) {
            uint8_t const* src = &file[offset + row_in * (size_t)y];
            uint8_t* dst = px + (size_t)(flip ? h - 1 - y : y) * w * 3;
            for (int x = 0; x < w; ++x) {
                int v;
                if (bpp == 8) v = src[x];
                else if (bpp == 4) v = (src[x >> 1] >> ((x & 1) ? 0 : 4)) & 15;
                else v = (src[x >> 3] >> (7 - (x & 7))) & 1;
                dst[3 * x] = pal[v][0]; dst[3 * x + 1] = pal[v][1]; dst[3 * x + 2] = pal[v][2];
            }
        }
    } else {
        int rshift = 0, gshift = 0, bshift = 0, ashift = 0, rcount = 0, gcount = 0, bcount = 0, acount = 0;
        bool const easy24 = bpp == 24, easy32 = bpp == 32 && mb == 0xffu && mg == 0xff00u && mr == 0x00ff0000u && ma == 0xff000000u;
        if (!easy24 && !easy32) {
            if (!mr || !mg || !mb) { image_free(px); bad("bad BMP masks"); }
            rshift = high_bit(mr) - 7; rcount = bit_count(mr);
            gshift = high_bit(mg) - 7; gcount = bit_count(mg);
            bshift = high_bit(mb) - 7; bcount = bit_count(mb);
            ashift = high_bit(ma) - 7; acount = bit_count(ma);
            if (rcount > 8 || gcount > 8 || bcount > 8 || acount > 8) { image_free(px); bad("bad BMP masks"); }
        }
        uint32_t all_a = 0;
        for (int y = 0; y < h; ++y) {
            uint8_t const* src = &file[offset + row_in * (size_t)y];
            uint8_t* dst = px + (size_t)(flip ? h - 1 - y : y) * w * out_ch;
            for (int x = 0; x < w; ++x) {
                if (easy24 || easy32) {
                    uint8_t const* s = src + (size_t)x * (bpp / 8);
                    dst[0] = s[2]; dst[1] = s[1]; dst[2] = s[0];
                    if (out_ch == 4) { dst[3] = easy32 ? s[3] : 255; all_a |= dst[3]; }
                } else {
                    uint32_t const v = bpp == 16 ? ((uint32_t)src[2 * x] | ((uint32_t)src[2 * x + 1] << 8))
                                                 : ((uint32_t)src[4 * x] | ((uint32_t)src[4 * x + 1] << 8) | ((uint32_t)src[4 * x + 2] << 16) | ((uint32_t)src[4 * x + 3] << 24));
                    dst[0] = (uint8_t)shift_signed(v & mr, rshift, rcount);
                    dst[1] = (uint8_t)shift_signed(v & mg, gshift, gcount);
                    dst[2] = (uint8_t)shift_signed(v & mb, bshift, bcount);
                    if (out_ch == 4) { dst[3] = ma ? (uint8_t)shift_signed(v & ma, ashift, acount) : 255; all_a |= dst[3]; }
                }
                dst += out_ch;
            }
        }
        if (out_ch == 4 && all_a_check && all_a == 0)  // plain 32-bit files usually leave the fourth byte at zero: opaque
            for (size_t i = 3; i < (size_t)w * h * 4; i += 4) px[i] = 255;
    }
    extent[0] = w; extent[1] = h; *channels = out_ch;
    return px;
}

bool looks_like_tga(std::vector<uint8_t> const& f) {
    if (f.size() < 18) return false;
    int const cmap = f[1], type = f[2], bpp = f[16];
    if (cmap > 1) return false;
    if (cmap == 1) {
        if (type != 1 && type != 9) return false;
        int const pb = f[7];
        if (pb != 8 && pb != 15 && pb != 16 && pb != 24 && pb != 32) return false;
    } else if (type != 2 && type != 3 && type != 10 && type != 11) {
        return false;
    }
    if ((f[12] | (f[13] << 8)) < 1 || (f[14] | (f[15] << 8)) < 1) return false;
    if (cmap == 1 && bpp != 8 && bpp != 16) return false;
    return bpp == 8 || bpp == 15 || bpp == 16 || bpp == 24 || bpp == 32;
}

uint8_t* load_tga(std::vector<uint8_t> const& file, char const* path, int* extent, int* channels) {
    auto bad = [&](char const* why) { fail(std::string("Failed to load image ") + path + ": " + why); };
    int const id_len = file[0], indexed = file[1];
    int type = file[2];
    int const pal_start = file[3] | (file[4] << 8), pal_len = file[5] | (file[6] << 8), pal_bits = file[7];
    int const w = file[12] | (file[13] << 8), h = file[14] | (file[15] << 8), bpp = file[16];
    bool const top_down = (file[17] >> 5) & 1;
    bool const rle = type >= 8;
    if (rle) type -= 8;
    auto comp_of = [&](int bits, bool grey) {
        switch (bits) {
        case 8: return 1;
        case 15: case 16: return grey ? 2 : 3;
        case 24: return 3;
        case 32: return 4;
        default: return 0;
        }
    };
    int const comp = indexed ? comp_of(pal_bits, false) : comp_of(bpp, type == 3);
    if (comp == 0) bad("bad TGA pixel format");
    if (comp == 2) bad("16-bit grey + alpha TGA has no counterpart among the image channel orders");
    if (indexed && bpp != 8 && bpp != 16) bad("bad TGA index size");
    bool const rgb16 = !indexed ? (bpp == 15 || bpp == 16) : (pal_bits == 15 || pal_bits == 16);
    size_t pos = 18 + (size_t)id_len;
    auto need = [&](size_t n) {
        if (pos + n > file.size()) bad("truncated TGA");
    };
    auto put_pixel = [&](uint8_t const* raw, uint8_t* out) {  // one stored pixel (file order) -> output order
        if (rgb16) {
            int const v = raw[0] | (raw[1] << 8);
            out[0] = (uint8_t)((((v >> 10) & 31) * 255) / 31);
            out[1] = (uint8_t)((((v >> 5) & 31) * 255) / 31);
            out[2] = (uint8_t)(((v & 31) * 255) / 31);
        } else if (comp == 1) {
            out[0] = raw[0];
        } else {
            out[0] = raw[2]; out[1] = raw[1]; out[2] = raw[0];
            if (comp == 4) out[3] = raw[3];
        }
    };
    int const pal_entry = indexed ? (pal_bits + 7) / 8 : 0;
    std::vector<uint8_t> palette;
    if (indexed) {
        if (pal_len == 0) bad("bad TGA palette");
        need((size_t)pal_len * pal_entry);
        palette.assign(file.begin() + (long)pos, file.begin() + (long)(pos + (size_t)pal_len * pal_entry));
        pos += (size_t)pal_len * pal_entry;
    }
    (void)pal_start;  // (entries are addressed from zero, like the reference's reader does)
    if ((uint64_t)w * h * comp > kMaxImageBytes) bad("image too large");
    uint8_t* px = image_alloc((size_t)w * h * comp);
    int const stored = indexed ? bpp / 8 : (bpp + 7) / 8;
    uint8_t raw[4] = {};
    uint8_t cur[4] = {};
    int rle_count = 0;
    bool rle_repeat = false, have = false;
    try {
        for (size_t i = 0; i < (size_t)w * h; ++i) {
            bool read_next = true;
            if (rle) {
                if (rle_count == 0) {
                    need(1);
                    int const cmd = file[pos++];
                    rle_count = 1 + (cmd & 127);
                    rle_repeat = (cmd >> 7) != 0;
                    have = false;
                } else if (rle_repeat && have) {
                    read_next = false;
                }
            }
            if (read_next) {
                need((size_t)stored);
                if (indexed) {
                    int idx = stored == 1 ? file[pos] : (file[pos] | (file[pos + 1] << 8));
                    if (idx >= pal_len) idx = 0;
                    std::memcpy(raw, &palette[(size_t)idx * pal_entry], (size_t)pal_entry);
                } else {
                    std::memcpy(raw, &file[pos], (size_t)stored);
                }
                pos += (size_t)stored;
                put_pixel(raw, cur);
                have = true;
            }
            size_t const y = i / (size_t)w, x = i % (size_t)w;
            std::memcpy(px + ((top_down ? y : (size_t)h - 1 - y) * (size_t)w + x) * comp, cur, (size_t)comp);
            if (rle) --rle_count;
        }
    } catch (...) {
        image_free(px);
        throw;
    }
    extent[0] = w; extent[1] = h; *channels = comp;
    return px;
}

void chunk(std::vector<uint8_t>& out, char const* type, std::vector<uint8_t> const& data) {
    put_be32(out, (uint32_t)data.size());
    size_t const start = out.size();
    out.insert(out.end(), type, type + 4);
    out.insert(out.end(), data.begin(), data.end());
    put_be32(out, crc32(&out[start], out.size() - start));
}

// zlib stream for the PNG writer: deflate with the FIXED Huffman code (RFC 1951 3.2.6) over a greedy LZ77 match search
// (hash of three bytes, chains of up to 24 candidates, 32 KiB window) -- the same class of encoder as the stb_image_write
// the reference saves with; masks shrink from megabytes to kilobytes.  Off the hot path.
struct BitWriter {
    std::vector<uint8_t>& out;
    uint32_t acc = 0;
    int n = 0;
    void bits(uint32_t v, int count) {  // LSB first
        acc |= v << n;
        n += count;
        while (n >= 8) {
            out.push_back((uint8_t)(acc & 0xFF));
            acc >>= 8;
            n -= 8;
        }
    }
    void code(uint32_t c, int count) {  // Huffman codes go out MSB first
        uint32_t r = 0;
        for (int i = 0; i < count; ++i) r |= ((c >> i) & 1u) << (count - 1 - i);
        bits(r, count);
    }
    void flush() {
        if (n > 0) bits(0, 8 - n);
    }
};

std::vector<uint8_t> zlib_compress(std::vector<uint8_t> const& raw) {
    static uint16_t const lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static uint8_t const lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static uint16_t const dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static uint8_t const dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    std::vector<uint8_t> z;
    z.reserve(raw.size() / 4 + 64);
    z.push_back(0x78);
    z.push_back(0x5E);  // 32 KiB window, "fast" level; (0x785E) % 31 == 0
    BitWriter bw{z};
    bw.bits(1, 1);  // final block
    bw.bits(1, 2);  // fixed Huffman codes
    auto literal = [&](int v) {
        if (v < 144) bw.code(0x30 + (uint32_t)v, 8);
        else bw.code(0x190 + (uint32_t)(v - 144), 9);
    };
    auto length_distance = [&](int len, int dist) {
        int lc = 28;
        while (lbase[lc] > len) --lc;
        int const sym = 257 + lc;
        if (sym < 280) bw.code((uint32_t)(sym - 256), 7);
        else bw.code(0xC0 + (uint32_t)(sym - 280), 8);
        if (lext[lc]) bw.bits((uint32_t)(len - lbase[lc]), lext[lc]);
        int dc = 29;
        while (dbase[dc] > dist) --dc;
        bw.code((uint32_t)dc, 5);
        if (dext[dc]) bw.bits((uint32_t)(dist - dbase[dc]), dext[dc]);
    };
    constexpr int kHashBits = 15, kWindow = 32768, kMaxChain = 24;
    std::vector<int32_t> head((size_t)1 << kHashBits, -1), prev(raw.size(), -1);
    size_t const n = raw.size();
    auto hash3 = [&](size_t i) { return (uint32_t)(((raw[i] << 16) | (raw[i + 1] << 8) | raw[i + 2]) * 2654435761u) >> (32 - kHashBits); };
    auto insert = [&](size_t i) {
        if (i + 2 < n) {
            uint32_t const h = hash3(i);
            prev[i] = head[h];
            head[h] = (int32_t)i;
        }
    };
    size_t i = 0;
    while (i < n) {
        int best_len = 0, best_dist = 0;
        if (i + 2 < n) {
            int32_t cand = head[hash3(i)];
            int const max_len = (int)std::min<size_t>(258, n - i);
            for (int chain = 0; cand >= 0 && chain < kMaxChain && i - (size_t)cand <= (size_t)kWindow; ++chain, cand = prev[(size_t)cand]) {
                if (raw[(size_t)cand + (size_t)best_len] != raw[i + (size_t)best_len]) continue;  // cannot beat the best so far
                int len = 0;
                while (len < max_len && raw[(size_t)cand + (size_t)len] == raw[i + (size_t)len]) ++len;
                if (len > best_len) {
                    best_len = len;
                    best_dist = (int)(i - (size_t)cand);
                    if (len == max_len) break;
                }
            }
        }
        if (best_len >= 3) {
            length_distance(best_len, best_dist);
            for (int k = 0; k < best_len; ++k) insert(i + (size_t)k);
            i += (size_t)best_len;
        } else {
            literal(raw[i]);
            insert(i);
            ++i;
        }
    }
    bw.code(0, 7);  // end of block
    bw.flush();
    if (z.size() > raw.size() + raw.size() / 65535 * 5 + 16) {
        // incompressible data (the fixed code spends nine bits on bytes >= 144): stored blocks instead
        z.resize(2);
        size_t pos = 0;
        do {
            size_t const len = std::min<size_t>(65535, raw.size() - pos);
            z.push_back(pos + len == raw.size() ? 1 : 0);
            z.push_back((uint8_t)(len & 0xFF)); z.push_back((uint8_t)(len >> 8));
            z.push_back((uint8_t)(~len & 0xFF)); z.push_back((uint8_t)((~len >> 8) & 0xFF));
            z.insert(z.end(), raw.begin() + (long)pos, raw.begin() + (long)(pos + len));
            pos += len;
        } while (pos < raw.size());
    }
    put_be32(z, adler32(raw.data(), raw.size()));
    return z;
}

}  // namespace

uint8_t* load_image(char const* filepath, int* out_extent, int* out_channels) {
    std::vector<uint8_t> const file = read_file(filepath);
    static uint8_t const png_sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (file.size() > 8 && !std::memcmp(file.data(), png_sig, 8)) return load_png(file, filepath, out_extent, out_channels);
    if (file.size() > 2 && file[0] == 'P' && (file[1] == '5' || file[1] == '6')) return load_pnm(file, filepath, out_extent, out_channels);
    if (file.size() > 4 && file[0] == 0xFF && file[1] == 0xD8) return load_jpeg(file, filepath, out_extent, out_channels);
    if (file.size() > 26 && file[0] == 'B' && file[1] == 'M') return load_bmp(file, filepath, out_extent, out_channels);
    if (looks_like_tga(file)) return load_tga(file, filepath, out_extent, out_channels);  // (no signature: tested last)
    fail(std::string("Failed to load image ") + filepath + ": unsupported format (PNG, JPEG, BMP, TGA and binary PGM/PPM are supported)");
}

// reference image.cpp:25-35: mask / rgb / rgba only, rows written packed (the reference also ignores stride)
void save_image(dlimg_ImageView const& img, char const* filepath) {
    if (!(img.channels == CH_MASK || img.channels == CH_RGB || img.channels == CH_RGBA))
        fail("Unsupported channel order [" + std::to_string(img.channels) + "]");
    if (img.width <= 0 || img.height <= 0 || !img.pixels) fail(std::string("Failed to save image ") + filepath);
    int const comp = bytes_per_pixel(img.channels);
    size_t const row = (size_t)img.width * comp;
    // scanlines with the PNG filter (none / sub / up / average / Paeth) that leaves the smallest sum of magnitudes
    std::vector<uint8_t> raw;
    raw.reserve((row + 1) * (size_t)img.height);
    std::vector<uint8_t> cand[5];
    for (auto& c : cand) c.resize(row);
    std::vector<uint8_t> const zeros(row, 0);
    for (int y = 0; y < img.height; ++y) {
        uint8_t const* src = img.pixels + (size_t)y * row;
        uint8_t const* up = y ? src - row : zeros.data();
        for (size_t x = 0; x < row; ++x) {
            int const a = x >= (size_t)comp ? src[x - (size_t)comp] : 0, b = up[x], c = x >= (size_t)comp ? up[x - (size_t)comp] : 0;
            cand[0][x] = src[x];
            cand[1][x] = (uint8_t)(src[x] - a);
            cand[2][x] = (uint8_t)(src[x] - b);
            cand[3][x] = (uint8_t)(src[x] - ((a + b) >> 1));
            cand[4][x] = (uint8_t)(src[x] - paeth(a, b, c));
        }
        int best = 0;
        uint64_t best_cost = UINT64_MAX;
        for (int f = 0; f < 5; ++f) {
            uint64_t cost = 0;
            for (size_t x = 0; x < row; ++x) cost += (uint64_t)std::abs((int)(int8_t)cand[f][x]);
            if (cost < best_cost) {
                best_cost = cost;
                best = f;
            }
        }
        raw.push_back((uint8_t)best);
        raw.insert(raw.end(), cand[best].begin(), cand[best].end());
    }
    std::vector<uint8_t> const z = zlib_compress(raw);

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
