// png_host.cuh — host-side PNG ingest / egress behind hg_png_decode / hg_png_encode (include/hgwarp.h).
//
// The step either side of the hot path.  In a browser the reference gets its RGBA bytes from a
// canvas (drawImage + getImageData, H.js:1071-1076) and hands results back as a PNG data URL (toDataURL, H.js:480-483);
// its Node smoke test reads test/testImgLogoBlack.png and its golden output is test/transformedImage.png
// (test/nodeTest.js:5-13).  This file is that file I/O for a host without a canvas: PNG (ISO/IEC 15948) chunks, zlib
// inflate / deflate from the system zlib, the five scanline filters, Adam7 interlacing, and expansion of every 8- or 16-bit
// colour type (and 1/2/4-bit grey / palette) to the RGBA8 layout of ImageData.  Pixel values are what any conforming
// decoder produces (16-bit samples keep their high byte, tRNS keys and palette alpha are honoured); no colour
// management, no gamma — exactly what getImageData returns for an untagged image.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>
#include <zlib.h>

namespace hg_png_detail {

inline uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline void put32(std::vector<uint8_t> &v, uint32_t x)
{
    v.push_back((uint8_t)(x >> 24));
    v.push_back((uint8_t)(x >> 16));
    v.push_back((uint8_t)(x >> 8));
    v.push_back((uint8_t)x);
}

inline int paeth(int a, int b, int c)
{
    const int p = a + b - c, pa = p > a ? p - a : a - p, pb = p > b ? p - b : b - p, pc = p > c ? p - c : c - p;
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

struct Header {
    uint32_t w = 0, h = 0;
    int depth = 0, color = 0, interlace = 0;
};

// 0 ok, 1 malformed / unsupported
inline int parse(const uint8_t *png, size_t n, Header &hd, std::vector<uint8_t> &idat, std::vector<uint8_t> &plte,
                 std::vector<uint8_t> &trns)
{
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (n < 8 || memcmp(png, sig, 8) != 0) return 1;
    size_t pos = 8;
    bool have_ihdr = false, end = false;
    while (!end && pos + 12 <= n) {
        const uint32_t len = be32(png + pos);
        const uint8_t *type = png + pos + 4, *data = png + pos + 8;
        if ((size_t)len > n - pos - 12) return 1;
        if (crc32(crc32(0L, type, 4), data, len) != be32(data + len)) return 1;
        if (!memcmp(type, "IHDR", 4)) {
            if (len != 13) return 1;
            hd.w = be32(data);
            hd.h = be32(data + 4);
            hd.depth = data[8];
            hd.color = data[9];
            hd.interlace = data[12];
            if (data[10] != 0 || data[11] != 0) return 1;
            have_ihdr = true;
        } else if (!memcmp(type, "PLTE", 4)) {
            plte.assign(data, data + len);
        } else if (!memcmp(type, "tRNS", 4)) {
            trns.assign(data, data + len);
        } else if (!memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!memcmp(type, "IEND", 4)) {
            end = true;
        }
        pos += 12 + (size_t)len;
    }
    if (!have_ihdr || !end || hd.w == 0 || hd.h == 0 || hd.w > 65536 || hd.h > 65536) return 1;
    if (hd.interlace != 0 && hd.interlace != 1) return 1;
    const int d = hd.depth;
    switch (hd.color) {
        case 0: if (d != 1 && d != 2 && d != 4 && d != 8 && d != 16) return 1; break;
        case 3: if (d != 1 && d != 2 && d != 4 && d != 8) return 1; if (plte.empty() || plte.size() % 3) return 1; break;
        case 2: case 4: case 6: if (d != 8 && d != 16) return 1; break;
        default: return 1;
    }
    return 0;
}

// rgba: w*h*4 bytes.  0 ok, 1 malformed / unsupported
inline int decode(const uint8_t *png, size_t n, Header &hd, uint8_t *rgba)
{
    std::vector<uint8_t> idat, plte, trns;
    if (parse(png, n, hd, idat, plte, trns)) return 1;
    if (!rgba) return 0;  // header only
    static const int channels_of[7] = {1, 0, 3, 1, 2, 0, 4};
    const int ch = channels_of[hd.color], bpp_bits = ch * hd.depth;
    const size_t bpp = (size_t)(bpp_bits + 7) / 8;
    // the image as one pass, or the seven Adam7 passes: (x0, y0, dx, dy) of each reduced image
    static const int adam7[7][4] = {{0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2}};
    static const int whole[1][4] = {{0, 0, 1, 1}};
    const int (*passes)[4] = hd.interlace ? adam7 : whole;
    const int n_pass = hd.interlace ? 7 : 1;
    size_t total = 0;
    for (int p = 0; p < n_pass; ++p) {
        const size_t pw = (hd.w + passes[p][2] - 1 - passes[p][0]) / passes[p][2], ph = (hd.h + passes[p][3] - 1 - passes[p][1]) / passes[p][3];
        if (pw && ph) total += (((size_t)pw * bpp_bits + 7) / 8 + 1) * ph;
    }
    // deflate cannot expand beyond ~1032 : 1: a header that promises more than the IDAT bytes can hold is malformed —
    // reject it before allocating what it asks for (a 65536 x 65536 header in a 100-byte file)
    if (total / 1040 > idat.size() + 1) return 1;
    std::vector<uint8_t> raw(total);
    uLongf out_len = (uLongf)raw.size();
    if (uncompress(raw.data(), &out_len, idat.data(), (uLong)idat.size()) != Z_OK || out_len != raw.size()) return 1;
    const int maxv = (1 << (hd.depth < 8 ? hd.depth : 8)) - 1;
    size_t off = 0;
    for (int p = 0; p < n_pass; ++p) {
        const uint32_t pw = (uint32_t)((hd.w + passes[p][2] - 1 - passes[p][0]) / passes[p][2]);
        const uint32_t ph = (uint32_t)((hd.h + passes[p][3] - 1 - passes[p][1]) / passes[p][3]);
        if (!pw || !ph) continue;
        const size_t stride = ((size_t)pw * bpp_bits + 7) / 8;
        uint8_t *pass = raw.data() + off;
        off += (stride + 1) * ph;
        // undo the scanline filters in place
        std::vector<uint8_t> zero(stride, 0);
        for (uint32_t y = 0; y < ph; ++y) {
            uint8_t *row = pass + (stride + 1) * y + 1;
            const uint8_t *up = y ? row - (stride + 1) : zero.data();
            const int ft = row[-1];
            for (size_t i = 0; i < stride; ++i) {
                const int a = i >= bpp ? row[i - bpp] : 0, b = up[i], c = i >= bpp ? up[i - bpp] : 0;
                int add;
                switch (ft) {
                    case 0: add = 0; break;
                    case 1: add = a; break;
                    case 2: add = b; break;
                    case 3: add = (a + b) >> 1; break;
                    case 4: add = paeth(a, b, c); break;
                    default: return 1;
                }
                row[i] = (uint8_t)(row[i] + add);
            }
        }
        // expand to RGBA8 at the pass's pixel positions
        for (uint32_t y = 0; y < ph; ++y) {
            const uint8_t *row = pass + (stride + 1) * y + 1;
            for (uint32_t x = 0; x < pw; ++x) {
                uint8_t *o = rgba + ((size_t)(passes[p][1] + y * passes[p][3]) * hd.w + (passes[p][0] + x * passes[p][2])) * 4;
                auto sample = [&](int c, uint16_t &full) -> uint8_t {  // c-th channel of pixel x: 8-bit value (+ raw sample)
                    if (hd.depth == 16) {
                        const uint8_t *q = row + ((size_t)x * ch + c) * 2;
                        full = (uint16_t)((q[0] << 8) | q[1]);
                        return q[0];
                    }
                    if (hd.depth == 8) {
                        full = row[(size_t)x * ch + c];
                        return (uint8_t)full;
                    }
                    const size_t bit = (size_t)x * hd.depth;
                    const int v = (row[bit >> 3] >> (8 - hd.depth - (int)(bit & 7))) & maxv;
                    full = (uint16_t)v;
                    return (uint8_t)v;
                };
                uint16_t f0 = 0, f1 = 0, f2 = 0, f3 = 0;
                switch (hd.color) {
                    case 0: {
                        uint8_t g = sample(0, f0);
                        if (hd.depth < 8) g = (uint8_t)(g * 255 / maxv);
                        o[0] = o[1] = o[2] = g;
                        o[3] = (trns.size() >= 2 && f0 == (uint16_t)((trns[0] << 8) | trns[1])) ? 0 : 255;
                        break;
                    }
                    case 2: {
                        o[0] = sample(0, f0);
                        o[1] = sample(1, f1);
                        o[2] = sample(2, f2);
                        o[3] = (trns.size() >= 6 && f0 == (uint16_t)((trns[0] << 8) | trns[1]) && f1 == (uint16_t)((trns[2] << 8) | trns[3]) &&
                                f2 == (uint16_t)((trns[4] << 8) | trns[5])) ? 0 : 255;
                        break;
                    }
                    case 3: {
                        const size_t i = sample(0, f0);
                        if (3 * i + 2 >= plte.size()) return 1;
                        o[0] = plte[3 * i];
                        o[1] = plte[3 * i + 1];
                        o[2] = plte[3 * i + 2];
                        o[3] = i < trns.size() ? trns[i] : 255;
                        break;
                    }
                    case 4:
                        o[0] = o[1] = o[2] = sample(0, f0);
                        o[3] = sample(1, f1);
                        break;
                    default:
                        o[0] = sample(0, f0);
                        o[1] = sample(1, f1);
                        o[2] = sample(2, f2);
                        o[3] = sample(3, f3);
                        break;
                }
            }
        }
    }
    return 0;
}

inline void chunk(std::vector<uint8_t> &out, const char *type, const uint8_t *data, size_t len)
{
    put32(out, (uint32_t)len);
    const size_t at = out.size();
    out.insert(out.end(), type, type + 4);
    if (len) out.insert(out.end(), data, data + len);
    put32(out, (uint32_t)crc32(0L, out.data() + at, (uInt)(4 + len)));
}

// RGBA8 -> PNG (colour type 6, 8 bit, per row the better of filter None / Up by the sum-of-absolute-values heuristic)
inline int encode(const uint8_t *rgba, uint32_t w, uint32_t h, std::vector<uint8_t> &out)
{
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    out.assign(sig, sig + 8);
    uint8_t ihdr[13];
    ihdr[0] = (uint8_t)(w >> 24); ihdr[1] = (uint8_t)(w >> 16); ihdr[2] = (uint8_t)(w >> 8); ihdr[3] = (uint8_t)w;
    ihdr[4] = (uint8_t)(h >> 24); ihdr[5] = (uint8_t)(h >> 16); ihdr[6] = (uint8_t)(h >> 8); ihdr[7] = (uint8_t)h;
    ihdr[8] = 8; ihdr[9] = 6; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;
    chunk(out, "IHDR", ihdr, 13);
    const size_t stride = (size_t)w * 4;
    std::vector<uint8_t> raw((stride + 1) * h);
    for (uint32_t y = 0; y < h; ++y) {
        const uint8_t *row = rgba + stride * y, *up = y ? row - stride : nullptr;
        uint8_t *o = raw.data() + (stride + 1) * y;
        unsigned long s_none = 0, s_up = 0;
        for (size_t i = 0; i < stride; ++i) {
            const int8_t a = (int8_t)row[i], b = (int8_t)(uint8_t)(row[i] - (up ? up[i] : 0));
            s_none += (unsigned long)(a < 0 ? -a : a);
            s_up += (unsigned long)(b < 0 ? -b : b);
        }
        if (up && s_up < s_none) {
            o[0] = 2;
            for (size_t i = 0; i < stride; ++i) o[1 + i] = (uint8_t)(row[i] - up[i]);
        } else {
            o[0] = 0;
            memcpy(o + 1, row, stride);
        }
    }
    uLongf zlen = compressBound((uLong)raw.size());
    std::vector<uint8_t> z(zlen);
    if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return 1;
    // a chunk length is a 31-bit field: large images leave as several IDAT chunks
    const size_t piece = (size_t)1 << 30;
    size_t at = 0;
    do {
        const size_t len = (size_t)zlen - at < piece ? (size_t)zlen - at : piece;
        chunk(out, "IDAT", z.data() + at, len);
        at += len;
    } while (at < (size_t)zlen);
    chunk(out, "IEND", nullptr, 0);
    return 0;
}

}  // namespace hg_png_detail
