// jpeg_host.cuh — host-side JPEG ingest (baseline and progressive) behind hg_jpeg_decode (include/hgwarp.h).
//
// The step before the hot path when there is no canvas: in a browser the reference gets the RGBA bytes
// of ANY image the browser can decode through drawImage + getImageData (H.js:1071-1076), and for JPEG files every major
// browser (and Pillow, the checker used by the tests) decodes with libjpeg-turbo's defaults.  This file restates that
// pipeline from the published algorithms (ITU-T T.81 and the Independent JPEG Group's reference arithmetic) so that the
// bytes are the ones getImageData returns:
//   * sequential AND progressive DCT (SOF0 / SOF1 / SOF2: spectral selection and successive approximation, end-of-band runs),
//     Huffman coding, 8-bit samples, 1 or 3 components, any number of scans, restart intervals, 8- and 16-bit quantisation
//     tables;
//   * dequantisation + the accurate integer inverse DCT ("islow": 13-bit constants, two passes, the zero-AC shortcuts give
//     the same numbers as the general path);
//   * chroma upsampling as libjpeg does by default: the triangle ("fancy") filters for 2:1 horizontal and 2:1 x 2:1 when the
//     component is wider than two samples, with the edge rows / columns replicated; pixel replication for other integer ratios;
//   * YCbCr -> RGB with the 16-bit fixed-point tables (1.40200, 0.34414, 0.71414, 1.77200), or RGB passed through when an
//     Adobe marker / the component ids say so; grey -> R = G = B; alpha = 255;
//   * the Exif Orientation tag (APP1) is applied, as browsers do when an image is drawn (CSS image-orientation: from-image is
//     the default): the returned width / height are those of the picture as shown.
// Not decoded (HG_ERR_UNSUPPORTED, never a wrong image): lossless and hierarchical modes, arithmetic coding, 12-bit samples,
// four-component (CMYK / YCCK) files, 1:2 vertical-only subsampling, fractional sampling ratios.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace hg_jpeg_detail {

enum { OK = 0, MALFORMED = 1, UNSUPPORTED = 2 };

struct Huff {
    // canonical Huffman table: for each code length 1..16 the smallest code, the largest code (or -1) and the index of its
    // first symbol
    int mincode[17], maxcode[18], valptr[17];
    uint8_t vals[256];
    bool present = false;
};

struct Component {
    int id = 0, hs = 1, vs = 1, tq = 0;
    int td = 0, ta = 0;          // Huffman table selectors of the current scan
    int blocks_w = 0, blocks_h = 0;  // plane size in 8x8 blocks (padded to whole MCUs)
    int real_w = 0, real_h = 0;  // downsampled_width / downsampled_height: ceil(W * hs / max_h), ceil(H * vs / max_v)
    int pred = 0;
    std::vector<uint8_t> plane;  // blocks_w*8 x blocks_h*8 samples
    std::vector<short> coefs;    // progressive mode: blocks_w*blocks_h blocks of 64 coefficients (natural order), undequantised
    bool dc_seen = false;        // progressive mode: a first DC scan covered this component
};

struct BitReader {
    const uint8_t *p, *end;
    uint32_t acc = 0;
    int bits = 0;
    bool hit_marker = false;  // a marker inside entropy-coded data: feed zero bits from here on (what libjpeg does)
    BitReader(const uint8_t *b, const uint8_t *e) : p(b), end(e) {}
    void fill()
    {
        while (bits <= 24) {
            int byte = 0;
            if (!hit_marker && p < end) {
                byte = *p;
                if (byte == 0xFF) {
                    if (p + 1 < end && p[1] == 0x00) {
                        p += 2;  // stuffed zero
                    } else {
                        hit_marker = true;  // leave p on the 0xFF
                        byte = 0;
                    }
                } else {
                    ++p;
                }
            } else {
                hit_marker = true;
            }
            acc |= (uint32_t)byte << (24 - bits);
            bits += 8;
        }
    }
    int get(int n)  // n <= 16
    {
        if (n == 0) return 0;
        if (bits < n) fill();
        const int v = (int)(acc >> (32 - n));
        acc <<= n;
        bits -= n;
        return v;
    }
    void align()
    {
        acc = 0;
        bits = 0;
        hit_marker = false;
    }
};

inline int build_huff(Huff &h, const uint8_t counts[16], const uint8_t *vals, int nvals)
{
    int code = 0, k = 0;
    for (int l = 1; l <= 16; ++l) {
        h.valptr[l] = k;
        h.mincode[l] = code;
        if (counts[l - 1]) {
            k += counts[l - 1];
            code += counts[l - 1];
            h.maxcode[l] = code - 1;
            if (code > (1 << l)) return MALFORMED;  // more codes than the length can hold
        } else {
            h.maxcode[l] = -1;
        }
        code <<= 1;
    }
    h.maxcode[17] = 0x7FFFFFFF;
    if (k != nvals || nvals > 256) return MALFORMED;
    memcpy(h.vals, vals, (size_t)nvals);
    h.present = true;
    return OK;
}

inline int decode_symbol(BitReader &br, const Huff &h)
{
    int code = 0;
    for (int l = 1; l <= 16; ++l) {
        code = (code << 1) | br.get(1);
        if (h.maxcode[l] >= 0 && code <= h.maxcode[l] && code >= h.mincode[l]) return h.vals[h.valptr[l] + code - h.mincode[l]];
    }
    return 0;  // corrupt data: libjpeg substitutes a zero symbol too
}

inline int extend(int v, int s) { return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }

static const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                    41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// clamp to a sample the way the inverse DCT's range-limit table does: the descaled value is taken modulo 1024 (10-bit
// index), centred on 128
inline uint8_t idct_limit(long long v)
{
    int x = (int)(v & 1023);
    if (x >= 512) x -= 1024;
    x += 128;
    return (uint8_t)(x < 0 ? 0 : (x > 255 ? 255 : x));
}

inline long long descale(long long x, int n) { return (x + (1LL << (n - 1))) >> n; }

// the accurate integer inverse DCT on one dequantised block (natural order) -> 8x8 samples with row pitch `pitch`
inline void idct_islow(const int *blk, uint8_t *out, size_t pitch)
{
    const int CB = 13, P1 = 2;
    const long long F0_298 = 2446, F0_390 = 3196, F0_541 = 4433, F0_765 = 6270, F0_899 = 7373, F1_175 = 9633, F1_501 = 12299,
                    F1_847 = 15137, F1_961 = 16069, F2_053 = 16819, F2_562 = 20995, F3_072 = 25172;
    long long ws[64];
    for (int c = 0; c < 8; ++c) {  // pass 1: columns
        const int *in = blk + c;
        long long *w = ws + c;
        if (!(in[8] | in[16] | in[24] | in[32] | in[40] | in[48] | in[56])) {
            const long long dc = (long long)in[0] * (1 << P1);
            for (int r = 0; r < 8; ++r) w[8 * r] = dc;
            continue;
        }
        long long z2 = in[16], z3 = in[48];
        long long z1 = (z2 + z3) * F0_541;
        long long tmp2 = z1 + z3 * (-F1_847), tmp3 = z1 + z2 * F0_765;
        z2 = in[0];
        z3 = in[32];
        long long tmp0 = (z2 + z3) * (1LL << CB), tmp1 = (z2 - z3) * (1LL << CB);
        const long long tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = in[56];
        tmp1 = in[40];
        tmp2 = in[24];
        tmp3 = in[8];
        z1 = tmp0 + tmp3;
        z2 = tmp1 + tmp2;
        z3 = tmp0 + tmp2;
        long long z4 = tmp1 + tmp3;
        const long long z5 = (z3 + z4) * F1_175;
        tmp0 *= F0_298; tmp1 *= F2_053; tmp2 *= F3_072; tmp3 *= F1_501;
        z1 *= -F0_899; z2 *= -F2_562; z3 *= -F1_961; z4 *= -F0_390;
        z3 += z5;
        z4 += z5;
        tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
        w[0] = descale(tmp10 + tmp3, CB - P1);  w[56] = descale(tmp10 - tmp3, CB - P1);
        w[8] = descale(tmp11 + tmp2, CB - P1);  w[48] = descale(tmp11 - tmp2, CB - P1);
        w[16] = descale(tmp12 + tmp1, CB - P1); w[40] = descale(tmp12 - tmp1, CB - P1);
        w[24] = descale(tmp13 + tmp0, CB - P1); w[32] = descale(tmp13 - tmp0, CB - P1);
    }
    for (int r = 0; r < 8; ++r) {  // pass 2: rows
        const long long *w = ws + 8 * r;
        uint8_t *o = out + pitch * r;
        if (!(w[1] | w[2] | w[3] | w[4] | w[5] | w[6] | w[7])) {
            const uint8_t dc = idct_limit(descale(w[0], P1 + 3));
            for (int c = 0; c < 8; ++c) o[c] = dc;
            continue;
        }
        long long z2 = w[2], z3 = w[6];
        long long z1 = (z2 + z3) * F0_541;
        long long tmp2 = z1 + z3 * (-F1_847), tmp3 = z1 + z2 * F0_765;
        long long tmp0 = (w[0] + w[4]) * (1LL << CB), tmp1 = (w[0] - w[4]) * (1LL << CB);
        const long long tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = w[7];
        tmp1 = w[5];
        tmp2 = w[3];
        tmp3 = w[1];
        z1 = tmp0 + tmp3;
        z2 = tmp1 + tmp2;
        z3 = tmp0 + tmp2;
        long long z4 = tmp1 + tmp3;
        const long long z5 = (z3 + z4) * F1_175;
        tmp0 *= F0_298; tmp1 *= F2_053; tmp2 *= F3_072; tmp3 *= F1_501;
        z1 *= -F0_899; z2 *= -F2_562; z3 *= -F1_961; z4 *= -F0_390;
        z3 += z5;
        z4 += z5;
        tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
        const int S = CB + P1 + 3;
        o[0] = idct_limit(descale(tmp10 + tmp3, S)); o[7] = idct_limit(descale(tmp10 - tmp3, S));
        o[1] = idct_limit(descale(tmp11 + tmp2, S)); o[6] = idct_limit(descale(tmp11 - tmp2, S));
        o[2] = idct_limit(descale(tmp12 + tmp1, S)); o[5] = idct_limit(descale(tmp12 - tmp1, S));
        o[3] = idct_limit(descale(tmp13 + tmp0, S)); o[4] = idct_limit(descale(tmp13 - tmp0, S));
    }
}

struct Decoder {
    const uint8_t *data;
    size_t n;
    int W = 0, H = 0, ncomp = 0, max_h = 1, max_v = 1;
    bool have_sof = false, jfif = false, adobe = false, progressive = false;
    int eobrun = 0;  // progressive AC scans: blocks still covered by the current end-of-band run
    int orientation = 1;  // Exif tag 0x0112 (1..8); browsers apply it when the image is drawn (image-orientation: from-image)
    int adobe_transform = 0;
    int restart_interval = 0;
    // the bytes are untrusted: work is bounded by what they can pay for (see parse_headers_and_scans)
    int n_scans = 0;
    unsigned long long block_visits = 0, visit_budget = 0;
    uint16_t qt[4][64];
    bool qt_present[4] = {false, false, false, false};
    Huff dc[4], ac[4];
    Component comp[3];
    bool comp_done[3] = {false, false, false};

    static int be16(const uint8_t *p) { return (p[0] << 8) | p[1]; }

    // one 8x8 block: Huffman-decoded, dequantised, natural order
    int decode_block(BitReader &br, Component &c, int (&coef)[64])
    {
        const Huff &hd = dc[c.td], &ha = ac[c.ta];
        memset(coef, 0, sizeof coef);
        const uint16_t *q = qt[c.tq];
        int s = decode_symbol(br, hd);
        if (s > 15) return MALFORMED;
        int diff = s ? extend(br.get(s), s) : 0;
        c.pred += diff;
        coef[0] = (int)(short)c.pred * q[0];
        for (int k = 1; k < 64;) {
            const int rs = decode_symbol(br, ha), r = rs >> 4;
            s = rs & 15;
            if (s) {
                k += r;
                const int v = extend(br.get(s), s);
                // a run past the end of the block (corrupt data) lands on the last coefficient, as in libjpeg, whose
                // order table is padded with 63s
                const int kk = k > 63 ? 63 : k;
                coef[kZigzag[kk]] = (int)(short)v * q[kk];
                ++k;
            } else {
                if (r != 15) break;  // EOB
                k += 16;
            }
        }
        return OK;
    }

    // ---- progressive mode (T.81 annex G): one call per block and scan; `b` = the block's 64 stored coefficients
    void prog_dc_first(BitReader &br, Component &c, short *b, int al)
    {
        const int s = decode_symbol(br, dc[c.td]);
        int diff = 0;
        if (s && s <= 15) diff = extend(br.get(s), s);
        c.pred += diff;
        b[0] = (short)(c.pred * (1 << al));
    }
    static void prog_dc_refine(BitReader &br, short *b, int al)
    {
        if (br.get(1)) b[0] = (short)(b[0] | (1 << al));
    }
    void prog_ac_first(BitReader &br, Component &c, short *b, int ss, int se, int al)
    {
        if (eobrun > 0) {
            --eobrun;
            return;
        }
        for (int k = ss; k <= se; ++k) {
            const int rs = decode_symbol(br, ac[c.ta]), r = rs >> 4, s = rs & 15;
            if (s) {
                k += r;
                const int v = extend(br.get(s), s);
                b[kZigzag[k > 63 ? 63 : k]] = (short)(v * (1 << al));
            } else if (r == 15) {
                k += 15;
            } else {
                eobrun = 1 << r;
                if (r) eobrun += br.get(r);
                --eobrun;
                break;
            }
        }
    }
    void prog_ac_refine(BitReader &br, Component &c, short *b, int ss, int se, int al)
    {
        const int p1 = 1 << al, m1 = -(1 << al);
        auto correct = [&](short &coef) {  // one correction bit for an already-nonzero coefficient
            if (br.get(1) && (coef & p1) == 0) coef = (short)(coef >= 0 ? coef + p1 : coef + m1);
        };
        int k = ss;
        if (eobrun == 0) {
            for (; k <= se; ++k) {
                const int rs = decode_symbol(br, ac[c.ta]);
                int r = rs >> 4, s = rs & 15;
                if (s) {
                    s = br.get(1) ? p1 : m1;  // a new coefficient is always +-1 at this bit position
                } else if (r != 15) {
                    eobrun = 1 << r;
                    if (r) eobrun += br.get(r);
                    break;  // the rest of the band is handled by the end-of-band branch below
                }
                // skip r still-zero coefficients, correcting the nonzero ones passed on the way
                do {
                    short &coef = b[kZigzag[k]];
                    if (coef != 0) correct(coef);
                    else if (--r < 0) break;
                    ++k;
                } while (k <= se);
                if (s && k <= 63) b[kZigzag[k]] = (short)s;
            }
        }
        if (eobrun > 0) {
            for (; k <= se; ++k) {
                short &coef = b[kZigzag[k]];
                if (coef != 0) correct(coef);
            }
            --eobrun;
        }
    }

    static constexpr int kMaxScans = 100;
    static constexpr unsigned long long kMaxPixels = 1ull << 28;  // 16384 x 16384: planes + coefficients stay below ~2.5 GB

    int scan(size_t &pos)
    {
        if (pos + 2 > n) return MALFORMED;
        const int len = be16(data + pos);
        if (len < 6 || pos + (size_t)len > n) return MALFORMED;
        const int ns = data[pos + 2];
        if (ns < 1 || ns > ncomp || len != 6 + 2 * ns) return MALFORMED;
        Component *sc[3];
        for (int i = 0; i < ns; ++i) {
            const int id = data[pos + 3 + 2 * i], tt = data[pos + 4 + 2 * i];
            Component *c = nullptr;
            for (int j = 0; j < ncomp; ++j)
                if (comp[j].id == id) c = &comp[j];
            if (!c) return MALFORMED;
            for (int j = 0; j < i; ++j)
                if (sc[j] == c) return MALFORMED;
            c->td = tt >> 4;
            c->ta = tt & 15;
            if (c->td > 3 || c->ta > 3 || !qt_present[c->tq]) return MALFORMED;
            if (!progressive) {
                if (!dc[c->td].present || !ac[c->ta].present) return MALFORMED;
                if (comp_done[c - comp]) return MALFORMED;  // sequential mode: one scan per component
            }
            sc[i] = c;
        }
        const int ss = data[pos + 3 + 2 * ns], se = data[pos + 4 + 2 * ns], ahal = data[pos + 5 + 2 * ns];
        const int ah = ahal >> 4, al = ahal & 15;
        if (!progressive) {
            if (ss != 0 || se != 63 || ahal != 0) return MALFORMED;  // sequential: full spectrum, no successive approximation
        } else {
            if (ss > se || se > 63 || al > 13 || ah > 13) return MALFORMED;
            if (ss == 0 ? se != 0 : ns != 1) return MALFORMED;      // DC scans carry only DC; AC scans one component
            for (int i = 0; i < ns; ++i) {
                if (ss == 0 && ah == 0 && !dc[sc[i]->td].present) return MALFORMED;
                if (ss != 0 && !ac[sc[i]->ta].present) return MALFORMED;
            }
        }
        // a scan walks every block of its components even when its entropy-coded segment is empty (the bit reader feeds
        // zeros after a marker, like libjpeg): bound the scans and the blocks they may walk by the size of the file, so a
        // few kilobytes cannot buy minutes of decoding (encoders emit about 10 scans; libjpeg-turbo's fuzzers stop at 500)
        if (++n_scans > kMaxScans) return MALFORMED;
        {
            unsigned long long blocks = 0;
            for (int i = 0; i < ns; ++i) blocks += (unsigned long long)sc[i]->blocks_w * (unsigned long long)sc[i]->blocks_h;
            block_visits += blocks;
            if (block_visits > visit_budget) return MALFORMED;
        }
        eobrun = 0;
        pos += (size_t)len;
        BitReader br(data + pos, data + n);
        for (int i = 0; i < ns; ++i) sc[i]->pred = 0;
        int mcus_x, mcus_y;
        if (ns == 1) {  // non-interleaved: the component's own block raster
            mcus_x = (sc[0]->real_w + 7) / 8;
            mcus_y = (sc[0]->real_h + 7) / 8;
        } else {
            mcus_x = (W + 8 * max_h - 1) / (8 * max_h);
            mcus_y = (H + 8 * max_v - 1) / (8 * max_v);
        }
        int until_restart = restart_interval, next_rst = 0;
        int blk[64];
        for (int my = 0; my < mcus_y; ++my) {
            for (int mx = 0; mx < mcus_x; ++mx) {
                if (restart_interval && until_restart == 0) {
                    // byte-align, expect RSTn, reset the predictors
                    br.align();
                    const uint8_t *p = br.p;
                    while (p + 1 < data + n && !(p[0] == 0xFF && p[1] != 0x00 && p[1] != 0xFF)) ++p;  // resynchronise
                    if (p + 1 < data + n && p[1] == 0xD0 + next_rst) p += 2;
                    else if (p + 1 < data + n && p[1] >= 0xD0 && p[1] <= 0xD7) p += 2;  // out-of-sequence restart: accept
                    br.p = p;
                    next_rst = (next_rst + 1) & 7;
                    for (int i = 0; i < ns; ++i) sc[i]->pred = 0;
                    eobrun = 0;
                    until_restart = restart_interval;
                }
                for (int i = 0; i < ns; ++i) {
                    Component &c = *sc[i];
                    const int bh = ns == 1 ? 1 : c.hs, bv = ns == 1 ? 1 : c.vs;
                    for (int by = 0; by < bv; ++by)
                        for (int bx = 0; bx < bh; ++bx) {
                            const int X = mx * bh + bx, Y = my * bv + by;
                            if (progressive) {
                                short dummy[64] = {0};
                                short *b = (X < c.blocks_w && Y < c.blocks_h) ? c.coefs.data() + ((size_t)Y * c.blocks_w + X) * 64 : dummy;
                                if (ss == 0) {
                                    if (ah == 0) prog_dc_first(br, c, b, al);
                                    else prog_dc_refine(br, b, al);
                                } else {
                                    if (ah == 0) prog_ac_first(br, c, b, ss, se, al);
                                    else prog_ac_refine(br, c, b, ss, se, al);
                                }
                                continue;
                            }
                            const int r = decode_block(br, c, blk);
                            if (r) return r;
                            if (X < c.blocks_w && Y < c.blocks_h)
                                idct_islow(blk, c.plane.data() + ((size_t)Y * 8 * c.blocks_w + X) * 8, (size_t)c.blocks_w * 8);
                        }
                }
                if (restart_interval) --until_restart;
            }
        }
        for (int i = 0; i < ns; ++i) {
            if (!progressive) comp_done[sc[i] - comp] = true;
            else if (ss == 0 && ah == 0) sc[i]->dc_seen = true;
        }
        // the scan's entropy-coded segment ends at the next marker that is not a restart
        const uint8_t *p = br.p;
        while (p + 1 < data + n && !(p[0] == 0xFF && p[1] != 0x00 && p[1] != 0xFF && !(p[1] >= 0xD0 && p[1] <= 0xD7))) ++p;
        pos = (size_t)(p - data);
        return OK;
    }

    // APP1 "Exif\0\0" + TIFF header + IFD0: the Orientation entry (tag 0x0112, SHORT).  Anything odd leaves orientation 1.
    void parse_exif_orientation(const uint8_t *p, int body)
    {
        if (body < 14 || memcmp(p, "Exif\0\0", 6) != 0) return;
        const uint8_t *t = p + 6;
        const int tl = body - 6;
        const bool le = t[0] == 'I' && t[1] == 'I', be = t[0] == 'M' && t[1] == 'M';
        if (!le && !be) return;
        auto u16 = [&](int o) { return le ? (t[o] | (t[o + 1] << 8)) : ((t[o] << 8) | t[o + 1]); };
        auto u32 = [&](int o) {
            return le ? ((uint32_t)t[o] | ((uint32_t)t[o + 1] << 8) | ((uint32_t)t[o + 2] << 16) | ((uint32_t)t[o + 3] << 24))
                      : (((uint32_t)t[o] << 24) | ((uint32_t)t[o + 1] << 16) | ((uint32_t)t[o + 2] << 8) | (uint32_t)t[o + 3]);
        };
        if (u16(2) != 42) return;
        const uint32_t ifd = u32(4);
        if (ifd > (uint32_t)tl || (uint32_t)tl - ifd < 2) return;
        const int count = u16((int)ifd);
        for (int i = 0; i < count; ++i) {
            const uint32_t e = ifd + 2 + 12u * (uint32_t)i;
            if (e + 12 > (uint32_t)tl) return;
            if (u16((int)e) == 0x0112) {
                if (u16((int)e + 2) == 3 && u32((int)e + 4) == 1) {
                    const int v = u16((int)e + 8);
                    if (v >= 1 && v <= 8) orientation = v;
                }
                return;
            }
        }
    }

    int parse_headers_and_scans(bool header_only)
    {
        if (n < 4 || data[0] != 0xFF || data[1] != 0xD8) return MALFORMED;
        size_t pos = 2;
        memset(qt, 0, sizeof qt);
        for (;;) {
            // next marker (fill bytes 0xFF may precede it)
            while (pos < n && data[pos] != 0xFF) ++pos;
            while (pos < n && data[pos] == 0xFF) ++pos;
            if (pos >= n) break;  // no EOI: a truncated file is accepted if every component was decoded (checked below)
            const int m = data[pos++];
            if (m == 0xD9) break;                        // EOI
            if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;  // TEM / stray RSTn: no payload
            if (pos + 2 > n) return MALFORMED;
            const int len = be16(data + pos);
            if (len < 2 || pos + (size_t)len > n) return MALFORMED;
            const uint8_t *p = data + pos + 2;
            const int body = len - 2;
            if (m == 0xC0 || m == 0xC1 || m == 0xC2) {
                progressive = (m == 0xC2);
                if (have_sof || body < 6) return MALFORMED;
                if (p[0] != 8) return UNSUPPORTED;
                H = be16(p + 1);
                W = be16(p + 3);
                ncomp = p[5];
                if (W == 0 || H == 0) return MALFORMED;  // (a zero height means DNL: not supported by libjpeg either)
                if (ncomp == 4) return UNSUPPORTED;
                if ((ncomp != 1 && ncomp != 3) || body != 6 + 3 * ncomp) return MALFORMED;
                for (int i = 0; i < ncomp; ++i) {
                    comp[i].id = p[6 + 3 * i];
                    comp[i].hs = p[7 + 3 * i] >> 4;
                    comp[i].vs = p[7 + 3 * i] & 15;
                    comp[i].tq = p[8 + 3 * i];
                    if (comp[i].hs < 1 || comp[i].hs > 4 || comp[i].vs < 1 || comp[i].vs > 4 || comp[i].tq > 3) return MALFORMED;
                    if (comp[i].hs > max_h) max_h = comp[i].hs;
                    if (comp[i].vs > max_v) max_v = comp[i].vs;
                }
                if (ncomp == 1) {  // a single component is never subsampled: its factors only scale the MCU
                    comp[0].hs = comp[0].vs = 1;
                    max_h = max_v = 1;
                }
                have_sof = true;
                if ((unsigned long long)W * (unsigned long long)H > kMaxPixels) return UNSUPPORTED;
                if (header_only) {   // keep reading markers up to the first scan: an Exif segment may follow the frame header
                    pos += (size_t)len;
                    continue;
                }
                // a block costs at least two bits per component: a frame far larger than the bytes that follow could
                // encode is refused before its planes are allocated
                if ((unsigned long long)W * (unsigned long long)H / 4096ull > (unsigned long long)n) return MALFORMED;
                // every scan of an honest file pays for the blocks it walks with entropy-coded bytes, or skips them in long
                // end-of-band runs: 256 block visits per file byte (+ a floor for tiny files) is far beyond any encoder
                visit_budget = 256ull * (unsigned long long)n + (1ull << 20);
                const int mcus_x = (W + 8 * max_h - 1) / (8 * max_h), mcus_y = (H + 8 * max_v - 1) / (8 * max_v);
                for (int i = 0; i < ncomp; ++i) {
                    Component &c = comp[i];
                    if (max_h % c.hs != 0 || max_v % c.vs != 0) return UNSUPPORTED;  // fractional ratios
                    c.blocks_w = mcus_x * c.hs;
                    c.blocks_h = mcus_y * c.vs;
                    c.real_w = (W * c.hs + max_h - 1) / max_h;
                    c.real_h = (H * c.vs + max_v - 1) / max_v;
                    c.plane.assign((size_t)c.blocks_w * 8 * c.blocks_h * 8, 128);
                    if (progressive) c.coefs.assign((size_t)c.blocks_w * c.blocks_h * 64, 0);
                }
            } else if (m == 0xC3 || (m >= 0xC5 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC)) {
                return UNSUPPORTED;  // lossless, differential, arithmetic-coded frames
            } else if (m == 0xCC) {
                return UNSUPPORTED;  // arithmetic conditioning
            } else if (m == 0xC4) {  // DHT
                int off = 0;
                while (off < body) {
                    if (off + 17 > body) return MALFORMED;
                    const int tc = p[off] >> 4, th = p[off] & 15;
                    if (tc > 1 || th > 3) return MALFORMED;
                    int nv = 0;
                    for (int i = 0; i < 16; ++i) nv += p[off + 1 + i];
                    if (off + 17 + nv > body) return MALFORMED;
                    const int r = build_huff(tc ? ac[th] : dc[th], p + off + 1, p + off + 17, nv);
                    if (r) return r;
                    off += 17 + nv;
                }
            } else if (m == 0xDB) {  // DQT
                int off = 0;
                while (off < body) {
                    const int pq = p[off] >> 4, tq = p[off] & 15;
                    if (pq > 1 || tq > 3) return MALFORMED;
                    const int bytes = pq ? 128 : 64;
                    if (off + 1 + bytes > body) return MALFORMED;
                    for (int k = 0; k < 64; ++k) qt[tq][k] = pq ? (uint16_t)be16(p + off + 1 + 2 * k) : p[off + 1 + k];
                    qt_present[tq] = true;
                    off += 1 + bytes;
                }
            } else if (m == 0xDD) {  // DRI
                if (body != 2) return MALFORMED;
                restart_interval = be16(p);
            } else if (m == 0xE0) {
                if (body >= 5 && !memcmp(p, "JFIF\0", 5)) jfif = true;
            } else if (m == 0xE1) {
                parse_exif_orientation(p, body);
            } else if (m == 0xEE) {
                if (body >= 12 && !memcmp(p, "Adobe", 5)) {
                    adobe = true;
                    adobe_transform = p[11];
                }
            } else if (m == 0xDA) {  // SOS
                if (!have_sof) return MALFORMED;
                if (header_only) return OK;  // size and orientation are known once the first scan begins
                const int r = scan(pos);
                if (r) return r;
                continue;  // pos already sits on the next marker
            }
            pos += (size_t)len;
        }
        if (!have_sof) return MALFORMED;
        if (header_only) return OK;
        if (progressive) {
            // every scan has refined the stored coefficients: dequantise and transform each block once.  (libjpeg's
            // inter-block smoothing only acts while low-frequency coefficients are still missing precision, i.e. on files
            // whose scans have not all arrived.)
            for (int i = 0; i < ncomp; ++i) {
                Component &c = comp[i];
                if (!c.dc_seen || !qt_present[c.tq]) return MALFORMED;
                int blk[64];
                for (int Y = 0; Y < c.blocks_h; ++Y)
                    for (int X = 0; X < c.blocks_w; ++X) {
                        const short *b = c.coefs.data() + ((size_t)Y * c.blocks_w + X) * 64;
                        for (int k = 0; k < 64; ++k) blk[kZigzag[k]] = (int)b[kZigzag[k]] * qt[c.tq][k];
                        idct_islow(blk, c.plane.data() + ((size_t)Y * 8 * c.blocks_w + X) * 8, (size_t)c.blocks_w * 8);
                    }
            }
            return OK;
        }
        for (int i = 0; i < ncomp; ++i)
            if (!comp_done[i]) return MALFORMED;
        return OK;
    }

    // one full-resolution row of component c (W samples) into `row`, y = output row
    void upsampled_row(const Component &c, int y, uint8_t *row) const
    {
        const size_t pitch = (size_t)c.blocks_w * 8;
        const int hx = max_h / c.hs, vx = max_v / c.vs;
        const bool fancy = c.real_w > 2;
        auto src_row = [&](int r) { return c.plane.data() + pitch * (size_t)(r < 0 ? 0 : (r >= c.real_h ? c.real_h - 1 : r)); };
        if (hx == 1 && vx == 1) {
            memcpy(row, src_row(y), (size_t)W);
        } else if (hx == 2 && vx == 1 && fancy) {  // h2v1 triangle filter
            const uint8_t *in = src_row(y);
            const int n_in = c.real_w;
            for (int x = 0; x < W; ++x) {
                const int i = x >> 1;
                int v;
                if (x & 1) v = (i == n_in - 1) ? in[i] : (3 * in[i] + in[i + 1] + 2) >> 2;
                else v = (i == 0) ? in[0] : (3 * in[i] + in[i - 1] + 1) >> 2;
                row[x] = (uint8_t)v;
            }
        } else if (hx == 2 && vx == 2 && fancy) {  // h2v2 triangle filter: 3/4 nearer row + 1/4 further row, then columns
            const int r = y >> 1;
            const uint8_t *in0 = src_row(r), *in1 = src_row((y & 1) ? r + 1 : r - 1);
            const int n_in = c.real_w;
            for (int x = 0; x < W; ++x) {
                const int i = x >> 1;
                const int cur = 3 * in0[i] + in1[i];
                int v;
                if (x & 1) v = (i == n_in - 1) ? (cur * 4 + 7) >> 4 : (cur * 3 + (3 * in0[i + 1] + in1[i + 1]) + 7) >> 4;
                else v = (i == 0) ? (cur * 4 + 8) >> 4 : (cur * 3 + (3 * in0[i - 1] + in1[i - 1]) + 8) >> 4;
                row[x] = (uint8_t)v;
            }
        } else {  // integer replication (also 2:1 ratios of components at most two samples wide)
            const uint8_t *in = src_row(y / vx);
            for (int x = 0; x < W; ++x) row[x] = in[x / hx];
        }
    }

    int output(uint8_t *rgba) const
    {
        for (int i = 0; i < ncomp; ++i) {
            const int hx = max_h / comp[i].hs, vx = max_v / comp[i].vs;
            if (hx == 1 && vx == 2) return UNSUPPORTED;  // 1:2 vertical-only: libjpeg's h1v2 filter is not restated here
        }
        // which colour space the three components are in (libjpeg's default_decompress_parms)
        bool ycc = true;
        if (ncomp == 3) {
            if (jfif) ycc = true;
            else if (adobe) ycc = adobe_transform != 0;
            else ycc = !(comp[0].id == 'R' && comp[1].id == 'G' && comp[2].id == 'B');
        }
        std::vector<uint8_t> r0((size_t)W), r1((size_t)W), r2((size_t)W);
        for (int y = 0; y < H; ++y) {
            uint8_t *o = rgba + (size_t)y * W * 4;
            upsampled_row(comp[0], y, r0.data());
            if (ncomp == 1) {
                for (int x = 0; x < W; ++x) {
                    o[4 * x] = o[4 * x + 1] = o[4 * x + 2] = r0[x];
                    o[4 * x + 3] = 255;
                }
                continue;
            }
            upsampled_row(comp[1], y, r1.data());
            upsampled_row(comp[2], y, r2.data());
            for (int x = 0; x < W; ++x) {
                int R, G, B;
                if (ycc) {
                    const int Y = r0[x], cb = r1[x] - 128, cr = r2[x] - 128;
                    // 16-bit fixed point: FIX(1.40200) = 91881, FIX(1.77200) = 116130, FIX(0.71414) = 46802, FIX(0.34414) = 22554
                    R = Y + (int)((91881LL * cr + 32768) >> 16);
                    B = Y + (int)((116130LL * cb + 32768) >> 16);
                    G = Y + (int)(((-22554LL * cb + 32768) + (-46802LL * cr)) >> 16);
                } else {
                    R = r0[x]; G = r1[x]; B = r2[x];
                }
                o[4 * x] = (uint8_t)(R < 0 ? 0 : (R > 255 ? 255 : R));
                o[4 * x + 1] = (uint8_t)(G < 0 ? 0 : (G > 255 ? 255 : G));
                o[4 * x + 2] = (uint8_t)(B < 0 ? 0 : (B > 255 ? 255 : B));
                o[4 * x + 3] = 255;
            }
        }
        return OK;
    }
};

// rgba == nullptr: only the size.  0 ok, 1 malformed, 2 unsupported
inline int decode(const uint8_t *jpg, size_t n, int &w, int &h, uint8_t *rgba)
{
    Decoder d;
    d.data = jpg;
    d.n = n;
    const int r = d.parse_headers_and_scans(rgba == nullptr);
    const bool turned = d.orientation >= 5;  // orientations 5..8 exchange the axes
    w = turned ? d.H : d.W;
    h = turned ? d.W : d.H;
    if (r || !rgba) return r;
    if (d.orientation == 1) return d.output(rgba);
    // decode upright-as-stored, then lay the pixels out the way the Exif orientation says the picture is to be shown
    std::vector<uint8_t> tmp((size_t)d.W * d.H * 4);
    const int r2 = d.output(tmp.data());
    if (r2) return r2;
    const int W = d.W, H = d.H;
    const uint32_t *in = reinterpret_cast<const uint32_t *>(tmp.data());
    for (int oy = 0; oy < h; ++oy)
        for (int ox = 0; ox < w; ++ox) {
            int x, y;
            switch (d.orientation) {
                case 2: x = W - 1 - ox; y = oy; break;           // mirrored left-right
                case 3: x = W - 1 - ox; y = H - 1 - oy; break;   // rotated 180
                case 4: x = ox; y = H - 1 - oy; break;           // mirrored top-bottom
                case 5: x = oy; y = ox; break;                   // transposed
                case 6: x = oy; y = H - 1 - ox; break;           // to be shown rotated 90 clockwise
                case 7: x = W - 1 - oy; y = H - 1 - ox; break;   // transversed
                default: x = W - 1 - oy; y = ox; break;          // 8: to be shown rotated 90 counter-clockwise
            }
            memcpy(rgba + ((size_t)oy * w + ox) * 4, &in[(size_t)y * W + x], 4);
        }
    return OK;
}

}  // namespace hg_jpeg_detail
