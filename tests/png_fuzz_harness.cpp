// Sanitizer harness for csrc/png_host.cuh (host-only code): built by tests/test_png_fuzz.py with
// g++ -fsanitize=address,undefined.  Reads base PNG files, mutates bytes INSIDE chunk payloads / lengths / types with a
// seeded generator, re-seals every chunk CRC (so the damage reaches the IHDR checks, zlib, the filters and the pixel
// expansion instead of dying at the first CRC), and decodes.  Any out-of-bounds access, overflow or leak aborts.
#include <cstdio>
#include <cstdlib>
#include <string>
#include "../homography.js_b200/csrc/png_host.cuh"

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd()
{
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return (uint32_t)(rng_state >> 16);
}

static void reseal(std::vector<uint8_t> &f)
{
    size_t pos = 8;
    while (pos + 12 <= f.size()) {
        const uint32_t len = hg_png_detail::be32(f.data() + pos);
        if ((size_t)len > f.size() - pos - 12) break;
        const uint32_t c = (uint32_t)crc32(crc32(0L, f.data() + pos + 4, 4), f.data() + pos + 8, len);
        uint8_t *q = f.data() + pos + 8 + len;
        q[0] = (uint8_t)(c >> 24); q[1] = (uint8_t)(c >> 16); q[2] = (uint8_t)(c >> 8); q[3] = (uint8_t)c;
        pos += 12 + (size_t)len;
    }
}

int main(int argc, char **argv)
{
    const int iters = argc > 1 ? atoi(argv[1]) : 1000;
    long ok = 0, bad = 0;
    for (int a = 2; a < argc; ++a) {
        FILE *fp = fopen(argv[a], "rb");
        if (!fp) return 2;
        std::vector<uint8_t> base;
        uint8_t buf[65536];
        size_t n;
        while ((n = fread(buf, 1, sizeof buf, fp)) > 0) base.insert(base.end(), buf, buf + n);
        fclose(fp);
        for (int it = 0; it < iters; ++it) {
            std::vector<uint8_t> f = base;
            const int n_mut = 1 + (int)(rnd() % 4);
            for (int k = 0; k < n_mut; ++k) {
                const uint32_t r = rnd() % 10;
                if (r < 4 && f.size() > 33) f[8 + rnd() % 25] = (uint8_t)rnd();                  // IHDR fields
                else if (r < 8) f[8 + rnd() % (f.size() - 8)] ^= (uint8_t)(1u << (rnd() % 8));   // a bit anywhere
                else if (r == 8 && f.size() > 64) f.resize(f.size() - 1 - rnd() % 32);           // truncation
                else f.insert(f.begin() + 8 + rnd() % (f.size() - 8), (uint8_t)rnd());           // shifted bytes
            }
            if (rnd() % 4) reseal(f);
            hg_png_detail::Header hd;
            if (hg_png_detail::decode(f.data(), f.size(), hd, nullptr)) { ++bad; continue; }
            if ((size_t)hd.w * hd.h > (size_t)1 << 24) { ++bad; continue; }   // what a caller's capacity check would refuse
            std::vector<uint8_t> rgba((size_t)hd.w * hd.h * 4);
            if (hg_png_detail::decode(f.data(), f.size(), hd, rgba.data())) ++bad;
            else ++ok;
        }
    }
    printf("decoded %ld, rejected %ld\n", ok, bad);
    return 0;
}
