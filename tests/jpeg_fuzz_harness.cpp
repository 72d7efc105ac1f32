// Sanitizer harness for csrc/jpeg_host.cuh (host-only code): built by tests/test_jpeg_io.py with
// g++ -fsanitize=address,undefined.  Reads base JPEG files, mutates bytes (headers, tables, entropy-coded data), truncates and
// shifts them with a seeded generator, and decodes.  Any out-of-bounds access, overflow or leak aborts.
#include <cstdio>
#include <cstdlib>
#include "../homography.js_b200/csrc/jpeg_host.cuh"

static uint64_t rng_state = 0xD1B54A32D192ED03ull;
static uint32_t rnd()
{
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return (uint32_t)(rng_state >> 16);
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
                const size_t hdr = f.size() < 700 ? f.size() : 700;   // markers and tables live at the front
                if (r < 4) f[2 + rnd() % (hdr - 2)] = (uint8_t)rnd();
                else if (r < 8) f[2 + rnd() % (f.size() - 2)] ^= (uint8_t)(1u << (rnd() % 8));
                else if (r == 8 && f.size() > 64) f.resize(f.size() - 1 - rnd() % 48);
                else f.insert(f.begin() + 2 + rnd() % (f.size() - 2), (uint8_t)rnd());
            }
            int w = 0, h = 0;
            if (hg_jpeg_detail::decode(f.data(), f.size(), w, h, nullptr)) { ++bad; continue; }
            if ((size_t)w * h > (size_t)1 << 22) { ++bad; continue; }   // what a caller's capacity check would refuse
            std::vector<uint8_t> rgba((size_t)w * h * 4);
            if (hg_jpeg_detail::decode(f.data(), f.size(), w, h, rgba.data())) ++bad;
            else ++ok;
        }
    }
    printf("decoded %ld, rejected %ld\n", ok, bad);
    return 0;
}
