// Fuzz harness for te_counter_b200/csrc/fast_inflate.h under AddressSanitizer / UBSan: exact-size heap
// buffers, so any read outside the input or write outside the output aborts.  Every accepted
// stream must decode to what zlib decodes.  Build and run:
//   g++ -O1 -g -std=c++17 -fsanitize=address,undefined tools/inflate_fuzz.cpp -lz -o /tmp/inflate_fuzz && /tmp/inflate_fuzz 20000
#include "../te_counter_b200/csrc/fast_inflate.h"

#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

static std::vector<uint8_t> deflate_raw(const std::vector<uint8_t> &d, int level, int strategy) {
    z_stream zs{};
    deflateInit2(&zs, level, Z_DEFLATED, -15, 8, strategy);
    std::vector<uint8_t> out(deflateBound(&zs, d.size()) + 16);
    zs.next_in = const_cast<Bytef *>(d.data());
    zs.avail_in = uInt(d.size());
    zs.next_out = out.data();
    zs.avail_out = uInt(out.size());
    deflate(&zs, Z_FINISH);
    out.resize(zs.total_out);
    deflateEnd(&zs);
    return out;
}

static bool zlib_inflate(const uint8_t *in, size_t n_in, uint8_t *out, size_t n_out) {
    z_stream zs{};
    inflateInit2(&zs, -15);
    zs.next_in = const_cast<Bytef *>(in);
    zs.avail_in = uInt(n_in);
    zs.next_out = out;
    zs.avail_out = uInt(n_out);
    int rc = inflate(&zs, Z_FINISH);
    bool ok = rc == Z_STREAM_END && zs.avail_out == 0;
    inflateEnd(&zs);
    return ok;
}

int main(int argc, char **argv) {
    long iters = argc > 1 ? atol(argv[1]) : 5000;
    std::mt19937_64 rng(12345);
    fast_inflate::Inflater inf;
    long accepted = 0, clean_ok = 0, clean = 0;
    for (long it = 0; it < iters; it++) {
        size_t n = size_t(rng() % (it % 7 == 0 ? 65537 : 3000));
        std::vector<uint8_t> d(n);
        int alphabet = 1 + int(rng() % 200);
        for (size_t i = 0; i < n; i++) d[i] = uint8_t(rng() % alphabet);
        if (it % 3 == 0 && n > 64)
            for (size_t i = 64; i < n; i++)
                if (rng() % 4) d[i] = d[i - 1 - rng() % 60];                        // plenty of short matches
        static const int strategies[] = {Z_DEFAULT_STRATEGY, Z_FIXED, Z_HUFFMAN_ONLY, Z_RLE, Z_FILTERED};
        std::vector<uint8_t> c = deflate_raw(d, int(rng() % 10), strategies[rng() % 5]);
        bool damaged = it % 2;
        if (damaged && !c.empty()) {
            int k = 1 + int(rng() % 3);
            while (k--) {
                switch (rng() % 3) {
                case 0: c[rng() % c.size()] ^= uint8_t(1u << (rng() % 8)); break;
                case 1: c.resize(rng() % (c.size() + 1)); break;
                default: c[rng() % c.size()] = uint8_t(rng()); break;
                }
                if (c.empty()) break;
            }
        }
        // exact-size heap copies: ASan guards both ends
        uint8_t *in = (uint8_t *)malloc(c.size() ? c.size() : 1);
        if (!c.empty()) memcpy(in, c.data(), c.size());
        uint8_t *o1 = (uint8_t *)malloc(n ? n : 1), *o2 = (uint8_t *)malloc(n ? n : 1);
        bool a = inf.run(in, c.size(), o1, n);
        bool z = zlib_inflate(in, c.size(), o2, n);
        if (a) {
            accepted++;
            if (!z || memcmp(o1, o2, n) != 0) {
                fprintf(stderr, "MISMATCH at iteration %ld (zlib ok=%d)\n", it, int(z));
                return 1;
            }
        }
        if (!damaged) {
            clean++;
            clean_ok += a;
            if (!z || memcmp(o2, d.data(), n) != 0) {
                fprintf(stderr, "zlib round trip failed?! %ld\n", it);
                return 1;
            }
        }
        free(in); free(o1); free(o2);
    }
    printf("iterations %ld, accepted %ld, undamaged streams handled %ld / %ld\n", iters, accepted, clean_ok, clean);
    return clean_ok * 100 < clean * 99;
}
