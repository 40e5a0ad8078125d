// Builds the bulk cell table for index arrays dumped by python (chrom_off i64, L/R i32, slot u32,
// type u8), prints footprint statistics and verifies stab_lookup against brute force.
#include "../te_counter_b200/csrc/stab_build.h"
#include <cstdio>
#include <cstdlib>
#include <random>
#include <set>
template <class T> std::vector<T> rd(const char* dir, const char* name) {
    std::string p = std::string(dir) + "/" + name; FILE* f = fopen(p.c_str(), "rb"); if (!f) { perror(p.c_str()); exit(1); }
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET); std::vector<T> v(n / sizeof(T)); if (fread(v.data(), 1, n, f) != (size_t)n) exit(1); fclose(f); return v; }
int main(int argc, char** argv) {
    if (argc < 3) { printf("usage: stab_stats DIR SHIFT\n"); return 2; }
    const char* d = argv[1]; int shift = atoi(argv[2]);
    auto off = rd<int64_t>(d, "chrom_off"); auto L = rd<int32_t>(d, "L"); auto R = rd<int32_t>(d, "R"); auto slot = rd<uint32_t>(d, "slot"); auto type = rd<uint8_t>(d, "type");
    int n_chrom = (int)off.size() - 1; uint32_t ns = 0; for (auto s : slot) ns = std::max(ns, s + 1);
    StabTable t; stab_build(t, n_chrom, off.data(), L.data(), R.data(), slot.data(), type.data(), (int)ns, shift);
    if (!t.why_not.empty()) { printf("not built: %s\n", t.why_not.c_str()); return 0; }
    printf("shift %d: features %ld merged %ld entries %ld primary %ld overflow %ld max_chain %ld dup-sectors %ld total %.1f MB\n", shift, (long)off[n_chrom],
           (long)t.n_merged, (long)t.n_entries, (long)t.n_primary, (long)t.n_overflow, (long)t.max_chain, (long)t.n_dup, t.bytes() * 1e-6);
    std::mt19937_64 rng(1); int bad = 0; long follow = 0, total = 0;
    for (int it = 0; it < 200000; ++it) {
        int c = rng() % n_chrom; int64_t lo = off[c], hi = off[c + 1]; if (hi == lo) continue;
        int64_t f = lo + rng() % (hi - lo); int64_t x = (it & 1) ? L[f] + (int64_t)(rng() % 200) - 100 : R[f] + (int64_t)(rng() % 5) - 2; if (it % 7 == 0) x = rng() % 250000000; if (it % 11 == 0) x = (x >> shift) << shift;
        if (it >= 30000) {  // follow-rate sample only
            int ns = 0; stab_lookup(t, c, x, &ns);
            if (ns) { total++; if (ns > 1) follow++; }
            continue;
        }
        std::set<uint32_t> want;
        for (int64_t i = lo; i < hi && L[i] <= x; ++i) if (x < R[i]) want.insert(slot[i]);
        auto got = stab_lookup(t, c, x);
        if (std::vector<uint32_t>(want.begin(), want.end()) != got) { if (bad++ < 5) printf("MISMATCH c=%d x=%ld want %zu got %zu\n", c, (long)x, want.size(), got.size()); }
    }
    printf("verify: %d mismatches of 30000; overflow followed by %.1f%% of feature-biased queries\n", bad, 100.0 * follow / std::max(total, 1L));
}
