// CPU self-test of the bulk cell table, layout 2 (stab2_build.h is plain C++): random small indices, random and
// exhaustive point pairs checked against brute force -- the set of distinct ensg stabbed by x1 or x2 -- through
// stab2_unit(), the scalar statement of what the two kernels of bulk2.cuh do (fast sector test with in-place
// twin rule, second pass over sector chains).  Units routed to the exact search must be exactly those that
// touch an EDGE cell or hit too many ensg.  Every deferred unit also goes through stab2_pair_unit(), the scalar
// statement of bulk2_pair_kernel (two sectors, completeness rules, 5 + 5 slot comparison): what it answers must be the
// brute-force set, and it must leave EDGE units alone.  Exit code 0 = all good.  Run by tests/test_cell_table_cpu.py.
#include "../te_counter_b200/csrc/stab2_build.h"
#include <cstdio>
#include <cstdlib>
#include <random>
#include <set>

static std::vector<uint32_t> brute2(const std::vector<int64_t>& off, const std::vector<int32_t>& L, const std::vector<int32_t>& R,
                                    const std::vector<uint32_t>& slot, int c, int64_t xa, int64_t xb) {
    std::set<uint32_t> s;
    for (int64_t i = off[c]; i < off[c + 1]; ++i)
        if ((L[i] <= xa && xa < R[i]) || (L[i] <= xb && xb < R[i])) s.insert(slot[i]);
    return std::vector<uint32_t>(s.begin(), s.end());
}

int main(int argc, char** argv) {
    const int rounds = argc > 1 ? atoi(argv[1]) : 6;
    std::mt19937_64 rng(777);
    long checked = 0, fast = 0, second = 0, exact = 0, twin_sectors = 0, forced = 0, pair_done = 0, pair_left = 0;
    for (int round = 0; round < rounds; ++round) {
        const int shift = 8 + round % 4;                       // 8..11
        const int n_chrom = 1 + round % 3;
        const int64_t len = 3000 + (int64_t)(rng() % 30000);
        const int n_slots = 2 + (int)(rng() % (round % 2 ? 6 : 40));      // few slots: many twins
        const int bs = (round % 2) ? 1000 : 10000;             // small buckets: many EDGE cells
        std::vector<int64_t> off(1, 0);
        std::vector<int32_t> L, R;
        std::vector<uint32_t> slot;
        std::vector<uint8_t> type;
        for (int c = 0; c < n_chrom; ++c) {
            const int nf = (int)(rng() % (round % 2 ? 500 : 80));
            std::vector<std::pair<int32_t, int32_t>> iv;
            for (int i = 0; i < nf; ++i) {
                int32_t a = (int32_t)(rng() % len);
                if (rng() % 40 == 0) a = a / bs * bs;              // start on a bucket edge
                int32_t w = (rng() % 10 == 0) ? (int32_t)(rng() % 5000) : (int32_t)(rng() % 300);
                if (rng() % 40 == 0) w = std::max(0, (a + w) / bs * bs + bs - 1 - a);      // (R + 1) % bs == 0
                iv.push_back({a, a + w});
            }
            std::sort(iv.begin(), iv.end());
            for (auto& p : iv) {
                L.push_back(p.first); R.push_back(p.second);
                slot.push_back((uint32_t)(rng() % n_slots));
            }
            off.push_back((int64_t)L.size());
        }
        type.assign(L.size(), 2);
        StabTable2 t;
        stab2_build(t, n_chrom, off.data(), L.data(), R.data(), slot.data(), type.data(), n_slots, shift, bs);
        if (!t.why_not.empty()) { printf("round %d: table not built: %s\n", round, t.why_not.c_str()); return 1; }
        twin_sectors += t.n_twin_sectors; forced += t.n_force;
        const int64_t S = (int64_t)1 << shift;
        std::vector<uint32_t> got;
        for (int c = 0; c < n_chrom; ++c) {
            const int64_t ncc = t.cell_base[c + 1] - t.cell_base[c];
            const int64_t hi = ncc * S + 300;
            auto check = [&](int64_t xa, int64_t xb) -> bool {
                const int how = stab2_unit(t, c, xa, xb, got, 8);
                const auto want = brute2(off, L, R, slot, c, xa, xb);
                ++checked;
                if (how != 0) {
                    // the two-sector kernel goes first on every deferred unit: what it answers must be the brute-force set
                    // (each ensg once), and it must leave every unit of an EDGE cell to the passes behind it
                    std::vector<uint32_t> pg;
                    if (stab2_pair_unit(t, c, xa, xb, pg)) {
                        ++pair_done;
                        if (pg != want) {
                            printf("round %d: two-sector kernel mismatch chrom %d xa %ld xb %ld shift %d: want %zu got %zu\n", round, c, (long)xa, (long)xb, shift, want.size(), pg.size());
                            return false;
                        }
                        if (how == 2 && want.size() <= 8) { printf("round %d: two-sector kernel answered an EDGE unit c %d xa %ld xb %ld\n", round, c, (long)xa, (long)xb); return false; }
                    } else {
                        ++pair_left;
                    }
                }
                fast += how == 0; second += how == 1; exact += how == 2;
                if (how == 2) {
                    // legitimate only for EDGE cells or large sets
                    bool edge = false;
                    for (int64_t x : {xa, xb}) {
                        if (x < 0) continue;
                        for (int64_t k : {x >> shift, (x >> shift) - 1}) {
                            if (k < 0 || k >= ncc) continue;
                            if (t.sectors[(size_t)(t.cell_base[c] + k) * 8 + 2] >> 16 & S2_H_EDGE) edge = true;
                        }
                    }
                    if (!edge && want.size() <= 8) { printf("round %d: needless exact c %d xa %ld xb %ld\n", round, c, (long)xa, (long)xb); return false; }
                    return true;
                }
                if (got != want) {
                    printf("round %d: mismatch (how %d) chrom %d xa %ld xb %ld shift %d: want %zu got %zu\n", round, how, c, (long)xa, (long)xb, shift, want.size(), got.size());
                    return false;
                }
                return true;
            };
            // every position paired with a mate a short random distance away (both orders), and as a single point
            for (int64_t x = -2; x < hi; ++x) {
                const int64_t d = (int64_t)(rng() % 400);
                if (!check(x, x + d)) return 1;
                if (!check(x + (int64_t)(rng() % 300), x)) return 1;
                if (x % 3 == 0 && !check(x, x)) return 1;
            }
            for (int it = 0; it < 20000; ++it) {                  // far pairs
                const int64_t xa = (int64_t)(rng() % (hi + 50)) - 20, xb = (int64_t)(rng() % (hi + 50)) - 20;
                if (!check(xa, xb)) return 1;
            }
            // the EDGE flag must cover every position where the candidate rule can bite
            for (int64_t i = off[c]; i < off[c + 1]; ++i) {
                if (R[i] <= L[i]) continue;
                for (int64_t p : {(L[i] % bs == 0) ? (int64_t)L[i] : -1, ((R[i] + 1) % bs == 0) ? (int64_t)R[i] - 1 : -1}) {
                    if (p < 0) continue;
                    for (int64_t other : {p, p + 10, p - 10})
                        if (stab2_unit(t, c, p, other, got, 1 << 20) != 2 || stab2_unit(t, c, other, p, got, 1 << 20) != 2) {
                            printf("round %d: EDGE position %ld not routed to the exact search\n", round, (long)p);
                            return 1;
                        }
                }
            }
        }
    }
    printf("cell table 2 self-test ok: %ld units (%ld fast, %ld second pass, %ld exact), %ld twin sectors, %ld forced; two-sector kernel: %ld answered, %ld left\n",
           checked, fast, second, exact, twin_sectors, forced, pair_done, pair_left);
    return 0;
}
