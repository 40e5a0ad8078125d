// CPU self-test of the cell table builder (stab_build.h is plain C++): random small indices, every
// position of every chromosome checked against brute force, plus the two-point-in-one-sector
// property of the STAB_EXT extension.  Exit code 0 = all good.  Built and run by
// tests/test_cell_table_cpu.py.
#include "../te_counter_b200/csrc/stab_build.h"
#include <cstdio>
#include <cstdlib>
#include <random>
#include <set>

static std::vector<uint32_t> brute(const std::vector<int64_t>& off, const std::vector<int32_t>& L, const std::vector<int32_t>& R,
                                   const std::vector<uint32_t>& slot, int c, int64_t x) {
    std::set<uint32_t> s;
    for (int64_t i = off[c]; i < off[c + 1]; ++i) if (L[i] <= x && x < R[i]) s.insert(slot[i]);
    return std::vector<uint32_t>(s.begin(), s.end());
}

// all slots of the sector chain of cell `prim` whose entry contains relative position r (r may lie in the extension)
static std::vector<uint32_t> chain_at(const StabTable& t, int64_t prim, uint32_t r) {
    std::vector<uint32_t> s;
    int64_t sec = prim;
    for (;;) {
        const uint32_t* w = &t.sectors[(size_t)sec * 8];
        const uint32_t header = w[2] >> 16;
        for (int i = 0; i < STAB_ENTRIES; ++i) {
            const uint32_t st = (w[i >> 1] >> (16 * (i & 1))) & 0xFFFFu, en = (w[3 + (i >> 1)] >> (16 * (i & 1))) & 0xFFFFu;
            if (st <= r && r <= en) s.push_back(i < 4 ? (w[6 + (i >> 1)] >> (16 * (i & 1))) & 0xFFFFu : w[5] >> 16);
        }
        if (!(header & 1u) || r < (w[2] & 0xFFFFu)) break;
        sec = (sec == prim) ? (int64_t)t.ovf_base[(size_t)(prim >> STAB_BLOCK_SHIFT)] + (header >> 6) : sec + 1;
    }
    std::sort(s.begin(), s.end());
    s.erase(std::unique(s.begin(), s.end()), s.end());
    return s;
}

int main(int argc, char** argv) {
    const int rounds = argc > 1 ? atoi(argv[1]) : 6;
    std::mt19937_64 rng(12345);
    long checked = 0;
    for (int round = 0; round < rounds; ++round) {
        const int shift = 8 + round % 4;                       // 8..11
        const int n_chrom = 1 + round % 3;
        const int64_t len = 3000 + (int64_t)(rng() % 30000);
        const int n_slots = 3 + (int)(rng() % 40);
        std::vector<int64_t> off(1, 0);
        std::vector<int32_t> L, R;
        std::vector<uint32_t> slot;
        std::vector<uint8_t> type;
        for (int c = 0; c < n_chrom; ++c) {
            const int nf = (int)(rng() % (round % 2 ? 400 : 60));
            std::vector<std::pair<int32_t, int32_t>> iv;
            for (int i = 0; i < nf; ++i) {
                const int32_t a = (int32_t)(rng() % len);
                const int32_t w = (rng() % 10 == 0) ? (int32_t)(rng() % 5000) : (int32_t)(rng() % 300);
                iv.push_back({a, a + w});                          // w == 0: empty interval
            }
            std::sort(iv.begin(), iv.end());
            for (auto& p : iv) {
                L.push_back(p.first); R.push_back(p.second);
                const uint32_t sl = (uint32_t)(rng() % n_slots);
                slot.push_back(sl); type.push_back((uint8_t)(1 + sl % 2));
            }
            off.push_back((int64_t)L.size());
        }
        StabTable t;
        stab_build(t, n_chrom, off.data(), L.data(), R.data(), slot.data(), type.data(), n_slots, shift);
        if (!t.why_not.empty()) { printf("round %d: table not built: %s\n", round, t.why_not.c_str()); return 1; }
        const int64_t csize = (int64_t)1 << shift;
        for (int c = 0; c < n_chrom; ++c) {
            const int64_t n_cells = t.cell_base[c + 1] - t.cell_base[c];
            for (int64_t x = -2; x < n_cells * csize + 5; ++x) {
                const auto want = (x < 0) ? std::vector<uint32_t>() : brute(off, L, R, slot, c, x);
                const auto got = stab_lookup(t, c, x);
                if (want != got) { printf("round %d: point mismatch chrom %d x %ld (shift %d)\n", round, c, (long)x, shift); return 1; }
                ++checked;
                // the extension: x seen from the previous cell
                const int64_t k = x >> shift;
                if (x >= 0 && k >= 1 && k < n_cells && (x & (csize - 1)) < STAB_EXT) {
                    const auto ext = chain_at(t, t.cell_base[c] + k - 1, (uint32_t)((x & (csize - 1)) + csize));
                    if (ext != want) { printf("round %d: extension mismatch chrom %d x %ld\n", round, c, (long)x); return 1; }
                }
            }
        }
    }
    printf("cell table self-test ok: %ld positions\n", checked);
    return 0;
}
