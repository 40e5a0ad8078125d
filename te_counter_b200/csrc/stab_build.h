// Host-side builder of the bulk "cell table" (plain C++, no CUDA).
//
// For a point x on chromosome c the bulk rules need S(x) = { ensg of f : L_f <= x < R_f }
// (bulk.cuh header).  The genome is cut into cells of 2^shift bp and every cell owns exactly one
// 32-byte sector at a fixed address, so a point query is ONE random L2 sector read with no
// directory in front of it:
//
//   sector of cell k of chromosome c  =  sectors[cell_base[c] + k]          (8 x u32)
//   entry i (at most 5 per sector)    =  one interval [s, e] of one ensg, clipped to the cell,
//                                        cell-relative, both ends inclusive (11-bit values)
//        w0 = s0 | s1 << 16     w1 = s2 | s3 << 16     w2 = s4 | header << 16
//        w3 = e0 | e1 << 16     w4 = e2 | e3 << 16     w5 = e4 | slot4  << 16
//        w6 = slot0 | slot1 << 16                       w7 = slot2 | slot3 << 16
//   header (16 bits): bit 0 "more" (the cell's list continues in an overflow sector),
//                     bits 1..5 "twin" mask: entry i has the same ensg as another entry of this
//                     sector; bit order e4, e2, e0, e3, e1 (the order the kernel's hit mask
//                     compacts to),
//                     bits 6..15 link: first overflow sector of this cell, relative to
//                     ovf_base[cell >> 7] (primary sectors only; an overflow sector's
//                     continuation is simply the next sector)
//   unused entries have s = 2047, e = 0 (never contain a point).
//
// The 16-bit lanes are what makes the kernel's test cheap: with X = r * 0x10001 + 0x80008000,
// bit 15 of each lane of (X - w_s) says r >= s and of (w_e + 0x80008000 - r * 0x10001) says e >= r,
// i.e. one add per word tests two entries (SIMD within a register; the adds run on either the
// integer or the FMA pipe, which matters because the kernel is integer-ALU bound).
//
// Every cell's list also covers the first STAB_EXT bp of the next cell (relative positions up to
// 2^shift + STAB_EXT - 1 still fit the 16-bit lanes), so that the two points of a read pair that
// lies across a cell border are answered by ONE sector; a single point is always looked up in the
// cell it falls in, where the extension is invisible.
//
// Intervals of the same ensg are merged per chromosome first (union of [L, R)), so one ensg never
// has two overlapping or touching entries; entries are sorted by start.  A cell with more than 5
// entries continues in overflow sectors (same format, stored after the primary cells, consecutive
// per cell); a query only follows when its point is >= the start of the 5th entry, i.e. when the
// overflow can actually contain a hit.
//
// "slot" is the rank of an ensg in hotness order (most feature rows first): the kernel keeps the
// counters of the first slots in shared memory.  Feature types are looked up per slot
// (slot_type), which requires ensg -> type to be a function (checked; otherwise no table).
#pragma once
#include <stdint.h>
#include <algorithm>
#include <string>
#include <atomic>
#include <thread>
#include <vector>

#define STAB_ENTRIES 5
#ifndef STAB_EXT
#define STAB_EXT 256                     // a cell's entries also cover the first STAB_EXT bp of the next cell
#endif                                   // (A/B on the 500 M-record config, 1 kbp cells: 128 -> 121.8e9, 256 -> 125.5e9, 384 -> 124.7e9 records/s)
#define STAB_MAX_SHIFT 11
#define STAB_MAX_SLOTS 65535
#define STAB_BLOCK_SHIFT 7               // ovf_base granularity: 128 primary sectors
#define STAB_MAX_LINK 1023u

struct StabTable {
    int shift = 11;
    int n_slots = 0;
    int all_counted = 1;                 // every type is gene / TE / snRNA: the type rule is always true
    std::vector<int64_t> cell_base;      // n_chrom + 1
    std::vector<uint32_t> sectors;       // 8 words per sector: primary cells, then overflow sectors
    std::vector<uint32_t> ovf_base;      // per block of 256 primary sectors: first overflow sector
    std::vector<uint8_t> slot_type;      // n_slots
    int64_t n_primary = 0, n_overflow = 0, n_entries = 0, n_merged = 0, max_chain = 0, n_dup = 0;
    std::string why_not;                 // non-empty: no table (limits) -> the exact kernel is used
    size_t bytes() const { return (sectors.size() + ovf_base.size()) * 4 + cell_base.size() * 8 + slot_type.size(); }
};

struct StabEntry {
    int64_t cell;        // global cell index
    uint32_t s, len, slot;
};

// Entries of the cell table: per chromosome, per slot the union of [L, R) (so one ensg never has two
// overlapping or touching entries), clipped to cells of 2^shift bp that also cover the first `ext` bp of
// the next cell.  Sorted by (cell, start, slot); cell_base[c] = first cell of chromosome c.
inline void stab_collect_entries(std::vector<int64_t>& cell_base, int64_t& n_merged, std::vector<StabEntry>& ent,
                                 int n_chrom, const int64_t* chrom_off, const int32_t* L, const int32_t* R,
                                 const uint32_t* slot, int shift, int ext) {
    const int64_t csize = (int64_t)1 << shift;
    cell_base.assign((size_t)n_chrom + 1, 0);
    for (int c = 0; c < n_chrom; ++c) {
        int32_t maxc = 0;
        for (int64_t i = chrom_off[c]; i < chrom_off[c + 1]; ++i) maxc = std::max(maxc, R[i]);
        cell_base[(size_t)c + 1] = cell_base[(size_t)c] + (((int64_t)maxc >> shift) + 1);
    }
    // Chromosomes own disjoint, ascending cell ranges, so each one is built and sorted on its own (a
    // thread per chromosome at a time) and the concatenation in chromosome order is globally sorted.
    struct Iv { uint32_t slot; int32_t L, R; };
    std::vector<std::vector<StabEntry>> per((size_t)n_chrom);
    std::vector<int64_t> merged((size_t)n_chrom, 0);
    auto build_chrom = [&](int c) {
        std::vector<Iv> iv;
        std::vector<StabEntry>& e = per[(size_t)c];
        iv.reserve((size_t)(chrom_off[c + 1] - chrom_off[c]));
        for (int64_t i = chrom_off[c]; i < chrom_off[c + 1]; ++i)
            if (R[i] > L[i]) iv.push_back({slot[i], L[i], R[i]});    // [L, R) empty: never stabbed
        std::sort(iv.begin(), iv.end(), [](const Iv& a, const Iv& b) { return a.slot != b.slot ? a.slot < b.slot : a.L < b.L; });
        e.reserve(iv.size() + iv.size() / 2);
        const int64_t n_cells_c = cell_base[(size_t)c + 1] - cell_base[(size_t)c];
        size_t i = 0;
        while (i < iv.size()) {
            const uint32_t s = iv[i].slot;
            int64_t a = iv[i].L, b = iv[i].R;
            size_t j = i + 1;
            while (j < iv.size() && iv[j].slot == s && iv[j].L <= b) { b = std::max<int64_t>(b, iv[j].R); ++j; }
            merged[(size_t)c]++;
            for (int64_t k = std::max<int64_t>(0, (a - ext) >> shift); k <= (b - 1) >> shift && k < n_cells_c; ++k) {
                const int64_t c0 = k << shift;
                const int64_t lo = std::max(a, c0), hi = std::min(b, c0 + csize + ext);
                if (hi <= lo) continue;
                e.push_back({cell_base[(size_t)c] + k, (uint32_t)(lo - c0), (uint32_t)(hi - lo), s});
            }
            i = j;
        }
        std::sort(e.begin(), e.end(), [](const StabEntry& a, const StabEntry& b) {
            if (a.cell != b.cell) return a.cell < b.cell;
            if (a.s != b.s) return a.s < b.s;
            return a.slot < b.slot;
        });
    };
    {
        const int n_threads = (int)std::max(1u, std::min<unsigned>({std::thread::hardware_concurrency(), 16u, (unsigned)std::max(n_chrom, 1)}));
        std::atomic<int> next{0};
        std::vector<std::thread> pool;
        auto work = [&] {
            for (int c; (c = next.fetch_add(1)) < n_chrom;) build_chrom(c);
        };
        for (int k = 1; k < n_threads; ++k) pool.emplace_back(work);
        work();
        for (auto& th : pool) th.join();
    }
    size_t total = 0;
    for (const auto& v : per) total += v.size();
    ent.clear();
    ent.reserve(total);
    n_merged = 0;
    for (int c = 0; c < n_chrom; ++c) {
        n_merged += merged[(size_t)c];
        ent.insert(ent.end(), per[(size_t)c].begin(), per[(size_t)c].end());
        std::vector<StabEntry>().swap(per[(size_t)c]);
    }
}

inline void stab_pack_sector(uint32_t* w, const StabEntry* e, int n, bool more, uint32_t link) {
    uint32_t sv[6], ev[6], sl[5];
    for (int i = 0; i < 6; ++i) { sv[i] = 2047u; ev[i] = 0u; }
    for (int i = 0; i < 5; ++i) sl[i] = 0xFFFFu;
    static const int twin_bit[5] = {2, 4, 1, 3, 0};          // entry -> bit of the compacted hit mask
    uint32_t twin = 0;
    for (int i = 0; i < n; ++i) {
        sv[i] = e[i].s; ev[i] = e[i].s + e[i].len - 1; sl[i] = e[i].slot & 0xFFFFu;
        for (int j = 0; j < n; ++j) if (j != i && e[i].slot == e[j].slot) twin |= 1u << twin_bit[i];
    }
    const uint32_t header = (more ? 1u : 0u) | (twin << 1) | (link << 6);
    w[0] = sv[0] | sv[1] << 16; w[1] = sv[2] | sv[3] << 16; w[2] = sv[4] | header << 16;
    w[3] = ev[0] | ev[1] << 16; w[4] = ev[2] | ev[3] << 16; w[5] = ev[4] | sl[4] << 16;
    w[6] = sl[0] | sl[1] << 16; w[7] = sl[2] | sl[3] << 16;
}

// L, R sorted by L inside each chromosome; slot[f] < n_slots and type[f] < 8 per feature.
inline void stab_build(StabTable& t, int n_chrom, const int64_t* chrom_off, const int32_t* L, const int32_t* R,
                       const uint32_t* slot, const uint8_t* type, int n_slots, int shift) {
    t = StabTable();
    t.shift = shift;
    if (shift < 8 || shift > STAB_MAX_SHIFT) { t.why_not = "cell shift out of range"; return; }
    if (n_slots > STAB_MAX_SLOTS) { t.why_not = "more than 65535 ensg (16-bit slots)"; return; }
    t.n_slots = n_slots;
    t.slot_type.assign((size_t)std::max(n_slots, 1), 0xFF);
    for (int64_t i = 0; i < chrom_off[n_chrom]; ++i) {
        uint8_t& ty = t.slot_type[slot[i]];
        if (ty == 0xFF) ty = type[i];
        else if (ty != type[i]) { t.why_not = "an ensg carries more than one feature type"; return; }
    }
    for (auto& ty : t.slot_type) {
        if (ty == 0xFF) ty = 0;
        if (!(ty == 1 || ty == 2 || ty == 3)) t.all_counted = 0;     // TEC_T_GENE / TE / SNRNA
    }
    std::vector<StabEntry> ent;
    stab_collect_entries(t.cell_base, t.n_merged, ent, n_chrom, chrom_off, L, R, slot, shift, STAB_EXT);
    t.n_primary = t.cell_base[(size_t)n_chrom];
    if ((uint64_t)t.n_primary >= 0xFFFFFFF0ull) { t.why_not = "too many cells"; return; }
    t.n_entries = (int64_t)ent.size();
    // overflow sectors: consecutive per cell, cells in order; links are relative to the block base
    int64_t n_over = 0;
    t.ovf_base.assign((size_t)((t.n_primary >> STAB_BLOCK_SHIFT) + 2), 0);
    {
        int64_t blk = -1;
        for (size_t i = 0; i < ent.size();) {
            size_t j = i;
            while (j < ent.size() && ent[j].cell == ent[i].cell) ++j;
            const int64_t n = (int64_t)(j - i), b = ent[i].cell >> STAB_BLOCK_SHIFT;
            while (blk < b) t.ovf_base[(size_t)++blk] = (uint32_t)(t.n_primary + n_over);
            if (n > STAB_ENTRIES) n_over += (n - 1) / STAB_ENTRIES;
            t.max_chain = std::max(t.max_chain, (n + STAB_ENTRIES - 1) / STAB_ENTRIES);
            i = j;
        }
        while (blk + 1 < (int64_t)t.ovf_base.size()) t.ovf_base[(size_t)++blk] = (uint32_t)(t.n_primary + n_over);
    }
    t.n_overflow = n_over;
    if ((uint64_t)(t.n_primary + n_over) >= 0xFFFFFFF0ull) { t.why_not = "too many sectors"; return; }
    t.sectors.assign((size_t)(t.n_primary + n_over) * 8, 0);
    for (int64_t c = 0; c < t.n_primary; ++c) stab_pack_sector(&t.sectors[(size_t)c * 8], nullptr, 0, false, 0);
    int64_t next_over = t.n_primary;
    for (size_t i = 0; i < ent.size();) {
        size_t j = i;
        while (j < ent.size() && ent[j].cell == ent[i].cell) ++j;
        int64_t sec = ent[i].cell;
        const uint32_t base = t.ovf_base[(size_t)(ent[i].cell >> STAB_BLOCK_SHIFT)];
        for (size_t k = i; k < j; k += STAB_ENTRIES) {
            const int n = (int)std::min<size_t>(STAB_ENTRIES, j - k);
            const bool more = k + STAB_ENTRIES < j;
            uint32_t link = 0;
            if (more && k == i) {                    // primary sector: relative link
                if ((uint64_t)next_over - base > STAB_MAX_LINK) { t.why_not = "overflow link out of range"; t.sectors.clear(); return; }
                link = (uint32_t)(next_over - base);
            }
            stab_pack_sector(&t.sectors[(size_t)sec * 8], &ent[k], n, more, link);
            if (t.sectors[(size_t)sec * 8 + 2] & 0x3E0000u) t.n_dup++;
            if (more) sec = next_over++;
        }
        i = j;
    }
}

// S(x) as sorted distinct slots (host-side reader mirroring the kernel; tests and tools)
inline std::vector<uint32_t> stab_lookup(const StabTable& t, int c, int64_t x, int* n_sectors = nullptr) {
    std::vector<uint32_t> s;
    if (n_sectors) *n_sectors = 0;
    if (x < 0) return s;
    const int64_t cell = x >> t.shift;
    if (cell >= t.cell_base[(size_t)c + 1] - t.cell_base[(size_t)c]) return s;
    const uint32_t r = (uint32_t)(x & (((int64_t)1 << t.shift) - 1));
    const int64_t prim = t.cell_base[(size_t)c] + cell;
    int64_t sec = prim;
    for (;;) {
        const uint32_t* w = &t.sectors[(size_t)sec * 8];
        if (n_sectors) ++*n_sectors;
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
