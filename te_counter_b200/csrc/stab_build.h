// Host-side builder of the bulk "cell table" (plain C++, no CUDA).
//
// For a point x on chromosome c the bulk rules need S(x) = { ensg of f : L_f <= x < R_f }
// (bulk.cuh header).  The genome is cut into cells of 2^shift bp and every cell owns exactly one
// 32-byte sector at a fixed address, so a point query is ONE random L2 sector read with no
// directory in front of it:
//
//   sector of cell k of chromosome c  =  sectors[cell_base[c] + k]          (8 x u32)
//   entry i (at most 6 per sector)    =  one interval of one ensg, clipped to the cell:
//        slot_i   16 bits   w[i/2] >> 16*(i&1)                              (w0..w2)
//        pos_i    22 bits   bit string w3..w7 at bit offset 22*i:  s | (len-1) << 11
//                           s = interval start - cell start, len = clipped length
//   header   28 bits   w7 >> 4:   code (3 bits: 0..6 = number of entries, 7 = 6 entries + link)
//                                 | dup (1 bit: two entries of this sector carry the same ensg)
//                                 | link (24 bits: sector index of the overflow sector)
//
// Intervals of the same ensg are merged per chromosome first (union of [L, R)), so one ensg never
// has two overlapping or touching entries; entries are sorted by start.  A cell with more than 6
// entries keeps its first 6 in the primary sector and links to overflow sectors (same format,
// stored after the primary cells); a query only follows the link when its point is >= the start
// of the last entry of the sector, i.e. when the overflow can actually contain a hit.
//
// "slot" is the rank of an ensg in hotness order (most feature rows first): the kernel keeps the
// counters of the first slots in shared memory.  Feature types are looked up per slot
// (slot_type), which requires ensg -> type to be a function (checked; otherwise no table).
//
// Footprint on the hg38-like synthetic index (5.9 M features, shift 11): 1.51 M primary sectors
// (48 MB) + overflow; sized to stay L2 resident on B200 beside the streamed records.
#pragma once
#include <stdint.h>
#include <algorithm>
#include <string>
#include <vector>

#define STAB_ENTRIES 6
#define STAB_POS_BITS 22
#define STAB_MAX_SHIFT 11
#define STAB_MAX_SLOTS 65535
#define STAB_MAX_SECTORS (1u << 24)

struct StabTable {
    int shift = 11;
    int n_slots = 0;
    int all_counted = 1;                 // every type is gene / TE / snRNA: the type rule is always true
    std::vector<int64_t> cell_base;      // n_chrom + 1
    std::vector<uint32_t> sectors;       // 8 words per sector: primary cells, then overflow sectors
    std::vector<uint8_t> slot_type;      // n_slots
    int64_t n_primary = 0, n_overflow = 0, n_entries = 0, n_merged = 0, max_chain = 0, n_dup = 0;
    std::string why_not;                 // non-empty: no table (limits) -> the exact kernel is used
    size_t bytes() const { return sectors.size() * 4 + cell_base.size() * 8 + slot_type.size(); }
};

struct StabEntry {
    int64_t cell;        // global cell index
    uint32_t s, len, slot;
};

inline void stab_pack_sector(uint32_t* w, const StabEntry* e, int n, bool has_link, uint32_t link) {
    for (int i = 0; i < 8; ++i) w[i] = 0;
    uint64_t lo = 0, hi = 0, top = 0;            // 160-bit string w3..w7 as lo (64) | hi (64) | top (32)
    for (int i = 0; i < n; ++i) {
        w[i >> 1] |= (e[i].slot & 0xFFFFu) << (16 * (i & 1));
        const uint64_t pos = (uint64_t)e[i].s | ((uint64_t)(e[i].len - 1) << 11);
        const int off = STAB_POS_BITS * i;
        if (off < 64) { lo |= pos << off; if (off + STAB_POS_BITS > 64) hi |= pos >> (64 - off); }
        else if (off < 128) { hi |= pos << (off - 64); if (off + STAB_POS_BITS > 128) top |= pos >> (128 - off); }
        else top |= pos << (off - 128);
    }
    bool dup = false;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < i; ++j) dup |= e[i].slot == e[j].slot;
    const uint32_t header = (has_link ? 7u : (uint32_t)n) | (dup ? 8u : 0u) | (link << 4);
    top |= (uint64_t)header << 4;                // bit 132 of the string = bit 4 of w7
    w[3] = (uint32_t)lo; w[4] = (uint32_t)(lo >> 32); w[5] = (uint32_t)hi; w[6] = (uint32_t)(hi >> 32);
    w[7] = (uint32_t)top;
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
    const int64_t csize = (int64_t)1 << shift;
    // cells per chromosome
    t.cell_base.assign((size_t)n_chrom + 1, 0);
    for (int c = 0; c < n_chrom; ++c) {
        int32_t maxc = 0;
        for (int64_t i = chrom_off[c]; i < chrom_off[c + 1]; ++i) maxc = std::max(maxc, R[i]);
        t.cell_base[(size_t)c + 1] = t.cell_base[(size_t)c] + (((int64_t)maxc >> shift) + 1);
    }
    t.n_primary = t.cell_base[(size_t)n_chrom];
    if ((uint64_t)t.n_primary >= STAB_MAX_SECTORS) { t.why_not = "too many cells for the 24-bit link"; return; }
    // entries: per chromosome, per slot union of [L, R), clipped to cells
    struct Iv { uint32_t slot; int32_t L, R; };
    std::vector<Iv> iv;
    std::vector<StabEntry> ent;
    ent.reserve((size_t)chrom_off[n_chrom] + (size_t)chrom_off[n_chrom] / 4);
    for (int c = 0; c < n_chrom; ++c) {
        iv.clear();
        for (int64_t i = chrom_off[c]; i < chrom_off[c + 1]; ++i)
            if (R[i] > L[i]) iv.push_back({slot[i], L[i], R[i]});    // [L, R) empty: never stabbed
        std::sort(iv.begin(), iv.end(), [](const Iv& a, const Iv& b) { return a.slot != b.slot ? a.slot < b.slot : a.L < b.L; });
        size_t i = 0;
        while (i < iv.size()) {
            const uint32_t s = iv[i].slot;
            int64_t a = iv[i].L, b = iv[i].R;
            size_t j = i + 1;
            while (j < iv.size() && iv[j].slot == s && iv[j].L <= b) { b = std::max<int64_t>(b, iv[j].R); ++j; }
            t.n_merged++;
            for (int64_t k = a >> shift; k <= (b - 1) >> shift; ++k) {
                const int64_t c0 = k << shift;
                const int64_t lo = std::max(a, c0), hi = std::min(b, c0 + csize);
                ent.push_back({t.cell_base[(size_t)c] + k, (uint32_t)(lo - c0), (uint32_t)(hi - lo), s});
            }
            i = j;
        }
    }
    std::sort(ent.begin(), ent.end(), [](const StabEntry& a, const StabEntry& b) {
        if (a.cell != b.cell) return a.cell < b.cell;
        if (a.s != b.s) return a.s < b.s;
        return a.slot < b.slot;
    });
    t.n_entries = (int64_t)ent.size();
    // overflow sectors needed
    int64_t n_over = 0;
    for (size_t i = 0; i < ent.size();) {
        size_t j = i;
        while (j < ent.size() && ent[j].cell == ent[i].cell) ++j;
        const int64_t n = (int64_t)(j - i);
        if (n > STAB_ENTRIES) n_over += (n - 1) / STAB_ENTRIES;
        t.max_chain = std::max(t.max_chain, (n + STAB_ENTRIES - 1) / STAB_ENTRIES);
        i = j;
    }
    t.n_overflow = n_over;
    if ((uint64_t)(t.n_primary + n_over) >= STAB_MAX_SECTORS) { t.why_not = "too many sectors for the 24-bit link"; return; }
    t.sectors.assign((size_t)(t.n_primary + n_over) * 8, 0);
    int64_t next_over = t.n_primary;
    for (size_t i = 0; i < ent.size();) {
        size_t j = i;
        while (j < ent.size() && ent[j].cell == ent[i].cell) ++j;
        int64_t sec = ent[i].cell;
        for (size_t k = i; k < j; k += STAB_ENTRIES) {
            const int n = (int)std::min<size_t>(STAB_ENTRIES, j - k);
            const bool more = k + STAB_ENTRIES < j;
            const int64_t link = more ? next_over++ : 0;
            stab_pack_sector(&t.sectors[(size_t)sec * 8], &ent[k], n, more, (uint32_t)link);
            if (t.sectors[(size_t)sec * 8 + 7] & 0x80u) t.n_dup++;
            sec = link;
        }
        i = j;
    }
}

// S(x) as sorted distinct slots (host-side reader mirroring the kernel; tests and tools)
inline std::vector<uint32_t> stab_lookup(const StabTable& t, int c, int64_t x) {
    std::vector<uint32_t> s;
    if (x < 0) return s;
    const int64_t cell = x >> t.shift;
    if (cell >= t.cell_base[(size_t)c + 1] - t.cell_base[(size_t)c]) return s;
    const uint32_t r = (uint32_t)(x & (((int64_t)1 << t.shift) - 1));
    int64_t sec = t.cell_base[(size_t)c] + cell;
    for (;;) {
        const uint32_t* w = &t.sectors[(size_t)sec * 8];
        const uint32_t header = w[7] >> 4;
        const bool has_link = (header & 7u) == 7u;
        const int n = has_link ? STAB_ENTRIES : (int)(header & 7u);
        uint32_t last_s = 0;
        for (int i = 0; i < n; ++i) {
            const int off = STAB_POS_BITS * i;
            uint64_t bits = 0;                       // 64-bit window of the string starting at word 3 + off / 32
            const int wi = 3 + off / 32;
            bits = (uint64_t)w[wi] | (wi + 1 < 8 ? (uint64_t)w[wi + 1] << 32 : 0);
            const uint32_t pos = (uint32_t)(bits >> (off % 32)) & ((1u << STAB_POS_BITS) - 1);
            const uint32_t st = pos & 2047u, lm1 = pos >> 11;
            last_s = st;
            if (r - st <= lm1) s.push_back((w[i >> 1] >> (16 * (i & 1))) & 0xFFFFu);
        }
        if (!has_link || r < last_s) break;
        sec = header >> 4;
    }
    std::sort(s.begin(), s.end());
    s.erase(std::unique(s.begin(), s.end()), s.end());
    return s;
}
