// Host-side builder of the bulk cell table, layout 2 (plain C++, no CUDA), and a scalar reader of it.
//
// Same idea as stab_build.h -- the genome is cut into cells of 2^shift bp, every cell owns one 32-byte
// sector at a fixed address and also covers the first `ext` bp of the next cell, entries are one ensg's
// merged interval clipped to the cell in 16-bit lanes -- but the sector header is laid out so that the
// counting kernel (bulk2.cuh) answers a unit with ONE test of the form `rmax >= thr`:
//
//        w0 = s0 | s1 << 16     w1 = s2 | s3 << 16     w2 = s4 | header << 16
//        w3 = e0 | e1 << 16     w4 = e2 | e3 << 16     w5 = e4 | slot4  << 16
//        w6 = slot0 | slot1 << 16                       w7 = slot2 | slot3 << 16
//   header bits 0..11  thr: a unit whose larger cell-relative point is >= thr cannot be answered from this
//                      sector alone.  0xFFF = never (no point reaches it), 0 = always.
//               bit 12 EDGE  the cell covers a position where the reference's two-bucket candidate set can
//                      differ from a plain stab query (a feature with L % bs == 0 or (R + 1) % bs == 0,
//                      bulk.cuh header): such units take the exact search.  thr = 0.
//               bit 13 MORE  the cell's list continues in overflow sectors
//               bit 14 TW23  entries 2 and 3 carry the same ensg        (w2 bit 30)
//               bit 15 TW01  entries 0 and 1 carry the same ensg        (w2 bit 31)
//   unused entries: s = 0xFFF, e = 0 (never contain a point), slot = 0xFFFF.  Cell-relative positions are below
//   2^shift + ext <= 2304, so they fit the 12 bits.
//
// The entries with the smallest starts stay in the primary sector (at most five), so with thr = start of
// the first entry that did not fit a point below thr cannot touch anything in the overflow sectors.
// Entries of one ensg never overlap (they were merged), so when two of them share a sector a unit can hit
// both only with its two points; the builder moves such twins to positions (0, 1) and (2, 3) and the kernel
// counts a pair once.  The primary prefix ends early where that cannot be expressed (a third entry of one
// ensg, a third pair): stab2_primary_prefix().
// Overflow sectors (same format, consecutive per cell) are only read by the second-pass kernel, which keeps
// a register set of distinct ensg and therefore needs no twin flags; their thr is the start of the first
// entry of the next overflow sector.  ovf_first[primary sector] = first overflow sector of the cell.
#pragma once
#include "stab_build.h"

#define S2_ENTRIES 5
#define S2_THR_NEVER 0xFFFu
#define S2_THR_MASK 0xFFFu
#define S2_H_EDGE (1u << 12)
#define S2_H_MORE (1u << 13)
#define S2_H_TW23 (1u << 14)
#define S2_H_TW01 (1u << 15)
#define S2_R_NONE 0x7FFFu            // "no point": fails the e >= r test of every entry

struct StabTable2 {
    int shift = 10, ext = STAB_EXT;
    int n_slots = 0;
    int all_counted = 1;
    std::vector<int64_t> cell_base;      // n_chrom + 1
    std::vector<uint32_t> sectors;       // 8 words per sector: primary cells, then overflow sectors
    std::vector<uint32_t> ovf_first;     // per primary sector: first overflow sector (0 = none)
    std::vector<uint8_t> slot_type;
    int64_t n_primary = 0, n_overflow = 0, n_entries = 0, n_merged = 0, max_chain = 0;
    int64_t n_edge_cells = 0, n_force = 0, n_twin_sectors = 0;
    std::string why_not;
    size_t bytes() const { return (sectors.size() + ovf_first.size()) * 4 + cell_base.size() * 8 + slot_type.size(); }
};

inline void stab2_pack(uint32_t* w, const StabEntry* const* e, int n, uint32_t header) {
    uint32_t sv[5], ev[5], sl[5];
    for (int i = 0; i < 5; ++i) { sv[i] = 0xFFFu; ev[i] = 0u; sl[i] = 0xFFFFu; }
    for (int i = 0; i < n; ++i) {
        if (!e[i]) continue;
        sv[i] = e[i]->s; ev[i] = e[i]->s + e[i]->len - 1; sl[i] = e[i]->slot & 0xFFFFu;
    }
    w[0] = sv[0] | sv[1] << 16; w[1] = sv[2] | sv[3] << 16; w[2] = sv[4] | header << 16;
    w[3] = ev[0] | ev[1] << 16; w[4] = ev[2] | ev[3] << 16; w[5] = ev[4] | sl[4] << 16;
    w[6] = sl[0] | sl[1] << 16; w[7] = sl[2] | sl[3] << 16;
}

// how many entries of a cell (sorted by start) stay in its primary sector: the longest prefix of at most
// five in which no ensg appears three times and at most two ensg appear twice
inline int stab2_primary_prefix(const StabEntry* e, int n) {
    int np = 0, n_pairs = 0;
    for (; np < std::min(n, S2_ENTRIES); ++np) {
        int same = 0;
        for (int b = 0; b < np; ++b) same += e[b].slot == e[np].slot;
        if (same >= 2 || (same == 1 && n_pairs == 2)) break;
        if (same == 1) ++n_pairs;
    }
    return np;
}

// L, R sorted by L inside each chromosome; slot[f] < n_slots and type[f] < 8 per feature; bs = bucket size.
inline void stab2_build(StabTable2& t, int n_chrom, const int64_t* chrom_off, const int32_t* L, const int32_t* R,
                        const uint32_t* slot, const uint8_t* type, int n_slots, int shift, int bs) {
    t = StabTable2();
    t.shift = shift;
    if (shift < 8 || shift > 11) { t.why_not = "cell shift out of range (8..11)"; return; }   // relative positions < 2^shift + ext < 0xFFF
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
    stab_collect_entries(t.cell_base, t.n_merged, ent, n_chrom, chrom_off, L, R, slot, shift, t.ext);
    t.n_primary = t.cell_base[(size_t)n_chrom];
    if ((uint64_t)t.n_primary >= 0x7FFFFFF0ull) { t.why_not = "too many cells"; return; }
    t.n_entries = (int64_t)ent.size();
    // cells that cover a position at which the candidate rule can bite (bulk.cuh header): point A at
    // L when L % bs == 0, point B (x = loc2 - 1) at R - 1 when (R + 1) % bs == 0
    std::vector<uint8_t> edge((size_t)std::max<int64_t>(t.n_primary, 1), 0);
    const int64_t cmask = ((int64_t)1 << shift) - 1;
    for (int c = 0; c < n_chrom; ++c) {
        const int64_t ncc = t.cell_base[(size_t)c + 1] - t.cell_base[(size_t)c];
        auto mark = [&](int64_t p) {
            if (p < 0) return;
            const int64_t k = p >> shift;
            if (k < ncc) edge[(size_t)(t.cell_base[(size_t)c] + k)] = 1;
            if (k >= 1 && k - 1 < ncc && (p & cmask) < t.ext) edge[(size_t)(t.cell_base[(size_t)c] + k - 1)] = 1;
        };
        for (int64_t i = chrom_off[c]; i < chrom_off[c + 1]; ++i) {
            if (R[i] <= L[i]) continue;
            if (L[i] % bs == 0) mark(L[i]);
            if (((int64_t)R[i] + 1) % bs == 0) mark((int64_t)R[i] - 1);
        }
    }
    // overflow sectors: consecutive per cell, cells in order.  How many entries stay in the primary sector
    // depends on the twin rule, so the chains are sized here with the same prefix rule as below.
    t.ovf_first.assign((size_t)std::max<int64_t>(t.n_primary, 1), 0);
    int64_t n_over = 0;
    for (size_t i = 0; i < ent.size();) {
        size_t j = i;
        while (j < ent.size() && ent[j].cell == ent[i].cell) ++j;
        const int n = (int)(j - i);
        const int np = stab2_primary_prefix(&ent[i], n);
        if (n > np) {
            t.ovf_first[(size_t)ent[i].cell] = (uint32_t)(t.n_primary + n_over);
            n_over += (n - np + S2_ENTRIES - 1) / S2_ENTRIES;
        }
        t.max_chain = std::max<int64_t>(t.max_chain, 1 + (n - np + S2_ENTRIES - 1) / S2_ENTRIES);
        i = j;
    }
    t.n_overflow = n_over;
    if ((uint64_t)(t.n_primary + n_over) >= 0x7FFFFFF0ull) { t.why_not = "too many sectors"; return; }
    t.sectors.assign((size_t)std::max<int64_t>(t.n_primary + n_over, 1) * 8, 0);
    for (int64_t c = 0; c < t.n_primary; ++c) {
        const uint32_t header = edge[(size_t)c] ? (S2_H_EDGE | 0u) : S2_THR_NEVER;
        stab2_pack(&t.sectors[(size_t)c * 8], nullptr, 0, header);
        t.n_edge_cells += edge[(size_t)c];
    }
    for (size_t i = 0; i < ent.size();) {
        size_t j = i;
        while (j < ent.size() && ent[j].cell == ent[i].cell) ++j;
        const int64_t cell = ent[i].cell;
        const int n = (int)(j - i);
        // ---- primary sector: the longest prefix (by start) of at most five entries in which no ensg appears
        //      three times and at most two ensg appear twice; twins go to positions (0, 1) and (2, 3)
        const int np = stab2_primary_prefix(&ent[i], n);
        if (np < std::min(n, S2_ENTRIES)) t.n_force++;
        const StabEntry* pos[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
        uint32_t header = 0;
        {
            bool used[5] = {false, false, false, false, false};
            int pairs = 0;
            for (int a = 0; a < np; ++a) {
                if (used[a]) continue;
                for (int b = a + 1; b < np; ++b)
                    if (!used[b] && ent[i + b].slot == ent[i + a].slot) {
                        pos[2 * pairs] = &ent[i + a];
                        pos[2 * pairs + 1] = &ent[i + b];
                        used[a] = used[b] = true;
                        header |= pairs ? S2_H_TW23 : S2_H_TW01;
                        ++pairs;
                        break;
                    }
            }
            int at = 2 * pairs;
            for (int a = 0; a < np; ++a)
                if (!used[a]) pos[at++] = &ent[i + a];
            if (pairs) t.n_twin_sectors++;
        }
        uint32_t thr = (n > np) ? ent[i + np].s : S2_THR_NEVER;
        if (n > np) header |= S2_H_MORE;
        if (edge[(size_t)cell]) { thr = 0; header |= S2_H_EDGE; }
        stab2_pack(&t.sectors[(size_t)cell * 8], pos, 5, header | thr);
        // ---- overflow sectors: plain runs of five, thr = start of the next run
        int64_t sec = t.ovf_first[(size_t)cell];
        for (int k = np; k < n; k += S2_ENTRIES, ++sec) {
            const int m = std::min(S2_ENTRIES, n - k);
            const StabEntry* q[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
            for (int a = 0; a < m; ++a) q[a] = &ent[i + k + a];
            const bool more = k + S2_ENTRIES < n;
            stab2_pack(&t.sectors[(size_t)sec * 8], q, m, (more ? (S2_H_MORE | ent[i + k + S2_ENTRIES].s) : S2_THR_NEVER));
        }
        i = j;
    }
}

// ---------------------------------------------------------------------------------------------------
// Scalar reader (tests and tools): what the two kernels of bulk2.cuh compute for one unit, written with
// plain loops.  out: distinct slots hit by point xa or xb (sorted).  Returns 0 answered by the fast kernel,
// 1 answered by the second pass, 2 the unit needs the exact search (EDGE cell or more than max_distinct ensg).
struct Stab2Probe { int64_t prim; uint32_t ra, rb; };

inline bool stab2_entry_hit(const uint32_t* w, int i, uint32_t r) {
    const uint32_t st = (w[i >> 1] >> (16 * (i & 1))) & 0xFFFFu, en = (w[3 + (i >> 1)] >> (16 * (i & 1))) & 0xFFFFu;
    return st <= r && r <= en;
}
inline uint32_t stab2_entry_slot(const uint32_t* w, int i) {
    return i < 4 ? (w[6 + (i >> 1)] >> (16 * (i & 1))) & 0xFFFFu : w[5] >> 16;
}

inline int stab2_unit(const StabTable2& t, int c, int64_t xa, int64_t xb, std::vector<uint32_t>& out, int max_distinct = 8) {
    out.clear();
    const int64_t ncc = t.cell_base[(size_t)c + 1] - t.cell_base[(size_t)c];
    const int64_t S = (int64_t)1 << t.shift;
    // ---- fast kernel: one sector holds both points
    const int64_t mn = std::min(xa, xb), k = mn >> t.shift;
    const int64_t base = k << t.shift;
    const int64_t rmax = std::max(xa, xb) - base;
    if (k >= 0 && k < ncc && rmax < S + t.ext) {
        const uint32_t* w = &t.sectors[(size_t)(t.cell_base[(size_t)c] + k) * 8];
        const uint32_t header = w[2] >> 16;
        if (rmax < (int64_t)(header & S2_THR_MASK)) {
            bool hit[5];
            for (int i = 0; i < 5; ++i) hit[i] = stab2_entry_hit(w, i, (uint32_t)(xa - base)) || stab2_entry_hit(w, i, (uint32_t)(xb - base));
            if ((header & S2_H_TW01) && hit[0]) hit[1] = false;
            if ((header & S2_H_TW23) && hit[2]) hit[3] = false;
            for (int i = 0; i < 5; ++i) if (hit[i]) out.push_back(stab2_entry_slot(w, i));
            std::sort(out.begin(), out.end());
            return 0;
        }
    }
    // ---- second pass: one or two chains, register set of distinct ensg
    Stab2Probe pr[2];
    int np = 0;
    if (k >= 0 && k < ncc && rmax < S + t.ext) {
        pr[np++] = {t.cell_base[(size_t)c] + k, (uint32_t)(xa - base), (uint32_t)(xb - base)};
    } else {
        if (xa >= 0 && (xa >> t.shift) < ncc) pr[np++] = {t.cell_base[(size_t)c] + (xa >> t.shift), (uint32_t)(xa & (S - 1)), S2_R_NONE};
        if (xb >= 0 && (xb >> t.shift) < ncc) pr[np++] = {t.cell_base[(size_t)c] + (xb >> t.shift), S2_R_NONE, (uint32_t)(xb & (S - 1))};
    }
    bool exact = false;
    for (int p = 0; p < np; ++p) {
        const int rm = std::max(pr[p].ra == S2_R_NONE ? -1 : (int)pr[p].ra, pr[p].rb == S2_R_NONE ? -1 : (int)pr[p].rb);
        int64_t sec = pr[p].prim;
        for (;;) {
            const uint32_t* w = &t.sectors[(size_t)sec * 8];
            const uint32_t header = w[2] >> 16;
            if (sec == pr[p].prim && (header & S2_H_EDGE)) exact = true;
            for (int i = 0; i < 5; ++i)
                if (stab2_entry_hit(w, i, pr[p].ra) || stab2_entry_hit(w, i, pr[p].rb)) out.push_back(stab2_entry_slot(w, i));
            if (!(header & S2_H_MORE)) break;
            const bool forced = sec == pr[p].prim && (header & S2_H_EDGE);
            if (!forced && rm < (int)(header & S2_THR_MASK)) break;
            sec = (sec == pr[p].prim) ? (int64_t)t.ovf_first[(size_t)sec] : sec + 1;
        }
    }
    std::sort(out.begin(), out.end());
    out.erase(std::unique(out.begin(), out.end()), out.end());
    if (exact || (int)out.size() > max_distinct) return 2;
    return 1;
}

// Scalar statement of bulk2_pair_kernel (bulk2.cuh) for a unit the fast kernel deferred: the two sectors it reads, its
// completeness rules, the twin rule by flag bits and the 5 + 5 slot comparison.  Returns false when the kernel leaves the
// unit to bulk2_second_kernel; else out = the slots it counts (sorted), each once.
inline bool stab2_pair_unit(const StabTable2& t, int c, int64_t xa, int64_t xb, std::vector<uint32_t>& out) {
    out.clear();
    const int64_t ncc = t.cell_base[(size_t)c + 1] - t.cell_base[(size_t)c];
    const int64_t S = (int64_t)1 << t.shift;
    const int64_t mn = std::min(xa, xb), k = mn >> t.shift, base = k << t.shift;
    const int64_t rmax = std::max(xa, xb) - base;
    int64_t secA, secB;
    uint32_t raA, rbA, raB, rbB;
    bool two_chains;
    if (k >= 0 && k < ncc && rmax < S + t.ext) {               // one chain, both points: primary + first overflow sector
        secA = t.cell_base[(size_t)c] + k;
        secB = (int64_t)t.ovf_first[(size_t)secA];
        raA = raB = (uint32_t)(xa - base); rbA = rbB = (uint32_t)(xb - base);
        two_chains = false;
    } else {
        Stab2Probe pr[2];
        int np = 0;
        if (xa >= 0 && (xa >> t.shift) < ncc) pr[np++] = {t.cell_base[(size_t)c] + (xa >> t.shift), (uint32_t)(xa & (S - 1)), S2_R_NONE};
        if (xb >= 0 && (xb >> t.shift) < ncc) pr[np++] = {t.cell_base[(size_t)c] + (xb >> t.shift), S2_R_NONE, (uint32_t)(xb & (S - 1))};
        if (np == 0) return true;                               // no probe: nothing to count
        if (np == 1) return false;                              // left to the second pass
        secA = pr[0].prim; raA = pr[0].ra; rbA = pr[0].rb;
        secB = pr[1].prim; raB = pr[1].ra; rbB = pr[1].rb;
        two_chains = true;
    }
    const uint32_t* wA = &t.sectors[(size_t)secA * 8];
    const uint32_t* wB = &t.sectors[(size_t)secB * 8];
    const uint32_t hA = wA[2] >> 16, hB = wB[2] >> 16;
    const int rmA = std::max(raA == S2_R_NONE ? -1 : (int)raA, rbA == S2_R_NONE ? -1 : (int)rbA);
    const int rmB = std::max(raB == S2_R_NONE ? -1 : (int)raB, rbB == S2_R_NONE ? -1 : (int)rbB);
    const bool onA = (hA & S2_H_MORE) && rmA >= (int)(hA & S2_THR_MASK);
    const bool onB = (hB & S2_H_MORE) && rmB >= (int)(hB & S2_THR_MASK);
    bool useB = true;
    if (hA & S2_H_EDGE) return false;
    if (two_chains) { if (onA || onB || (hB & S2_H_EDGE)) return false; }
    else { useB = onA; if (onA && onB) return false; }
    bool hitA[5], hitB[5];
    for (int i = 0; i < 5; ++i) {
        hitA[i] = stab2_entry_hit(wA, i, raA) || stab2_entry_hit(wA, i, rbA);
        hitB[i] = useB && (stab2_entry_hit(wB, i, raB) || stab2_entry_hit(wB, i, rbB));
    }
    if ((hA & S2_H_TW01) && hitA[0]) hitA[1] = false;
    if ((hA & S2_H_TW23) && hitA[2]) hitA[3] = false;
    if ((hB & S2_H_TW01) && hitB[0]) hitB[1] = false;
    if ((hB & S2_H_TW23) && hitB[2]) hitB[3] = false;
    uint32_t a[5], b[5];
    for (int i = 0; i < 5; ++i) {
        a[i] = hitA[i] ? stab2_entry_slot(wA, i) : 0x10000u + (uint32_t)i;
        b[i] = hitB[i] ? stab2_entry_slot(wB, i) : 0x20000u + (uint32_t)i;
    }
    for (int i = 0; i < 5; ++i) if (hitA[i]) out.push_back(a[i]);
    for (int j = 0; j < 5; ++j) {
        if (!hitB[j]) continue;
        bool dup = false;
        for (int i = 0; i < 5; ++i) dup |= b[j] == a[i];
        for (int i = 0; i < j; ++i) dup |= b[j] == b[i];
        if (!dup) out.push_back(b[j]);
    }
    std::sort(out.begin(), out.end());
    return true;
}
