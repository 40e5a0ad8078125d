/*
 * TEST INFRASTRUCTURE -- C++ restatement of the SINGLE-CELL loop of te_counter (oracle), for parity
 * checks at sizes the Python oracle cannot reach (tens of millions of records, real 1e7-key bundles).
 * Same rules as oracle/te_oracle.py, which it is checked against (tests/test_oracle_c.py): only tests,
 * smoke() and bench.py's parity legs may load it, as the checker.
 *
 * Literal restatement of te_count.py:298-707 with the canonical first-inserted rule:
 *   Part 1  dict (cell, umi) -> insertion-ordered list of distinct fragments, flushed into a sorted
 *           "bundle" when it holds bundle_keys keys (:372-491)
 *   Part 2  top maxcells + pad cells by raw count (stable), held-line scan over the bundles, first
 *           bundle wins (:494-575)
 *   Part 3  per line: later fragment per (chrom, strand) wins, ALL buckets of the range, inclusive
 *           point tests, distinct (ensg, strand), strand rule (:594-686)
 * It uses hash maps and the bucket hash, not the sort / scan / cell-table formulation of the kernels.
 */
#include <stdint.h>
#include <algorithm>
#include <map>
#include <unordered_map>
#include <vector>
#include <thread>
#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace {
struct Frag { uint32_t chrom; uint32_t strand; int32_t left, rite; };
struct Key { uint32_t cell; uint64_t umi; bool operator==(const Key& o) const { return cell == o.cell && umi == o.umi; } };
struct KeyHash { size_t operator()(const Key& k) const { uint64_t h = k.umi * 0x9E3779B97F4A7C15ULL ^ ((uint64_t)k.cell * 0xC2B2AE3D27D4EB4FULL); return (size_t)(h ^ (h >> 29)); } };
struct Line { uint32_t cell; uint64_t umi; std::vector<Frag> frags; };
inline int64_t fdiv(int64_t a, int64_t b) { int64_t q = a / b; return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q; }

struct Buckets {                     // genelist.py:367-380; bucket b*bs of chromosome c = list [off[c][b], off[c][b+1]) of ids
    int bs;
    std::vector<std::vector<int64_t>> off;       // per chromosome, n_buckets + 1
    std::vector<std::vector<int32_t>> ids;       // per chromosome, feature indices in linearData order
    std::vector<uint8_t> has;
};
}

struct teo_sc_result {
    std::vector<int32_t> t_ensg;
    std::vector<uint32_t> t_cell;
    std::vector<int64_t> t_count;
    std::vector<uint32_t> h_cell;       // insertion order of self.barcodes after Part 3
    std::vector<int64_t> h_count;
    int64_t stats[12];                   // total_reads, invalid_barcode, already_seen, lowq, qcfail, valid, assigned, raw_barcodes, n_bundles, crash_strand, zero_division
};

extern "C" {

void* teo_sc_count(int64_t n_feat, const int32_t* f_chrom, const int32_t* f_L, const int32_t* f_R, const int32_t* f_ensg,
                   const uint8_t* f_type, const uint8_t* f_strand, int n_chrom, int bs,
                   int qual, int strand_mode, int64_t bundle_keys, int64_t maxcells, int64_t pad,
                   int64_t n, const int32_t* start, const int32_t* end, const uint16_t* chrom, const uint8_t* mapq,
                   const uint8_t* flag, const uint32_t* cell, const uint64_t* umi) {
    teo_sc_result* res = new teo_sc_result();
    for (auto& s : res->stats) s = 0;
    Buckets B;
    B.bs = bs;
    B.off.resize((size_t)n_chrom);
    B.ids.resize((size_t)n_chrom);
    B.has.assign((size_t)n_chrom, 0);
    {
        std::vector<int64_t> nb((size_t)n_chrom, 0);
        for (int64_t i = 0; i < n_feat; ++i) {
            const int c = f_chrom[i];
            B.has[(size_t)c] = 1;
            const int64_t lb = fdiv(f_L[i], bs), rb = fdiv((int64_t)f_R[i] + bs, bs);     // buckets lb .. rb-1
            if (rb > lb && rb > nb[(size_t)c]) nb[(size_t)c] = rb;
        }
        for (int c = 0; c < n_chrom; ++c) B.off[(size_t)c].assign((size_t)nb[(size_t)c] + 2, 0);
        for (int64_t i = 0; i < n_feat; ++i) {
            const int c = f_chrom[i];
            const int64_t lb = fdiv(f_L[i], bs), rb = fdiv((int64_t)f_R[i] + bs, bs);
            for (int64_t b = std::max<int64_t>(lb, 0); b < rb; ++b) B.off[(size_t)c][(size_t)b + 1]++;
        }
        std::vector<std::vector<int64_t>> fill((size_t)n_chrom);
        for (int c = 0; c < n_chrom; ++c) {
            auto& o = B.off[(size_t)c];
            for (size_t b = 1; b < o.size(); ++b) o[b] += o[b - 1];
            B.ids[(size_t)c].resize((size_t)o.back());
            fill[(size_t)c].assign(o.begin(), o.end());
        }
        for (int64_t i = 0; i < n_feat; ++i) {
            const int c = f_chrom[i];
            const int64_t lb = fdiv(f_L[i], bs), rb = fdiv((int64_t)f_R[i] + bs, bs);
            for (int64_t b = std::max<int64_t>(lb, 0); b < rb; ++b) B.ids[(size_t)c][(size_t)fill[(size_t)c][(size_t)b]++] = (int32_t)i;
        }
    }
    auto T0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!getenv("TEO_TIMING")) return;
        auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "teo_sc %s %.2f s\n", what, std::chrono::duration<double>(t - T0).count());
        T0 = t;
    };
    lap("index");
    // ---- Part 1
    int64_t lowq = 0, qcfail = 0, invalid_barcode = 0, already_seen = 0, units = 0;
    std::vector<std::pair<uint32_t, int64_t>> barcodes;          // insertion ordered
    std::unordered_map<uint32_t, size_t> bc_pos;
    std::vector<std::vector<Line>> bundles;
    std::unordered_map<Key, std::vector<Frag>, KeyHash> umis;
    auto save_bundle = [&]() {
        std::vector<Line> L;
        L.reserve(umis.size());
        for (auto& kv : umis) L.push_back(Line{kv.first.cell, kv.first.umi, std::move(kv.second)});
        std::sort(L.begin(), L.end(), [](const Line& a, const Line& b) { return a.cell != b.cell ? a.cell < b.cell : a.umi < b.umi; });
        bundles.push_back(std::move(L));
        umis.clear();
    };
    auto bump = [&](uint32_t bc) {
        auto it = bc_pos.find(bc);
        if (it == bc_pos.end()) { bc_pos.emplace(bc, barcodes.size()); barcodes.push_back({bc, 1}); }
        else barcodes[it->second].second++;
    };
    int64_t pos = 0;
    for (;;) {
        units++;                                                         // :373
        if ((int64_t)umis.size() >= bundle_keys) save_bundle();          // :377
        if (pos >= n) break;                                             // :393 StopIteration
        const int64_t r = pos++;
        if (flag[r] & 7u) { qcfail++; continue; }                        // :394
        if ((int)mapq[r] < qual) { lowq++; continue; }                   // :398
        if (cell[r] == 0xFFFFFFFFu) { invalid_barcode++; continue; }     // :412
        if (chrom[r] == 0xFFFE) continue;                                // :432
        const Key key{cell[r], umi[r]};
        const uint32_t sc = strand_mode ? ((flag[r] & 8u) ? 1u : 0u) : 2u;      // :437-438
        const Frag f{chrom[r], sc, start[r], end[r]};
        auto it = umis.find(key);
        if (it != umis.end()) {                                          // :444
            const Frag& first = it->second.front();                      // :452 canonical: first inserted
            if (first.chrom == f.chrom && first.strand == f.strand) { already_seen++; continue; }
            bool present = false;
            for (const Frag& g : it->second) present |= g.chrom == f.chrom && g.strand == f.strand && g.left == f.left && g.rite == f.rite;
            if (!present) it->second.push_back(f);                       // :459 set.add
            bump(key.cell);                                              // :460-462
        } else {
            umis.emplace(key, std::vector<Frag>{f});                     // :470
            bump(key.cell);                                              // :471-473
        }
    }
    if (!umis.empty()) save_bundle();                                    // :479
    lap("part1");
    // ---- Part 2
    const int64_t n_raw = (int64_t)barcodes.size();
    std::vector<std::pair<uint32_t, int64_t>> order = barcodes;
    std::stable_sort(order.begin(), order.end(), [](const std::pair<uint32_t, int64_t>& a, const std::pair<uint32_t, int64_t>& b) { return a.second > b.second; });
    if ((int64_t)order.size() > maxcells + pad) order.resize((size_t)(maxcells + pad));
    std::vector<uint32_t> todo;
    for (auto& p : order) todo.push_back(p.first);
    std::sort(todo.begin(), todo.end());                                 // processed in ascending id (:503, pop from the end of the reversed list)
    struct Handle { size_t at; bool open; };
    std::vector<Handle> hs(bundles.size());
    for (size_t b = 0; b < bundles.size(); ++b) hs[b] = Handle{0, true};    // line = first line (:512-513)
    std::vector<Line> merged;
    int64_t umi_count = 0;
    for (uint32_t cur : todo) {
        std::vector<const Line*> this_data;
        for (size_t b = 0; b < bundles.size(); ++b) {
            Handle& h = hs[b];
            const std::vector<Line>& L = bundles[b];
            while (L[h.at].cell <= cur) {                                // :528
                if (!h.open) break;
                if (h.at + 1 < L.size()) {                               // next(): the held line is never kept
                    h.at++;
                    if (L[h.at].cell == cur) this_data.push_back(&L[h.at]);
                } else h.open = false;                                   // StopIteration
            }
        }
        umi_count += (int64_t)this_data.size();                          // :545
        std::unordered_map<uint64_t, size_t> seen;                       // umi -> index in merged (first bundle wins, :552-555)
        for (const Line* ln : this_data) {
            if (seen.find(ln->umi) == seen.end()) { seen.emplace(ln->umi, merged.size()); merged.push_back(*ln); }
        }
    }
    lap("part2");
    // ---- Part 3: the lines are independent; they are split over threads and the per-thread increments are
    //      merged afterwards (self.barcodes gets its keys in the order of the merged file = ascending cell id)
    struct P3 { std::vector<uint64_t> keys; std::vector<std::pair<uint32_t, int64_t>> hits; int64_t assigned = 0, crash = 0; };
    const int n_thr = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)std::thread::hardware_concurrency(), (int64_t)merged.size() / 4096 + 1));
    std::vector<P3> parts((size_t)n_thr);
    auto work = [&](int t) {
        P3& out = parts[(size_t)t];
        std::vector<uint32_t> stamp((size_t)n_feat + 1, 0);
        uint32_t tick = 0;
        const size_t lo = merged.size() * (size_t)t / (size_t)n_thr, hi = merged.size() * ((size_t)t + 1) / (size_t)n_thr;
        std::vector<Frag> reads;
        std::vector<int32_t> result;
        std::vector<std::pair<int32_t, uint8_t>> ensgs;
        for (size_t li = lo; li < hi; ++li) {
            const Line& ln = merged[li];
            reads.clear();                                           // dict keyed (chrom, strand): later one wins, first position kept
            for (const Frag& f : ln.frags) {
                bool rep = false;
                for (Frag& g : reads) if (g.chrom == f.chrom && g.strand == f.strand) { g.left = f.left; g.rite = f.rite; rep = true; }
                if (!rep) reads.push_back(f);
            }
            for (const Frag& f : reads) {
                if (f.chrom >= (uint32_t)n_chrom || !B.has[f.chrom]) continue;      // :614
                const int64_t left = f.left, rite = f.rite;
                const int64_t lbk = fdiv(left - 1, bs) * bs, rbk = fdiv(rite, bs) * bs;     // :619-620
                result.clear();
                if (++tick == 0) { std::fill(stamp.begin(), stamp.end(), 0u); tick = 1; }
                const auto& off = B.off[f.chrom];
                const auto& ids = B.ids[f.chrom];
                const int64_t n_b = (int64_t)off.size() - 2;            // bucket numbers 0 .. n_b may exist
                for (int64_t b = lbk / bs; b <= rbk / bs; ++b) {        // range(left_buck, right_buck + bs, bs)
                    if (b < 0 || b >= n_b + 1) continue;
                    for (int64_t q = off[(size_t)b]; q < off[(size_t)b + 1]; ++q) {
                        const int32_t i = ids[(size_t)q];
                        if (stamp[(size_t)i] == tick) continue;         // loc_ids is a set
                        stamp[(size_t)i] = tick;
                        if (left + 1 >= f_L[i] && left <= f_R[i]) result.push_back(i);          // :645
                        if (rite >= f_L[i] && rite - 1 <= f_R[i]) result.push_back(i);          // :648
                    }
                }
                if (result.empty()) continue;
                if (!out.hits.empty() && out.hits.back().first == ln.cell) out.hits.back().second++;    // :653-655
                else out.hits.push_back({ln.cell, 1});
                unsigned types = 0;
                bool missing = false;
                ensgs.clear();
                for (int32_t i : result) {
                    types |= 1u << f_type[i];
                    missing |= f_strand[i] == 255;
                    const std::pair<int32_t, uint8_t> e{f_ensg[i], f_strand[i]};
                    if (std::find(ensgs.begin(), ensgs.end(), e) == ensgs.end()) ensgs.push_back(e);
                }
                if (missing) { out.crash++; continue; }                  // :661 KeyError (the reference stops here)
                if (types & (1u << 1)) {                                 // gene branch
                    for (auto& e : ensgs) {
                        if (strand_mode && f.strand != e.second) continue;   // :665
                        out.keys.push_back(((uint64_t)(uint32_t)e.first << 32) | ln.cell);
                    }
                } else if (types & ((1u << 2) | (1u << 4))) {            // TE (:673) / enhancer (:679)
                    for (auto& e : ensgs) out.keys.push_back(((uint64_t)(uint32_t)e.first << 32) | ln.cell);
                } else continue;                                         // :684
                out.assigned++;
            }
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < n_thr; ++t) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
    }
    int64_t assigned = 0, crash = 0;
    std::unordered_map<uint64_t, int64_t> triples;           // ensg << 32 | cell
    for (auto& p : parts) {
        assigned += p.assigned; crash += p.crash;
        for (uint64_t k : p.keys) triples[k]++;
        for (auto& h : p.hits) {                              // merged is ordered by cell: a cell can only straddle two neighbours
            if (!res->h_cell.empty() && res->h_cell.back() == h.first) res->h_count.back() += h.second;
            else { res->h_cell.push_back(h.first); res->h_count.push_back(h.second); }
        }
    }
    lap("part3");
    {
        std::vector<std::pair<uint64_t, int64_t>> tv(triples.begin(), triples.end());
        std::sort(tv.begin(), tv.end());
        for (auto& kv : tv) { res->t_ensg.push_back((int32_t)(kv.first >> 32)); res->t_cell.push_back((uint32_t)kv.first); res->t_count.push_back(kv.second); }
    }
    res->stats[0] = units; res->stats[1] = invalid_barcode; res->stats[2] = already_seen; res->stats[3] = lowq; res->stats[4] = qcfail;
    res->stats[5] = umi_count; res->stats[6] = assigned; res->stats[7] = n_raw; res->stats[8] = (int64_t)bundles.size();
    res->stats[9] = crash; res->stats[10] = umi_count == 0;
    return res;
}

int64_t teo_sc_n_triples(void* p) { return (int64_t)((teo_sc_result*)p)->t_ensg.size(); }
int64_t teo_sc_n_hit(void* p) { return (int64_t)((teo_sc_result*)p)->h_cell.size(); }
void teo_sc_fetch(void* p, int32_t* ensg, uint32_t* cell, int64_t* count, uint32_t* hcell, int64_t* hcount, int64_t* stats) {
    teo_sc_result* r = (teo_sc_result*)p;
    std::copy(r->t_ensg.begin(), r->t_ensg.end(), ensg);
    std::copy(r->t_cell.begin(), r->t_cell.end(), cell);
    std::copy(r->t_count.begin(), r->t_count.end(), count);
    std::copy(r->h_cell.begin(), r->h_cell.end(), hcell);
    std::copy(r->h_count.begin(), r->h_count.end(), hcount);
    for (int i = 0; i < 12; ++i) stats[i] = r->stats[i];
}
void teo_sc_free(void* p) { delete (teo_sc_result*)p; }

}
