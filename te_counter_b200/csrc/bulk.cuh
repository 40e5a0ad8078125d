// Bulk counting kernels: filter (+ mate merge), point overlap, type rule, tally.
//
// Reference semantics (te_counter, te_count/te_count.py):
//   filter SE :203-218, filter + mate merge PE :76-102, candidate buckets :106-116 / :222-231,
//   point tests :118-126 / :233-241, type rule + tally :128-149 / :243-261.
//
// Closed forms used here, with bs = bucket size:
//   point A  loc1 in [L, R-1]   <=>  x1 = loc1     stabs the half-open interval [L, R)
//   point B  loc2 in [L+1, R]   <=>  x2 = loc2 - 1 stabs [L, R)
//   candidate(f) <=> L//bs <= b1 <= R//bs  or  L//bs <= b2 <= R//bs,
//                    b1 = (loc1-1)//bs, b2 = (loc2+1)//bs   (the two buckets of :106-108)
//   it can only be false when loc1 == L (A) or loc2 == R (B), so the divisions are off the
//   common path.
#pragma once
#include "common.cuh"

// number of features of chromosome [lo, lo+n) with L <= x, through the coarse directory
__device__ __forceinline__ int upper_bound_L(const IndexView& iv, int c, int64_t lo, int n, int x) {
    if (x < 0) return 0;                              // L >= 0 is enforced at upload
    const int64_t doff = iv.dir_off[c];
    const int ncell = (int)(iv.dir_off[c + 1] - doff);
    const int k = x >> iv.shift;
    if (k >= ncell - 1) return n;
    u32 a = __ldg(iv.dir + doff + k), b = __ldg(iv.dir + doff + k + 1);
    while (a < b) {
        u32 mid = (a + b) >> 1;
        if (__ldg(iv.L + lo + mid) <= x) a = mid + 1; else b = mid;
    }
    return (int)a;
}

__device__ __forceinline__ bool bulk_candidate(int Lk, int Rk, int loc1, int loc2, int b1, int b2,
                                               bool hitA, bool hitB, int bs) {
    if ((hitA && loc1 != Lk) || (hitB && loc2 != Rk)) return true;
    const int lb = Lk / bs, rb = Rk / bs;
    return (lb <= b1 && b1 <= rb) || (lb <= b2 && b2 <= rb);
}

// Calls f(feature_index) exactly once for every feature the reference appends to `result`
// (te_count.py:118-126); returns false as soon as f returns false.
template <class F>
__device__ __forceinline__ bool bulk_for_each_hit(const IndexView& iv, int c, int loc1, int loc2, F f) {
    const int64_t lo = iv.chrom_off[c];
    const int n = (int)(iv.chrom_off[c + 1] - lo);
    const int bs = iv.bs;
    const int b1 = floordiv(loc1 - 1, bs), b2 = floordiv(loc2 + 1, bs);
    {   // point A
        const int x = loc1;
        for (int k = upper_bound_L(iv, c, lo, n, x) - 1; k >= 0; --k) {
            if (__ldg(iv.pmaxR + lo + k) <= x) break;
            const int Rk = __ldg(iv.R + lo + k);
            if (Rk > x) {
                const int Lk = __ldg(iv.L + lo + k);
                const bool hitB = (Lk < loc2 && loc2 <= Rk);
                if (bulk_candidate(Lk, Rk, loc1, loc2, b1, b2, true, hitB, bs))
                    if (!f(lo + k)) return false;
            }
        }
    }
    {   // point B
        const int x = loc2 - 1;
        for (int k = upper_bound_L(iv, c, lo, n, x) - 1; k >= 0; --k) {
            if (__ldg(iv.pmaxR + lo + k) <= x) break;
            const int Rk = __ldg(iv.R + lo + k);
            if (Rk > x) {
                const int Lk = __ldg(iv.L + lo + k);
                if (Lk <= loc1 && loc1 < Rk) continue;          // enumerated under point A
                if (bulk_candidate(Lk, Rk, loc1, loc2, b1, b2, false, true, bs))
                    if (!f(lo + k)) return false;
            }
        }
    }
    return true;
}

#define BULK_MAX_DISTINCT 12

struct BulkStatsLocal {
    u32 units, assigned, lowq, badchrom, qcfail, crash_enh, crash_name;
};

// One thread per unit (record in SE, pair in PE), grid-stride.  Per-feature counters are int64 in
// global memory (L2-resident, 8 B x n_ensg); statistics are reduced per warp, then per CTA.
template <bool PAIRED>
__global__ void __launch_bounds__(256)
bulk_count_kernel(IndexView iv, int64_t n_units, int qual,
                  const int32_t* __restrict__ start, const int32_t* __restrict__ end,
                  const uint16_t* __restrict__ chrom, const uint8_t* __restrict__ mapq,
                  const uint8_t* __restrict__ flag, u64* __restrict__ counts, u64* __restrict__ stats) {
    BulkStatsLocal st = {0, 0, 0, 0, 0, 0, 0};
    const u32 reject = TEC_F_UNMAPPED | TEC_F_DUP | TEC_F_QCFAIL;
    for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < n_units;
         u += (int64_t)gridDim.x * blockDim.x) {
        st.units++;
        int c, loc1, loc2;
        if (PAIRED) {
            const uchar2 f2 = *reinterpret_cast<const uchar2*>(flag + 2 * u);
            if ((f2.x & reject) || (f2.y & reject)) { st.qcfail++; continue; }       // :81-86
            if ((int)mapq[2 * u] < qual) { st.lowq++; continue; }                      // :88 read1 only
            if (f2.x & TEC_F_NAME_MISMATCH) { st.crash_name++; continue; }             // :92-94
            c = chrom[2 * u];                                                          // :96 read1 only
            const int2 s2 = *reinterpret_cast<const int2*>(start + 2 * u);
            loc1 = s2.x;                                                               // :97
            loc2 = s2.y;                                                               // :98 mate START
        } else {
            if (flag[u] & reject) { st.qcfail++; continue; }                           // :204
            if ((int)mapq[u] < qual) { st.lowq++; continue; }                          // :208
            c = chrom[u];
            loc1 = start[u];                                                           // :213
            loc2 = end[u];                                                             // :214
        }
        if (c >= iv.n_chrom) { st.badchrom++; continue; }                              // :100 / :216

        u32 typemask = 0, nd = 0, dist[BULK_MAX_DISTINCT];
        bool overflow = false;
        bulk_for_each_hit(iv, c, loc1, loc2, [&](int64_t fi) {
            const u32 w = __ldg(iv.info + fi);
            typemask |= 1u << info_type(w);
            const u32 e = info_ensg(w);
            bool found = false;
            for (u32 i = 0; i < nd; ++i) found |= (dist[i] == e);
            if (!found) {
                if (nd < BULK_MAX_DISTINCT) dist[nd++] = e; else overflow = true;
            }
            return true;
        });
        if (!typemask) continue;                                                       // :128 no result
        st.assigned++;                                                                 // :149
        const u32 counted = (1u << TEC_T_GENE) | (1u << TEC_T_TE) | (1u << TEC_T_SNRNA);
        if (!(typemask & counted)) {
            if (typemask & (1u << TEC_T_ENHANCER)) st.crash_enh++;                     // :145-147
            continue;
        }
        if (!overflow) {
            for (u32 i = 0; i < nd; ++i) atomicAdd(counts + dist[i], 1ULL);            // one per distinct ensg
        } else {
            // more distinct ensg than the register list holds: count a hit iff no earlier hit
            // (in enumeration order) carries the same ensg -- O(h^2) re-walks, no storage
            int h = 0;
            bulk_for_each_hit(iv, c, loc1, loc2, [&](int64_t fi) {
                const u32 e = info_ensg(__ldg(iv.info + fi));
                int j = 0;
                bool dup = false;
                bulk_for_each_hit(iv, c, loc1, loc2, [&](int64_t fj) {
                    if (j++ >= h) return false;
                    if (info_ensg(__ldg(iv.info + fj)) == e) { dup = true; return false; }
                    return true;
                });
                if (!dup) atomicAdd(counts + e, 1ULL);
                ++h;
                return true;
            });
        }
    }
    // statistics: warp shuffle -> shared -> one global atomic per CTA per counter
    __shared__ u64 s_stats[TEC_BULK_NSTATS];
    if (threadIdx.x < TEC_BULK_NSTATS) s_stats[threadIdx.x] = 0;
    __syncthreads();
    u64 v[7] = {st.units, st.assigned, st.lowq, st.badchrom, st.qcfail, st.crash_enh, st.crash_name};
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        const u64 s = warp_sum(v[i]);
        if ((threadIdx.x & 31) == 0 && s) atomicAdd(&s_stats[i], s);
    }
    __syncthreads();
    if (threadIdx.x < 7 && s_stats[threadIdx.x]) atomicAdd(stats + threadIdx.x, s_stats[threadIdx.x]);
}

// =====================================================================================================
// Fast path: cell table (stab_build.h) + shared-memory counters for the hottest ensg.
//
// Per unit: 16 B (PE; `end` is never read) / 12 B (SE) of streamed records (evict-first, no L1
// allocate), ONE 32-byte sector of the L2-resident cell table fetched with a single 256-bit load
// (two when the unit's points fall in different cells, one more when a crowded cell links to an
// overflow sector the point can reach), and one shared-memory or global atomic per distinct ensg.
// Units whose candidate test can fail (loc1 % bs == 0 or (loc2 + 1) % bs == 0, see the header of
// this file) or that hit more than STAB_MAXD distinct ensg are only flagged here (one ballot word per
// warp) and counted by bulk_slow_kernel with the exact search above; both are rare.
// =====================================================================================================
struct StabView {
    const u32* sectors;          // 8 words per sector, 32-byte aligned
    const uint2* cells;          // per chromosome: {first sector, number of cells}
    const uint8_t* slot_type;    // n_slots
    int shift;
    int all_counted;
};

#define STAB_MAXD 4
#define TEC_HOT_SLOTS 4096
#define BULK_THREADS 512
#ifndef BULK_MIN_CTAS
#define BULK_MIN_CTAS 2
#endif
#define BULK_WARPS (BULK_THREADS / 32)
#define BULK_QCAP 64             // deferred-unit ring per warp (entries)
#define R_NONE 0xFFFFu           // "no point in this sector": (R_NONE - start) > any length

struct Sector { u32 w[8]; };

__device__ __forceinline__ Sector ld_sector(const u32* sectors, u32 idx) {
    Sector r;
    const u32* p = sectors + (size_t)idx * 8;
    asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
                 : "l"(p));
    return r;
}

// streamed record loads: read once, keep them out of L1 and first in line for L2 eviction
__device__ __forceinline__ u64 make_evict_first_policy() {
    u64 pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ int2 ld_stream_int2(const int32_t* p, u64 pol) {
    int2 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.s32 {%0,%1}, [%2], %3;" : "=r"(r.x), "=r"(r.y) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ int ld_stream_int(const int32_t* p, u64 pol) {
    int r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ u32 ld_stream_u32(const void* p, u64 pol) {
    u32 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ u32 ld_stream_u16(const void* p, u64 pol) {
    unsigned short r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u16 %0, [%1], %2;" : "=h"(r) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ u32 ld_stream_u8(const uint8_t* p, u64 pol) {
    u32 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
    return r;
}

// bit i set <=> entry i of the sector contains point ra or rb (cell-relative, R_NONE = no point).
// last_s = start of entry 5 (the link test of stab_build.h).
__device__ __forceinline__ u32 sector_hits(const Sector& s, u32 ra, u32 rb, u32& last_s) {
    const u32 M = (1u << 22) - 1;
    u32 pos[6];
    pos[0] = s.w[3] & M;
    pos[1] = __funnelshift_r(s.w[3], s.w[4], 22) & M;
    pos[2] = __funnelshift_r(s.w[4], s.w[5], 12) & M;
    pos[3] = (s.w[5] >> 2) & M;
    pos[4] = __funnelshift_r(s.w[5], s.w[6], 24) & M;
    pos[5] = __funnelshift_r(s.w[6], s.w[7], 14) & M;
    u32 hit = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const u32 st = pos[i] & 2047u, lm1 = pos[i] >> 11;
        hit |= (u32)((ra - st <= lm1) | (rb - st <= lm1)) << i;
    }
    last_s = pos[5] & 2047u;
    const u32 code = (s.w[7] >> 4) & 7u;
    return hit & ((code == 7u) ? 63u : ((1u << code) - 1u));
}
__device__ __forceinline__ u32 sector_slot(const Sector& s, int i) {
    const u32 pair = (i < 2) ? s.w[0] : ((i < 4) ? s.w[1] : s.w[2]);
    return (pair >> ((i & 1) << 4)) & 0xFFFFu;
}
template <int I>
__device__ __forceinline__ u32 sector_slot_c(const Sector& s) {
    return (I & 1) ? (s.w[I >> 1] >> 16) : (s.w[I >> 1] & 0xFFFFu);
}
__device__ __forceinline__ bool sector_has_link(const Sector& s) { return ((s.w[7] >> 4) & 7u) == 7u; }
__device__ __forceinline__ bool sector_has_dup(const Sector& s) { return (s.w[7] >> 7) & 1u; }
__device__ __forceinline__ u32 sector_link(const Sector& s) { return s.w[7] >> 8; }

// distinct ensg slots of one unit (registers only)
struct SlotSet {
    u32 v0, v1, v2, v3;
    int n;
    __device__ __forceinline__ void clear() { v0 = v1 = v2 = v3 = 0xFFFFFFFFu; n = 0; }
    // branch-free: `on` gates the whole insertion
    __device__ __forceinline__ void add(u32 w, bool on) {
        const bool fresh = on & !((v0 == w) | (v1 == w) | (v2 == w) | (v3 == w));
        v0 = (fresh & (n == 0)) ? w : v0;
        v1 = (fresh & (n == 1)) ? w : v1;
        v2 = (fresh & (n == 2)) ? w : v2;
        v3 = (fresh & (n == 3)) ? w : v3;
        n += fresh;                                      // n > STAB_MAXD: overflow
    }
};

// x % 10000 == 0 for x >= 0 (negative x answers true: the exact kernel then decides):
// 10000 = 16 * 625, and for odd d  n % d == 0  <=>  n * d^-1 (mod 2^32) <= (2^32 - 1) / d
__device__ __forceinline__ bool mult_of_10000(int x) {
    const u32 ux = (u32)x;
    return (x < 0) | (((ux & 15u) == 0) & ((ux >> 4) * 0x3AFB7E91u <= 0xFFFFFFFFu / 625u));
}

struct BulkRec {
    u32 fl, q;
    int c, loc1, loc2;
};

template <bool PAIRED>
__device__ __forceinline__ BulkRec bulk_load(int64_t u, const int32_t* __restrict__ start, const int32_t* __restrict__ end,
                                             const uint16_t* __restrict__ chrom, const uint8_t* __restrict__ mapq,
                                             const uint8_t* __restrict__ flag, u64 pol) {
    BulkRec r;
    if (PAIRED) {
        r.fl = ld_stream_u16(flag + 2 * u, pol);                    // both mates' flag bytes
        r.q = ld_stream_u16(mapq + 2 * u, pol) & 0xFFu;             // read1 only (:88)
        r.c = (int)(ld_stream_u32(chrom + 2 * u, pol) & 0xFFFFu);   // read1 only (:96)
        const int2 s2 = ld_stream_int2(start + 2 * u, pol);
        r.loc1 = s2.x;                                              // :97
        r.loc2 = s2.y;                                              // :98 mate START
    } else {
        r.fl = ld_stream_u8(flag + u, pol);
        r.q = ld_stream_u8(mapq + u, pol);
        r.c = (int)ld_stream_u16(chrom + u, pol);
        r.loc1 = ld_stream_int(start + u, pol);                     // :213
        r.loc2 = ld_stream_int(end + u, pol);                       // :214
    }
    return r;
}

// a unit that needs a second sector (its points straddle two cells, or it can reach the overflow
// sector), or that hit two entries of a sector holding one ensg twice: parked in the warp's ring
// and looked up 32 at a time by bulk_deferred(), so the common path stays branch-light
struct QEnt {
    u32 secA;       // sector | kind << 24   (kind 0: cell A only, 1: a second cell B)
    u32 secB;
    u32 pa;         // points in A:  ra | rb << 16   (R_NONE = none)
    u32 pb;         // points in B
};

struct BulkShared {
    u32 hot[TEC_HOT_SLOTS];
    u64 stats[TEC_BULK_NSTATS];
    QEnt q[BULK_WARPS][BULK_QCAP];
    u32 qu[BULK_WARPS][BULK_QCAP];       // unit index inside the launch (for the slow flag)
};

// +1 for entry I of the sector when its hit bit is set: predicated reductions, no branches
template <int I>
__device__ __forceinline__ void bump_entry(u32 hit, const Sector& s, u32 hot_addr, u64* __restrict__ counts, u32 one) {
    const u32 slot = sector_slot_c<I>(s);
    asm volatile("{\n\t.reg .pred p, q, r;\n\t.reg .b32 t;\n\t"
                 "and.b32 t, %0, %1;\n\t"
                 "setp.ne.u32 p, t, 0;\n\t"
                 "setp.lt.u32 q, %2, %3;\n\t"
                 "and.pred r, p, q;\n\t"
                 "@r red.shared.add.u32 [%4], %6;\n\t"
                 "not.pred q, q;\n\t"
                 "and.pred r, p, q;\n\t"
                 "@r red.global.add.u64 [%5], %7;\n\t}"
                 :: "r"(hit), "n"(1 << I), "r"(slot), "n"(TEC_HOT_SLOTS), "r"(hot_addr + slot * 4u), "l"(counts + slot),
                    "r"(one), "l"((u64)one) : "memory");
}

__device__ __forceinline__ void bulk_bump(BulkShared& sh, u64* __restrict__ counts, u32 slot) {
    if (slot < TEC_HOT_SLOTS) atomicAdd(&sh.hot[slot], 1u);
    else atomicAdd(counts + slot, 1ULL);
}

// slow list: word 0 = number of flagged units, then their indices (capacity = units of the launch)
__device__ __forceinline__ void flag_slow(u32* __restrict__ slow_list, u32 u) { slow_list[1 + atomicAdd(slow_list, 1u)] = u; }

// type rule of te_count.py:134-147 over the distinct slots of a unit; returns "count it"
__device__ __forceinline__ bool bulk_type_rule(const StabView& sv, u32 typemask, u64* __restrict__ stats) {
    const u32 counted = (1u << TEC_T_GENE) | (1u << TEC_T_TE) | (1u << TEC_T_SNRNA);
    if (typemask & counted) return true;
    if (typemask & (1u << TEC_T_ENHANCER)) atomicAdd(stats + TEC_BS_CRASH_ENHANCER, 1ULL);       // :145-147
    return false;
}

// all entries of a sector chain that contain point ra or rb go into S
__device__ __forceinline__ void chain_collect(const StabView& sv, u32 sec, u32 ra, u32 rb, SlotSet& S) {
    const int rmax = max((int)(short)ra, (int)(short)rb);                              // R_NONE -> -1
    for (;;) {
        const Sector s = ld_sector(sv.sectors, sec);
        u32 last_s;
        const u32 hit = sector_hits(s, ra, rb, last_s);
        if (hit) {
            S.add(sector_slot_c<0>(s), hit & 1u);
            S.add(sector_slot_c<1>(s), hit & 2u);
            S.add(sector_slot_c<2>(s), hit & 4u);
            S.add(sector_slot_c<3>(s), hit & 8u);
            S.add(sector_slot_c<4>(s), hit & 16u);
            S.add(sector_slot_c<5>(s), hit & 32u);
        }
        if (!sector_has_link(s) || rmax < (int)last_s) break;
        sec = sector_link(s);
    }
}

__device__ __forceinline__ void bulk_deferred(const StabView& sv, BulkShared& sh, const QEnt e, u32 u, bool live,
                                              u64* __restrict__ counts, u64* __restrict__ stats,
                                              u32* __restrict__ slow_list, u32& n_assigned) {
    if (!live) return;
    SlotSet S;
    S.clear();
    chain_collect(sv, e.secA & 0xFFFFFFu, e.pa & 0xFFFFu, e.pa >> 16, S);
    if (e.secA >> 24) chain_collect(sv, e.secB, e.pb & 0xFFFFu, e.pb >> 16, S);
    const bool slow = false;
    if (slow || S.n > STAB_MAXD) { flag_slow(slow_list, u); return; }
    if (!S.n) return;                                                                  // :128 no result
    n_assigned++;                                                                      // :149
    if (!sv.all_counted) {
        u32 typemask = 1u << __ldg(sv.slot_type + S.v0);
        if (S.n > 1) typemask |= 1u << __ldg(sv.slot_type + S.v1);
        if (S.n > 2) typemask |= 1u << __ldg(sv.slot_type + S.v2);
        if (S.n > 3) typemask |= 1u << __ldg(sv.slot_type + S.v3);
        if (!bulk_type_rule(sv, typemask, stats)) return;
    }
    bulk_bump(sh, counts, S.v0);
    if (S.n > 1) bulk_bump(sh, counts, S.v1);
    if (S.n > 2) bulk_bump(sh, counts, S.v2);
    if (S.n > 3) bulk_bump(sh, counts, S.v3);
}

// One warp per 32 consecutive units, grid-stride; the next warp-tile's records are requested before
// the current one is looked up.
template <bool PAIRED>
__global__ void __launch_bounds__(BULK_THREADS, BULK_MIN_CTAS)
bulk_count_cell_kernel(IndexView iv, StabView sv, int64_t n_units, int qual,
                       const int32_t* __restrict__ start, const int32_t* __restrict__ end,
                       const uint16_t* __restrict__ chrom, const uint8_t* __restrict__ mapq,
                       const uint8_t* __restrict__ flag, u64* __restrict__ counts, u64* __restrict__ stats,
                       u32* __restrict__ slow_list) {
    __shared__ BulkShared sh;
    for (int i = threadIdx.x; i < TEC_HOT_SLOTS; i += blockDim.x) sh.hot[i] = 0;
    if (threadIdx.x < TEC_BULK_NSTATS) sh.stats[threadIdx.x] = 0;
    __syncthreads();
    u32 n_assigned = 0, n_lowq = 0, n_badchrom = 0, n_qcfail = 0;
    const u32 reject = TEC_F_UNMAPPED | TEC_F_DUP | TEC_F_QCFAIL;
    const u32 reject2 = PAIRED ? (reject | (reject << 8)) : reject;
    const u64 pol = make_evict_first_policy();
    const int bs = iv.bs;
    const int shift = sv.shift;
    const u32 cmask = (1u << shift) - 1;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const u32 lt_mask = (1u << lane) - 1u;
    const u32 hot_addr = (u32)__cvta_generic_to_shared(&sh.hot[0]);
    // the increment as a run-time value: a literal 1 makes ptxas pick ATOMS.POPC.INC, which needs a
    // converged warp and therefore a branch around every reduction
    const u32 one = (u32)(n_units > 0);
    QEnt* const ring = sh.q[wib];
    u32* const ring_u = sh.qu[wib];
    u32 q_head = 0, q_count = 0;                     // warp-uniform
    const int64_t n_tiles = (n_units + 31) >> 5;
    const int64_t tile_stride = (int64_t)gridDim.x * BULK_WARPS;
    int64_t tile = (int64_t)blockIdx.x * BULK_WARPS + wib;
    BulkRec cur, nxt;
    cur.fl = cur.q = 0; cur.c = cur.loc1 = cur.loc2 = 0;
    nxt = cur;
    if (tile < n_tiles && tile * 32 + lane < n_units) cur = bulk_load<PAIRED>(tile * 32 + lane, start, end, chrom, mapq, flag, pol);
    for (; tile < n_tiles; tile += tile_stride) {
        const int64_t u = tile * 32 + lane;
        const int64_t un = u + tile_stride * 32;
        if (un < n_units) nxt = bulk_load<PAIRED>(un, start, end, chrom, mapq, flag, pol);
        // ---- filter (te_count.py:78-102 / :203-218)
        const int c = cur.c, loc1 = cur.loc1, loc2 = cur.loc2;
        bool look = false;
        if (u < n_units) {
            if (cur.fl & reject2) n_qcfail++;                                              // :81-86 / :204
            else if ((int)cur.q < qual) n_lowq++;                                          // :88 / :208
            else if (PAIRED && (cur.fl & TEC_F_NAME_MISMATCH)) atomicAdd(stats + TEC_BS_CRASH_NAME, 1ULL);   // :92-94
            else if (c >= iv.n_chrom) n_badchrom++;                                        // :100 / :216
            else look = true;
        }
        // ---- which sector(s)
        QEnt qe;
        qe.secA = qe.secB = 0; qe.pa = qe.pb = R_NONE | (R_NONE << 16);
        bool defer = false, single = false;
        if (look) {
            const bool edge = (bs == 10000) ? (mult_of_10000(loc1) | mult_of_10000(loc2 + 1))
                                            : ((loc1 % bs == 0) || ((loc2 + 1) % bs == 0));
            if (edge) flag_slow(slow_list, (u32)u);
            else {
                const uint2 cell = __ldg(sv.cells + c);
                const int xa = loc1, xb = loc2 - 1;
                const int ka = xa >> shift, kb = xb >> shift;                    // arithmetic shift: negative stays negative
                const bool va = (u32)ka < cell.y, vb = (u32)kb < cell.y;         // negative -> huge -> false
                const u32 ra = (u32)xa & cmask, rb = (u32)xb & cmask;
                if (va && vb && ka != kb) {
                    qe.secA = (cell.x + (u32)ka) | (1u << 24);
                    qe.secB = cell.x + (u32)kb;
                    qe.pa = ra | (R_NONE << 16);
                    qe.pb = R_NONE | (rb << 16);
                    defer = true;
                } else if (va || vb) {
                    qe.secA = cell.x + (u32)(va ? ka : kb);
                    qe.pa = (va ? ra : R_NONE) | ((vb ? rb : R_NONE) << 16);
                    single = true;
                }
            }
        }
        // ---- the common case: one sector
        if (single) {
            const Sector s = ld_sector(sv.sectors, qe.secA);
            u32 last_s;
            u32 hit = sector_hits(s, qe.pa & 0xFFFFu, qe.pa >> 16, last_s);
            const int rmax = max((int)(short)(qe.pa & 0xFFFFu), (int)(short)(qe.pa >> 16));
            if (sector_has_link(s) && rmax >= (int)last_s) {
                defer = true;                                                              // the chain is walked again from A
            } else if (sector_has_dup(s) && (hit & (hit - 1))) {
                defer = true;
            } else if (hit) {                                                              // :128 result not empty
                n_assigned++;                                                              // :149
                bool count_it = true;
                if (!sv.all_counted) {
                    u32 typemask = 0, h = hit;
                    while (h) { const int i = __ffs(h) - 1; h &= h - 1; typemask |= 1u << __ldg(sv.slot_type + sector_slot(s, i)); }
                    count_it = bulk_type_rule(sv, typemask, stats);
                }
                if (count_it) {
                    bump_entry<0>(hit, s, hot_addr, counts, one);
                    bump_entry<1>(hit, s, hot_addr, counts, one);
                    bump_entry<2>(hit, s, hot_addr, counts, one);
                    bump_entry<3>(hit, s, hot_addr, counts, one);
                    bump_entry<4>(hit, s, hot_addr, counts, one);
                    bump_entry<5>(hit, s, hot_addr, counts, one);
                }
            }
        }
        // ---- park deferred units; look a full warp of them up when there are 32
        const u32 dm = __ballot_sync(0xFFFFFFFFu, defer);
        if (dm) {
            if (defer) {
                const u32 at = (q_head + q_count + __popc(dm & lt_mask)) & (BULK_QCAP - 1);
                ring[at] = qe;
                ring_u[at] = (u32)u;
            }
            q_count += __popc(dm);
            __syncwarp();
            if (q_count >= 32) {
                const u32 at = (q_head + lane) & (BULK_QCAP - 1);
                bulk_deferred(sv, sh, ring[at], ring_u[at], true, counts, stats, slow_list, n_assigned);
                q_head = (q_head + 32) & (BULK_QCAP - 1);
                q_count -= 32;
                __syncwarp();
            }
        }
        cur = nxt;
    }
    if (q_count) {
        const u32 at = (q_head + lane) & (BULK_QCAP - 1);
        bulk_deferred(sv, sh, ring[at], ring_u[at], (u32)lane < q_count, counts, stats, slow_list, n_assigned);
    }
    u64 v[4] = {n_assigned, n_lowq, n_badchrom, n_qcfail};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const u64 s = warp_sum(v[i]);
        if (lane == 0 && s) atomicAdd(&sh.stats[TEC_BS_ASSIGNED + i], s);
    }
    __syncthreads();
    if (threadIdx.x >= TEC_BS_ASSIGNED && threadIdx.x < TEC_BS_ASSIGNED + 4 && sh.stats[threadIdx.x])
        atomicAdd(stats + threadIdx.x, sh.stats[threadIdx.x]);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(stats + TEC_BS_UNITS, (u64)n_units);
    for (int i = threadIdx.x; i < TEC_HOT_SLOTS; i += blockDim.x) {
        const u32 x = sh.hot[i];
        if (x) atomicAdd(counts + i, (u64)x);
    }
}

// Units flagged by the fast kernel: bucket-edge candidates take the exact search (counters in slot
// space); units with more than STAB_MAXD distinct ensg walk the cell table again with a larger set.
// One thread per flagged unit.
#define SLOW_MAXD 32
#define SLOW_MAXD_EXACT 96
template <bool PAIRED>
__global__ void __launch_bounds__(256)
bulk_slow_kernel(IndexView iv, StabView sv, int has_stab, const int32_t* __restrict__ start, const int32_t* __restrict__ end,
                 const uint16_t* __restrict__ chrom, u64* __restrict__ counts, u64* __restrict__ stats,
                 const u32* __restrict__ slow_list) {
    const u32 n = slow_list[0];
    const u32 counted = (1u << TEC_T_GENE) | (1u << TEC_T_TE) | (1u << TEC_T_SNRNA);
    for (u32 t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const int64_t u = slow_list[1 + t];
        int c, loc1, loc2;
        if (PAIRED) { c = chrom[2 * u]; loc1 = start[2 * u]; loc2 = start[2 * u + 1]; }
        else { c = chrom[u]; loc1 = start[u]; loc2 = end[u]; }
        const bool edge = loc1 < 0 || loc2 + 1 < 0 || (loc1 % iv.bs == 0) || ((loc2 + 1) % iv.bs == 0);
        if (has_stab && !edge) {
            // same lookup as the fast kernel, distinct slots in a local list
            u32 nd = 0, dist[SLOW_MAXD];
            bool overflow = false;
            const uint2 cell = __ldg(sv.cells + c);
            const int x[2] = {loc1, loc2 - 1};
            for (int p = 0; p < 2; ++p) {
                const int k = x[p] >> sv.shift;
                if ((u32)k >= cell.y) continue;
                const u32 r = (u32)x[p] & ((1u << sv.shift) - 1);
                u32 sec = cell.x + (u32)k;
                for (;;) {
                    const Sector s = ld_sector(sv.sectors, sec);
                    u32 last_s;
                    u32 hit = sector_hits(s, r, R_NONE, last_s);
                    while (hit) {
                        const int i = __ffs(hit) - 1;
                        hit &= hit - 1;
                        const u32 e = sector_slot(s, i);
                        bool found = false;
                        for (u32 j = 0; j < nd; ++j) found |= (dist[j] == e);
                        if (!found) {
                            if (nd < SLOW_MAXD) dist[nd++] = e; else overflow = true;
                        }
                    }
                    if (!sector_has_link(s) || r < last_s) break;
                    sec = sector_link(s);
                }
            }
            if (!overflow) {
                if (!nd) continue;                                                     // :128 no result
                atomicAdd(stats + TEC_BS_ASSIGNED, 1ULL);                              // :149
                u32 typemask = sv.all_counted ? counted : 0u;
                if (!sv.all_counted)
                    for (u32 j = 0; j < nd; ++j) typemask |= 1u << __ldg(sv.slot_type + dist[j]);
                if (!(typemask & counted)) {
                    if (typemask & (1u << TEC_T_ENHANCER)) atomicAdd(stats + TEC_BS_CRASH_ENHANCER, 1ULL);
                    continue;
                }
                for (u32 j = 0; j < nd; ++j) atomicAdd(counts + dist[j], 1ULL);
                continue;
            }
        }
        // one walk: type mask + the first SLOW_MAXD_EXACT distinct ensg
        u32 typemask = 0, nd = 0, dist[SLOW_MAXD_EXACT];
        bool overflow = false;
        bulk_for_each_hit(iv, c, loc1, loc2, [&](int64_t fi) {
            const u32 w = __ldg(iv.info + fi);
            typemask |= 1u << info_type(w);
            const u32 e = info_ensg(w);
            bool found = false;
            for (u32 i = 0; i < nd; ++i) found |= (dist[i] == e);
            if (!found) {
                if (nd < SLOW_MAXD_EXACT) dist[nd++] = e; else overflow = true;
            }
            return true;
        });
        if (!typemask) continue;                                                       // :128 no result
        atomicAdd(stats + TEC_BS_ASSIGNED, 1ULL);                                      // :149
        if (!(typemask & counted)) {
            if (typemask & (1u << TEC_T_ENHANCER)) atomicAdd(stats + TEC_BS_CRASH_ENHANCER, 1ULL);
            continue;
        }
        if (!overflow) {
            for (u32 i = 0; i < nd; ++i) atomicAdd(counts + dist[i], 1ULL);            // one per distinct ensg
            continue;
        }
        // more distinct ensg than the list holds: count a hit iff no earlier hit (in enumeration
        // order) carries the same ensg -- O(h^2) re-walks, no storage
        int h = 0;
        bulk_for_each_hit(iv, c, loc1, loc2, [&](int64_t fi) {
            const u32 e = info_ensg(__ldg(iv.info + fi));
            int j = 0;
            bool dup = false;
            bulk_for_each_hit(iv, c, loc1, loc2, [&](int64_t fj) {
                if (j++ >= h) return false;
                if (info_ensg(__ldg(iv.info + fj)) == e) { dup = true; return false; }
                return true;
            });
            if (!dup) atomicAdd(counts + e, 1ULL);
            ++h;
            return true;
        });
    }
}
