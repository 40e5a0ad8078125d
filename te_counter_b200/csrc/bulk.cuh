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
    const int64_t* cell_base;    // n_chrom + 1
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

struct Sector { u32 w[8]; };

__device__ __forceinline__ Sector ld_sector(const u32* sectors, int64_t idx) {
    Sector r;
    const u32* p = sectors + idx * 8;
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

// distinct ensg slots of one unit (registers only)
struct SlotSet {
    u32 v0, v1, v2, v3;
    int n;
    bool ovf;
    __device__ __forceinline__ void clear() { v0 = v1 = v2 = v3 = 0xFFFFFFFFu; n = 0; ovf = false; }
    __device__ __forceinline__ void add(u32 w) {
        if ((v0 == w) | (v1 == w) | (v2 == w) | (v3 == w)) return;
        if (n == 0) v0 = w; else if (n == 1) v1 = w; else if (n == 2) v2 = w; else if (n == 3) v3 = w; else ovf = true;
        ++n;
    }
};

// every entry of the cell's sector chain that contains point ra or rb (cell-relative; 0xFFFFFFFF =
// no point) goes into S
__device__ __forceinline__ void stab_cell(const StabView& sv, int64_t sec, u32 ra, u32 rb, SlotSet& S) {
    const u32 M = (1u << 22) - 1;
    const int rmax = max((int)ra, (int)rb);                 // 0xFFFFFFFF -> -1
    for (;;) {
        const Sector s = ld_sector(sv.sectors, sec);
        const u32 header = s.w[7] >> 4;
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
        hit &= (1u << (header & 7u)) - 1u;
        while (hit) {
            const int i = __ffs(hit) - 1;
            hit &= hit - 1;
            const u32 pair = (i < 2) ? s.w[0] : ((i < 4) ? s.w[1] : s.w[2]);
            S.add((pair >> ((i & 1) << 4)) & 0xFFFFu);
        }
        if (!(header & 8u) || rmax < (int)(pos[5] & 2047u)) break;
        sec = header >> 4;
    }
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

// One warp per 32 consecutive units, grid-stride; the next warp-tile's records are requested before
// the current one is looked up.
template <bool PAIRED>
__global__ void __launch_bounds__(BULK_THREADS, BULK_MIN_CTAS)
bulk_count_cell_kernel(IndexView iv, StabView sv, int64_t n_units, int qual,
                       const int32_t* __restrict__ start, const int32_t* __restrict__ end,
                       const uint16_t* __restrict__ chrom, const uint8_t* __restrict__ mapq,
                       const uint8_t* __restrict__ flag, u64* __restrict__ counts, u64* __restrict__ stats,
                       u32* __restrict__ slow_bits) {
    __shared__ u32 s_hot[TEC_HOT_SLOTS];
    __shared__ u64 s_stats[TEC_BULK_NSTATS];
    for (int i = threadIdx.x; i < TEC_HOT_SLOTS; i += blockDim.x) s_hot[i] = 0;
    if (threadIdx.x < TEC_BULK_NSTATS) s_stats[threadIdx.x] = 0;
    __syncthreads();
    u32 n_assigned = 0, n_lowq = 0, n_badchrom = 0, n_qcfail = 0;
    const u32 reject = TEC_F_UNMAPPED | TEC_F_DUP | TEC_F_QCFAIL;
    const u32 reject2 = PAIRED ? (reject | (reject << 8)) : reject;
    const u64 pol = make_evict_first_policy();
    const int bs = iv.bs;
    const u32 counted = (1u << TEC_T_GENE) | (1u << TEC_T_TE) | (1u << TEC_T_SNRNA);
    const u32 cmask = (1u << sv.shift) - 1;
    const int lane = threadIdx.x & 31;
    const int64_t n_tiles = (n_units + 31) >> 5;
    const int64_t tile_stride = (int64_t)gridDim.x * (blockDim.x >> 5);
    int64_t tile = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    BulkRec cur, nxt;
    cur.fl = cur.q = 0; cur.c = cur.loc1 = cur.loc2 = 0;
    nxt = cur;
    if (tile < n_tiles && tile * 32 + lane < n_units) cur = bulk_load<PAIRED>(tile * 32 + lane, start, end, chrom, mapq, flag, pol);
    for (; tile < n_tiles; tile += tile_stride) {
        const int64_t u = tile * 32 + lane;
        const int64_t un = u + tile_stride * 32;
        if (un < n_units) nxt = bulk_load<PAIRED>(un, start, end, chrom, mapq, flag, pol);
        bool slow = false;
        if (u < n_units) {
            const int c = cur.c, loc1 = cur.loc1, loc2 = cur.loc2;
            if (cur.fl & reject2) n_qcfail++;                                              // :81-86 / :204
            else if ((int)cur.q < qual) n_lowq++;                                          // :88 / :208
            else if (PAIRED && (cur.fl & TEC_F_NAME_MISMATCH)) atomicAdd(stats + TEC_BS_CRASH_NAME, 1ULL);   // :92-94
            else if (c >= iv.n_chrom) n_badchrom++;                                        // :100 / :216
            else {
                const bool edge = (bs == 10000) ? ((loc1 % 10000 == 0) || ((loc2 + 1) % 10000 == 0))
                                                : ((loc1 % bs == 0) || ((loc2 + 1) % bs == 0));
                if (edge) slow = true;
                else {
                    SlotSet S;
                    S.clear();
                    const int64_t cb = __ldg(sv.cell_base + c);
                    const int64_t n_cells = __ldg(sv.cell_base + c + 1) - cb;
                    const int xa = loc1, xb = loc2 - 1;
                    const int64_t ka = xa >> sv.shift, kb = xb >> sv.shift;            // arithmetic shift: negative stays negative
                    const bool va = xa >= 0 && ka < n_cells, vb = xb >= 0 && kb < n_cells;
                    const u32 ra = va ? ((u32)xa & cmask) : 0xFFFFFFFFu, rb = vb ? ((u32)xb & cmask) : 0xFFFFFFFFu;
                    if (va && vb && ka == kb) stab_cell(sv, cb + ka, ra, rb, S);
                    else {
                        if (va) stab_cell(sv, cb + ka, ra, 0xFFFFFFFFu, S);
                        if (vb) stab_cell(sv, cb + kb, 0xFFFFFFFFu, rb, S);
                    }
                    if (S.ovf) slow = true;
                    else if (S.n) {                                                        // :128 result not empty
                        n_assigned++;                                                      // :149
                        bool count_it = true;
                        if (!sv.all_counted) {
                            u32 typemask = 1u << __ldg(sv.slot_type + S.v0);
                            if (S.n > 1) typemask |= 1u << __ldg(sv.slot_type + S.v1);
                            if (S.n > 2) typemask |= 1u << __ldg(sv.slot_type + S.v2);
                            if (S.n > 3) typemask |= 1u << __ldg(sv.slot_type + S.v3);
                            count_it = (typemask & counted) != 0;
                            if (!count_it && (typemask & (1u << TEC_T_ENHANCER))) atomicAdd(stats + TEC_BS_CRASH_ENHANCER, 1ULL);   // :145-147
                        }
                        if (count_it) {
                            auto bump = [&](u32 slot) {
                                if (slot < TEC_HOT_SLOTS) atomicAdd(&s_hot[slot], 1u);
                                else atomicAdd(counts + slot, 1ULL);
                            };
                            bump(S.v0);
                            if (S.n > 1) bump(S.v1);
                            if (S.n > 2) bump(S.v2);
                            if (S.n > 3) bump(S.v3);
                        }
                    }
                }
            }
        }
        const u32 sb = __ballot_sync(0xFFFFFFFFu, slow);
        if (lane == 0) slow_bits[tile] = sb;
        cur = nxt;
    }
    u64 v[4] = {n_assigned, n_lowq, n_badchrom, n_qcfail};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const u64 s = warp_sum(v[i]);
        if (lane == 0 && s) atomicAdd(&s_stats[TEC_BS_ASSIGNED + i], s);
    }
    __syncthreads();
    if (threadIdx.x >= TEC_BS_ASSIGNED && threadIdx.x < TEC_BS_ASSIGNED + 4 && s_stats[threadIdx.x])
        atomicAdd(stats + threadIdx.x, s_stats[threadIdx.x]);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(stats + TEC_BS_UNITS, (u64)n_units);
    for (int i = threadIdx.x; i < TEC_HOT_SLOTS; i += blockDim.x) {
        const u32 x = s_hot[i];
        if (x) atomicAdd(counts + i, (u64)x);
    }
}

// Units flagged by the fast kernel (bucket-edge candidates, more than STAB_MAXD distinct ensg): exact
// search, counters in slot space.  One thread per ballot word.
template <bool PAIRED>
__global__ void __launch_bounds__(256)
bulk_slow_kernel(IndexView iv, int64_t n_units,
                 const int32_t* __restrict__ start, const int32_t* __restrict__ end,
                 const uint16_t* __restrict__ chrom, u64* __restrict__ counts, u64* __restrict__ stats,
                 const u32* __restrict__ slow_bits) {
    const int64_t n_tiles = (n_units + 31) >> 5;
    const u32 counted = (1u << TEC_T_GENE) | (1u << TEC_T_TE) | (1u << TEC_T_SNRNA);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_tiles; t += (int64_t)gridDim.x * blockDim.x) {
        u32 bits = slow_bits[t];
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            const int64_t u = t * 32 + b;
            int c, loc1, loc2;
            if (PAIRED) { c = chrom[2 * u]; loc1 = start[2 * u]; loc2 = start[2 * u + 1]; }
            else { c = chrom[u]; loc1 = start[u]; loc2 = end[u]; }
            u32 typemask = 0;
            bulk_for_each_hit(iv, c, loc1, loc2, [&](int64_t fi) {
                typemask |= 1u << info_type(__ldg(iv.info + fi));
                return true;
            });
            if (!typemask) continue;                                                       // :128 no result
            atomicAdd(stats + TEC_BS_ASSIGNED, 1ULL);                                      // :149
            if (!(typemask & counted)) {
                if (typemask & (1u << TEC_T_ENHANCER)) atomicAdd(stats + TEC_BS_CRASH_ENHANCER, 1ULL);
                continue;
            }
            // count a hit iff no earlier hit (in enumeration order) carries the same ensg -- O(h^2)
            // re-walks, no storage
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
}
