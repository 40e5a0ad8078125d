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
