// Bulk counting, two passes over the cell table of stab2_build.h (the default bulk path).
//
// Reference semantics (te_counter, te_count/te_count.py): filter SE :203-218, filter + mate merge PE
// :76-102, candidate buckets :106-116 / :222-231, point tests :118-126 / :233-241, type rule + tally
// :128-149 / :243-261; closed forms in the header of bulk.cuh.
//
//   bulk2_fast_kernel    one pass over the records, straight-line code, two units per thread.  A unit whose
//                        two points are covered by ONE primary sector (the cell of the smaller point also
//                        covers `ext` bp of the next cell) and lie below the sector's threshold is filtered,
//                        tested against the five entries (16-bit lanes, both points at once), de-duplicated
//                        (twin flags) and tallied into shared-memory counters: the ensg of the hit entries go
//                        through a per-warp queue in shared memory and are added 32 at a time (a shared-memory
//                        reduction costs ~18 SM cycles per warp instruction whatever the number of active
//                        lanes, tools/microbench2.cu, so ten sparse ones per tile become about three full ones).
//                        Every other unit
//                        that passed the filter is appended to the warp's own segment of the deferred list
//                        (no atomics: the list is partitioned by warp).
//   bulk2_pair_kernel    the deferred units (5 % of the pairs of a paired-end workload, 13 % of spliced single-end
//                        reads) that two sectors answer, straight-line; leaves the rest in place for
//   bulk2_second_kernel  one or two sector chains,
//                        a register set of distinct ensg, hot counters in shared memory.  Units in EDGE cells
//                        (the reference's two-bucket candidate rule can bite there) or with more distinct ensg
//                        than the register set go on to bulk_slow_kernel's exact search.
//
// Streamed per unit: 16 B (PE; `end` is never read) / 12 B (SE) of records; one random 32-byte sector of the
// L2-resident table; one shared-memory reduction per entry.
#pragma once
#include "common.cuh"
#include "bulk.cuh"
#include "stab2_build.h"

struct Stab2View {
    const u32* sectors;          // 8 words per sector, 32-byte aligned
    const uint2* cells;          // per chromosome {first sector, number of cells}; [n_chrom] = {0, 0} sentinel
    const u32* ovf_first;        // per primary sector: first overflow sector
    const uint8_t* slot_type;    // n_slots
    int shift, ext;
    int all_counted;
};

#define B2_UPT 2                 // units per thread in the fast kernel
#define B2_MAXD 8                // register set of the second pass

__device__ __forceinline__ u32 ld_stream_v2u32(const void* p, u64 pol, u32& y) {
    u32 x;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(x), "=r"(y) : "l"(p), "l"(pol));
    return x;
}
__device__ __forceinline__ int4 ld_stream_int4(const int32_t* p, u64 pol) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
    return r;
}

// the records of a thread's two consecutive units, as loaded (unpacked at use)
template <bool PAIRED> struct B2Raw;
template <> struct B2Raw<true> { u32 fl, mq, c0, c1; int4 st; };            // 2 pairs = 4 records
template <> struct B2Raw<false> { u32 fl, mq, ch; int2 st, en; };          // 2 records

template <bool PAIRED, bool FULL>
__device__ __forceinline__ void b2_load(B2Raw<PAIRED>& r, u32 u0, u32 n_units, const int32_t* __restrict__ start,
                                        const int32_t* __restrict__ end, const uint16_t* __restrict__ chrom,
                                        const uint8_t* __restrict__ mapq, const uint8_t* __restrict__ flag, u64 pol) {
    if (FULL || u0 + 2 <= n_units) {
        if constexpr (PAIRED) {
            r.fl = ld_stream_u32(flag + 2 * (size_t)u0, pol);
            r.mq = ld_stream_u32(mapq + 2 * (size_t)u0, pol);
            r.c0 = ld_stream_v2u32(chrom + 2 * (size_t)u0, pol, r.c1);
            r.st = ld_stream_int4(start + 2 * (size_t)u0, pol);
        } else {
            r.fl = ld_stream_u16(flag + u0, pol);
            r.mq = ld_stream_u16(mapq + u0, pol);
            r.ch = ld_stream_u32(chrom + u0, pol);
            r.st = ld_stream_int2(start + u0, pol);
            r.en = ld_stream_int2(end + u0, pol);
        }
    } else if (u0 < n_units) {                       // the last, odd unit of the launch
        if constexpr (PAIRED) {
            r.fl = ld_stream_u16(flag + 2 * (size_t)u0, pol);
            r.mq = ld_stream_u16(mapq + 2 * (size_t)u0, pol);
            r.c0 = ld_stream_u32(chrom + 2 * (size_t)u0, pol);
            r.c1 = 0;
            const int2 s = ld_stream_int2(start + 2 * (size_t)u0, pol);
            r.st = make_int4(s.x, s.y, 0, 0);
        } else {
            r.fl = ld_stream_u8(flag + u0, pol);
            r.mq = ld_stream_u8(mapq + u0, pol);
            r.ch = ld_stream_u16(chrom + u0, pol);
            r.st = make_int2(ld_stream_int(start + u0, pol), 0);
            r.en = make_int2(ld_stream_int(end + u0, pol), 0);
        }
    }
}

template <bool PAIRED>
__device__ __forceinline__ void b2_unpack(const B2Raw<PAIRED>& r, int j, u32& fl, u32& q, u32& c, int& loc1, int& loc2) {
    if constexpr (PAIRED) {
        fl = j ? (r.fl >> 16) : (r.fl & 0xFFFFu);             // both mates' flag bytes
        q = j ? ((r.mq >> 16) & 0xFFu) : (r.mq & 0xFFu);      // read1 only (:88)
        c = (j ? r.c1 : r.c0) & 0xFFFFu;                      // read1 only (:96)
        loc1 = j ? r.st.z : r.st.x;                           // :97
        loc2 = j ? r.st.w : r.st.y;                           // :98 mate START
    } else {
        fl = j ? (r.fl >> 8) : (r.fl & 0xFFu);
        q = j ? (r.mq >> 8) : (r.mq & 0xFFu);
        c = j ? (r.ch >> 16) : (r.ch & 0xFFFFu);
        loc1 = j ? r.st.y : r.st.x;                           // :213
        loc2 = j ? r.en.y : r.en.x;                           // :214
    }
}

// +1 for an entry when its hit bit is set (ALLHOT: every counter in shared memory; the lanes that did not
// hit add to a scratch word of their own behind the counters, so there is no branch)
template <bool ALLHOT, bool HITONLY>
__device__ __forceinline__ void b2_bump(bool hit, u32 slot, u32 hot_addr, u32 scratch_addr, u32 n_hot, u64* __restrict__ counts, u32 one) {
    if (HITONLY) {
        // only the lanes that hit take part: a branch per entry, but fewer shared-memory wavefronts (the scratch
        // word of a lane shares its bank with the counters that map there)
        if (hit) {
            if (ALLHOT || slot < n_hot) asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(hot_addr + slot * 4u), "r"(one) : "memory");
            else atomicAdd(counts + slot, 1ULL);
        }
    } else if (ALLHOT) {
        const u32 addr = hit ? (hot_addr + slot * 4u) : scratch_addr;
        asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(addr), "r"(one) : "memory");
    } else {
        const bool hot = slot < n_hot;
        const u32 addr = (hit && hot) ? (hot_addr + slot * 4u) : scratch_addr;
        asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(addr), "r"(one) : "memory");
        if (hit && !hot) atomicAdd(counts + slot, 1ULL);
    }
}
__device__ __forceinline__ u32 b2_lo16(u32 w) { return __byte_perm(w, 0u, 0x4410u); }
__device__ __forceinline__ u32 b2_hi16(u32 w) { return __byte_perm(w, 0u, 0x4432u); }

// what a thread of the fast kernel carries across tiles
struct B2Thread {
    u32 n_assigned, n_lowq, n_badchrom, n_qcfail;
    uint4* wp;                    // next free entry of the warp's segment of the deferred list (warp-uniform)
    u32 q_head, q_tail;           // hit queue of the warp (warp-uniform): entries [q_head, q_tail) are pending
};
struct B2Const {
    u32 reject2, lim, lt_mask, hot_addr, scratch_addr, one, n_hot, n_units, mode;
    int shift, cmask, qual, n_chrom;
    u64 pol_table;                // L2 policy of the table loads (evict_last when mode & B2_MODE_KEEP)
    u32 q_addr;                   // shared-memory address of the warp's hit queue (B2_QCAP 16-bit entries)
};
#define B2_QCAP 512u              // per warp; a tile adds at most 10 x 32 entries between two drains (31 may be pending)
#define B2_DEF_GATHER 0x80000000u
#define B2_DEF_NAME 0x40000000u
#define B2_MODE_KEEP 1u           // table sectors: L2 evict_last (the records stream through with evict_first)
#define B2_MODE_PREFETCH 2u       // request the next tile's sectors into L2 one turn ahead
#define B2_MODE_QUEUE 4u          // tally through the per-warp hit queue instead of one reduction per entry
#define B2_MODE_DEEP 8u           // 512-thread CTAs with 128 registers per thread: three tiles in flight per warp
#define B2_MODE_SCAN 16u          // hit queue filled once per tile (both units of a lane): one warp prefix sum instead of ten ballots
#define B2_MODE_768 32u           // with B2_MODE_DEEP: 768-thread CTAs (85 registers), two tiles in flight per warp
#define B2_MODE_TILE_DRAIN 64u    // ballot queue drained once per tile instead of once per unit (default: 1 + 4 + 8 + 64 = 77)

__device__ __forceinline__ Sector ld_sector_pol(const u32* sectors, u32 idx, u64 pol) {
    Sector r;
    const u32* p = sectors + (size_t)idx * 8;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
                 : "l"(p), "l"(pol));
    return r;
}

// what phase A leaves for phase B: per unit the sector and the two cell-relative points; per tile the flag bits
struct B2Stage {
    u32 sec[B2_UPT];              // sector index (0 when the unit is not answered in place)
    u32 pts[B2_UPT];              // ra | rb << 16
    u32 flags;                    // bit j: unit j passed the filter and one sector covers both points
};

// Phase A of a warp's turn (64 consecutive units, two per lane): filter (te_count.py:78-102 / :203-218) and
// the sector each unit needs.  FULL: every unit of the tile exists.
template <bool PAIRED, bool FULL>
__device__ __forceinline__ void b2_phase_a(const B2Raw<PAIRED>& cur, const u32 u0, const Stab2View& sv, const B2Const& k,
                                           B2Thread& t, B2Stage& st) {
    st.flags = 0;
#pragma unroll
    for (int j = 0; j < B2_UPT; ++j) {
        u32 fl, q, c;
        int loc1, loc2;
        b2_unpack<PAIRED>(cur, j, fl, q, c, loc1, loc2);
        // first failing test wins, as in the reference's if / continue chain; branch-free counters
        const bool live = FULL || (u0 + j < k.n_units);
        const bool f_qc = (fl & k.reject2) != 0;                                           // :81-86 / :204
        const bool f_lq = (int)q < k.qual;                                                 // :88 / :208
        const bool f_nm = PAIRED && (fl & TEC_F_NAME_MISMATCH);                            // :92-94
        // chromosomes that are not keys of the bucket hash (no feature row) have zero cells
        const uint2 cell = __ldg(sv.cells + min(c, (u32)k.n_chrom));
        const bool f_bc = cell.y == 0;                                                     // :100 / :216
        t.n_qcfail += live & f_qc;
        t.n_lowq += live & !f_qc & f_lq;
        t.n_badchrom += live & !(f_qc | f_lq | f_nm) & f_bc;
        // a name mismatch (the reference dies there, :92-94) is left to the second pass, like every rare case
        const bool look = live & !(f_qc | f_lq) & (f_nm | !f_bc);
        // point A x = loc1, point B x = loc2 - 1 (bulk.cuh header); the cell of the smaller one
        const int xa = loc1, xb = loc2 - 1;
        const int mn = min(xa, xb);
        const int kc = mn >> k.shift;                         // arithmetic shift: negative stays negative
        const int base = mn & ~k.cmask;
        const u32 ra = (u32)(xa - base), rb = (u32)(xb - base);
        const u32 rmax = (u32)(max(xa, xb) - base);
        const bool ok = look & !f_nm & ((u32)kc < cell.y) & (rmax < k.lim);
        st.sec[j] = ok ? cell.x + (u32)kc : 0u;
        asm volatile("" : "+r"(st.sec[j]));                   // select the index, not the 64-bit address
        st.pts[j] = ra | (rb << 16);                          // garbage unless ok
        st.flags |= (u32)ok << j;
        // units that no single sector covers (points far apart or outside the cells, name mismatch) go to the
        // second pass with their coordinates: {unit, chromosome | B2_DEF_GATHER | name mismatch, loc1, loc2}
        const bool far = look & !ok;
        const u32 fm = __ballot_sync(0xFFFFFFFFu, far);
        if (fm) {
            if (far) t.wp[__popc(fm & k.lt_mask)] = make_uint4(u0 + j, c | B2_DEF_GATHER | (f_nm ? B2_DEF_NAME : 0u), (u32)loc1, (u32)loc2);
            t.wp += __popc(fm);
        }
    }
}

__device__ __forceinline__ void b2_prefetch(const B2Stage& st, const Stab2View& sv, const B2Const& k) {
#pragma unroll
    for (int j = 0; j < B2_UPT; ++j) {
        const u32* p = sv.sectors + (size_t)st.sec[j] * 8;
        if (k.mode & B2_MODE_KEEP) asm volatile("prefetch.global.L2::evict_last [%0];" :: "l"(p));
        else asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
    }
}

// hit queue: the lanes whose entry hit append its ensg slot (ballot + popc, no atomics)
__device__ __forceinline__ void b2_q_push(bool hit, u32 slot, const B2Const& k, B2Thread& t) {
    const u32 m = __ballot_sync(0xFFFFFFFFu, hit);
    if (hit) {
        const u32 at = (t.q_tail + __popc(m & k.lt_mask)) & (B2_QCAP - 1u);
        asm volatile("st.shared.u16 [%0], %1;" :: "r"(k.q_addr + at * 2u), "h"((unsigned short)slot) : "memory");
    }
    t.q_tail += __popc(m);
}
// add the pending entries to the counters, 32 at a time (all == false: only full groups)
template <bool ALLHOT>
__device__ __forceinline__ void b2_q_drain(const B2Const& k, B2Thread& t, u64* __restrict__ counts, int lane, bool all) {
    __syncwarp();
    while (t.q_tail - t.q_head >= (all ? 1u : 32u)) {
        const bool on = (u32)lane < t.q_tail - t.q_head;
        unsigned short sl = 0;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(sl) : "r"(k.q_addr + ((t.q_head + (u32)lane) & (B2_QCAP - 1u)) * 2u) : "memory");
        const u32 slot = sl;
        if (on) {
            if (ALLHOT || slot < k.n_hot) asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(k.hot_addr + slot * 4u), "r"(k.one) : "memory");
            else atomicAdd(counts + slot, 1ULL);
        }
        t.q_head += min(32u, t.q_tail - t.q_head);
    }
    __syncwarp();
}

// Scan variant of the hit queue (B2_MODE_SCAN): the queue is linear (head = 0 after every drain, the < 32 pending entries
// are moved to its front), a lane's hits of BOTH units of the tile are placed behind one warp prefix sum.
// y: hit bits of unit 0 at 15, 31, 14, 30, 13 (entries 0..4), of unit 1 three bits lower.
__device__ __forceinline__ void b2_q_put(u32& addr, bool hit, u32 slot) {
    if (hit) {
        asm volatile("st.shared.u16 [%0], %1;" :: "r"(addr), "h"((unsigned short)slot) : "memory");
        addr += 2u;
    }
}
template <bool ALLHOT>
__device__ __forceinline__ void b2_q_drain_linear(const B2Const& k, B2Thread& t, u64* __restrict__ counts, int lane, bool all) {
    __syncwarp();
    const u32 n = t.q_tail;
    u32 g = 0;
    while (n - g >= (all ? 1u : 32u)) {
        unsigned short sl = 0;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(sl) : "r"(k.q_addr + (g + (u32)lane) * 2u) : "memory");
        const u32 slot = sl;
        if ((u32)lane < n - g) {
            if (ALLHOT || slot < k.n_hot) asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(k.hot_addr + slot * 4u), "r"(k.one) : "memory");
            else atomicAdd(counts + slot, 1ULL);
        }
        g += min(32u, n - g);
    }
    if (g) {                                                  // pending entries to the front
        unsigned short sl = 0;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(sl) : "r"(k.q_addr + (g + (u32)lane) * 2u) : "memory");
        __syncwarp();
        if ((u32)lane < n - g) asm volatile("st.shared.u16 [%0], %1;" :: "r"(k.q_addr + (u32)lane * 2u), "h"(sl) : "memory");
        t.q_tail = n - g;
    }
    __syncwarp();
}
template <bool ALLHOT>
__device__ __forceinline__ void b2_q_push_tile(const u32 y, const Sector (&s)[B2_UPT], const B2Const& k, B2Thread& t, u64* __restrict__ counts, int lane) {
    const u32 cnt = __popc(y);
    u32 incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u32 v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += v;
    }
    const u32 total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    u32 addr = k.q_addr + (t.q_tail + incl - cnt) * 2u;
    b2_q_put(addr, y & (1u << 15), s[0].w[6]);
    b2_q_put(addr, y & (1u << 31), s[0].w[6] >> 16);
    b2_q_put(addr, y & (1u << 14), s[0].w[7]);
    b2_q_put(addr, y & (1u << 30), s[0].w[7] >> 16);
    b2_q_put(addr, y & (1u << 13), s[0].w[5] >> 16);
    b2_q_put(addr, y & (1u << 12), s[1].w[6]);
    b2_q_put(addr, y & (1u << 28), s[1].w[6] >> 16);
    b2_q_put(addr, y & (1u << 11), s[1].w[7]);
    b2_q_put(addr, y & (1u << 27), s[1].w[7] >> 16);
    b2_q_put(addr, y & (1u << 10), s[1].w[5] >> 16);
    t.q_tail += total;
    b2_q_drain_linear<ALLHOT>(k, t, counts, lane, false);
}

// Phase B: the sector test, the tally and the deferred list.
__device__ __forceinline__ void b2_load_sectors(Sector (&s)[B2_UPT], const B2Stage& st, const Stab2View& sv, const B2Const& k) {
#pragma unroll
    for (int j = 0; j < B2_UPT; ++j) s[j] = ld_sector_pol(sv.sectors, st.sec[j], k.pol_table);
}

template <bool ALLHOT, int QUEUE>
__device__ __forceinline__ void b2_phase_b(const B2Stage& st, const Sector (&s)[B2_UPT], const u32 u0, const Stab2View& sv, const B2Const& k,
                                           B2Thread& t, u64* __restrict__ counts, u64* __restrict__ stats) {
    u32 y = 0;                                                // QUEUE == 2: the hit bits of both units
#pragma unroll
    for (int j = 0; j < B2_UPT; ++j) {
        const bool ok = (st.flags >> j) & 1u;
        const u32 ra = st.pts[j] & 0xFFFFu, rb = st.pts[j] >> 16;
        const u32 rmax = max(ra, rb);
        const u32 w2 = s[j].w[2];
        const u32 thr = (w2 >> 16) & S2_THR_MASK;
        const bool inplace = ok & (rmax < thr);
        const bool defer = ok & !inplace;
        // ---- five interval tests, both points at once (16-bit lanes, stab_build.h)
        const PointK pa = make_point(ra), pb = make_point(rb);
        const u32 m = inplace ? 0x80008000u : 0u;
        u32 a0 = (((pa.xg - s[j].w[0]) & (s[j].w[3] + pa.kg)) | ((pb.xg - s[j].w[0]) & (s[j].w[3] + pb.kg))) & m;
        u32 a1 = (((pa.xg - s[j].w[1]) & (s[j].w[4] + pa.kg)) | ((pb.xg - s[j].w[1]) & (s[j].w[4] + pb.kg))) & m;
        u32 a2 = (((pa.xg - w2) & (s[j].w[5] + pa.kg)) | ((pb.xg - w2) & (s[j].w[5] + pb.kg))) & m & 0x8000u;
        // twins: entries (0, 1) / (2, 3) carry the same ensg -> the pair counts once
        a0 &= ~((a0 << 16) & w2 & 0x80000000u);
        a1 &= ~((a1 << 16) & (w2 << 1) & 0x80000000u);
        if (!sv.all_counted) {
            // type rule of te_count.py:134-147 over the hit entries
            u32 typemask = 0;
            if (a0 & 0x8000u) typemask |= 1u << __ldg(sv.slot_type + (s[j].w[6] & 0xFFFFu));
            if (a0 & 0x80000000u) typemask |= 1u << __ldg(sv.slot_type + (s[j].w[6] >> 16));
            if (a1 & 0x8000u) typemask |= 1u << __ldg(sv.slot_type + (s[j].w[7] & 0xFFFFu));
            if (a1 & 0x80000000u) typemask |= 1u << __ldg(sv.slot_type + (s[j].w[7] >> 16));
            if (a2) typemask |= 1u << __ldg(sv.slot_type + (s[j].w[5] >> 16));
            if (typemask) {
                t.n_assigned++;                                                            // :149
                const u32 counted = (1u << TEC_T_GENE) | (1u << TEC_T_TE) | (1u << TEC_T_SNRNA);
                if (!(typemask & counted)) {
                    if (typemask & (1u << TEC_T_ENHANCER)) atomicAdd(stats + TEC_BS_CRASH_ENHANCER, 1ULL);   // :145-147
                    a0 = a1 = a2 = 0;
                }
            }
        } else {
            t.n_assigned += (a0 | a1 | a2) != 0;                                           // :128, :149
        }
        if (QUEUE == 2) {
            static_assert(B2_UPT == 2, "b2_q_push_tile places the hits of two units");
            y |= (a0 | (a1 >> 1) | (a2 >> 2)) >> (3 * j);
        } else if (QUEUE) {
            b2_q_push(a0 & 0x8000u, s[j].w[6], k, t);
            b2_q_push(a0 & 0x80000000u, s[j].w[6] >> 16, k, t);
            b2_q_push(a1 & 0x8000u, s[j].w[7], k, t);
            b2_q_push(a1 & 0x80000000u, s[j].w[7] >> 16, k, t);
            b2_q_push(a2 != 0, s[j].w[5] >> 16, k, t);
            if (QUEUE == 1) b2_q_drain<ALLHOT>(k, t, counts, (int)(threadIdx.x & 31), false);    // QUEUE 3: once per tile, below
        } else {
            b2_bump<ALLHOT, false>(a0 & 0x8000u, b2_lo16(s[j].w[6]), k.hot_addr, k.scratch_addr, k.n_hot, counts, k.one);
            b2_bump<ALLHOT, false>(a0 & 0x80000000u, b2_hi16(s[j].w[6]), k.hot_addr, k.scratch_addr, k.n_hot, counts, k.one);
            b2_bump<ALLHOT, false>(a1 & 0x8000u, b2_lo16(s[j].w[7]), k.hot_addr, k.scratch_addr, k.n_hot, counts, k.one);
            b2_bump<ALLHOT, false>(a1 & 0x80000000u, b2_hi16(s[j].w[7]), k.hot_addr, k.scratch_addr, k.n_hot, counts, k.one);
            b2_bump<ALLHOT, false>(a2 != 0, b2_hi16(s[j].w[5]), k.hot_addr, k.scratch_addr, k.n_hot, counts, k.one);
        }
        // ---- everything else goes to the second pass: the warp's own segment of the list, no atomics
        // {unit, sector, points}: the chain of that sector is walked from there (phase A wrote the other kind)
        const u32 dm = __ballot_sync(0xFFFFFFFFu, defer);
        if (defer) t.wp[__popc(dm & k.lt_mask)] = make_uint4(u0 + j, st.sec[j], st.pts[j], 0u);
        t.wp += __popc(dm);
    }
    if (QUEUE == 2) b2_q_push_tile<ALLHOT>(y, s, k, t, counts, (int)(threadIdx.x & 31));
    if (QUEUE == 3) b2_q_drain<ALLHOT>(k, t, counts, (int)(threadIdx.x & 31), false);           // at most 31 + 10 x 32 entries < B2_QCAP
}

template <bool PAIRED, int NT, bool ALLHOT, int QUEUE, int DEPTH>
__global__ void __launch_bounds__(NT, DEPTH ? 1 : 2048 / NT / 2)
bulk2_fast_kernel(Stab2View sv, int n_chrom, u32 n_units, int qual,
                  const int32_t* __restrict__ start, const int32_t* __restrict__ end,
                  const uint16_t* __restrict__ chrom, const uint8_t* __restrict__ mapq,
                  const uint8_t* __restrict__ flag, u64* __restrict__ counts, u64* __restrict__ stats,
                  uint4* __restrict__ defer_list, u32* __restrict__ defer_count, u32 seg_cap, u32 n_hot, u32 mode) {
    constexpr int WARPS = NT / 32;
    __shared__ u64 s_stats[TEC_BULK_NSTATS];
    extern __shared__ __align__(16) u32 s_hot_dyn[];
    // dynamic shared memory: n_hot counters, 32 scratch words, then one hit queue of B2_QCAP 16-bit entries per warp
    for (u32 i = threadIdx.x; i < n_hot + 32; i += blockDim.x) s_hot_dyn[i] = 0;
    if (threadIdx.x < TEC_BULK_NSTATS) s_stats[threadIdx.x] = 0;
    __syncthreads();
    const u32 reject = TEC_F_UNMAPPED | TEC_F_DUP | TEC_F_QCFAIL;
    const u64 pol = make_evict_first_policy();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    B2Const k;
    k.reject2 = PAIRED ? (reject | (reject << 8)) : reject;
    k.shift = sv.shift;
    k.cmask = (1 << sv.shift) - 1;
    k.lim = (u32)(k.cmask + 1 + sv.ext);                      // relative positions a sector covers
    k.lt_mask = (1u << lane) - 1u;
    k.hot_addr = (u32)__cvta_generic_to_shared(&s_hot_dyn[0]);
    asm volatile("" : "+r"(k.hot_addr));                      // keep it in a register (not recomputed per use)
    k.scratch_addr = k.hot_addr + (n_hot + (u32)lane) * 4u;
    k.q_addr = k.hot_addr + (n_hot + 32u) * 4u + (u32)wib * (B2_QCAP * 2u);
    // the increment as a run-time value: a literal 1 makes ptxas pick ATOMS.POPC.INC, which needs a
    // converged warp and therefore a branch around every reduction
    k.one = (u32)(n_units > 0);
    asm volatile("" : "+r"(k.one));
    k.n_hot = n_hot;
    k.n_units = n_units;
    k.qual = qual;
    k.n_chrom = n_chrom;
    k.mode = mode;
    if (mode & B2_MODE_KEEP) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(k.pol_table));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(k.pol_table));
    const u32 gw = blockIdx.x * WARPS + wib;
    uint4* const my_list = defer_list + (size_t)gw * seg_cap;
    B2Thread t;
    t.n_assigned = t.n_lowq = t.n_badchrom = t.n_qcfail = 0;
    t.wp = my_list;
    t.q_head = t.q_tail = 0;
    const u32 n_full = n_units >> 6;
    const u32 stride = gridDim.x * WARPS;
    const bool pf = (mode & B2_MODE_PREFETCH) != 0;
    u32 tile = gw;
    B2Raw<PAIRED> raw;
    if constexpr (DEPTH != 0) {
        // Full tiles of 64 units, warp-strided, DEPTH tiles in flight per warp (one CTA per SM: 512 threads with 128 registers
        // and three tiles, or 768 threads with 85 registers and two): while tile i is tested and tallied, the sectors of the
        // tiles behind it and the records of tile i + DEPTH are on their way -- the warp's own arithmetic hides its own
        // L2 / HBM latency.
        B2Stage st[DEPTH];
        Sector sec[DEPTH][B2_UPT];
        auto fill = [&](B2Stage& stg, Sector (&sc)[B2_UPT], u32 tl) {       // records of `tl` are in raw
            b2_phase_a<PAIRED, true>(raw, tl * 64 + 2 * lane, sv, k, t, stg);
            if (tl + stride < n_full && tl + stride >= tl)
                b2_load<PAIRED, true>(raw, (tl + stride) * 64 + 2 * lane, n_units, start, end, chrom, mapq, flag, pol);
            b2_load_sectors(sc, stg, sv, k);
        };
        if (tile < n_full) {
            b2_load<PAIRED, true>(raw, tile * 64 + 2 * lane, n_units, start, end, chrom, mapq, flag, pol);
#pragma unroll
            for (int d = 0; d < DEPTH - 1; ++d)
                if (tile + (u32)d * stride < n_full) fill(st[d], sec[d], tile + (u32)d * stride);
        }
        while (tile < n_full) {
#pragma unroll
            for (int r = 0; r < DEPTH; ++r) {
                if (tile < n_full) {
                    const u32 t2 = tile + (u32)(DEPTH - 1) * stride;
                    if (t2 < n_full) fill(st[(r + DEPTH - 1) % DEPTH], sec[(r + DEPTH - 1) % DEPTH], t2);
                    b2_phase_b<ALLHOT, QUEUE>(st[r], sec[r], tile * 64 + 2 * lane, sv, k, t, counts, stats);
                    tile += stride;
                }
            }
        }
    } else {
    // Full tiles of 64 units, warp-strided, software-pipelined over three turns: the records of tile i+2 are
    // requested, tile i+1 is filtered and its sectors are requested into L2, tile i is tested and tallied.
    B2Stage sa, sb;
    Sector sec[B2_UPT];
    if (tile < n_full) {
        b2_load<PAIRED, true>(raw, tile * 64 + 2 * lane, n_units, start, end, chrom, mapq, flag, pol);
        b2_phase_a<PAIRED, true>(raw, tile * 64 + 2 * lane, sv, k, t, sa);
        if (pf) b2_prefetch(sa, sv, k);
        if (tile + stride < n_full) b2_load<PAIRED, true>(raw, (tile + stride) * 64 + 2 * lane, n_units, start, end, chrom, mapq, flag, pol);
    }
    while (tile < n_full) {
        // here: sa = stage of `tile`; raw = records of tile + stride (if it exists)
        u32 nx = tile + stride;
        if (nx < n_full) {
            b2_phase_a<PAIRED, true>(raw, nx * 64 + 2 * lane, sv, k, t, sb);
            if (pf) b2_prefetch(sb, sv, k);
            if (nx + stride < n_full) b2_load<PAIRED, true>(raw, (nx + stride) * 64 + 2 * lane, n_units, start, end, chrom, mapq, flag, pol);
        }
        b2_load_sectors(sec, sa, sv, k);
        b2_phase_b<ALLHOT, QUEUE>(sa, sec, tile * 64 + 2 * lane, sv, k, t, counts, stats);
        tile = nx;
        if (tile >= n_full) break;
        nx = tile + stride;
        if (nx < n_full) {
            b2_phase_a<PAIRED, true>(raw, nx * 64 + 2 * lane, sv, k, t, sa);
            if (pf) b2_prefetch(sa, sv, k);
            if (nx + stride < n_full) b2_load<PAIRED, true>(raw, (nx + stride) * 64 + 2 * lane, n_units, start, end, chrom, mapq, flag, pol);
        }
        b2_load_sectors(sec, sb, sv, k);
        b2_phase_b<ALLHOT, QUEUE>(sb, sec, tile * 64 + 2 * lane, sv, k, t, counts, stats);
        tile = nx;
    }
    }
    // the last, partial tile belongs to the warp whose turn it would be
    if ((n_units & 63u) && (n_full % stride) == gw) {
        const u32 u0 = n_full * 64 + 2 * lane;
        raw = B2Raw<PAIRED>();
        B2Stage sl;
        Sector secl[B2_UPT];
        b2_load<PAIRED, false>(raw, u0, n_units, start, end, chrom, mapq, flag, pol);
        b2_phase_a<PAIRED, false>(raw, u0, sv, k, t, sl);
        b2_load_sectors(secl, sl, sv, k);
        b2_phase_b<ALLHOT, QUEUE>(sl, secl, u0, sv, k, t, counts, stats);
    }
    if (QUEUE == 2) b2_q_drain_linear<ALLHOT>(k, t, counts, lane, true);
    else if (QUEUE) b2_q_drain<ALLHOT>(k, t, counts, lane, true);
    static_assert(B2_QCAP >= 31u + 10u * 32u, "a tile's hits must fit the ring between two drains");
    if (lane == 0) defer_count[gw] = (u32)(t.wp - my_list);
    u64 v[4] = {t.n_assigned, t.n_lowq, t.n_badchrom, t.n_qcfail};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const u64 sum = warp_sum(v[i]);
        if (lane == 0 && sum) atomicAdd(&s_stats[TEC_BS_ASSIGNED + i], sum);
    }
    __syncthreads();
    if (threadIdx.x >= TEC_BS_ASSIGNED && threadIdx.x < TEC_BS_ASSIGNED + 4 && s_stats[threadIdx.x])
        atomicAdd(stats + threadIdx.x, s_stats[threadIdx.x]);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(stats + TEC_BS_UNITS, (u64)n_units);
    for (u32 i = threadIdx.x; i < n_hot; i += blockDim.x) {
        const u32 x = s_hot_dyn[i];
        if (x) atomicAdd(counts + i, (u64)x);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Second pass: the units the fast kernel deferred (they passed the filter).  One thread per unit.  The deferred
// list is partitioned by the warps of the fast kernel (n_seg segments of seg_cap entries, defer_count[] used);
// here `parts` warps share a segment, 32 entries at a time.
__device__ __forceinline__ u32 b2_sector_hits(const Sector& s, const PointK a, const PointK b) {
    const u32 a0 = ((a.xg - s.w[0]) & (s.w[3] + a.kg)) | ((b.xg - s.w[0]) & (s.w[3] + b.kg));
    const u32 a1 = ((a.xg - s.w[1]) & (s.w[4] + a.kg)) | ((b.xg - s.w[1]) & (s.w[4] + b.kg));
    const u32 a2 = ((a.xg - s.w[2]) & (s.w[5] + a.kg)) | ((b.xg - s.w[2]) & (s.w[5] + b.kg));
    return (a0 & 0x80008000u) | ((a1 & 0x80008000u) >> 1) | ((a2 & 0x8000u) >> 2);     // HB0..HB4 of bulk.cuh
}

#define B2_QCAP2 512u             // per warp: at most B2_MAXD x 32 entries are added between two drains

// The probes of a deferred entry: one sector chain holding both points, or one chain per point.  prim = primary sector,
// pts = ra | rb << 16 (S2_R_NONE = no point).  A unit with a name mismatch has no probe (the reference dies there).
template <bool PAIRED>
__device__ __forceinline__ int b2_probes(const uint4 rec, const Stab2View& sv, int shift, int cmask, u32 lim, u32 n_chrom, u32 (&prim)[2], u32 (&pts)[2]) {
    if (!(rec.y & B2_DEF_GATHER)) { prim[0] = rec.y; pts[0] = rec.z; return 1; }
    if (rec.y & B2_DEF_NAME) return 0;                       // (set for paired-end units only)
    const u32 c = rec.y & 0xFFFFu;
    const int loc1 = (int)rec.z, loc2 = (int)rec.w;
    const uint2 cell = __ldg(sv.cells + min(c, n_chrom));
    const int xa = loc1, xb = loc2 - 1;
    const int mn = min(xa, xb), k = mn >> shift, base = mn & ~cmask;
    const u32 rm = (u32)(max(xa, xb) - base);
    if ((u32)k < cell.y && rm < lim) {
        prim[0] = cell.x + (u32)k;
        pts[0] = (u32)(xa - base) | ((u32)(xb - base) << 16);
        return 1;
    }
    int np = 0;
    if (xa >= 0 && (u32)(xa >> shift) < cell.y) { prim[np] = cell.x + (u32)(xa >> shift); pts[np] = (u32)(xa & cmask) | (S2_R_NONE << 16); ++np; }
    if (xb >= 0 && (u32)(xb >> shift) < cell.y) { prim[np] = cell.x + (u32)(xb >> shift); pts[np] = S2_R_NONE | ((u32)(xb & cmask) << 16); ++np; }
    return np;
}

// Deferred units that TWO sectors answer, straight-line and converged (no hit loop, no chain loop): an overflow unit
// whose chain ends in the first overflow sector (primary + ovf_first[primary], both points), or a unit whose points lie
// in two cells whose primary sectors answer one point each (spliced single-end reads).  Duplicates: twins inside a
// primary sector by its flag bits; an ensg hit in both sectors, or twice inside an overflow sector, by comparing the at
// most 5 + 5 slots.  Everything else (EDGE cells, longer chains, name mismatches, single-probe gathers) is left IN PLACE
// for bulk2_second_kernel: the m-th entry a warp leaves goes where the m-th entry of its walk was; part_count[w] = how
// many.  Only used when every ensg is of a counted type (sv.all_counted).
template <bool PAIRED>
__global__ void __launch_bounds__(256, 5)
bulk2_pair_kernel(Stab2View sv, u64* __restrict__ counts, u64* __restrict__ stats,
                  uint4* __restrict__ defer_list, const u32* __restrict__ defer_count, u32 seg_cap, u32 n_seg, u32 parts,
                  u32* __restrict__ part_count, u32 sv_n_chrom) {
    __shared__ u32 s_hot[TEC_HOT_SLOTS];
    __shared__ unsigned short s_q[8][B2_QCAP2];
    for (int i = threadIdx.x; i < TEC_HOT_SLOTS; i += blockDim.x) s_hot[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    unsigned short* const q = s_q[threadIdx.x >> 5];
    u32 q_head = 0, q_tail = 0;                                    // warp-uniform
    const u32 lt_mask = (1u << lane) - 1u;
    const int shift = sv.shift;
    const int cmask = (1 << shift) - 1;
    const u32 lim = (u32)(cmask + 1 + sv.ext);
    u32 n_assigned = 0;
    const u32 n_w = n_seg * parts;
    auto drain = [&](bool all) {
        __syncwarp();
        while (q_tail - q_head >= (all ? 1u : 32u)) {
            const u32 slot = q[(q_head + (u32)lane) & (B2_QCAP2 - 1u)];
            if ((u32)lane < q_tail - q_head) {
                if (slot < TEC_HOT_SLOTS) atomicAdd(&s_hot[slot], 1u);
                else atomicAdd(counts + slot, 1ULL);
            }
            q_head += min(32u, q_tail - q_head);
        }
        __syncwarp();
    };
    for (u32 w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < n_w; w += gridDim.x * (blockDim.x >> 5)) {
        const u32 seg = w % n_seg, part = w / n_seg;
        const u32 cnt = __ldg(defer_count + seg);
        uint4* const list = defer_list + (size_t)seg * seg_cap;
        u32 m_out = 0;                                             // entries left for the second pass (warp-uniform)
        for (u32 i0 = part * 32; i0 < cnt; i0 += parts * 32) {
            const bool live = i0 + lane < cnt;
            uint4 rec = make_uint4(0u, 0u, 0u, 0u);
            if (live) rec = list[i0 + lane];                       // plain load: the list is rewritten in place below
            bool left = false, use = false, two_chains = false;
            u32 secA = 0, secB = 0, ptsA = 0, ptsB = 0;
            if (live) {
                if (!(rec.y & B2_DEF_GATHER)) {                    // one chain, both points: primary + first overflow sector
                    secA = rec.y; ptsA = ptsB = rec.z;
                    secB = __ldg(sv.ovf_first + secA);
                    use = true;
                } else if (rec.y & B2_DEF_NAME) {
                    left = true;                                   // the second pass counts the crash (:92-94)
                } else {
                    u32 prim[2], pts[2];
                    const int np = b2_probes<PAIRED>(rec, sv, shift, cmask, lim, sv_n_chrom, prim, pts);
                    if (np == 2) { secA = prim[0]; ptsA = pts[0]; secB = prim[1]; ptsB = pts[1]; use = two_chains = true; }
                    else if (np == 1) left = true;                 // (no probe at all: nothing to count)
                }
            }
            const Sector sA = ld_sector(sv.sectors, secA);
            const Sector sB = ld_sector(sv.sectors, secB);
            const u32 hA = sA.w[2] >> 16, hB = sB.w[2] >> 16;
            const u32 raA = ptsA & 0xFFFFu, rbA = ptsA >> 16, raB = ptsB & 0xFFFFu, rbB = ptsB >> 16;
            const int rmA = max(raA == S2_R_NONE ? -1 : (int)raA, rbA == S2_R_NONE ? -1 : (int)rbA);
            const int rmB = max(raB == S2_R_NONE ? -1 : (int)raB, rbB == S2_R_NONE ? -1 : (int)rbB);
            // does the chain go on behind this sector for this point?
            const bool onA = ((hA & S2_H_MORE) != 0) & (rmA >= (int)(hA & S2_THR_MASK));
            const bool onB = ((hB & S2_H_MORE) != 0) & (rmB >= (int)(hB & S2_THR_MASK));
            bool useB = use;
            if (use) {
                if (hA & S2_H_EDGE) left = true;
                if (two_chains) { if (onA || onB || (hB & S2_H_EDGE)) left = true; }
                else { useB = onA; if (onA && onB) left = true; }
            }
            const bool go = use && !left;
            u32 hitA = go ? b2_sector_hits(sA, make_point(raA), make_point(rbA)) : 0u;
            u32 hitB = (go && useB) ? b2_sector_hits(sB, make_point(raB), make_point(rbB)) : 0u;
            // twins of a primary sector count once (the flag bits are zero in overflow sectors)
            hitA &= ~(((hitA & HB0) << 16) & sA.w[2] & 0x80000000u);
            hitA &= ~(((hitA & HB2) << 16) & sA.w[2] & 0x40000000u);
            hitB &= ~(((hitB & HB0) << 16) & sB.w[2] & 0x80000000u);
            hitB &= ~(((hitB & HB2) << 16) & sB.w[2] & 0x40000000u);
            // slots; an entry that was not hit gets a value no slot has
            const u32 a0 = (hitA & HB0) ? (sA.w[6] & 0xFFFFu) : 0x10000u, a1 = (hitA & HB1) ? (sA.w[6] >> 16) : 0x10001u;
            const u32 a2 = (hitA & HB2) ? (sA.w[7] & 0xFFFFu) : 0x10002u, a3 = (hitA & HB3) ? (sA.w[7] >> 16) : 0x10003u;
            const u32 a4 = (hitA & HB4) ? (sA.w[5] >> 16) : 0x10004u;
            const u32 b0 = (hitB & HB0) ? (sB.w[6] & 0xFFFFu) : 0x20000u, b1 = (hitB & HB1) ? (sB.w[6] >> 16) : 0x20001u;
            const u32 b2 = (hitB & HB2) ? (sB.w[7] & 0xFFFFu) : 0x20002u, b3 = (hitB & HB3) ? (sB.w[7] >> 16) : 0x20003u;
            const u32 b4 = (hitB & HB4) ? (sB.w[5] >> 16) : 0x20004u;
            // (| and &, not || and &&: no short-circuit branches, the warp stays converged)
            const bool k0 = ((hitB & HB0) != 0) & !((b0 == a0) | (b0 == a1) | (b0 == a2) | (b0 == a3) | (b0 == a4));
            const bool k1 = ((hitB & HB1) != 0) & !((b1 == a0) | (b1 == a1) | (b1 == a2) | (b1 == a3) | (b1 == a4) | (b1 == b0));
            const bool k2 = ((hitB & HB2) != 0) & !((b2 == a0) | (b2 == a1) | (b2 == a2) | (b2 == a3) | (b2 == a4) | (b2 == b0) | (b2 == b1));
            const bool k3 = ((hitB & HB3) != 0) & !((b3 == a0) | (b3 == a1) | (b3 == a2) | (b3 == a3) | (b3 == a4) | (b3 == b0) | (b3 == b1) | (b3 == b2));
            const bool k4 = ((hitB & HB4) != 0) & !((b4 == a0) | (b4 == a1) | (b4 == a2) | (b4 == a3) | (b4 == a4) | (b4 == b0) | (b4 == b1) | (b4 == b2) | (b4 == b3));
            const u32 nd = (u32)__popc(hitA) + (u32)k0 + (u32)k1 + (u32)k2 + (u32)k3 + (u32)k4;
            n_assigned += nd != 0;                                                     // :128, :149
            // the warp is converged: one prefix sum places every lane's ensg in the queue
            u32 incl = nd;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const u32 v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl += v;
            }
            const u32 total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            u32 at = q_tail + incl - nd;
            if (hitA & HB0) q[at++ & (B2_QCAP2 - 1u)] = (unsigned short)a0;
            if (hitA & HB1) q[at++ & (B2_QCAP2 - 1u)] = (unsigned short)a1;
            if (hitA & HB2) q[at++ & (B2_QCAP2 - 1u)] = (unsigned short)a2;
            if (hitA & HB3) q[at++ & (B2_QCAP2 - 1u)] = (unsigned short)a3;
            if (hitA & HB4) q[at++ & (B2_QCAP2 - 1u)] = (unsigned short)a4;
            if (k0) q[at++ & (B2_QCAP2 - 1u)] = (unsigned short)b0;
            if (k1) q[at++ & (B2_QCAP2 - 1u)] = (unsigned short)b1;
            if (k2) q[at++ & (B2_QCAP2 - 1u)] = (unsigned short)b2;
            if (k3) q[at++ & (B2_QCAP2 - 1u)] = (unsigned short)b3;
            if (k4) q[at++ & (B2_QCAP2 - 1u)] = (unsigned short)b4;
            q_tail += total;
            drain(false);
            // what is left goes back into the list, densely in this warp's walk order
            const u32 lm = __ballot_sync(0xFFFFFFFFu, left);
            if (left) {
                const u32 m = m_out + (u32)__popc(lm & lt_mask);
                list[part * 32u + (m >> 5) * parts * 32u + (m & 31u)] = rec;
            }
            m_out += (u32)__popc(lm);
        }
        if (lane == 0) part_count[w] = m_out;
    }
    drain(true);
    const u64 a_sum = warp_sum((u64)n_assigned);
    if (lane == 0 && a_sum) atomicAdd(stats + TEC_BS_ASSIGNED, a_sum);
    __syncthreads();
    for (int i = threadIdx.x; i < TEC_HOT_SLOTS; i += blockDim.x) {
        const u32 x = s_hot[i];
        if (x) atomicAdd(counts + i, (u64)x);
    }
}

// SET 0: the distinct ensg of a unit in registers, a new one stored at position nd (one compare + select per position).
// SET 1: a new one enters at position 0 and the others move up (eight moves under one predicate), the slot of a hit entry
// is picked by word and half instead of a chain of selects.
// (Requesting the next turn's sectors into L2 one turn ahead was measured and dropped: the pass is bound by the divergent
// instruction stream of the hit loop, not by the latency of its loads -- profiles/r02_bulk_se_ncu_summary.md.)
template <bool PAIRED, int SET>
__global__ void __launch_bounds__(256, 6)
bulk2_second_kernel(Stab2View sv, u64* __restrict__ counts, u64* __restrict__ stats,
                    const uint4* __restrict__ defer_list, const u32* __restrict__ defer_count, u32 seg_cap, u32 n_seg, u32 parts,
                    u32* __restrict__ slow_list, u32 sv_n_chrom, const u32* __restrict__ part_count) {
    // hot ensg counters privatised per CTA (a Zipf-hot TE name would otherwise serialise in one L2 slice)
    __shared__ u32 s_hot[TEC_HOT_SLOTS];
    __shared__ unsigned short s_q[8][B2_QCAP2];                    // per warp: ensg slots waiting to be counted
    for (int i = threadIdx.x; i < TEC_HOT_SLOTS; i += blockDim.x) s_hot[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    unsigned short* const q = s_q[threadIdx.x >> 5];
    u32 q_head = 0, q_tail = 0;                                    // warp-uniform
    const u32 lt_mask = (1u << lane) - 1u;
    const int shift = sv.shift;
    const int cmask = (1 << shift) - 1;
    const u32 lim = (u32)(cmask + 1 + sv.ext);
    u32 n_assigned = 0;
    const u32 n_w = n_seg * parts;
    // add pending queue entries to the counters, 32 at a time (a shared-memory reduction costs the same with 1 or 32
    // active lanes); all: also the last, partial group
    auto drain = [&](bool all) {
        __syncwarp();
        while (q_tail - q_head >= (all ? 1u : 32u)) {
            const u32 slot = q[(q_head + (u32)lane) & (B2_QCAP2 - 1u)];
            if ((u32)lane < q_tail - q_head) {
                if (slot < TEC_HOT_SLOTS) atomicAdd(&s_hot[slot], 1u);
                else atomicAdd(counts + slot, 1ULL);
            }
            q_head += min(32u, q_tail - q_head);
        }
        __syncwarp();
    };
    for (u32 w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < n_w; w += gridDim.x * (blockDim.x >> 5)) {
        const u32 seg = w % n_seg, part = w / n_seg;
        // part_count: bulk2_pair_kernel went first and left, in place, the entries it could not answer -- the m-th one of
        // (segment, part) at the position the m-th entry of that part's walk had
        const u32 seg_cnt = __ldg(defer_count + seg);
        const u32 cnt = part_count ? part * 32u + ((__ldg(part_count + w) + 31u) / 32u) * parts * 32u : seg_cnt;
        const u32 left_n = part_count ? __ldg(part_count + w) : 0u;
        const uint4* const list = defer_list + (size_t)seg * seg_cap;
        u32 turn = 0;
        for (u32 i0 = part * 32; i0 < cnt; i0 += parts * 32, ++turn) {
            const bool live = part_count ? (turn * 32u + (u32)lane < left_n) : (i0 + lane < seg_cnt);
            bool exact = false;
            u32 u = 0, nd = 0;
            u32 dist[B2_MAXD];
#pragma unroll
            for (int i = 0; i < B2_MAXD; ++i) dist[i] = 0xFFFFFFFFu;
            if (live) {
                const uint4 rec = __ldg(list + i0 + lane);
                u = rec.x;
                u32 prim[2], pts[2];
                if (PAIRED && (rec.y & B2_DEF_GATHER) && (rec.y & B2_DEF_NAME)) atomicAdd(stats + TEC_BS_CRASH_NAME, 1ULL);   // :92-94
                const int np = b2_probes<PAIRED>(rec, sv, shift, cmask, lim, sv_n_chrom, prim, pts);
                for (int p = 0; p < np; ++p) {
                    const u32 ra = pts[p] & 0xFFFFu, rb = pts[p] >> 16;
                    const PointK pa = make_point(ra), pb = make_point(rb);
                    const int rm = max(ra == S2_R_NONE ? -1 : (int)ra, rb == S2_R_NONE ? -1 : (int)rb);
                    u32 sec = prim[p];
                    const u32 ovf = __ldg(sv.ovf_first + sec);             // requested together with the primary sector
                    for (;;) {
                        const Sector s = ld_sector(sv.sectors, sec);
                        const u32 header = s.w[2] >> 16;
                        const bool first = sec == prim[p];
                        if (first && (header & S2_H_EDGE)) exact = true;
                        u32 hit = b2_sector_hits(s, pa, pb);
                        while (hit) {
                            const u32 low = hit & (0u - hit);
                            hit ^= low;
                            u32 w;
                            if (SET) {
                                // HB0 / HB1 = bits 15 / 31 (slots in w6), HB2 / HB3 = bits 14 / 30 (w7), HB4 = bit 13 (high half of w5)
                                const u32 word = (low & (HB0 | HB1)) ? s.w[6] : (low & (HB2 | HB3)) ? s.w[7] : s.w[5];
                                w = (low & (HB1 | HB3 | HB4)) ? (word >> 16) : (word & 0xFFFFu);
                            } else {
                                w = (low == HB0) ? sector_slot_c<0>(s) : (low == HB1) ? sector_slot_c<1>(s) : (low == HB2) ? sector_slot_c<2>(s)
                                    : (low == HB3) ? sector_slot_c<3>(s) : sector_slot_c<4>(s);
                            }
                            bool found = false;
#pragma unroll
                            for (int j = 0; j < B2_MAXD; ++j) found |= (dist[j] == w);
                            if (!found) {
                                if (SET) {
#pragma unroll
                                    for (int j = B2_MAXD - 1; j > 0; --j) dist[j] = dist[j - 1];
                                    dist[0] = w;                                       // (the one that falls off the end: nd > B2_MAXD below)
                                } else {
#pragma unroll
                                    for (int j = 0; j < B2_MAXD; ++j) if ((u32)j == nd) dist[j] = w;
                                }
                                ++nd;                                                  // nd > B2_MAXD: overflow
                            }
                        }
                        if (!(header & S2_H_MORE)) break;
                        if (!(first && (header & S2_H_EDGE)) && rm < (int)(header & S2_THR_MASK)) break;
                        sec = first ? ovf : sec + 1;
                    }
                }
                if (nd > B2_MAXD) exact = true;
                if (exact) nd = 0;
                if (nd) {                                                              // :128 result not empty
                    n_assigned++;                                                      // :149
                    if (!sv.all_counted) {
                        u32 typemask = 0;
#pragma unroll
                        for (int j = 0; j < B2_MAXD; ++j) if ((u32)j < nd) typemask |= 1u << __ldg(sv.slot_type + dist[j]);
                        const u32 counted = (1u << TEC_T_GENE) | (1u << TEC_T_TE) | (1u << TEC_T_SNRNA);
                        if (!(typemask & counted)) {
                            if (typemask & (1u << TEC_T_ENHANCER)) atomicAdd(stats + TEC_BS_CRASH_ENHANCER, 1ULL);   // :145-147
                            nd = 0;
                        }
                    }
                }
            }
            // the warp is converged here: the ensg to count go through the queue
#pragma unroll
            for (int j = 0; j < B2_MAXD; ++j) {
                const bool on = (u32)j < nd;
                const u32 m = __ballot_sync(0xFFFFFFFFu, on);
                if (!m) break;
                if (on) q[(q_tail + __popc(m & lt_mask)) & (B2_QCAP2 - 1u)] = (unsigned short)dist[j];
                q_tail += __popc(m);
            }
            drain(false);
            flag_slow_warp(slow_list, exact, u, lt_mask);
        }
    }
    drain(true);
    const u64 a_sum = warp_sum((u64)n_assigned);
    if (lane == 0 && a_sum) atomicAdd(stats + TEC_BS_ASSIGNED, a_sum);
    __syncthreads();
    for (int i = threadIdx.x; i < TEC_HOT_SLOTS; i += blockDim.x) {
        const u32 x = s_hot[i];
        if (x) atomicAdd(counts + i, (u64)x);
    }
}
