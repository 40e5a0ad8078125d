// Stable LSD radix sort of (key, u32 value) pairs with digits of up to 11 bits, written for sm_100a.
//
// Used by the single-cell path (sc.cuh): the (cell, UMI) key groups of te_count.py:440-473 are found by
// sorting the survivors by a packed key, value = position in the file.  A library sort moves the pairs once
// per 8 bits of key; here a pass takes 11 bits, so the 41 key bits of the 10x configuration (17 bits of cell,
// 24 of UMI) need four passes instead of six.
//
// One pass = three kernels, no inter-CTA waiting (nothing can hang):
//   rdx_hist_kernel     every CTA owns a contiguous chunk of the input (tiles_per_cta tiles of RDX_TILE items) and
//                       counts its digits in shared memory: counts[digit][cta]
//   rdx_scan_kernel     exclusive prefix sum over (digit major, CTA minor): where each CTA's run of each digit starts
//   rdx_scatter_kernel  the same CTA walks its chunk tile by tile.  A tile (6144 pairs, 24 per thread, 72 KB of shared
//                       memory, two CTAs per SM) is sorted by its digit inside shared memory with two stable
//                       split steps (low 5 bits, then the high bits; ranks from ballots and per-lane running counts,
//                       rdx_split), so equal digits sit together and leave the SM as runs of consecutive addresses.
// Bytes per pass and pair: key read twice (histogram, scatter), value read once, both written once.
#pragma once
#include "common.cuh"

#define RDX_THREADS 256
#define RDX_WARPS (RDX_THREADS / 32)
#define RDX_ITEMS 24
#define RDX_TILE (RDX_THREADS * RDX_ITEMS)          // 6144
#define RDX_WARP_ITEMS (32 * RDX_ITEMS)             // 768: a warp's items are tile positions [w * 768, w * 768 + 768)
#define RDX_MAX_BITS 11
#define RDX_MAX_BINS (1 << RDX_MAX_BITS)

template <class K> __device__ __forceinline__ u32 rdx_digit(K key, int shift, u32 mask) { return (u32)(key >> shift) & mask; }

template <class K>
__global__ void __launch_bounds__(1024)
rdx_hist_kernel(const K* __restrict__ keys, int64_t n, int shift, int bits, int64_t tiles_per_cta, u32* __restrict__ counts) {
    __shared__ u32 hist[RDX_MAX_BINS];
    const u32 nb = 1u << bits, mask = nb - 1u;
    for (u32 i = threadIdx.x; i < nb; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const int64_t lo = (int64_t)blockIdx.x * tiles_per_cta * RDX_TILE;
    const int64_t hi = min(n, lo + tiles_per_cta * RDX_TILE);
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) atomicAdd(&hist[rdx_digit(keys[i], shift, mask)], 1u);
    __syncthreads();
    for (u32 i = threadIdx.x; i < nb; i += blockDim.x) counts[(size_t)i * gridDim.x + blockIdx.x] = hist[i];
}

// exclusive prefix sum of m counters, one CTA of 1024 threads (m = bins x CTAs, a few hundred thousand)
__global__ void __launch_bounds__(1024)
rdx_scan_kernel(u32* __restrict__ v, int64_t m) {
    __shared__ u32 warp_tot[32];
    const int64_t per = (m + 1023) / 1024;
    const int64_t a = min(m, (int64_t)threadIdx.x * per), b = min(m, a + per);
    u32 sum = 0;
    for (int64_t i = a; i < b; ++i) sum += v[i];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    u32 incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    if (w == 0) {
        u32 x = warp_tot[lane], y = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u32 t = __shfl_up_sync(0xFFFFFFFFu, y, d);
            if (lane >= d) y += t;
        }
        warp_tot[lane] = y - x;
    }
    __syncthreads();
    u32 run = warp_tot[w] + incl - sum;
    for (int64_t i = a; i < b; ++i) { const u32 c = v[i]; v[i] = run; run += c; }
}

// One stable split step of a tile: dig[r] < 2^NB is the bin of the thread's r-th item (tile position
// p0 + 32 r, p0 = warp * 768 + lane); returns pos[r] = its position after a stable sort of the tile by bin.  Items
// at tile positions >= n_valid are padding (the last tile of the input): they keep their position, the others are
// ranked among themselves and land below n_valid.
// Ranks inside the warp need neither shared memory nor match.any: lane L keeps the running count of bin L (and of
// bin L + 32 when NB == 6) in a register; per round NB ballots give every lane both the lanes that share its item's
// bin (AND of the ballots or their complements by the item's bits) and the lanes whose item falls in ITS bin (the
// same by the lane's bits); the base of an item's bin comes from the lane that owns the bin by one shuffle.
// cnt: shared [RDX_WARPS][64], base: shared [64].  Ends with the CTA synchronised.
template <int NB>
__device__ __forceinline__ void rdx_split(const u32 (&dig)[RDX_ITEMS], u32 (&pos)[RDX_ITEMS], u32 p0, u32 n_valid, u32 (*cnt)[64], u32* base) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const u32 lt = (1u << lane) - 1u;
    u32 run0 = 0, run1 = 0;
#pragma unroll
    for (int r = 0; r < RDX_ITEMS; ++r) {
        const u32 d = dig[r];
        u32 peers = __ballot_sync(0xFFFFFFFFu, p0 + 32u * (u32)r < n_valid);
        u32 mem0 = peers, mem1 = peers;
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            const u32 b = __ballot_sync(0xFFFFFFFFu, (d >> k) & 1u);
            peers &= ((d >> k) & 1u) ? b : ~b;
            if (k < 5) { const u32 x = ((lane >> k) & 1) ? b : ~b; mem0 &= x; mem1 &= x; }
            else { mem0 &= ~b; mem1 &= b; }
        }
        u32 c = __shfl_sync(0xFFFFFFFFu, run0, (int)(d & 31u));
        if (NB == 6) { const u32 c1 = __shfl_sync(0xFFFFFFFFu, run1, (int)(d & 31u)); c = (d & 32u) ? c1 : c; }
        pos[r] = c + __popc(peers & lt);
        run0 += __popc(mem0);
        if (NB == 6) run1 += __popc(mem1);
    }
    cnt[w][lane] = run0;
    cnt[w][lane + 32] = NB == 6 ? run1 : 0u;
    __syncthreads();
    // per bin: exclusive sum over the warps, bin totals
    if (threadIdx.x < 64) {
        u32 run = 0;
#pragma unroll
        for (int x = 0; x < RDX_WARPS; ++x) { const u32 c = cnt[x][threadIdx.x]; cnt[x][threadIdx.x] = run; run += c; }
        base[threadIdx.x] = run;
    }
    __syncthreads();
    if (w == 0) {                                   // exclusive sum of the 64 bin totals: two per lane
        const u32 v0 = base[lane * 2], v1 = base[lane * 2 + 1];
        u32 incl = v0 + v1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += t;
        }
        base[lane * 2] = incl - v0 - v1;
        base[lane * 2 + 1] = incl - v1;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RDX_ITEMS; ++r) {
        const u32 p = p0 + 32u * (u32)r;
        pos[r] = p < n_valid ? pos[r] + base[dig[r]] + cnt[w][dig[r]] : p;
    }
    __syncthreads();
}

template <class K> struct RdxSmem {
    K keys[RDX_TILE];
    u32 vals[RDX_TILE];
    u32 goff[RDX_MAX_BINS];                         // where the CTA's next item of each digit goes
    u32 delta[RDX_MAX_BINS];                        // per tile: goff[d] - (tile position of the digit's first item)
    u32 cnt[RDX_WARPS][64];
    u32 base[64];
};

template <class K, bool HAS_VALUES>
__global__ void __launch_bounds__(RDX_THREADS, 2)
rdx_scatter_kernel(const K* __restrict__ keys_in, const u32* __restrict__ vals_in, K* __restrict__ keys_out, u32* __restrict__ vals_out,
                   int64_t n, int shift, int bits, int64_t tiles_per_cta, const u32* __restrict__ starts) {
    extern __shared__ __align__(16) unsigned char rdx_smem_raw[];
    RdxSmem<K>& sm = *reinterpret_cast<RdxSmem<K>*>(rdx_smem_raw);
    const u32 nb = 1u << bits, mask = nb - 1u;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (u32 i = threadIdx.x; i < nb; i += RDX_THREADS) sm.goff[i] = starts[(size_t)i * gridDim.x + blockIdx.x];
    const int64_t n_tiles = (n + RDX_TILE - 1) / RDX_TILE;
    const int64_t t0 = (int64_t)blockIdx.x * tiles_per_cta, t1 = min(n_tiles, t0 + tiles_per_cta);
    const int lo_bits = bits < 5 ? bits : 5;
    const u32 lo_mask = (1u << lo_bits) - 1u;
    const u32 p0 = (u32)w * RDX_WARP_ITEMS + (u32)lane;
    __syncthreads();
    for (int64_t tile = t0; tile < t1; ++tile) {
        const int64_t g0 = tile * RDX_TILE;
        // the items of a partial tile (the last one of the input) are padded; padding always sits at tile positions
        // >= n_valid: it starts there and the split steps leave it where it is
        const u32 n_valid = (u32)min((int64_t)RDX_TILE, n - g0);
        K key[RDX_ITEMS];
        u32 val[RDX_ITEMS], dig[RDX_ITEMS], pos[RDX_ITEMS];
#pragma unroll
        for (int r = 0; r < RDX_ITEMS; ++r) {
            const u32 p = p0 + (u32)r * 32u;
            key[r] = p < n_valid ? keys_in[g0 + p] : (K)0;
            if (HAS_VALUES) val[r] = p < n_valid ? vals_in[g0 + p] : 0u;
        }
        // ---- split 1: low bits of the digit
#pragma unroll
        for (int r = 0; r < RDX_ITEMS; ++r) dig[r] = rdx_digit(key[r], shift, mask) & lo_mask;
        rdx_split<5>(dig, pos, p0, n_valid, sm.cnt, sm.base);
#pragma unroll
        for (int r = 0; r < RDX_ITEMS; ++r) {
            sm.keys[pos[r]] = key[r];
            if (HAS_VALUES) sm.vals[pos[r]] = val[r];
        }
        __syncthreads();
        // ---- split 2: high bits of the digit
#pragma unroll
        for (int r = 0; r < RDX_ITEMS; ++r) {
            const u32 p = p0 + (u32)r * 32u;
            key[r] = sm.keys[p];
            if (HAS_VALUES) val[r] = sm.vals[p];
            dig[r] = rdx_digit(key[r], shift, mask) >> lo_bits;
        }
        __syncthreads();                                               // everything is read before anything is overwritten
        rdx_split<6>(dig, pos, p0, n_valid, sm.cnt, sm.base);
#pragma unroll
        for (int r = 0; r < RDX_ITEMS; ++r) {
            sm.keys[pos[r]] = key[r];
            if (HAS_VALUES) sm.vals[pos[r]] = val[r];
        }
        __syncthreads();
        // ---- the tile is sorted by digit: the first item of every run fixes where the run goes
#pragma unroll
        for (int r = 0; r < RDX_ITEMS; ++r) {
            const u32 p = p0 + (u32)r * 32u;
            key[r] = sm.keys[p];
            dig[r] = rdx_digit(key[r], shift, mask);
            if (p < n_valid && (p == 0 || rdx_digit(sm.keys[p - 1], shift, mask) != dig[r])) sm.delta[dig[r]] = sm.goff[dig[r]] - p;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < RDX_ITEMS; ++r) {
            const u32 p = p0 + (u32)r * 32u;
            if (p < n_valid) {
                const u32 dst = sm.delta[dig[r]] + p;
                keys_out[dst] = key[r];
                if (HAS_VALUES) vals_out[dst] = sm.vals[p];
                // the last item of a run moves the digit's write position past the run
                if (p + 1 == n_valid || rdx_digit(sm.keys[p + 1], shift, mask) != dig[r]) sm.goff[dig[r]] = dst + 1u;
            }
        }
        __syncthreads();
    }
}

// ---- host side -----------------------------------------------------------------------------------------------------
struct RdxPlan {
    int n_ctas = 0;
    int64_t tiles_per_cta = 0;
    size_t counts_bytes = 0;                        // scratch: u32 [2048][n_ctas]
};

static inline RdxPlan rdx_plan(int64_t n, int n_sm) {
    RdxPlan p;
    const int64_t n_tiles = std::max<int64_t>(1, (n + RDX_TILE - 1) / RDX_TILE);
    p.n_ctas = (int)std::min<int64_t>(n_tiles, (int64_t)n_sm * 2);
    p.tiles_per_cta = (n_tiles + p.n_ctas - 1) / p.n_ctas;
    p.n_ctas = (int)((n_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta);
    p.counts_bytes = (size_t)RDX_MAX_BINS * p.n_ctas * 4;
    return p;
}

// one stable pass on key bits [shift, shift + bits), bits <= 11; scratch: plan.counts_bytes
template <class K, bool HAS_VALUES>
static cudaError_t rdx_pass(const K* keys_in, const u32* vals_in, K* keys_out, u32* vals_out, int64_t n, int shift, int bits,
                            const RdxPlan& plan, u32* scratch, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    rdx_hist_kernel<K><<<plan.n_ctas, 1024, 0, st>>>(keys_in, n, shift, bits, plan.tiles_per_cta, scratch);
    rdx_scan_kernel<<<1, 1024, 0, st>>>(scratch, (int64_t)(1 << bits) * plan.n_ctas);
    auto kfn = rdx_scatter_kernel<K, HAS_VALUES>;
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RdxSmem<K>));
    if (e != cudaSuccess) return e;
    kfn<<<plan.n_ctas, RDX_THREADS, sizeof(RdxSmem<K>), st>>>(keys_in, vals_in, keys_out, vals_out, n, shift, bits, plan.tiles_per_cta, scratch);
    return cudaGetLastError();
}

// stable sort on key bits [bit_lo, bit_hi) in ceil((bit_hi - bit_lo) / 11) passes of equal width, ping-pong between
// (keys_a, vals_a) and (keys_b, vals_b); *in_b says where the result is.  n < 2^32 - 2^13.
template <class K, bool HAS_VALUES>
static cudaError_t rdx_sort(K* keys_a, u32* vals_a, K* keys_b, u32* vals_b, int64_t n, int bit_lo, int bit_hi, int n_sm,
                            u32* scratch, cudaStream_t st, bool* in_b, int* n_passes = nullptr) {
    const int total = std::max(0, bit_hi - bit_lo);
    const int passes = (total + RDX_MAX_BITS - 1) / RDX_MAX_BITS;
    const RdxPlan plan = rdx_plan(n, n_sm);
    bool flip = false;
    int done = 0;
    for (int i = 0; i < passes; ++i) {
        const int bits = (total - done + (passes - i) - 1) / (passes - i);
        cudaError_t e = flip ? rdx_pass<K, HAS_VALUES>(keys_b, vals_b, keys_a, vals_a, n, bit_lo + done, bits, plan, scratch, st)
                             : rdx_pass<K, HAS_VALUES>(keys_a, vals_a, keys_b, vals_b, n, bit_lo + done, bits, plan, scratch, st);
        if (e != cudaSuccess) return e;
        done += bits;
        flip = !flip;
    }
    *in_b = flip;
    if (n_passes) *n_passes = passes;
    return cudaSuccess;
}
