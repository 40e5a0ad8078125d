// Shared declarations for libtecount (sm_100a).  See include/tecount.h for the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/tecount.h"

typedef unsigned long long u64;
typedef unsigned int u32;

// packed feature word: ensg id (24 bits) | type code (3 bits) | strand code (3 bits; 7 = missing)
#define TEC_ENSG_BITS 24
#define TEC_ENSG_MASK 0xFFFFFFu
__host__ __device__ __forceinline__ u32 info_pack(u32 ensg, u32 type, u32 strand) {
    return (ensg & TEC_ENSG_MASK) | ((type & 7u) << 24) | (((strand > 6u) ? 7u : strand) << 27);
}
__host__ __device__ __forceinline__ u32 info_ensg(u32 w) { return w & TEC_ENSG_MASK; }
__host__ __device__ __forceinline__ u32 info_type(u32 w) { return (w >> 24) & 7u; }
__host__ __device__ __forceinline__ u32 info_strand(u32 w) { return (w >> 27) & 7u; }

// Device view of the annotation index: per chromosome, features sorted by (L, R).
//   pmaxR[i] = max(R[lo..i]) inside the chromosome (the "max-end prefix" that bounds the sweep)
//   dir      = coarse directory: dir[dir_off[c] + k] = #features of chromosome c with L < (k << shift)
struct IndexView {
    const int32_t* L;
    const int32_t* R;
    const int32_t* pmaxR;
    const u32* info;
    const int64_t* chrom_off;   // n_chrom + 1
    const u32* dir;
    const int64_t* dir_off;     // n_chrom + 1
    const uint8_t* chrom_valid; // n_chrom: the chromosome is a key of genelist.buckets (it has a feature row)
    int n_chrom;
    int shift;
    int bs;                     // bucket size (10000)
    int n_ensg;
};

// te_count.py:100 / :216 / :614  `chrom not in buckets`
__device__ __forceinline__ bool chrom_in_index(const IndexView& iv, int c) { return c < iv.n_chrom && __ldg(iv.chrom_valid + c) != 0; }

// Python-style floor division by a positive divisor (te_count.py:106 `(loc1-1)//bucket_size`).
__host__ __device__ __forceinline__ int floordiv(int a, int b) {
    int q = a / b;
    return (a % b < 0) ? q - 1 : q;
}

__device__ __forceinline__ u64 warp_sum(u64 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

#define TEC_CUDA(call)                                                                   \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);               \
            if (e_ == cudaErrorMemoryAllocation) { cudaGetLastError(); return TEC_ERR_NOMEM; }   /* not sticky: clear it */ \
            return TEC_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)

#define TEC_FAIL(code, msg)  \
    do {                     \
        ctx->err = (msg);    \
        return (code);       \
    } while (0)
