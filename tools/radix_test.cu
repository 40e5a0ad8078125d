// Correctness and speed of csrc/radix.cuh against cub::DeviceRadixSort (both stable: results must be identical).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I te_counter_b200/csrc -o tools/radix_test tools/radix_test.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cub/cub.cuh>
#include "radix.cuh"
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ u64 mix64(u64 x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
template <class K> __global__ void fill_kernel(K* k, u32* v, int64_t n, u64 seed, int skew) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        u64 x = mix64(i * 0x9E3779B97F4A7C15ULL + seed);
        if (skew) x = (x & 0xFFFFFFull) | ((mix64(x) % 1000) * (mix64(x) % 1000) / 10) << 24;    // skewed high part (cells), uniform low part (UMIs)
        k[i] = (K)x;
        v[i] = (u32)i;
    }
}
template <class T> __global__ void diff_kernel(const T* a, const T* b, int64_t n, unsigned long long* bad) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (a[i] != b[i]) atomicAdd(bad, 1ULL);
}

template <class K> int run(int64_t n, int b0, int b1, int skew, int n_sm) {
    K *ka, *kb, *kc, *kd; u32 *va, *vb, *vc, *vd, *scratch; unsigned long long* bad;
    CK(cudaMalloc(&ka, n * sizeof(K))); CK(cudaMalloc(&kb, n * sizeof(K))); CK(cudaMalloc(&kc, n * sizeof(K))); CK(cudaMalloc(&kd, n * sizeof(K)));
    CK(cudaMalloc(&va, n * 4)); CK(cudaMalloc(&vb, n * 4)); CK(cudaMalloc(&vc, n * 4)); CK(cudaMalloc(&vd, n * 4));
    CK(cudaMalloc(&bad, 8)); CK(cudaMemset(bad, 0, 8));
    const RdxPlan plan = rdx_plan(n, n_sm);
    CK(cudaMalloc(&scratch, plan.counts_bytes));
    fill_kernel<K><<<1024, 256>>>(ka, va, n, 12345 + n, skew);
    CK(cudaMemcpy(kc, ka, n * sizeof(K), cudaMemcpyDeviceToDevice)); CK(cudaMemcpy(vc, va, n * 4, cudaMemcpyDeviceToDevice));
    size_t tb = 0; void* tmp = nullptr;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tb, kc, kd, vc, vd, (int)n, b0, b1));
    CK(cudaMalloc(&tmp, tb));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms_cub = 0, ms_own = 0;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        CK(cub::DeviceRadixSort::SortPairs(tmp, tb, kc, kd, vc, vd, (int)n, b0, b1));
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms_cub, e0, e1);
    }
    bool in_b = false; int passes = 0;
    for (int rep = 0; rep < 2; ++rep) {
        if (rep) { fill_kernel<K><<<1024, 256>>>(ka, va, n, 12345 + n, skew); CK(cudaDeviceSynchronize()); }
        cudaEventRecord(e0);
        CK((rdx_sort<K, true>(ka, va, kb, vb, n, b0, b1, n_sm, scratch, 0, &in_b, &passes)));
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms_own, e0, e1);
    }
    CK(cudaGetLastError());
    diff_kernel<K><<<1024, 256>>>(in_b ? kb : ka, kd, n, bad);
    diff_kernel<u32><<<1024, 256>>>(in_b ? vb : va, vd, n, bad);
    unsigned long long h = 0; CK(cudaMemcpy(&h, bad, 8, cudaMemcpyDeviceToHost));
    const double bytes = (double)n * (sizeof(K) + 4) * 2;
    printf("K=%zu n=%lld bits[%d,%d) skew=%d: %s  own %d passes %.3f ms (%.0f GB/s per pass moved)  cub %.3f ms  mismatches %llu\n", sizeof(K), (long long)n, b0, b1, skew,
           h ? "FAIL" : "ok", passes, ms_own, passes ? bytes * passes / ms_own / 1e6 : 0.0, ms_cub, h);
    cudaFree(ka); cudaFree(kb); cudaFree(kc); cudaFree(kd); cudaFree(va); cudaFree(vb); cudaFree(vc); cudaFree(vd); cudaFree(scratch); cudaFree(tmp); cudaFree(bad);
    return h ? 1 : 0;
}

int main(int argc, char** argv) {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int n_sm = p.multiProcessorCount;
    int fails = 0;
    if (getenv("RDX_BITS")) g_rdx_max_bits = atoi(getenv("RDX_BITS"));
    if (getenv("RDX_CHUNK")) g_rdx_chunk_tiles = atoi(getenv("RDX_CHUNK"));
    if (getenv("RDX_BIG_ONLY")) {
        const int64_t big = argc > 1 ? atoll(argv[1]) : 400000000;
        fails += run<u64>(big, 18, 59, 1, n_sm);
        fails += run<u64>(big, 18, 59, 0, n_sm);
        printf(fails ? "FAILED %d\n" : "all ok\n", fails);
        return fails ? 1 : 0;
    }
    const int64_t sizes[] = {1, 31, 6143, 6144, 6145, 100000, 1000003, 20000000};
    for (int64_t n : sizes) {
        fails += run<u64>(n, 0, 11, 0, n_sm);
        fails += run<u64>(n, 18, 59, 1, n_sm);
        fails += run<u64>(n, 3, 10, 0, n_sm);
        fails += run<u32>(n, 0, 24, 0, n_sm);
        fails += run<u32>(n, 20, 30, 0, n_sm);
    }
    const int64_t big = argc > 1 ? atoll(argv[1]) : 400000000;
    fails += run<u64>(big, 18, 59, 1, n_sm);
    fails += run<u64>(big, 18, 59, 0, n_sm);
    fails += run<u32>(big, 0, 24, 0, n_sm);
    fails += run<u32>(big, 20, 30, 0, n_sm);
    printf(fails ? "FAILED %d\n" : "all ok\n", fails);
    return fails ? 1 : 0;
}
