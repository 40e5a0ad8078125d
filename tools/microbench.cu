// Microbenchmarks that size the bulk kernel's design on B200 (run through gpurun):
//   1. random 32-byte-sector reads over a working set W (is the index L2-resident at W? what rate?)
//      optionally with a concurrent 24 B/thread record stream, with/without evict-first hints
//   2. atomics: spread / hot-address global RED, shared-memory privatised
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

__device__ __forceinline__ int2 ld_stream(const int2* p, int hint) {
    int2 r;
    if (hint == 0) { r = *p; }
    else if (hint == 1) asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    else {
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.s32 {%0,%1}, [%2], %3;" : "=r"(r.x), "=r"(r.y) : "l"(p), "l"(pol));
    }
    return r;
}
__device__ __forceinline__ uint32_t ld_index(const uint32_t* p, int hint) {
    uint32_t r;
    if (hint < 3) { r = __ldg(p); }
    else {
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
    }
    return r;
}

// each thread: per unit, stream 3 x int2 (24 B, coalesced) if stream != null, then Q dependent-free random reads
__global__ void k_random(const uint32_t* __restrict__ tab, uint32_t n_sectors, const int2* __restrict__ stream,
                         int64_t n_units, int q, int hint, unsigned long long* out) {
    uint32_t acc = 0;
    for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < n_units; u += (int64_t)gridDim.x * blockDim.x) {
        uint32_t h = mix((uint32_t)u * 2654435761u + 12345u);
        if (stream) {
            int2 a = ld_stream(stream + u * 3, hint), b = ld_stream(stream + u * 3 + 1, hint), c = ld_stream(stream + u * 3 + 2, hint);
            h ^= (uint32_t)(a.x + b.y + c.x);
        }
        for (int i = 0; i < q; ++i) {
            uint32_t s = mix(h + i * 0x9e3779b9u) % n_sectors;
            acc += ld_index(tab + (size_t)s * 8, hint);
        }
    }
    if (acc == 0xdeadbeef) atomicAdd(out, 1ULL);
}

// dependent chain of 2 random reads (directory -> entries)
__global__ void k_chain(const uint32_t* __restrict__ tab, uint32_t n_sectors, int64_t n_units, unsigned long long* out) {
    uint32_t acc = 0;
    for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < n_units; u += (int64_t)gridDim.x * blockDim.x) {
        uint32_t h = mix((uint32_t)u * 2654435761u + 777u);
        uint32_t a = __ldg(tab + (size_t)(h % n_sectors) * 8);
        uint32_t b = __ldg(tab + (size_t)(mix(h ^ a) % n_sectors) * 8);
        acc += b;
    }
    if (acc == 0xdeadbeef) atomicAdd(out, 1ULL);
}

// atomics: mode 0 spread over n_addr; mode 1 zipf-ish hot (20% to addr 0, 10% to 1, ...) ; mode 2 = mode 1 with smem privatised hot set (first 4096)
__global__ void k_atomic(unsigned long long* counts, uint32_t n_addr, int64_t n_units, int mode) {
    extern __shared__ uint32_t sh[];
    if (mode == 2) { for (int i = threadIdx.x; i < 4096; i += blockDim.x) sh[i] = 0; __syncthreads(); }
    for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < n_units; u += (int64_t)gridDim.x * blockDim.x) {
        uint32_t h = mix((uint32_t)u * 2654435761u + 99u);
        uint32_t a;
        if (mode == 0) a = h % n_addr;
        else {
            // p(k) ~ 1/k: pick k = floor(exp(U * ln(1200)))
            float U = (h >> 8) * (1.0f / 16777216.0f);
            uint32_t k = (uint32_t)__expf(U * 7.09f) - 1;
            a = (h & 1) ? k : (mix(h) % n_addr);      // half the adds go to the hot TE-like set, half to genes
        }
        if (mode == 2 && a < 4096) atomicAdd(&sh[a], 1u);
        else atomicAdd(counts + a, 1ULL);
    }
    if (mode == 2) {
        __syncthreads();
        for (int i = threadIdx.x; i < 4096; i += blockDim.x) if (sh[i]) atomicAdd(counts + i, (unsigned long long)sh[i]);
    }
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s SMs %d L2 %d MB persistL2max %d MB\n", p.name, p.multiProcessorCount, p.l2CacheSize >> 20, p.persistingL2CacheMaxSize >> 20);
    const size_t maxW = size_t(512) << 20;
    uint32_t* tab; CK(cudaMalloc(&tab, maxW)); CK(cudaMemset(tab, 1, maxW));
    const int64_t n_units = 200000000;
    int2* stream; CK(cudaMalloc(&stream, n_units * 24)); CK(cudaMemset(stream, 0, n_units * 24));
    unsigned long long* out; CK(cudaMalloc(&out, 8 * 65536)); CK(cudaMemset(out, 0, 8 * 65536));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = p.multiProcessorCount * 8, threads = 256;
    float ms;
    int Ws[] = {8, 16, 32, 48, 64, 80, 96, 128, 192, 256, 512};
    for (int with_stream = 0; with_stream < 2; ++with_stream)
        for (int hint = 0; hint < 4; ++hint) {
            if (!with_stream && (hint == 1 || hint == 2)) continue;
            for (int wi = 0; wi < 11; ++wi) {
                uint32_t n_sec = (uint32_t)((size_t(Ws[wi]) << 20) / 32);
                for (int rep = 0; rep < 2; ++rep) {
                    cudaEventRecord(e0);
                    k_random<<<blocks, threads>>>(tab, n_sec, with_stream ? stream : nullptr, n_units, 2, hint, out);
                    cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
                }
                printf("random stream=%d hint=%d W=%3d MB  q=2  %.3f ms  %.1f Gsectors/s  %.2f Gunits/s\n", with_stream, hint, Ws[wi], ms, n_units * 2 / ms / 1e6, n_units / ms / 1e6);
            }
        }
    for (int wi = 0; wi < 11; wi += 2) {
        uint32_t n_sec = (uint32_t)((size_t(Ws[wi]) << 20) / 32);
        for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); k_chain<<<blocks, threads>>>(tab, n_sec, n_units, out); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1); }
        printf("chain2 W=%3d MB %.3f ms %.2f Gunits/s\n", Ws[wi], ms, n_units / ms / 1e6);
    }
    for (int mode = 0; mode < 3; ++mode) {
        for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); k_atomic<<<blocks, threads, 16384>>>(out, 40000, n_units, mode); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1); }
        printf("atomic mode=%d %.3f ms %.2f Gatom/s\n", mode, ms, n_units / ms / 1e6);
    }
    return 0;
}
