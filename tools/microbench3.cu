// Microbenchmarks of the SM's memory pipe, with the address arithmetic OUT of the timed loop (tools/microbench2.cu
// computed a hash and two run-time modulos per operation, which is 15 issue cycles per warp instruction on its own):
// every thread draws its addresses once, keeps them in registers and rotates through them.
// Prints SM cycles per warp-level instruction when all 32 warps of a 1024-thread CTA per SM issue the same instruction.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
typedef unsigned int u32;
__device__ __forceinline__ u32 mix(u32 x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
#define NA 8

// mode 0 random slot, all lanes; 1 random slot, a third of the lanes; 2 bank == lane (conflict free), all lanes;
// 3 one address for the whole warp; 4 random among 64 hot slots; 5 conflict free, a third of the lanes
__global__ void __launch_bounds__(1024, 1) k_red(int iters, int mode, u32 one, u32 n_slots, u32* out) {
    extern __shared__ u32 sh[];
    for (u32 i = threadIdx.x; i < n_slots + 32; i += 1024) sh[i] = 0;
    __syncthreads();
    const u32 base = (u32)__cvta_generic_to_shared(sh);
    const int lane = threadIdx.x & 31;
    u32 a[NA]; bool on[NA];
    u32 h = mix(blockIdx.x * 1024 + threadIdx.x + 77);
#pragma unroll
    for (int j = 0; j < NA; ++j) {
        h = mix(h + j);
        u32 slot = h % n_slots;
        if (mode == 2 || mode == 5) slot = (slot & ~31u) | lane;
        if (mode == 3) slot = (mix(threadIdx.x >> 5) + j * 131) % n_slots;
        if (mode == 4) slot = h % 64;
        a[j] = base + slot * 4;
        on[j] = (mode == 1 || mode == 5) ? ((h >> 20) % 3 == 0) : true;
    }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < NA; ++j)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p red.shared.add.u32 [%0], %1;\n\t}" :: "r"(a[j]), "r"(one), "r"((u32)on[j]) : "memory");
    }
    __syncthreads();
    u32 s = 0;
    for (u32 i = threadIdx.x; i < n_slots; i += 1024) s += sh[i];
    if (s == 0xdeadbeef) out[0] = s;
}

// sparse 16-bit stores to consecutive shared addresses (the hit queue's push) and full-warp 16-bit loads
__global__ void __launch_bounds__(1024, 1) k_sts(int iters, int mode, u32* out) {
    __shared__ unsigned short q[32][512];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const u32 base = (u32)__cvta_generic_to_shared(&q[w][0]);
    u32 h = mix(blockIdx.x * 1024 + threadIdx.x + 5);
    bool on[NA];
#pragma unroll
    for (int j = 0; j < NA; ++j) { h = mix(h + j); on[j] = mode == 0 ? (h % 3 == 0) : true; }
    u32 acc = 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < NA; ++j) {
            const u32 at = base + (((u32)(i * NA + j) * 11u + (u32)lane) & 511u) * 2u;
            if (mode < 2) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.u16 [%0], %1;\n\t}" :: "r"(at), "h"((unsigned short)lane), "r"((u32)on[j]) : "memory");
            else { unsigned short v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(at) : "memory"); acc += v; }
        }
    }
    if (acc == 0xdeadbeef) out[0] = acc;
}

// divergent global loads of one 32-byte sector per lane, sector numbers drawn once (8 per thread, L2-resident table)
// mode 0 ld.global.nc.L1::no_allocate.v8.b32; 1 two v4 halves; 2 plain ld.global.v8 (L1 allocate); 3 every lane pair
// shares a sector (16 distinct sectors per warp instruction); 4 v8 with only a third of the lanes active
__global__ void __launch_bounds__(1024, 1) k_gather(const u32* __restrict__ tab, u32 n_sec, int iters, int mode, u32* out) {
    u32 acc = 0, h = mix(blockIdx.x * 1024 + threadIdx.x);
    const u32* p[NA]; bool on[NA];
#pragma unroll
    for (int j = 0; j < NA; ++j) {
        h = mix(h + j);
        u32 s = h % n_sec;
        if (mode == 3) s = __shfl_sync(0xFFFFFFFFu, s, (threadIdx.x & 31) & ~1);
        p[j] = tab + (size_t)s * 8;
        on[j] = mode == 4 ? ((h >> 20) % 3 == 0) : true;
    }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < NA; ++j) {
            u32 w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            const u32* q = p[j] + (size_t)((i & 63) * 8 * 97);       // moves inside the table: no line is read twice in a row
            if (mode == 1) {
                asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "l"(q));
                asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(q + 4));
            } else if (mode == 2) {
                asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(q));
            } else if (on[j]) {
                asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(q));
            }
            acc += w[0] + w[3] + w[7];
        }
    }
    if (acc == 0xdeadbeef) out[0] = acc;
}

// coalesced streaming loads, 16 bytes per lane (the record columns)
__global__ void __launch_bounds__(1024, 1) k_stream(const uint4* __restrict__ src, size_t n16, int iters, u32* out) {
    u32 acc = 0;
    size_t i0 = (size_t)blockIdx.x * 1024 + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * 1024;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < NA; ++j) {
            uint4 v;
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src + (i0 % n16)));
            acc += v.x + v.w;
            i0 += stride;
        }
    }
    if (acc == 0xdeadbeef) out[0] = acc;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s SMs %d\n", p.name, p.multiProcessorCount);
    const u32 n_sec = (48u << 20) / 32;
    u32* tab; CK(cudaMalloc(&tab, (size_t)n_sec * 32 + 64 * 8 * 97 * 4 + 64)); CK(cudaMemset(tab, 1, (size_t)n_sec * 32 + 64 * 8 * 97 * 4 + 64));
    u32* out; CK(cudaMalloc(&out, 64));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    const int blocks = p.multiProcessorCount;
    const double hz = 1.965e9;
#define RUN(name, launch, iters)                                                              \
    for (int rep = 0; rep < 3; ++rep) { cudaEventRecord(e0); launch; cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError()); cudaEventElapsedTime(&ms, e0, e1); } \
    printf("%-58s %.3f ms  %6.2f SM-cycles per warp instruction\n", name, ms, ms * 1e-3 * hz / ((double)(iters) * NA * 32));
    const u32 n_slots = 39200;
    const int dyn = (n_slots + 32) * 4;
    CK(cudaFuncSetAttribute(k_red, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    const int IT = 4000;
    RUN("red.shared random of 39200 words, 32 lanes", (k_red<<<blocks, 1024, dyn>>>(IT, 0, 1, n_slots, out)), IT)
    RUN("red.shared random, a third of the lanes", (k_red<<<blocks, 1024, dyn>>>(IT, 1, 1, n_slots, out)), IT)
    RUN("red.shared bank == lane, 32 lanes", (k_red<<<blocks, 1024, dyn>>>(IT, 2, 1, n_slots, out)), IT)
    RUN("red.shared bank == lane, a third of the lanes", (k_red<<<blocks, 1024, dyn>>>(IT, 5, 1, n_slots, out)), IT)
    RUN("red.shared one address per warp", (k_red<<<blocks, 1024, dyn>>>(IT, 3, 1, n_slots, out)), IT)
    RUN("red.shared random of 64 words", (k_red<<<blocks, 1024, dyn>>>(IT, 4, 1, n_slots, out)), IT)
    RUN("st.shared.u16 consecutive, a third of the lanes", (k_sts<<<blocks, 1024>>>(IT, 0, out)), IT)
    RUN("st.shared.u16 consecutive, 32 lanes", (k_sts<<<blocks, 1024>>>(IT, 1, out)), IT)
    RUN("ld.shared.u16 consecutive, 32 lanes", (k_sts<<<blocks, 1024>>>(IT, 2, out)), IT)
    const int IG = 1000;
    RUN("gather 32 B/lane v8 nc no_allocate", (k_gather<<<blocks, 1024>>>(tab, n_sec, IG, 0, out)), IG)
    RUN("gather 32 B/lane as 2 x v4", (k_gather<<<blocks, 1024>>>(tab, n_sec, IG, 1, out)), IG)
    RUN("gather 32 B/lane v8 plain ld.global", (k_gather<<<blocks, 1024>>>(tab, n_sec, IG, 2, out)), IG)
    RUN("gather 32 B/lane v8, lane pairs share a sector", (k_gather<<<blocks, 1024>>>(tab, n_sec, IG, 3, out)), IG)
    RUN("gather 32 B/lane v8, a third of the lanes", (k_gather<<<blocks, 1024>>>(tab, n_sec, IG, 4, out)), IG)
    const size_t n16 = (size_t)(1u << 30) / 16;
    uint4* src; CK(cudaMalloc(&src, n16 * 16)); CK(cudaMemset(src, 1, n16 * 16));
    RUN("stream 16 B/lane coalesced (1 GiB, HBM)", (k_stream<<<blocks, 1024>>>(src, n16, 200, out)), 200)
    printf("stream: %.0f GB/s\n", (double)blocks * 1024 * 200 * NA * 16 / ms / 1e6);
    return 0;
}
