// Microbenchmarks behind the round-2 bulk kernel design (run through gpurun):
//   gather   divergent global loads of 4 / 16 / 32 bytes per lane from an L2-resident table: cost per warp instruction
//   atoms    shared-memory reductions: every lane active (misses go to a per-lane scratch word) vs only the hit lanes
//   ldsg     32 bytes per lane gathered from shared memory
// Prints SM cycles per warp-level operation (all warps of the SM together), from the kernel's elapsed time.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
typedef unsigned int u32;
__device__ __forceinline__ u32 mix(u32 x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template <int BYTES>
__global__ void __launch_bounds__(1024, 1) k_gather(const u32* __restrict__ tab, u32 n_sec, int iters, u32* out) {
    u32 acc = 0, h = mix(blockIdx.x * 1024 + threadIdx.x);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            h = h * 1664525u + 1013904223u;
            const u32* p = tab + (size_t)(mix(h) % n_sec) * 8;
            if (BYTES == 4) { u32 r; asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p)); acc += r; }
            else if (BYTES == 16) { u32 a, b, c, d; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p)); acc += a + b + c + d; }
            else if (BYTES == 32) { u32 w[8]; asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p)); acc += w[0] + w[3] + w[7]; }
            else { u32 a, b, c, d, e, f, g, hh; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
                   asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(e), "=r"(f), "=r"(g), "=r"(hh) : "l"(p + 4)); acc += a + d + e + hh; }
        }
    }
    if (acc == 0xdeadbeef) out[0] = acc;
}

// mode 0: 5 reductions per unit, every lane active, misses to a per-lane scratch word; mode 1: only the hit lanes (branch);
// mode 2: hit lanes only, through a predicated red (inline PTX)
__global__ void __launch_bounds__(1024, 1) k_atoms(int iters, int mode, u32 one, u32 n_slots, u32* out) {
    extern __shared__ u32 sh[];
    for (u32 i = threadIdx.x; i < n_slots + 32; i += 1024) sh[i] = 0;
    __syncthreads();
    const u32 base = (u32)__cvta_generic_to_shared(sh), scratch = base + (n_slots + (threadIdx.x & 31)) * 4;
    u32 h = mix(blockIdx.x * 1024 + threadIdx.x);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            h = h * 1664525u + 1013904223u;
            const u32 m = mix(h);
            const bool hit = (m & 0xFFFF) < 22282;                      // 34 % of the lanes
            const u32 slot = (m >> 16) % 1200 + ((m >> 8) & 1 ? 0 : (m >> 12) % n_slots) % n_slots;
            const u32 addr = base + (slot % n_slots) * 4;
            if (mode == 0) { const u32 a = hit ? addr : scratch; asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(a), "r"(one) : "memory"); }
            else if (mode == 1) { if (hit) asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(addr), "r"(one) : "memory"); }
            else { asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p red.shared.add.u32 [%0], %1;\n\t}" :: "r"(addr), "r"(one), "r"((u32)hit) : "memory"); }
        }
    }
    __syncthreads();
    u32 s = 0;
    for (u32 i = threadIdx.x; i < n_slots; i += 1024) s += sh[i];
    if (s == 0xdeadbeef) out[0] = s;
}

__global__ void __launch_bounds__(1024, 1) k_ldsg(int iters, u32 n_sec, u32* out) {
    extern __shared__ uint4 sh4[];
    for (u32 i = threadIdx.x; i < n_sec * 2; i += 1024) sh4[i] = make_uint4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    u32 acc = 0, h = mix(blockIdx.x * 1024 + threadIdx.x);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            h = h * 1664525u + 1013904223u;
            const u32 s = mix(h) % n_sec;
            const uint4 a = sh4[s * 2], b = sh4[s * 2 + 1];
            acc += a.x + a.w + b.y + b.w;
        }
    }
    if (acc == 0xdeadbeef) out[0] = acc;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("device %s SMs %d clock %d kHz\n", p.name, p.multiProcessorCount, clk);
    const u32 n_sec = (32u << 20) / 32;
    u32* tab; CK(cudaMalloc(&tab, (size_t)n_sec * 32)); CK(cudaMemset(tab, 1, (size_t)n_sec * 32));
    u32* out; CK(cudaMalloc(&out, 64));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    const int blocks = p.multiProcessorCount;
    const double hz = 1.965e9;
#define RUN(name, launch, ops_per_thread_iter, iters)                                                              \
    for (int rep = 0; rep < 3; ++rep) { cudaEventRecord(e0); launch; cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1); } \
    printf("%-34s %.3f ms  %.2f SM-cycles per warp-op (at 1.965 GHz)  %.1f G lane-ops/s\n", name, ms, ms * 1e-3 * hz / ((double)(iters) * (ops_per_thread_iter) * 32),  \
           (double)blocks * 1024 * (iters) * (ops_per_thread_iter) / ms / 1e6);
    RUN("gather 4 B/lane (L2-resident 32 MB)", (k_gather<4><<<blocks, 1024>>>(tab, n_sec, 2000, out)), 4, 2000)
    RUN("gather 16 B/lane", (k_gather<16><<<blocks, 1024>>>(tab, n_sec, 2000, out)), 4, 2000)
    RUN("gather 32 B/lane (v8.b32)", (k_gather<32><<<blocks, 1024>>>(tab, n_sec, 2000, out)), 4, 2000)
    RUN("gather 2 x 16 B/lane", (k_gather<33><<<blocks, 1024>>>(tab, n_sec, 2000, out)), 4, 2000)
    const u32 n_slots = 39200;
    const int dyn = (n_slots + 32) * 4;
    CK(cudaFuncSetAttribute(k_atoms, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    RUN("red.shared all lanes + scratch", (k_atoms<<<blocks, 1024, dyn>>>(2000, 0, 1, n_slots, out)), 5, 2000)
    RUN("red.shared hit lanes (branch)", (k_atoms<<<blocks, 1024, dyn>>>(2000, 1, 1, n_slots, out)), 5, 2000)
    RUN("red.shared hit lanes (predicated)", (k_atoms<<<blocks, 1024, dyn>>>(2000, 2, 1, n_slots, out)), 5, 2000)
    const u32 ls = 3200;   // 100 KB of sectors
    CK(cudaFuncSetAttribute(k_ldsg, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(ls * 32)));
    RUN("LDS gather 32 B/lane (100 KB)", (k_ldsg<<<blocks, 1024, ls * 32>>>(2000, ls, out)), 4, 2000)
    return 0;
}
