// Micro-benchmark (profiling aid, not product code): random row gathers from a BIGSI-shaped matrix.
// Answers two questions for DESIGN.md: (1) how many DRAM bytes does one 8-byte row gather cost on
// B200 under different load flavours / L2 fetch granularities, (2) what random-row bandwidth is the
// ceiling for 128-byte rows (C3) and 8-byte rows (C1/C2).
//   gather_probe <gran 0|32|64|128> <rows_millions> <row_bytes 8|16|32|128> <iters>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t mix32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template <int FLAVOUR> __device__ __forceinline__ uint2 ld8(const uint2* p) {
    uint2 v;
    if (FLAVOUR == 0) v = __ldg(p);
    else if (FLAVOUR == 1) asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    else if (FLAVOUR == 2) v = __ldcs(p);
    else if (FLAVOUR == 3) v = __ldcv(p);
    else if (FLAVOUR == 4) asm volatile("ld.global.nc.L2::64B.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    else if (FLAVOUR == 5) asm volatile("ld.global.nc.L2::128B.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    else v = *p;
    return v;
}

// LPR lanes cooperate on one row (LPR * VB bytes per row; VB = 8 or 16 bytes per lane)
template <int FLAVOUR, int LPR, int VB>
__global__ void __launch_bounds__(256) gather_kernel(const uint8_t* __restrict__ m, uint32_t S, uint32_t iters, uint32_t* out) {
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t grp = gtid / LPR, sub = gtid % LPR;
    const uint32_t ngrp = gridDim.x * blockDim.x / LPR;
    uint32_t acc = 0;
    constexpr int ILP = 8;
    for (uint32_t it = 0; it < iters; it += ILP) {
        uint32_t r[ILP];
#pragma unroll
        for (int u = 0; u < ILP; u++) r[u] = (uint32_t)(((uint64_t)mix32((it + u) * ngrp + grp + 0x9e3779b9u) * S) >> 32);
        if (VB == 8) {
            uint2 v[ILP];
#pragma unroll
            for (int u = 0; u < ILP; u++) v[u] = ld8<FLAVOUR>((const uint2*)(m + ((uint64_t)r[u] * LPR + sub) * 8));
#pragma unroll
            for (int u = 0; u < ILP; u++) acc += v[u].x ^ v[u].y;
        } else {
            uint4 v[ILP];
#pragma unroll
            for (int u = 0; u < ILP; u++) v[u] = __ldg((const uint4*)(m + ((uint64_t)r[u] * LPR + sub) * 16));
#pragma unroll
            for (int u = 0; u < ILP; u++) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}

template <int FLAVOUR, int LPR, int VB>
static void run(const char* name, const uint8_t* m, uint32_t S, uint32_t iters, uint32_t* out, int sms) {
    const int grid = sms * 8;
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    gather_kernel<FLAVOUR, LPR, VB><<<grid, 256>>>(m, S, iters, out);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    gather_kernel<FLAVOUR, LPR, VB><<<grid, 256>>>(m, S, iters, out);
    CK(cudaEventRecord(b));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    const double rows = (double)grid * 256 / LPR * iters;
    printf("%-34s rows/s %8.2f G   algorithmic %8.1f GB/s   (%.3f ms, %.0f M rows of %d B)\n", name, rows / ms / 1e6,
           rows * LPR * VB / ms / 1e6, ms, rows / 1e6, LPR * VB);
}

int main(int argc, char** argv) {
    int gran = argc > 1 ? atoi(argv[1]) : 0;
    uint32_t S = (uint32_t)((argc > 2 ? atof(argv[2]) : 50.0) * 1e6);
    int rowb = argc > 3 ? atoi(argv[3]) : 8;
    uint32_t iters = argc > 4 ? (uint32_t)atoi(argv[4]) : 1024;
    if (gran) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran); printf("set L2 fetch granularity %d: %s\n", gran, cudaGetErrorString(e)); }
    size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity); printf("L2 fetch granularity now %zu\n", g);
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    uint8_t* m; uint32_t* out;
    CK(cudaMalloc(&m, (size_t)S * rowb)); CK(cudaMemset(m, 1, (size_t)S * rowb)); CK(cudaMalloc(&out, 64));
    printf("matrix %u rows x %d B = %.1f MB, %d SMs, L2 %d MB\n", S, rowb, (double)S * rowb / 1e6, p.multiProcessorCount, p.l2CacheSize >> 20);
    if (rowb == 8) {
        run<0, 1, 8>("8B  __ldg", m, S, iters, out, p.multiProcessorCount);
        run<1, 1, 8>("8B  nc.L1::no_allocate", m, S, iters, out, p.multiProcessorCount);
        run<2, 1, 8>("8B  __ldcs", m, S, iters, out, p.multiProcessorCount);
        run<3, 1, 8>("8B  __ldcv", m, S, iters, out, p.multiProcessorCount);
        run<4, 1, 8>("8B  nc.L2::64B", m, S, iters, out, p.multiProcessorCount);
        run<5, 1, 8>("8B  nc.L2::128B", m, S, iters, out, p.multiProcessorCount);
        run<6, 1, 8>("8B  plain ld", m, S, iters, out, p.multiProcessorCount);
    } else if (rowb == 16) {
        run<0, 1, 16>("16B uint4 per lane", m, S, iters, out, p.multiProcessorCount);
    } else if (rowb == 32) {
        run<0, 4, 8>("32B 4 lanes x 8B", m, S, iters, out, p.multiProcessorCount);
        run<0, 2, 16>("32B 2 lanes x 16B", m, S, iters, out, p.multiProcessorCount);
    } else if (rowb == 128) {
        run<0, 8, 16>("128B 8 lanes x 16B", m, S, iters, out, p.multiProcessorCount);
        run<0, 16, 8>("128B 16 lanes x 8B", m, S, iters, out, p.multiProcessorCount);
    } else if (rowb == 160) {
        run<0, 10, 16>("160B 10 lanes x 16B", m, S, iters, out, p.multiProcessorCount);
    }
    return 0;
}
