// Micro-benchmark (profiling aid, not product code): is a row-range-partitioned read_id vote worth building?
// (1) random 8-byte gathers confined to a window of W MB of the 400 MB matrix (L2-resident when W << 126 MB);
// (2) the same gathers fed by a streamed (row, slot) tuple list, each result AND-ed into a byte of a
//     streaming accumulator array with red.and.b32 (what a per-partition gather pass would do).
//   l2_window_probe <window_MB> <tuples_millions>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t mix32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

__global__ void __launch_bounds__(256) window_gather(const uint2* __restrict__ m, uint32_t rows, uint32_t iters, uint32_t* out) {
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    uint32_t acc = 0;
    constexpr int ILP = 8;
    for (uint32_t it = 0; it < iters; it += ILP) {
        uint2 v[ILP];
#pragma unroll
        for (int u = 0; u < ILP; u++) v[u] = __ldg(m + (uint32_t)(((uint64_t)mix32((it + u) * nthr + gtid + 0x9e3779b9u) * rows) >> 32));
#pragma unroll
        for (int u = 0; u < ILP; u++) acc += v[u].x ^ v[u].y;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

__global__ void make_tuples(uint2* t, uint64_t n, uint32_t rows, uint32_t slot_span) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t r = (uint32_t)(((uint64_t)mix32((uint32_t)i * 2654435761u + 17) * rows) >> 32);
        uint32_t slot = (uint32_t)((double)i / n * slot_span);
        t[i] = make_uint2(r, slot);
    }
}

// MODE 0: gather only; 1: + red.and.b32 into acc bytes; 2: + 64-bit atomicAnd into 8-byte accumulators
template <int MODE>
__global__ void __launch_bounds__(256) tuple_gather(const uint2* __restrict__ m, const uint2* __restrict__ t, uint64_t n, uint32_t* acc32,
                                                    unsigned long long* acc64, uint32_t* out) {
    constexpr int ILP = 8;
    uint32_t sink = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x * ILP; base < n; base += stride * ILP) {
        uint2 tp[ILP], v[ILP];
#pragma unroll
        for (int u = 0; u < ILP; u++) { uint64_t i = base + u * blockDim.x + threadIdx.x; tp[u] = i < n ? __ldcs(t + i) : make_uint2(0, 0xffffffffu); }
#pragma unroll
        for (int u = 0; u < ILP; u++) v[u] = __ldg(m + tp[u].x);
#pragma unroll
        for (int u = 0; u < ILP; u++) {
            if (tp[u].y == 0xffffffffu) continue;
            if (MODE == 0) sink += v[u].x ^ v[u].y;
            if (MODE == 1) {
                uint32_t b = ((v[u].x >> 3) & 1) | (((v[u].y >> 7) & 1) << 1) | 0xfc;   // two candidate colours
                uint32_t sh = (tp[u].y & 3) * 8;
                uint32_t mask = ~((0xffu ^ b) << sh);
                asm volatile("red.global.and.b32 [%0], %1;" :: "l"(acc32 + (tp[u].y >> 2)), "r"(mask) : "memory");
            }
            if (MODE == 2) atomicAnd(acc64 + tp[u].y, ((unsigned long long)v[u].y << 32) | v[u].x);
        }
    }
    if (sink == 0x12345678u) out[0] = sink;
}

int main(int argc, char** argv) {
    double wmb = argc > 1 ? atof(argv[1]) : 50.0;
    uint64_t nt = (uint64_t)((argc > 2 ? atof(argv[2]) : 48.0) * 1e6);
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const uint32_t S = 50000000;
    uint2* m; uint32_t* out; CK(cudaMalloc(&m, (size_t)S * 8)); CK(cudaMemset(m, 0xff, (size_t)S * 8)); CK(cudaMalloc(&out, 64));
    const uint32_t rows = (uint32_t)(wmb * 1e6 / 8);
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float ms;
    {
        const int grid = p.multiProcessorCount * 8; const uint32_t iters = 256;
        for (int rep = 0; rep < 2; rep++) {
            CK(cudaEventRecord(a));
            window_gather<<<grid, 256>>>(m, rows, iters, out);
            CK(cudaEventRecord(b)); CK(cudaDeviceSynchronize());
        }
        CK(cudaEventElapsedTime(&ms, a, b));
        double g = (double)grid * 256 * iters;
        printf("window %.1f MB: pure random 8-B gathers %8.2f G/s (%.3f ms, %.0f M)\n", wmb, g / ms / 1e6, ms, g / 1e6);
    }
    uint2* t; CK(cudaMalloc(&t, nt * 8));
    const uint32_t slot_span = (uint32_t)(nt * 2);    // ~ every other k-mer slot is touched by a partition pass (1 - (7/8)^4 = 41 %)
    uint32_t* acc32; unsigned long long* acc64;
    CK(cudaMalloc(&acc32, slot_span)); CK(cudaMemset(acc32, 0xff, slot_span));
    CK(cudaMalloc(&acc64, (size_t)slot_span * 8)); CK(cudaMemset(acc64, 0xff, (size_t)slot_span * 8));
    make_tuples<<<1024, 256>>>(t, nt, rows, slot_span); CK(cudaDeviceSynchronize());
    const int grid = p.multiProcessorCount * 8;
    for (int mode = 0; mode < 3; mode++) {
        for (int rep = 0; rep < 3; rep++) {
            // touch another window in between so the timed pass starts with a cold window, like consecutive partition passes
            window_gather<<<grid, 256>>>(m + (S - rows), rows, 64, out);
            CK(cudaEventRecord(a));
            if (mode == 0) tuple_gather<0><<<grid, 256>>>(m, t, nt, acc32, acc64, out);
            if (mode == 1) tuple_gather<1><<<grid, 256>>>(m, t, nt, acc32, acc64, out);
            if (mode == 2) tuple_gather<2><<<grid, 256>>>(m, t, nt, acc32, acc64, out);
            CK(cudaEventRecord(b)); CK(cudaDeviceSynchronize());
        }
        CK(cudaEventElapsedTime(&ms, a, b));
        printf("window %.1f MB: tuple-fed gathers, mode %d (%s): %8.2f G tuples/s (%.3f ms for %.0f M tuples)\n", wmb, mode,
               mode == 0 ? "gather only" : mode == 1 ? "+ red.and.b32 on a byte accumulator" : "+ atomicAnd on 8-byte accumulators", nt / ms / 1e6, ms, nt / 1e6);
    }
    return 0;
}
