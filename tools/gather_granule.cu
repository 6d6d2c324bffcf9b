// tools/gather_granule.cu -- what a random gather costs on B200: DRAM bytes moved and time per access
// for aligned items of 8..256 bytes fetched from a table much larger than L2, by one lane or by
// co-operating lanes, with the L2 prefetch-size qualifiers.  Evidence for the multilinear layouts
// (DESIGN.md 3.3): run it plain for the timings and under
//   ncu --metrics dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,gpu__time_duration.sum
// for the bytes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/gather_granule tools/gather_granule.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ double2 ldg16(const void *p)
{
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ double2 ldg16_l2_64(const void *p)
{
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ double2 ldg16_l2_128(const void *p)
{
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ double ldg8(const void *p)
{
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ldg32(const void *p)  // one 256-bit load
{
    double a, b, c, d;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
    return a + b + c + d;
}

// MODE 0: one lane fetches the whole item with 16-byte loads (8-byte load for S = 8)
// MODE 1: S/16 consecutive lanes fetch one item together (one 16-byte piece each)
// MODE 2: one lane, 256-bit loads (S >= 32)
// MODE 3: as 0 with L2::64B, MODE 4: as 0 with L2::128B prefetch-size qualifiers
// MODE 5: one lane fetches only the FIRST 16 bytes of an S-aligned item (what a partial use costs)
template <int S, int MODE>
__global__ void gather(const unsigned char *__restrict__ tab, uint64_t nitems, uint64_t naccess, double *sink, uint64_t seed)
{
    double acc = 0.0;
    const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x, nth = (uint64_t)gridDim.x * blockDim.x;
    if (MODE == 1) {
        constexpr int L = S / 16;
        const uint64_t grp = tid / L, ngrp = nth / L;
        const int piece = (int)(tid % L);
        for (uint64_t i = grp; i < naccess; i += ngrp) {
            const uint64_t item = mix(i ^ seed) % nitems;
            const double2 v = ldg16(tab + item * S + piece * 16);
            acc += v.x + v.y;
        }
    } else {
        for (uint64_t i = tid; i < naccess; i += nth) {
            const uint64_t item = mix(i ^ seed) % nitems;
            const unsigned char *p = tab + item * S;
            if (S == 8) acc += ldg8(p);
            else if (MODE == 5) { const double2 v = ldg16(p); acc += v.x + v.y; }
            else if (MODE == 2) {
#pragma unroll
                for (int k = 0; k < S / 32; ++k) acc += ldg32(p + 32 * k);
            } else {
#pragma unroll
                for (int k = 0; k < S / 16; ++k) {
                    const double2 v = MODE == 3 ? ldg16_l2_64(p + 16 * k) : MODE == 4 ? ldg16_l2_128(p + 16 * k) : ldg16(p + 16 * k);
                    acc += v.x + v.y;
                }
            }
        }
    }
    if (acc == 1.2345e-300) sink[0] = acc;
}

template <int S, int MODE>
void run(const unsigned char *tab, uint64_t bytes, uint64_t naccess, double *sink)
{
    const uint64_t nitems = bytes / S;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int grid = 148 * 8, block = 256;
    gather<S, MODE><<<grid, block>>>(tab, nitems, naccess / 8, sink, 1);
    cudaEventRecord(e0);
    gather<S, MODE><<<grid, block>>>(tab, nitems, naccess, sink, 2);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("{\"item_bytes\": %d, \"mode\": %d, \"accesses\": %llu, \"ms\": %.4f, \"accesses_per_s\": %.4g, \"useful_GBs\": %.1f, \"err\": \"%s\"}\n", S, MODE,
           (unsigned long long)naccess, ms, naccess / (ms * 1e-3), (MODE == 5 ? 16.0 : (double)S) * naccess / (ms * 1e-3) / 1e9,
           cudaGetErrorString(cudaGetLastError()));
    fflush(stdout);
}

__global__ void fill(double *t, uint64_t n)
{
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) t[i] = (double)(i & 1023);
}

int main(int argc, char **argv)
{
    const uint64_t bytes = (argc > 1 ? strtoull(argv[1], 0, 10) : 4096ull) << 20;  // table MiB
    const uint64_t naccess = argc > 2 ? strtoull(argv[2], 0, 10) : (1ull << 27);
    unsigned char *tab;
    double *sink;
    cudaMalloc(&tab, bytes);
    cudaMalloc(&sink, 8);
    fill<<<1184, 256>>>((double *)tab, bytes / 8);
    cudaDeviceSynchronize();
    run<8, 0>(tab, bytes, naccess, sink);
    run<16, 0>(tab, bytes, naccess, sink);
    run<32, 0>(tab, bytes, naccess, sink);
    run<32, 1>(tab, bytes, naccess, sink);
    run<32, 2>(tab, bytes, naccess, sink);
    run<32, 5>(tab, bytes, naccess, sink);
    run<64, 0>(tab, bytes, naccess, sink);
    run<64, 1>(tab, bytes, naccess, sink);
    run<64, 2>(tab, bytes, naccess, sink);
    run<64, 5>(tab, bytes, naccess, sink);
    run<128, 0>(tab, bytes, naccess, sink);
    run<128, 1>(tab, bytes, naccess, sink);
    run<128, 2>(tab, bytes, naccess, sink);
    run<128, 5>(tab, bytes, naccess, sink);
    run<256, 1>(tab, bytes, naccess / 2, sink);
    run<16, 3>(tab, bytes, naccess, sink);
    run<16, 4>(tab, bytes, naccess, sink);
    run<32, 3>(tab, bytes, naccess, sink);
    run<64, 3>(tab, bytes, naccess, sink);
    run<64, 4>(tab, bytes, naccess, sink);
    return 0;
}
