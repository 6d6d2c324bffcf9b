// tools/dmma_peak.cu -- FP64 tensor-core (DMMA) throughput on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_peak tools/dmma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma16816(double (&d)[4], const double (&a)[8], const double (&b)[4])
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
__device__ __forceinline__ void mma1684(double (&d)[4], const double (&a)[2], double b)
{
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}

template <int C>
__global__ void k884(double *sink, int iters)
{
    double d0[C], d1[C];
    for (int i = 0; i < C; ++i) { d0[i] = threadIdx.x; d1[i] = i; }
    double a = 1.0 + threadIdx.x * 1e-3, b = 0.5;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < C; ++i) mma884(d0[i], d1[i], a, b);
    double r = 0;
    for (int i = 0; i < C; ++i) r += d0[i] + d1[i];
    if (r == 123.456) sink[0] = r;
}
template <int C>
__global__ void k16816(double *sink, int iters)
{
    double d[C][4];
    for (int i = 0; i < C; ++i) for (int j = 0; j < 4; ++j) d[i][j] = threadIdx.x + j;
    double a[8], b[4];
    for (int j = 0; j < 8; ++j) a[j] = 1.0 + threadIdx.x * 1e-3 + j;
    for (int j = 0; j < 4; ++j) b[j] = 0.5 + j;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < C; ++i) mma16816(d[i], a, b);
    double r = 0;
    for (int i = 0; i < C; ++i) for (int j = 0; j < 4; ++j) r += d[i][j];
    if (r == 123.456) sink[0] = r;
}
template <int C>
__global__ void k1684(double *sink, int iters)
{
    double d[C][4];
    for (int i = 0; i < C; ++i) for (int j = 0; j < 4; ++j) d[i][j] = threadIdx.x + j;
    double a[2] = {1.0 + threadIdx.x * 1e-3, 2.0}, b = 0.5;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < C; ++i) mma1684(d[i], a, b);
    double r = 0;
    for (int i = 0; i < C; ++i) for (int j = 0; j < 4; ++j) r += d[i][j];
    if (r == 123.456) sink[0] = r;
}

template <class F>
void timeit(const char *name, F launch, double flop_per_launch)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-28s %8.2f TFLOP/s  (%.2f ms)  %s\n", name, flop_per_launch / (ms * 1e-3) / 1e12, ms, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    double *sink; cudaMalloc(&sink, 8);
    const int iters = 4000;
    for (int wps = 1; wps <= 4; wps *= 2) {
        int block = wps * 128, grid = 148 * (wps == 4 ? 2 : 1);
        double warps = (double)grid * block / 32;
        printf("-- %d warps/SMSP, grid %d\n", wps, grid);
        timeit("m8n8k4   x1 chain ", [&] { k884<1><<<grid, block>>>(sink, iters); }, warps * iters * 4 * 1 * 2.0 * 8 * 8 * 4);
        timeit("m8n8k4   x2 chains", [&] { k884<2><<<grid, block>>>(sink, iters); }, warps * iters * 4 * 2 * 2.0 * 8 * 8 * 4);
        timeit("m8n8k4   x3 chains", [&] { k884<3><<<grid, block>>>(sink, iters); }, warps * iters * 4 * 3 * 2.0 * 8 * 8 * 4);
        timeit("m8n8k4   x4 chains", [&] { k884<4><<<grid, block>>>(sink, iters); }, warps * iters * 4 * 4 * 2.0 * 8 * 8 * 4);
        timeit("m8n8k4   x8 chains", [&] { k884<8><<<grid, block>>>(sink, iters); }, warps * iters * 4 * 8 * 2.0 * 8 * 8 * 4);
        timeit("m16n8k4  x4 chains", [&] { k1684<4><<<grid, block>>>(sink, iters); }, warps * iters * 4 * 4 * 2.0 * 16 * 8 * 4);
        timeit("m16n8k16 x2 chains", [&] { k16816<2><<<grid, block>>>(sink, iters); }, warps * iters * 4 * 2 * 2.0 * 16 * 8 * 16);
        timeit("m16n8k16 x4 chains", [&] { k16816<4><<<grid, block>>>(sink, iters); }, warps * iters * 4 * 4 * 2.0 * 16 * 8 * 16);
    }
    return 0;
}
