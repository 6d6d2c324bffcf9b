// tools/dmma_mix.cu -- what does FP64 scalar work cost when it shares the SM with a DMMA stream?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_mix tools/dmma_mix.cu
// Per loop iteration a warp issues 40 DMMAs (4 chains) and NF DFMAs, the DFMAs either spread
// one by one between the DMMAs (MODE 0) or issued back to back after them (MODE 1).  Reported:
// SM cycles per iteration per scheduler (ideal = warps_per_scheduler * 40 * 16).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NF, int MODE>
__global__ void __launch_bounds__(512, 1) kmix(double *sink, int iters, long long *cyc)
{
    double d0[4], d1[4], f[8];
    for (int i = 0; i < 4; ++i) { d0[i] = threadIdx.x; d1[i] = i; }
    for (int i = 0; i < 8; ++i) f[i] = threadIdx.x * 0.25 + i;
    double a = 1.0 + threadIdx.x * 1e-3, b = 0.5, m = 1.0000001;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 10; ++u) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                mma884(d0[i], d1[i], a, b);
                if (MODE == 0) {
                    const int idx = u * 4 + i;  // 0..39
                    if (idx < NF) f[idx & 7] = fma(f[idx & 7], m, a);
                }
            }
        }
        if (MODE == 1) {
#pragma unroll
            for (int q = 0; q < NF; ++q) f[q & 7] = fma(f[q & 7], m, a);
        }
    }
    long long t1 = clock64();
    double r = 0;
    for (int i = 0; i < 4; ++i) r += d0[i] + d1[i];
    for (int i = 0; i < 8; ++i) r += f[i];
    if (r == 123.456) sink[0] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int NF, int MODE>
void run(int warps_per_sched, double *sink, long long *cyc)
{
    const int iters = 2000;
    kmix<NF, MODE><<<148, warps_per_sched * 128>>>(sink, iters, cyc);
    cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    const double per_it = (double)h / iters;
    printf("warps/sched=%d NF=%2d %s  cycles/iter=%8.1f  ideal=%5d  util=%.3f  %s\n", warps_per_sched, NF,
           MODE ? "batched" : "spread ", per_it, warps_per_sched * 640, warps_per_sched * 640 / per_it,
           cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    double *sink; long long *cyc;
    cudaMalloc(&sink, 8); cudaMalloc(&cyc, 8);
    for (int w = 1; w <= 4; w *= 2) {
        run<0, 0>(w, sink, cyc);
        run<8, 0>(w, sink, cyc);
        run<8, 1>(w, sink, cyc);
        run<16, 0>(w, sink, cyc);
        run<16, 1>(w, sink, cyc);
        run<40, 0>(w, sink, cyc);
        run<40, 1>(w, sink, cyc);
    }
    return 0;
}
