// tools/dfma_lat.cu -- DFMA issue/latency microbenchmark for sm_100a (B200).
// One CTA per SM; W warps per SMSP, each thread runs C independent DFMA chains.
// Prints cycles per DFMA warp-instruction per SMSP: ~2.0 means the pipe is full.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_lat tools/dfma_lat.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int C>
__global__ void k(double *sink, int iters, long long *cycles)
{
    double a[C];
#pragma unroll
    for (int i = 0; i < C; ++i) a[i] = 1.0 + threadIdx.x + i;
    const double m = 0.999999, c = 1e-9;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < C; ++i) a[i] = fma(a[i], m, c);
    }
    long long t1 = clock64();
    double r = 0;
#pragma unroll
    for (int i = 0; i < C; ++i) r += a[i];
    if (r == 123.456) sink[0] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int C>
void run(int warps_per_smsp)
{
    double *sink; long long *cyc, h;
    cudaMalloc(&sink, 8); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    k<C><<<148, warps_per_smsp * 128>>>(sink, iters, cyc);
    k<C><<<148, warps_per_smsp * 128>>>(sink, iters, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per = (double)h / ((double)iters * 8 * C * warps_per_smsp);
    printf("chains=%2d warps/SMSP=%d : %.2f cycles per DFMA warp-instr per SMSP (per-warp chain step %.1f cycles)\n", C,
           warps_per_smsp, per, (double)h / ((double)iters * 8));
    cudaFree(sink); cudaFree(cyc);
}

int main()
{
    for (int w = 1; w <= 4; w *= 2) {
        run<1>(w); run<2>(w); run<3>(w); run<4>(w); run<6>(w); run<8>(w); run<12>(w); run<16>(w);
    }
    return 0;
}
