// tools/mma_lab.cu -- the step loop of the DMMA contraction engine (csrc/contract_mma.cuh) in isolation:
// B fragments from a shared-memory slab, A fragments in registers, two 8-row N-tiles per step, the
// per-step fold (acc2 += bA * D) -- without the TMA ring, the per-pass prologue and the barriers.
// Separates what the loop itself costs on the FP64 pipe from what the pass structure costs
// (the FH 40^3 shape: KC = 10).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mma_lab tools/mma_lab.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// EPI: 0 none (D summed into acc at the end of a step with one add), 1 the engine's fold (bA from shared
// memory, 4 DFMA per step), 2 fold software-pipelined one step behind (as the engine does)
// ILV: 0 source order kc-major (two chains alternate), 1 chain-major (all of N-tile 0, then N-tile 1)
// SYNC: 0 none, 1 __syncwarp + lane-0 atomic + spin on a shared flag every SLAB_STEPS steps (the ring hand-shake)
template <int KC, int NW, int EPI, int ILV, int SYNC, int KEEP = 0, int BREG = 0>
__global__ void __launch_bounds__(NW * 32, 1) lab(double *sink, int passes, int steps_per_pass, int rs, int slab_steps, int nA)
{
    extern __shared__ __align__(128) unsigned char smem[];
    double *slab = reinterpret_cast<double *>(smem);            // [slab_steps*16][rs]
    double *Tb = slab + (size_t)slab_steps * 16 * rs;           // [nA][NW*8]
    int *cnt = reinterpret_cast<int *>(Tb + (size_t)nA * NW * 8);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    for (int i = tid; i < slab_steps * 16 * rs; i += NW * 32) slab[i] = 1.0 / (1 + (i % 97));
    for (int i = tid; i < nA * NW * 8; i += NW * 32) Tb[i] = 1.0 / (3 + (i % 31));
    if (tid == 0) cnt[0] = 0;
    __syncthreads();
    double A[KC];
#pragma unroll
    for (int kc = 0; kc < KC; ++kc) A[kc] = 1.0 + 1e-3 * (kc + t) + 1e-4 * g;
    const double *tcol = Tb + warp * 8 + g;
    double total = 0.0;
    for (int pass = 0; pass < passes; ++pass) {
        double acc2[2] = {0.0, 0.0};
        double PD[2][2] = {{0, 0}, {0, 0}};
        int piA = 0;
        int sstep = 0;
        for (int s = 0; s < steps_per_pass; ++s) {
            const int r16 = sstep * 16;
            const double *bp = slab + (size_t)(r16 + g) * rs + t;
            double D[2][2] = {{0, 0}, {0, 0}};
            if (KEEP) { D[0][0] = PD[0][0]; D[0][1] = PD[0][1]; D[1][0] = PD[1][0]; D[1][1] = PD[1][1]; }
            if (BREG) {
                const double b0 = bp[0], b1 = bp[(size_t)8 * rs];
#pragma unroll
                for (int kc = 0; kc < KC; ++kc) {
                    dmma884(D[0][0], D[0][1], A[kc], b0);
                    dmma884(D[1][0], D[1][1], A[kc], b1);
                }
            } else if (ILV == 0) {
#pragma unroll
                for (int kc = 0; kc < KC; ++kc) {
                    const double b0 = bp[4 * kc], b1 = bp[(size_t)8 * rs + 4 * kc];
                    dmma884(D[0][0], D[0][1], A[kc], b0);
                    dmma884(D[1][0], D[1][1], A[kc], b1);
                }
            } else {
#pragma unroll
                for (int kc = 0; kc < KC; ++kc) dmma884(D[0][0], D[0][1], A[kc], bp[4 * kc]);
#pragma unroll
                for (int kc = 0; kc < KC; ++kc) dmma884(D[1][0], D[1][1], A[kc], bp[(size_t)8 * rs + 4 * kc]);
            }
            const int iA = (2 * s) % nA;
            if (KEEP) {
                PD[0][0] = D[0][0]; PD[0][1] = D[0][1]; PD[1][0] = D[1][0]; PD[1][1] = D[1][1];
            } else if (EPI == 0) {
                acc2[0] += D[0][0] + D[1][0];
                acc2[1] += D[0][1] + D[1][1];
            } else if (EPI == 1) {
                const double bA0 = tcol[(size_t)iA * NW * 8], bA1 = tcol[(size_t)(iA + 1) * NW * 8];
                acc2[0] = fma(bA0, D[0][0], acc2[0]);
                acc2[1] = fma(bA0, D[0][1], acc2[1]);
                acc2[0] = fma(bA1, D[1][0], acc2[0]);
                acc2[1] = fma(bA1, D[1][1], acc2[1]);
            } else {
                const double bA0 = tcol[(size_t)piA * NW * 8], bA1 = tcol[(size_t)(piA + 1) * NW * 8];
                acc2[0] = fma(bA0, PD[0][0], acc2[0]);
                acc2[1] = fma(bA0, PD[0][1], acc2[1]);
                acc2[0] = fma(bA1, PD[1][0], acc2[0]);
                acc2[1] = fma(bA1, PD[1][1], acc2[1]);
                PD[0][0] = D[0][0]; PD[0][1] = D[0][1]; PD[1][0] = D[1][0]; PD[1][1] = D[1][1];
                piA = iA;
            }
            if (++sstep == slab_steps) {
                sstep = 0;
                if (SYNC == 1) {
                    __syncwarp();
                    if (lane == 0) {
                        __threadfence_block();
                        atomicAdd(&cnt[0], 1);
                    }
                }
            }
        }
        total += acc2[0] + acc2[1] + PD[0][0] + PD[1][1];
    }
    if (total == 1.2345e-300) sink[0] = total;
}

template <int KC, int NW, int EPI, int ILV, int SYNC, int KEEP = 0, int BREG = 0>
void run(const char *name, double *sink, int slab_steps, int rs, double peak)
{
    const int passes = 40, steps = 100, nA = 40;
    const size_t smem = (size_t)slab_steps * 16 * rs * 8 + (size_t)nA * NW * 8 * 8 + 64;
    auto k = lab<KC, NW, EPI, ILV, SYNC, KEEP, BREG>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<<<148, NW * 32, smem>>>(sink, 4, steps, rs, slab_steps, nA);
    cudaEventRecord(e0);
    k<<<148, NW * 32, smem>>>(sink, passes, steps, rs, slab_steps, nA);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flop = 148.0 * NW * passes * steps * 2 * KC * 2.0 * 8 * 8 * 4;
    const double tf = flop / (ms * 1e-3) / 1e12;
    printf("{\"variant\": \"%s\", \"KC\": %d, \"warps\": %d, \"epilogue\": %d, \"chain_major\": %d, \"slab_sync\": %d, \"slab_steps\": %d, \"ms\": %.4f, "
           "\"dmma_TFLOPs\": %.2f, \"frac_of_peak\": %.4f, \"err\": \"%s\"}\n",
           name, KC, NW, EPI, ILV, SYNC, slab_steps, ms, tf, tf / peak, cudaGetErrorString(cudaGetLastError()));
    fflush(stdout);
}


// ---- second family: NTB N-tiles per step, fold one step behind, options:
//   B128: B fragments of two consecutive k-chunks fetched by one LDS.128 (rows stored pair-interleaved)
//   A128: the NTB(=2) level-A basis values of a step fetched by one LDS.128 (level-A table stored [point][iA])
template <int KC, int NW, int NTB, int B128, int A128>
__global__ void __launch_bounds__(NW * 32, 1) lab2(double *sink, int passes, int steps_per_pass, int rs, int slab_steps, int nA)
{
    extern __shared__ __align__(128) unsigned char smem[];
    double *slab = reinterpret_cast<double *>(smem);                  // [slab_steps*8*NTB][rs]
    double *Tb = slab + (size_t)slab_steps * 8 * NTB * rs;            // [nA][NW*8] or [NW*8][nA+2]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    for (int i = tid; i < slab_steps * 8 * NTB * rs; i += NW * 32) slab[i] = 1.0 / (1 + (i % 97));
    for (int i = tid; i < (nA + 2) * NW * 8; i += NW * 32) Tb[i] = 1.0 / (3 + (i % 31));
    __syncthreads();
    double A[KC];
#pragma unroll
    for (int kc = 0; kc < KC; ++kc) A[kc] = 1.0 + 1e-3 * (kc + t) + 1e-4 * g;
    const int pt = warp * 8 + g;
    const double *tcol = A128 ? Tb + (size_t)pt * (nA + 2) : Tb + pt;
    const size_t tstride = A128 ? 1 : (size_t)NW * 8;
    double total = 0.0;
    for (int pass = 0; pass < passes; ++pass) {
        double acc2[2] = {0.0, 0.0};
        double PD[NTB][2];
#pragma unroll
        for (int j = 0; j < NTB; ++j) PD[j][0] = PD[j][1] = 0.0;
        int piA = 0, sstep = 0, iA = 0;
        for (int s = 0; s < steps_per_pass; ++s) {
            const double *bp = slab + (size_t)(sstep * 8 * NTB + g) * rs + (B128 ? 2 * t : t);
            double D[NTB][2];
#pragma unroll
            for (int j = 0; j < NTB; ++j) D[j][0] = D[j][1] = 0.0;
            if (B128) {
#pragma unroll
                for (int kc = 0; kc + 1 < KC; kc += 2) {
                    double2 b[NTB];
#pragma unroll
                    for (int j = 0; j < NTB; ++j) b[j] = *reinterpret_cast<const double2 *>(bp + (size_t)j * 8 * rs + 4 * kc);
#pragma unroll
                    for (int j = 0; j < NTB; ++j) dmma884(D[j][0], D[j][1], A[kc], b[j].x);
#pragma unroll
                    for (int j = 0; j < NTB; ++j) dmma884(D[j][0], D[j][1], A[kc + 1], b[j].y);
                }
                if (KC & 1) {
#pragma unroll
                    for (int j = 0; j < NTB; ++j) dmma884(D[j][0], D[j][1], A[KC - 1], bp[(size_t)j * 8 * rs + 4 * (KC - 1)]);
                }
            } else {
#pragma unroll
                for (int kc = 0; kc < KC; ++kc) {
                    double b[NTB];
#pragma unroll
                    for (int j = 0; j < NTB; ++j) b[j] = bp[(size_t)j * 8 * rs + 4 * kc];
#pragma unroll
                    for (int j = 0; j < NTB; ++j) dmma884(D[j][0], D[j][1], A[kc], b[j]);
                }
            }
            // fold the previous step
            double bA[NTB];
            if (A128 && NTB == 2) {
                const double2 v = *reinterpret_cast<const double2 *>(tcol + piA);
                bA[0] = v.x;
                bA[NTB - 1] = v.y;
            } else {
#pragma unroll
                for (int j = 0; j < NTB; ++j) bA[j] = tcol[(size_t)(piA + j) * tstride];
            }
#pragma unroll
            for (int j = 0; j < NTB; ++j) {
                acc2[0] = fma(bA[j], PD[j][0], acc2[0]);
                acc2[1] = fma(bA[j], PD[j][1], acc2[1]);
                PD[j][0] = D[j][0];
                PD[j][1] = D[j][1];
            }
            piA = iA;
            iA += NTB;
            if (iA + NTB > nA) iA = 0;
            if (++sstep == slab_steps) sstep = 0;
        }
        total += acc2[0] + acc2[1] + PD[0][0] + PD[NTB - 1][1];
    }
    if (total == 1.2345e-300) sink[0] = total;
}

template <int KC, int NW, int NTB, int B128, int A128>
void run2(const char *name, double *sink, int slab_steps, int rs, double peak)
{
    const int passes = 40, nA = 40;
    const int steps = 200 / NTB;
    const size_t smem = (size_t)slab_steps * 8 * NTB * rs * 8 + (size_t)(nA + 2) * NW * 8 * 8 + 64;
    auto k = lab2<KC, NW, NTB, B128, A128>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<<<148, NW * 32, smem>>>(sink, 4, steps, rs, slab_steps, nA);
    cudaEventRecord(e0);
    k<<<148, NW * 32, smem>>>(sink, passes, steps, rs, slab_steps, nA);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flop = 148.0 * NW * passes * steps * NTB * KC * 2.0 * 8 * 8 * 4;
    const double tf = flop / (ms * 1e-3) / 1e12;
    printf("{\"variant\": \"%s\", \"KC\": %d, \"warps\": %d, \"ntb\": %d, \"b128\": %d, \"a128\": %d, \"rs\": %d, \"ms\": %.4f, "
           "\"dmma_TFLOPs\": %.2f, \"frac_of_peak\": %.4f, \"err\": \"%s\"}\n",
           name, KC, NW, NTB, B128, A128, rs, ms, tf, tf / peak, cudaGetErrorString(cudaGetLastError()));
    fflush(stdout);
}

int main(int argc, char **argv)
{
    const double peak = argc > 1 ? atof(argv[1]) : 37.1;
    double *sink;
    cudaMalloc(&sink, 8);
    run<10, 16, 0, 0, 0>("skeleton, no epilogue", sink, 5, 44, peak);
    run<10, 16, 0, 1, 0>("skeleton, chain-major source", sink, 5, 44, peak);
    run<10, 16, 1, 0, 0>("fold in step", sink, 5, 44, peak);
    run<10, 16, 2, 0, 0>("fold one step behind", sink, 5, 44, peak);
    run<10, 16, 2, 0, 1>("fold behind + slab hand-shake", sink, 5, 44, peak);
    run<10, 8, 2, 0, 0>("fold behind, 8 warps", sink, 5, 44, peak);
    run<10, 12, 2, 0, 0>("fold behind, 12 warps", sink, 5, 44, peak);
    run<25, 16, 0, 0, 0>("KC 25 skeleton", sink, 4, 108, peak);
    run<25, 16, 2, 0, 0>("KC 25 fold behind", sink, 4, 108, peak);
    run<10, 16, 0, 0, 0, 1, 0>("skeleton, accumulators carried across steps (no chain restart)", sink, 5, 44, peak);
    run<10, 16, 0, 0, 0, 0, 1>("skeleton, B from registers (one LDS pair per step)", sink, 5, 44, peak);
    run<10, 16, 0, 0, 0, 1, 1>("carried accumulators + B from registers", sink, 5, 44, peak);
    run<5, 16, 0, 0, 0>("KC 5 skeleton", sink, 5, 28, peak);
    run<20, 16, 0, 0, 0>("KC 20 skeleton", sink, 4, 84, peak);
    run<40, 16, 0, 0, 0>("KC 40 skeleton", sink, 2, 164, peak);
    run<10, 16, 0, 0, 0>("skeleton, row stride 41 (bank conflicts)", sink, 5, 41, peak);
    run<10, 16, 0, 0, 0>("skeleton, row stride 52", sink, 5, 52, peak);
    run2<10, 16, 2, 0, 0>("lab2 baseline: 2 N-tiles, LDS.64", sink, 5, 44, peak);
    run2<10, 16, 2, 1, 0>("B by LDS.128, rs 40", sink, 5, 40, peak);
    run2<10, 16, 2, 1, 0>("B by LDS.128, rs 44", sink, 5, 44, peak);
    run2<10, 16, 2, 0, 1>("bA by LDS.128", sink, 5, 44, peak);
    run2<10, 16, 2, 1, 1>("B and bA by LDS.128, rs 40", sink, 5, 40, peak);
    run2<10, 16, 1, 0, 0>("1 N-tile per step", sink, 10, 44, peak);
    run2<10, 16, 3, 0, 0>("3 N-tiles per step", sink, 3, 44, peak);
    run2<10, 16, 4, 0, 0>("4 N-tiles per step", sink, 2, 44, peak);
    run2<10, 16, 4, 1, 0>("4 N-tiles per step, B by LDS.128 rs 40", sink, 2, 40, peak);
    run2<12, 16, 2, 0, 0>("KC 12 baseline", sink, 5, 52, peak);
    run2<12, 16, 2, 1, 0>("KC 12 B by LDS.128 rs 56", sink, 5, 56, peak);
    run2<16, 16, 2, 0, 0>("KC 16 baseline", sink, 4, 68, peak);
    run2<16, 16, 2, 1, 0>("KC 16 B by LDS.128 rs 72", sink, 4, 72, peak);
    run2<20, 16, 2, 0, 0>("KC 20 baseline", sink, 4, 84, peak);
    run2<20, 16, 2, 1, 0>("KC 20 B by LDS.128 rs 88", sink, 4, 88, peak);
    run2<25, 16, 2, 0, 0>("KC 25 baseline", sink, 4, 100, peak);
    run2<25, 16, 2, 1, 0>("KC 25 B by LDS.128 rs 104", sink, 4, 104, peak);
    return 0;
}
