// synth.cu -- counter-based synthetic point generator (measurement support).
//
// Coordinate d of GLOBAL point i is a pure function of (seed, i*rank + d), so any GPU or CPU slice
// of a batch regenerates identical values: a batch sharded over 2, 4 or 8 GPUs is the same batch,
// the CPU baseline and the parity checks rebuild the points they need instead of copying them, and
// a checksum over all results does not depend on how the batch was split.
// chebpol_b200/synth.py holds the numpy twin (bit-identical).
#include "common.cuh"

__host__ __device__ static inline uint64_t splitmix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256) synth_points_kernel(uint64_t seed, uint64_t first_elem, int64_t nelem, double lo,
                                                           double span, double *__restrict__ x)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nelem; e += stride) {
        const uint64_t z = splitmix64(seed + (first_elem + (uint64_t)e) * 0x9E3779B97F4A7C15ull);
        const double u = (double)(z >> 11) * 0x1.0p-53;  // [0, 1), exact
        x[e] = __dadd_rn(lo, __dmul_rn(span, u));        // two roundings, as numpy does it
    }
}

cudaError_t synth_fill_points(uint64_t seed, int64_t first_point, int64_t npts, int rank, double lo, double hi, double *d_x,
                              int num_sms, cudaStream_t s)
{
    const int64_t nelem = npts * rank;
    if (nelem <= 0) return cudaSuccess;
    int64_t nb = (nelem + 255) / 256;
    if (nb > (int64_t)num_sms * 16) nb = (int64_t)num_sms * 16;
    synth_points_kernel<<<(int)nb, 256, 0, s>>>(seed, (uint64_t)first_point * (uint64_t)rank, nelem, lo, hi - lo, d_x);
    return cudaGetLastError();
}
