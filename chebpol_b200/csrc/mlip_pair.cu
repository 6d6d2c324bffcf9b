// mlip_pair.cu -- fast multilinear (with blending) evaluation for sm_100a.
//
// Replaces C_evalmlip / binsearch / blendfun (src/chebpol.c:617-667, :543-556,
// src/chebpol.h:24-45) for rank <= 6.  This is a gather kernel bounded by the
// memory system, not by arithmetic: per point it reads `rank` coordinates,
// 2^rank table values and writes one double.
//
// Mapping: two adjacent lanes share one point.  The 2^rank corners of the
// cell come in 2^(rank-1) pairs that are adjacent along dim 0 (the
// reference's `vpos`, `vpos-1`, :655-660); lane h of the pair loads the
// dim-0 = h element of every pair, so each warp-level gather instruction
// touches one 16-byte span (usually one 32-byte sector) per point instead of
// two separate requests, halving the L1/L2 tag traffic that bounds a gather.
// The pair's partial sums are combined with one shuffle.  Knot vectors are
// staged in shared memory; the cell search is the reference's bisection
// (made NaN-safe); blend weights are computed once per dimension, not once
// per corner (same values, same product order g = 0..rank-1 as :654-662).
//
// Rounding: corner terms are accumulated in two interleaved partial sums
// instead of one (few-ulp differences); everything else is the reference's
// arithmetic.
#include "mlip_common.cuh"

namespace {

using mlip::blend_w;

constexpr int kBlock = 256;

template <int RANK, bool LINEAR>
__global__ void __launch_bounds__(kBlock) mlip_pair_kernel(const __grid_constant__ MlipArgs a, int grid_elems)
{
    extern __shared__ double sgrid[];  // all knot vectors, concatenated
    for (int i = threadIdx.x; i < grid_elems; i += kBlock) sgrid[i] = a.grid[i];
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int h = lane & 1;
    const unsigned full = 0xffffffffu;
    const int64_t pairs_per_grid = (int64_t)gridDim.x * (kBlock / 2);

    for (int64_t p0 = (int64_t)blockIdx.x * (kBlock / 2); p0 < a.npts; p0 += pairs_per_grid) {
        const int64_t p = p0 + (threadIdx.x >> 1);
        const bool ok = p < a.npts;
        // --- coordinates: lane h fetches the dims with (d & 1) == h
        double xd[RANK];
#pragma unroll
        for (int d = 0; d < RANK; ++d) xd[d] = 0.0;
        if (ok) {
            const double *xp = a.x + p * RANK;
            if constexpr ((RANK & 1) == 0) {
                // 16-byte aligned pairs of coordinates
#pragma unroll
                for (int d = 0; d < RANK; d += 2) {
                    if (((d >> 1) & 1) == h || RANK == 2) {
                        const double2 v = ldg_stream_f64x2(xp + d);
                        xd[d] = v.x;
                        xd[d + 1] = v.y;
                    }
                }
            } else {
#pragma unroll
                for (int d = 0; d < RANK; ++d)
                    if ((d & 1) == h || RANK == 1) xd[d] = ldg_stream_f64(xp + d);
            }
        }
        // --- cell search + weight for the dims this lane owns
        int gp[RANK];
        double w[RANK];
        bool isnan_ = false;
#pragma unroll
        for (int d = 0; d < RANK; ++d) {
            bool mine;
            if constexpr ((RANK & 1) == 0) mine = (((d >> 1) & 1) == h) || RANK == 2;
            else mine = ((d & 1) == h) || RANK == 1;
            gp[d] = 1;
            w[d] = 0.0;
            if (mine) {
                const double *g = sgrid + a.goff[d];
                const double v = xd[d];
                isnan_ |= (v != v);
                int lo = 0, hi = a.dims[d] - 1;
                while (lo + 1 < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (g[mid] >= v) hi = mid;
                    else lo = mid;
                }
                gp[d] = hi;
                const double g0 = g[hi - 1], g1 = g[hi];
                w[d] = (v - g0) / (g1 - g0);
            }
        }
        // --- exchange with the partner lane so both lanes know every dim
        if constexpr (RANK > 1 && RANK != 2) {
#pragma unroll
            for (int d = 0; d < RANK; ++d) {
                int owner;
                if constexpr ((RANK & 1) == 0) owner = (d >> 1) & 1;
                else owner = d & 1;
                const int src = (lane & ~1) | owner;
                gp[d] = __shfl_sync(full, gp[d], src);
                w[d] = __shfl_sync(full, w[d], src);
            }
        }
        isnan_ = __shfl_xor_sync(full, (int)isnan_, 1) | (int)isnan_;

        // --- blend weights, once per dimension
        double whi[RANK], wlo[RANK];
#pragma unroll
        for (int d = 0; d < RANK; ++d) {
            if constexpr (LINEAR) {
                whi[d] = w[d];
                wlo[d] = 1 - w[d];
            } else {
                whi[d] = blend_w(w[d], a.blend);
                wlo[d] = blend_w(1 - w[d], a.blend);
            }
        }
        // --- this lane's 2^(RANK-1) corners: dim 0 is fixed to h
        unsigned top = 0;
#pragma unroll
        for (int d = 0; d < RANK; ++d) top += (unsigned)a.stride[d] * (unsigned)gp[d];
        const unsigned base = top - (h ? 0u : 1u);
        const double w0 = h ? whi[0] : wlo[0];
        constexpr int NC = 1 << (RANK - 1);
        double vals[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            unsigned pos = base;
#pragma unroll
            for (int d = 1; d < RANK; ++d)
                if (!((c >> (d - 1)) & 1)) pos -= (unsigned)a.stride[d];
            vals[c] = ok ? __ldg(a.values + pos) : 0.0;
        }
        double wsum = 0.0, acc = 0.0;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            double cw = w0;  // == 1*w0, the reference's first factor
#pragma unroll
            for (int d = 1; d < RANK; ++d) cw *= ((c >> (d - 1)) & 1) ? whi[d] : wlo[d];
            wsum += cw;
            acc = fma(cw, vals[c], acc);
        }
        wsum += __shfl_xor_sync(full, wsum, 1);
        acc += __shfl_xor_sync(full, acc, 1);
        if (ok && h == 0) {
            double r = acc / wsum;
            if (isnan_) r = __longlong_as_double(0x7ff8000000000000LL);
            stg_stream_f64(a.out + p, r);
        }
    }
}

template <int RANK>
cudaError_t launch_rank(const MlipArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    int grid_elems = 0;
    for (int d = 0; d < RANK; ++d) grid_elems += a.dims[d];
    const int smem = grid_elems * 8;
    if (smem > 160 * 1024) return cudaErrorNotSupported;
    const int64_t need = (a.npts + kBlock / 2 - 1) / (kBlock / 2);
    int grid = num_sms * 8;
    if (need < grid) grid = (int)(need < 1 ? 1 : need);
    cudaError_t e;
    if (a.blend == 0) {
        auto k = mlip_pair_kernel<RANK, true>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        k<<<grid, kBlock, smem, s>>>(a, grid_elems);
    } else {
        auto k = mlip_pair_kernel<RANK, false>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        k<<<grid, kBlock, smem, s>>>(a, grid_elems);
    }
    if (li) { li->grid = grid; li->block = kBlock; li->smem = smem; li->kernel = K_MLIP_PAIR; }
    return cudaGetLastError();
}

}  // namespace

bool mlip_pair_supported(int rank, int blend)
{
    return rank >= 1 && rank <= 6 && blend >= 0 && blend <= 5;
}

cudaError_t launch_mlip_pair(const MlipArgs &a, int variant, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    (void)variant;
    switch (a.rank) {
    case 1: return launch_rank<1>(a, num_sms, s, li);
    case 2: return launch_rank<2>(a, num_sms, s, li);
    case 3: return launch_rank<3>(a, num_sms, s, li);
    case 4: return launch_rank<4>(a, num_sms, s, li);
    case 5: return launch_rank<5>(a, num_sms, s, li);
    case 6: return launch_rank<6>(a, num_sms, s, li);
    default: return cudaErrorNotSupported;
    }
}
