// mlip_bin.cu -- multilinear evaluation with the value table cache-blocked through shared memory.
//
// C_evalmlip (src/chebpol.c:617-667) on a table that does not fit L2 is a random gather: every
// point pulls 2^(D-1) separate 64-byte DRAM bursts to use 16 bytes of each (mlip_pair: 426 B/point
// of DRAM traffic against 168 algorithmic on the 64^4 table; the row-pair "quad" layout 2x the
// memory for 1.3x the speed; profiles/mlip_variants_r02.jsonl).  The cell-major table fixes the
// gather with 2^D times the memory.  This path fixes it with NO extra table memory by changing the
// order of the POINTS instead of the layout of the table -- a batch of 1e8 points meets every cell
// of a 64^4 table six times, so points that share a neighbourhood can share one copy of it:
//
//   1. scatter   every point is appended to the bin of its cell in the UPPER dimensions (dims >= m);
//                a bin's sub-table -- all of dims < m, two slices of every upper dimension -- is
//                2^(D-m) contiguous runs of the R array (131 KB for 64^4 with m = 2: 3969 bins);
//   2. evaluate  one CTA per bin: the runs travel to shared memory by TMA bulk copy, the bin's
//                points stream through, all 2^D corner reads are shared-memory loads;
//                every result goes straight to out[index of the point] (the index travels with the
//                point), so nothing has to be un-permuted afterwards.
//
// DRAM traffic per point for D = 4: 32 + 32 + 4 (scatter: read the point, write it and its index),
// 32 + 4 (evaluate: read them back) plus the scattered 8-byte result, and the table once per chunk
// of points (15 B/point at 2^25 points) -- streaming except for the result.  Bins have a fixed
// capacity (1.25 x the mean + 64); points that do not fit (a clustered batch) are listed by index
// and evaluated by a plain gather kernel -- so no counting pass and no host synchronisation.
// Workspace is proportional to the batch chunk (<= 2^25 points: 1.7 GB), not to the table.
//
// Arithmetic: weights, blend functions and the final division are the reference's; the corner terms
// are accumulated in the reference's corner order (bit d of the corner index = upper knot of dim d).
#include "mlip_common.cuh"

namespace {

using mlip::blend_w;
using mlip::cell_search;

constexpr int kTileBudget = 160 * 1024;  // bytes of value table per bin in shared memory
constexpr int kEvalThreads = 1024;
constexpr int kScatterThreads = 256;
constexpr int kCursorStride = 64;  // ints between two bins' cursors: one per 256 bytes, so the atomics spread over every L2 slice

struct BinGeom {
    int rank, mf;            // dims < mf are whole inside a bin
    int nbins, cap;          // bins, points per bin
    int nslots;              // 2^(rank - mf) runs per bin
    int low_elems;           // prod_{d < mf} dims[d]
    int tma;                 // runs are 16-byte multiples: TMA bulk copies; else a cooperative copy
    int bstride[CP_MAXR];    // bin id = sum_{d >= mf} cell_d * bstride[d]
    int ncell[CP_MAXR];      // dims[d] - 1
};

struct BinWork {
    double *bx;              // [nbins*cap][rank] binned points
    int *bidx;               // [nbins*cap] index (within the chunk) of each binned point
    int *cursor;             // [nbins] points appended to each bin (may exceed cap: the excess overflowed)
    int *counters;           // [0] overflow count, [1] work counter of the evaluate kernel
    int *oidx;               // [chunk] indices of the points that did not fit their bin
};

// one point, RANK doubles: a single 256-bit access for RANK = 4 on 32-byte aligned storage
// (LDG.E.ENL2.256 / STG.E.ENL2.256), 128-bit pairs for other even ranks, scalars otherwise
template <int RANK>
__device__ __forceinline__ void load_point(const double *p, double (&x)[RANK], bool al32)
{
    if constexpr (RANK == 4) {
        if (al32) {
            asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                         : "=d"(x[0]), "=d"(x[1]), "=d"(x[2]), "=d"(x[3])
                         : "l"(p));
            return;
        }
    }
    if constexpr ((RANK & 1) == 0) {
#pragma unroll
        for (int d = 0; d < RANK; d += 2) {
            const double2 v = ldg_stream_f64x2(p + d);
            x[d] = v.x;
            x[d + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int d = 0; d < RANK; ++d) x[d] = ldg_stream_f64(p + d);
    }
}

template <int RANK>
__device__ __forceinline__ void store_point(double *p, const double (&x)[RANK])
{
    if constexpr (RANK == 4) {
        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(x[0]), "d"(x[1]), "d"(x[2]), "d"(x[3]) : "memory");
    } else if constexpr ((RANK & 1) == 0) {
#pragma unroll
        for (int d = 0; d < RANK; d += 2) reinterpret_cast<double2 *>(p)[d / 2] = make_double2(x[d], x[d + 1]);
    } else {
#pragma unroll
        for (int d = 0; d < RANK; ++d) p[d] = x[d];
    }
}

// ------------------------------------------------------------------------------ 1. scatter
// U points per thread are in flight together: their loads, then their searches, then their slot
// reservations (one atomic each), then their stores -- the kernel is a chain of dependent memory
// round trips per point, so the only lever is how many chains overlap.
template <int RANK, int U>
__global__ void __launch_bounds__(kScatterThreads) mlip_bin_scatter(const __grid_constant__ MlipArgs a, const BinGeom gm,
                                                                    const BinWork w, int grid_elems, int lut_elems, int x32)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sgrid = reinterpret_cast<double *>(smem_raw);
    int *slut = reinterpret_cast<int *>(smem_raw + (size_t)grid_elems * 8);
    for (int i = threadIdx.x; i < grid_elems; i += blockDim.x) sgrid[i] = a.grid[i];
    for (int i = threadIdx.x; i < lut_elems; i += blockDim.x) slut[i] = a.lut[i];
    __syncthreads();
    const int64_t tile = (int64_t)kScatterThreads * U;
    for (int64_t base = (int64_t)blockIdx.x * tile; base < a.npts; base += (int64_t)gridDim.x * tile) {
        double x[U][RANK];
        int bin[U], slot[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + u * kScatterThreads + threadIdx.x;
            if (i < a.npts) load_point<RANK>(a.x + i * RANK, x[u], x32 != 0);
            else {
#pragma unroll
                for (int d = 0; d < RANK; ++d) x[u][d] = 0.0;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            int b = 0;
#pragma unroll
            for (int d = 0; d < RANK; ++d) {
                if (d < gm.mf) continue;
                double g0, g1;
                const int hi = cell_search(sgrid + a.goff[d], a.dims[d], slut + a.loff[d], a.nb[d], a.lscale[d], x[u][d], g0, g1);
                b += (hi - 1) * gm.bstride[d];
            }
            bin[u] = b;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + u * kScatterThreads + threadIdx.x;
            slot[u] = i < a.npts ? atomicAdd(w.cursor + (int64_t)bin[u] * kCursorStride, 1) : 0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + u * kScatterThreads + threadIdx.x;
            if (i >= a.npts) continue;
            if (slot[u] < gm.cap) {
                const int64_t where = (int64_t)bin[u] * gm.cap + slot[u];
                store_point<RANK>(w.bx + where * RANK, x[u]);
                w.bidx[where] = (int)i;
            } else {
                w.oidx[atomicAdd(w.counters, 1)] = (int)i;  // cannot exceed the chunk size
            }
        }
    }
}

// corner sum of one point whose 2^RANK corner values are read through `value(corner_bits)`
template <int RANK, bool LINEAR, class ValueFn>
__device__ __forceinline__ double corner_sum(const double (&wt)[RANK], int blend, bool nn, ValueFn value)
{
    double whi[RANK], wlo[RANK];
#pragma unroll
    for (int d = 0; d < RANK; ++d) {
        if constexpr (LINEAR) {
            whi[d] = wt[d];
            wlo[d] = 1 - wt[d];
        } else {
            whi[d] = blend_w(wt[d], blend);
            wlo[d] = blend_w(1 - wt[d], blend);
        }
    }
    // running products in the reference's order (g = 0 .. RANK-1), low half of the dims tabulated
    constexpr int LOWD = RANK < 3 ? RANK : 3;
    constexpr int NLOW = 1 << LOWD;
    double cw[NLOW];
    cw[0] = wlo[0];
    cw[1] = whi[0];
#pragma unroll
    for (int d = 1; d < LOWD; ++d) {
#pragma unroll
        for (int c = 0; c < (1 << d); ++c) {
            cw[c + (1 << d)] = cw[c] * whi[d];
            cw[c] = cw[c] * wlo[d];
        }
    }
    double wsum = 0.0, acc = 0.0;
#pragma unroll
    for (int h = 0; h < (1 << (RANK - LOWD)); ++h) {
        double f = 1.0;
#pragma unroll
        for (int d = LOWD; d < RANK; ++d) f *= ((h >> (d - LOWD)) & 1) ? whi[d] : wlo[d];
#pragma unroll
        for (int c = 0; c < NLOW; ++c) {
            double cc = cw[c];
            if constexpr (RANK > LOWD) cc *= f;
            wsum += cc;
            acc = fma(cc, value(h * NLOW + c), acc);
        }
    }
    double r = acc / wsum;
    if (nn) r = __longlong_as_double(0x7ff8000000000000LL);
    return r;
}

// ------------------------------------------------------------------------------ 2. evaluate
// One CTA per bin (dynamic queue).  Shared memory: knots + bucket tables, one mbarrier, the bin's
// 2^(RANK-MF) runs of the value table.
template <int RANK, int MF, bool LINEAR>
__global__ void __launch_bounds__(kEvalThreads, 1) mlip_bin_eval(const __grid_constant__ MlipArgs a, const BinGeom gm,
                                                                 const BinWork w, int grid_elems, int lut_elems)
{
    constexpr int NSLOT = 1 << (RANK - MF);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    int *sbin = reinterpret_cast<int *>(smem_raw + 8);
    double *sgrid = reinterpret_cast<double *>(smem_raw + 16);
    int *slut = reinterpret_cast<int *>(smem_raw + 16 + (size_t)grid_elems * 8);
    const int head = (16 + grid_elems * 8 + lut_elems * 4 + 127) & ~127;
    double *tile = reinterpret_cast<double *>(smem_raw + head);

    for (int i = threadIdx.x; i < grid_elems; i += blockDim.x) sgrid[i] = a.grid[i];
    for (int i = threadIdx.x; i < lut_elems; i += blockDim.x) slut[i] = a.lut[i];
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();

    const uint32_t run_bytes = (uint32_t)gm.low_elems * 8u;
    uint32_t phase = 0;
    for (;;) {
        if (threadIdx.x == 0) {
            const int b = atomicAdd(w.counters + 1, 1);
            *sbin = b;
            if (b < gm.nbins && gm.tma) {
                // the 2^(RANK-MF) runs of this bin: upper-dimension index c_d + bit_d
                int rem = b;
                int64_t base = 0;
                int64_t ustride[RANK];
#pragma unroll
                for (int d = RANK - 1; d >= MF; --d) {
                    const int c = rem / gm.bstride[d];
                    rem -= c * gm.bstride[d];
                    base += (int64_t)c * a.stride[d];
                    ustride[d] = a.stride[d];
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads of the tile are done
                mbar_expect_tx(bar, run_bytes * NSLOT);
#pragma unroll
                for (int j = 0; j < NSLOT; ++j) {
                    int64_t off = base;
#pragma unroll
                    for (int d = MF; d < RANK; ++d)
                        if ((j >> (d - MF)) & 1) off += ustride[d];
                    // bulk copies of at most 32 KB each
                    for (uint32_t done = 0; done < run_bytes; done += 32768u) {
                        const uint32_t n = run_bytes - done < 32768u ? run_bytes - done : 32768u;
                        tma_bulk_g2s(reinterpret_cast<unsigned char *>(tile + (size_t)j * gm.low_elems) + done,
                                     reinterpret_cast<const unsigned char *>(a.values + off) + done, n, bar);
                    }
                }
            }
        }
        __syncthreads();
        const int b = *sbin;
        if (b >= gm.nbins) break;
        int npt = w.cursor[(int64_t)b * kCursorStride];
        npt = npt < gm.cap ? npt : gm.cap;
        // knots of the bin's cell in the upper dimensions (the same for every point of the bin)
        double ug0[RANK], uinv[RANK];
        {
            int rem = b;
#pragma unroll
            for (int d = RANK - 1; d >= MF; --d) {
                const int c = rem / gm.bstride[d];
                rem -= c * gm.bstride[d];
                const double g0 = sgrid[a.goff[d] + c], g1 = sgrid[a.goff[d] + c + 1];
                ug0[d] = g0;
                uinv[d] = g1 - g0;
            }
        }
        const double *bx = w.bx + (int64_t)b * gm.cap * RANK;
        const int *bi = w.bidx + (int64_t)b * gm.cap;

        auto load_x = [&](int i, double (&x)[RANK], int &idx) {
            if (i < npt) {
                load_point<RANK>(bx + (int64_t)i * RANK, x, true);
                idx = __ldg(bi + i);
            } else {
#pragma unroll
                for (int d = 0; d < RANK; ++d) x[d] = 0.0;
                idx = 0;
            }
        };
        double xc[RANK], xn[RANK];
        int ic, in;
        load_x(threadIdx.x, xc, ic);
        load_x(threadIdx.x + kEvalThreads, xn, in);
        if (gm.tma) {
            mbar_wait(bar, phase);
            phase ^= 1u;
        } else {
            // odd run length (8-byte aligned only): the CTA copies the runs itself
            int rem = b;
            int64_t base = 0;
#pragma unroll
            for (int d = RANK - 1; d >= MF; --d) {
                const int c = rem / gm.bstride[d];
                rem -= c * gm.bstride[d];
                base += (int64_t)c * a.stride[d];
            }
#pragma unroll
            for (int j = 0; j < NSLOT; ++j) {
                int64_t off = base;
#pragma unroll
                for (int d = MF; d < RANK; ++d)
                    if ((j >> (d - MF)) & 1) off += a.stride[d];
                for (int e = threadIdx.x; e < gm.low_elems; e += kEvalThreads) tile[(size_t)j * gm.low_elems + e] = __ldg(a.values + off + e);
            }
            __syncthreads();
        }
        for (int i = threadIdx.x; i < npt; i += kEvalThreads) {
            double xnn[RANK];
            int inn;
            load_x(i + 2 * kEvalThreads, xnn, inn);  // two iterations ahead: the stream stays in flight
            double wt[RANK];
            bool nn = false;
            int lidx = 0;
#pragma unroll
            for (int d = 0; d < RANK; ++d) {
                const double v = xc[d];
                nn |= (v != v);
                if (d < MF) {
                    double g0, g1;
                    const int hi = cell_search(sgrid + a.goff[d], a.dims[d], slut + a.loff[d], a.nb[d], a.lscale[d], v, g0, g1);
                    wt[d] = (v - g0) / (g1 - g0);
                    lidx += (hi - 1) * (int)a.stride[d];
                } else {
                    wt[d] = (v - ug0[d]) / uinv[d];
                }
            }
            const double *tp = tile + lidx;
            const double r = corner_sum<RANK, LINEAR>(wt, a.blend, nn, [&](int corner) -> double {
                int off = 0;
#pragma unroll
                for (int d = 0; d < MF; ++d)
                    if ((corner >> d) & 1) off += (int)a.stride[d];
                return tp[(size_t)(corner >> MF) * gm.low_elems + off];
            });
            a.out[ic] = r;  // the caller's order, straight from here
#pragma unroll
            for (int d = 0; d < RANK; ++d) { xc[d] = xn[d]; xn[d] = xnn[d]; }
            ic = in;
            in = inn;
        }
        __syncthreads();  // everyone is done with the tile and with *sbin
    }
}

// ------------------------------------------------------------------- overflow: plain gather
template <int RANK, bool LINEAR>
__global__ void __launch_bounds__(256) mlip_bin_overflow(const __grid_constant__ MlipArgs a, const BinWork w, int grid_elems,
                                                         int lut_elems)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sgrid = reinterpret_cast<double *>(smem_raw);
    int *slut = reinterpret_cast<int *>(smem_raw + (size_t)grid_elems * 8);
    for (int i = threadIdx.x; i < grid_elems; i += blockDim.x) sgrid[i] = a.grid[i];
    for (int i = threadIdx.x; i < lut_elems; i += blockDim.x) slut[i] = a.lut[i];
    __syncthreads();
    const int n = w.counters[0];
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const int i = w.oidx[j];
        double wt[RANK];
        bool nn = false;
        int64_t base = 0;
#pragma unroll
        for (int d = 0; d < RANK; ++d) {
            const double v = a.x[(int64_t)i * RANK + d];
            nn |= (v != v);
            double g0, g1;
            const int hi = cell_search(sgrid + a.goff[d], a.dims[d], slut + a.loff[d], a.nb[d], a.lscale[d], v, g0, g1);
            wt[d] = (v - g0) / (g1 - g0);
            base += (int64_t)(hi - 1) * a.stride[d];
        }
        const double *vp = a.values + base;
        a.out[i] = corner_sum<RANK, LINEAR>(wt, a.blend, nn, [&](int corner) -> double {
            int64_t off = 0;
#pragma unroll
            for (int d = 0; d < RANK; ++d)
                if ((corner >> d) & 1) off += a.stride[d];
            return __ldg(vp + off);
        });
    }
}

bool bin_geometry(int rank, const int *dims, int64_t chunk, BinGeom &gm)
{
    if (rank < 2 || rank > 6) return false;
    gm.rank = rank;
    int64_t low = 1;
    int mf = 0;
    for (int m = 1; m < rank; ++m) {  // the most whole dimensions whose 2^(rank-m) runs fit
        low *= dims[m - 1];
        if (low * 8 * (int64_t(1) << (rank - m)) <= kTileBudget) mf = m;
    }
    if (mf == 0) return false;
    low = 1;
    for (int d = 0; d < mf; ++d) low *= dims[d];
    gm.tma = (low % 2) == 0;  // TMA bulk copies need 16-byte multiples
    gm.mf = mf;
    gm.low_elems = (int)low;
    gm.nslots = 1 << (rank - mf);
    int64_t nb = 1;
    for (int d = 0; d < rank; ++d) {
        gm.ncell[d] = dims[d] - 1;
        gm.bstride[d] = 0;
    }
    for (int d = mf; d < rank; ++d) {
        gm.bstride[d] = (int)nb;
        nb *= dims[d] - 1;
        if (nb > (1 << 22)) return false;
    }
    gm.nbins = (int)nb;
    int64_t cap = (chunk + nb - 1) / nb;
    cap = cap + cap / 4 + 64;
    cap = (cap + 3) & ~int64_t(3);
    if (nb * cap >= (int64_t(1) << 31)) return false;
    gm.cap = (int)cap;
    return true;
}

template <int RANK, int MF>
cudaError_t launch_eval_mf(const MlipArgs &a, const BinGeom &gm, const BinWork &w, int ge, int le, int num_sms, int smem,
                           cudaStream_t s)
{
    cudaError_t e;
    if (a.blend == 0) {
        auto k = mlip_bin_eval<RANK, MF, true>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        k<<<num_sms, kEvalThreads, smem, s>>>(a, gm, w, ge, le);
    } else {
        auto k = mlip_bin_eval<RANK, MF, false>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        k<<<num_sms, kEvalThreads, smem, s>>>(a, gm, w, ge, le);
    }
    return cudaGetLastError();
}

template <int RANK>
cudaError_t launch_eval(const MlipArgs &a, const BinGeom &gm, const BinWork &w, int ge, int le, int num_sms, int smem,
                        cudaStream_t s)
{
    switch (gm.mf) {
    case 1: if constexpr (RANK > 1) return launch_eval_mf<RANK, 1>(a, gm, w, ge, le, num_sms, smem, s); break;
    case 2: if constexpr (RANK > 2) return launch_eval_mf<RANK, 2>(a, gm, w, ge, le, num_sms, smem, s); break;
    case 3: if constexpr (RANK > 3) return launch_eval_mf<RANK, 3>(a, gm, w, ge, le, num_sms, smem, s); break;
    case 4: if constexpr (RANK > 4) return launch_eval_mf<RANK, 4>(a, gm, w, ge, le, num_sms, smem, s); break;
    case 5: if constexpr (RANK > 5) return launch_eval_mf<RANK, 5>(a, gm, w, ge, le, num_sms, smem, s); break;
    }
    return cudaErrorNotSupported;
}

template <int RANK>
cudaError_t run_chunk(const MlipArgs &a, const BinGeom &gm, const BinWork &w, int num_sms, cudaStream_t s, int *launches,
                      LaunchInfo *li)
{
    int ge = 0, le = 0;
    for (int d = 0; d < RANK; ++d) { ge += a.dims[d]; le += a.nb[d]; }
    const int head = ge * 8 + le * 4;
    cudaError_t e = cudaMemsetAsync(w.cursor, 0, sizeof(int) * (size_t)gm.nbins * kCursorStride, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(w.counters, 0, sizeof(int) * 2, s);
    if (e != cudaSuccess) return e;
    {
        constexpr int U = 4;
        auto k = mlip_bin_scatter<RANK, U>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, head);
        if (e != cudaSuccess) return e;
        int64_t nb = (a.npts + kScatterThreads * U - 1) / (kScatterThreads * U);
        if (nb > (int64_t)num_sms * 8) nb = (int64_t)num_sms * 8;
        k<<<(int)nb, kScatterThreads, head, s>>>(a, gm, w, ge, le, (((uintptr_t)a.x) & 31) == 0 ? 1 : 0);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    const int smem = ((16 + head + 127) & ~127) + gm.nslots * gm.low_elems * 8;
    if ((e = launch_eval<RANK>(a, gm, w, ge, le, num_sms, smem, s)) != cudaSuccess) return e;
    if (a.blend == 0) {
        auto k = mlip_bin_overflow<RANK, true>;
        if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, head)) != cudaSuccess) return e;
        k<<<num_sms * 4, 256, head, s>>>(a, w, ge, le);
    } else {
        auto k = mlip_bin_overflow<RANK, false>;
        if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, head)) != cudaSuccess) return e;
        k<<<num_sms * 4, 256, head, s>>>(a, w, ge, le);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    *launches += 3;
    if (li) { li->grid = num_sms; li->block = kEvalThreads; li->smem = smem; li->kernel = K_MLIP_BIN; }
    return cudaSuccess;
}

}  // namespace

// ---- host interface ------------------------------------------------------------------------
struct MlipBinState {
    BinGeom gm;
    BinWork w;
    int64_t chunk = 0;
    size_t bytes = 0;
    void *block = nullptr;
};

int64_t mlip_bin_chunk_points() { return int64_t(1) << 25; }

// can this shape take the binned path for a batch of npts points, and is it worth it (enough points
// per bin to pay for loading the bin's runs)?
bool mlip_bin_supported(int rank, const int *dims, int64_t npts, int blend, bool force)
{
    if (blend < 0 || blend > 5) return false;
    BinGeom gm;
    const int64_t chunk = npts < mlip_bin_chunk_points() ? npts : mlip_bin_chunk_points();
    if (!bin_geometry(rank, dims, chunk, gm)) return false;
    const int64_t tile_bytes = (int64_t)gm.nslots * gm.low_elems * 8;
    return force || chunk / gm.nbins * 64 >= tile_bytes;  // staging the runs costs at most 64 bytes per point
}

void mlip_bin_destroy(MlipBinState *st)
{
    if (!st) return;
    cudaFree(st->block);
    delete st;
}

int64_t mlip_bin_bytes(const MlipBinState *st) { return st ? (int64_t)st->bytes : 0; }

// (re)allocates the workspace for batches of up to `npts` points (capped at the chunk size)
cudaError_t mlip_bin_prepare(MlipBinState **pst, int rank, const int *dims, int64_t npts)
{
    const int64_t chunk = npts < mlip_bin_chunk_points() ? npts : mlip_bin_chunk_points();
    if (*pst && (*pst)->chunk >= chunk && (*pst)->chunk <= 4 * chunk) return cudaSuccess;
    mlip_bin_destroy(*pst);
    *pst = nullptr;
    MlipBinState *st = new MlipBinState();
    if (!bin_geometry(rank, dims, chunk, st->gm)) { delete st; return cudaErrorNotSupported; }
    st->chunk = chunk;
    const BinGeom &gm = st->gm;
    auto up = [](size_t b) { return (b + 255) & ~size_t(255); };
    const size_t slots = (size_t)gm.nbins * gm.cap;
    const size_t b_bx = up(slots * rank * 8), b_bidx = up(slots * 4), b_cur = up((size_t)gm.nbins * 4 * kCursorStride), b_cnt = 256,
                 b_oidx = up((size_t)chunk * 4);
    st->bytes = b_bx + b_bidx + b_cur + b_cnt + b_oidx;
    cudaError_t e = cudaMalloc(&st->block, st->bytes);
    if (e != cudaSuccess) { delete st; return e; }
    unsigned char *p = static_cast<unsigned char *>(st->block);
    st->w.bx = reinterpret_cast<double *>(p); p += b_bx;
    st->w.bidx = reinterpret_cast<int *>(p); p += b_bidx;
    st->w.cursor = reinterpret_cast<int *>(p); p += b_cur;
    st->w.counters = reinterpret_cast<int *>(p); p += b_cnt;
    st->w.oidx = reinterpret_cast<int *>(p);
    *pst = st;
    return cudaSuccess;
}

cudaError_t launch_mlip_bin(const MlipArgs &a_in, MlipBinState *st, int num_sms, cudaStream_t s, LaunchInfo *li, int *launches)
{
    if (!st || (((uintptr_t)a_in.x) & 15) != 0 || !a_in.lut) return cudaErrorNotSupported;
    *launches = 0;
    for (int64_t lo = 0; lo < a_in.npts; lo += st->chunk) {
        MlipArgs a = a_in;
        a.x = a_in.x + lo * a.rank;
        a.out = a_in.out + lo;
        a.npts = a_in.npts - lo < st->chunk ? a_in.npts - lo : st->chunk;
        BinGeom gm = st->gm;  // capacity was sized for a full chunk; a shorter tail chunk fits a fortiori
        cudaError_t e;
        switch (a.rank) {
        case 2: e = run_chunk<2>(a, gm, st->w, num_sms, s, launches, li); break;
        case 3: e = run_chunk<3>(a, gm, st->w, num_sms, s, launches, li); break;
        case 4: e = run_chunk<4>(a, gm, st->w, num_sms, s, launches, li); break;
        case 5: e = run_chunk<5>(a, gm, st->w, num_sms, s, launches, li); break;
        case 6: e = run_chunk<6>(a, gm, st->w, num_sms, s, launches, li); break;
        default: e = cudaErrorNotSupported;
        }
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}
