// cheb_slab.cu -- fast tensor-product Chebyshev evaluation for sm_100a.
//
// Replaces the per-point recursion of C_evalcheb (src/chebpol.c:314-338) for
// tensors whose innermost dimension is small (n0 <= 32), which covers the
// shapes chebpol is used with (BASELINE configs 1, 2 and 5).
//
// Work decomposition (points are the parallel axis):
//  * every consumer thread owns PT points for the whole pass over the tensor;
//  * the T_k(x_0) basis of the innermost dimension lives in registers
//    (PT*n0p doubles), so the dominant contraction -- prod(dims) fused
//    multiply-adds per point -- is one DFMA per coefficient with the
//    coefficient fetched as a warp-broadcast LDS.128 (two coefficients feed
//    2*PT DFMAs);
//  * dimensions 1..depth-1 (inside a slab) and depth..rank-1 (across slabs)
//    use the Clenshaw recurrence b <- 2x*b1 - b2 + v, whose cost is 1/n0 of
//    the total or less;
//  * the coefficient tensor stays in L2 and is streamed, slab by slab (a slab
//    = the first `depth` dims, <= 48 KB), into a shared-memory ring with the
//    TMA bulk-copy engine (cp.async.bulk + mbarrier complete_tx); there is no
//    producer warp: the last warp to finish a slab issues the refill of its
//    stage, so all warps are arithmetic warps and none waits on a producer;
//    when the whole tensor fits in the ring it is loaded once per CTA and
//    stays resident;
//  * one persistent CTA per SM walks point tiles of NW*32*PT points.
//
// Rounding: the innermost sum is sum_k T_k(x0)*c_k (ascending k, FMA) instead
// of Clenshaw, and the outer recurrences use FMA; differences from the
// reference are O(n*eps*sum|c|), far inside 1e-12*max|f| for fitted tensors
// (tests/test_parity_gpu.py states the bound).
#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int kMaxStages = 8;
constexpr int kBarBytes = 128;  // 2*kMaxStages mbarriers

template <int N0P, int PT>
struct RowEval {
    // T[p][k] = T_k(x0) of point p (k < n0), 0 for the pad column
    double T[PT][N0P];

    __device__ __forceinline__ void init(const double (&x0)[PT], int n0)
    {
#pragma unroll
        for (int p = 0; p < PT; ++p) {
            const double tx = 2.0 * x0[p];
            double tkm1 = 1.0, tk = x0[p];
#pragma unroll
            for (int k = 0; k < N0P; ++k) {
                double t;
                if (k == 0) t = 1.0;
                else if (k == 1) t = x0[p];
                else {
                    t = fma(tx, tk, -tkm1);
                    tkm1 = tk;
                    tk = t;
                }
                T[p][k] = (k < n0) ? t : 0.0;
            }
        }
    }

    // v[p] = sum_k T[p][k] * row[k]; row is 16-byte aligned shared memory
    __device__ __forceinline__ void dot(const double *row, double (&v)[PT]) const
    {
        const double2 *r2 = reinterpret_cast<const double2 *>(row);
        double2 c[N0P / 2];
#pragma unroll
        for (int k = 0; k < N0P / 2; ++k) c[k] = r2[k];
#pragma unroll
        for (int p = 0; p < PT; ++p) v[p] = T[p][0] * c[0].x;
#pragma unroll
        for (int p = 0; p < PT; ++p) v[p] = fma(T[p][1], c[0].y, v[p]);
#pragma unroll
        for (int k = 1; k < N0P / 2; ++k) {
#pragma unroll
            for (int p = 0; p < PT; ++p) v[p] = fma(T[p][2 * k], c[k].x, v[p]);
#pragma unroll
            for (int p = 0; p < PT; ++p) v[p] = fma(T[p][2 * k + 1], c[k].y, v[p]);
        }
    }

    // two rows at once: 2*PT independent accumulation chains per thread, so one warp
    // alone covers the DFMA dependent-issue latency
    __device__ __forceinline__ void dot2(const double *row_a, const double *row_b, double (&va)[PT],
                                         double (&vb)[PT]) const
    {
        const double2 *ra = reinterpret_cast<const double2 *>(row_a);
        const double2 *rb = reinterpret_cast<const double2 *>(row_b);
#pragma unroll
        for (int k = 0; k < N0P / 2; ++k) {
            const double2 ca = ra[k], cb = rb[k];
            if (k == 0) {
#pragma unroll
                for (int p = 0; p < PT; ++p) {
                    va[p] = T[p][0] * ca.x;
                    vb[p] = T[p][0] * cb.x;
                }
            } else {
#pragma unroll
                for (int p = 0; p < PT; ++p) {
                    va[p] = fma(T[p][2 * k], ca.x, va[p]);
                    vb[p] = fma(T[p][2 * k], cb.x, vb[p]);
                }
            }
#pragma unroll
            for (int p = 0; p < PT; ++p) {
                va[p] = fma(T[p][2 * k + 1], ca.y, va[p]);
                vb[p] = fma(T[p][2 * k + 1], cb.y, vb[p]);
            }
        }
    }
};

// Clenshaw step for PT points: b <- tx*b + (v - b1), b1 <- old b
template <int PT>
__device__ __forceinline__ void clenshaw_step(double (&b)[PT], double (&b1)[PT], const double (&tx)[PT],
                                              const double (&v)[PT])
{
#pragma unroll
    for (int p = 0; p < PT; ++p) {
        const double nb = fma(tx[p], b[p], v[p] - b1[p]);
        b1[p] = b[p];
        b[p] = nb;
    }
}

// result of a finished recurrence: b - x*b1 with x = tx/2
template <int PT>
__device__ __forceinline__ void clenshaw_finish(const double (&b)[PT], const double (&b1)[PT],
                                                const double (&tx)[PT], double (&v)[PT])
{
#pragma unroll
    for (int p = 0; p < PT; ++p) v[p] = fma(-0.5 * tx[p], b1[p], b[p]);
}

// Clenshaw over the n1 rows ending at row_top (row n1-1), descending; RB rows per step
template <int N0P, int PT, int RB>
__device__ __forceinline__ void level1_eval(const RowEval<N0P, PT> &re, const double *row_top, int n1,
                                            const double (&tx1)[PT], double (&v)[PT])
{
    double b[PT], b1[PT];
#pragma unroll
    for (int p = 0; p < PT; ++p) b[p] = b1[p] = 0.0;
    const double *row = row_top;
    int i = n1;
    if constexpr (RB == 2) {
        if (i & 1) {
            double r[PT];
            re.dot(row, r);
            clenshaw_step<PT>(b, b1, tx1, r);
            row -= N0P;
            --i;
        }
        for (; i > 0; i -= 2, row -= 2 * N0P) {
            double r0[PT], r1[PT];
            re.dot2(row, row - N0P, r0, r1);
            clenshaw_step<PT>(b, b1, tx1, r0);
            clenshaw_step<PT>(b, b1, tx1, r1);
        }
    } else {
#pragma unroll 2
        for (; i > 0; --i, row -= N0P) {
            double r[PT];
            re.dot(row, r);
            clenshaw_step<PT>(b, b1, tx1, r);
        }
    }
    clenshaw_finish<PT>(b, b1, tx1, v);
}

template <int N0P, int PT, int NW, int MINB, int RB>
__global__ void __launch_bounds__(NW * 32, MINB) cheb_slab_kernel(const __grid_constant__ SlabArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem);
    int *done_cnt = reinterpret_cast<int *>(smem + kMaxStages * 8);
    unsigned char *ring = smem + kBarBytes;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    constexpr int TP = NW * 32 * PT;  // points per tile
    const int64_t ntiles = (a.npts + TP - 1) / TP;
    const int ns = a.ns;
    const uint32_t slab_bytes = (uint32_t)a.slab_elems * 8u;
    // fills of the ring by this CTA: one per (tile, slab), slabs last-to-first
    const int64_t my_tiles = (ntiles > (int64_t)blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t total_fills = a.resident ? (int64_t)a.nslabs : my_tiles * a.nslabs;

    if (tid == 0) {
        for (int s = 0; s < ns; ++s) {
            mbar_init(&full_bar[s], 1);
            done_cnt[s] = 0;
        }
        mbar_fence_init();
        if (my_tiles > 0) {
            for (int f = 0; f < ns && f < total_fills; ++f) {
                const int s = a.resident ? f : a.nslabs - 1 - (f % a.nslabs);
                mbar_expect_tx(&full_bar[f], slab_bytes);
                tma_bulk_g2s(ring + (size_t)f * a.stage_stride, a.cf + (size_t)s * a.slab_elems, slab_bytes,
                             &full_bar[f]);
            }
        }
    }
    __syncthreads();

    const int rank = a.rank, depth = a.depth;
    const int n1 = (depth > 1) ? a.dims[1] : 1;
    const int n2 = (depth > 2) ? a.dims[2] : 1;
    int stage = 0;
    uint32_t round = 0;
    int64_t fill = 0;  // index of the fill being consumed

    // 1-D and 2-D tensors (the latency regime: a tile is a few microseconds of arithmetic, two warps per
    // scheduler): the next tile's coordinates are fetched while the current tile computes, so that a tile does
    // not begin with a DRAM round trip.  Only the instantiations with few points per thread carry the registers.
    constexpr bool kPrefetch = PT <= 3;
    const bool pre = kPrefetch && a.rank <= 2;
    double xn[kPrefetch ? PT : 1][2];
    auto fetch_x = [&](int64_t tile) {
        if constexpr (kPrefetch) {
#pragma unroll
            for (int p = 0; p < PT; ++p) {
                const int64_t idx = tile * TP + (int64_t)p * (NW * 32) + tid;
                const bool ok = tile < ntiles && idx < a.npts;
                xn[p][0] = ok ? ldg_stream_f64(a.x + idx * a.rank) : 0.0;
                xn[p][1] = (ok && a.rank > 1) ? ldg_stream_f64(a.x + idx * a.rank + 1) : 0.0;
            }
        }
    };
    if (pre) fetch_x(blockIdx.x);

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        RowEval<N0P, PT> re;
        double tx1[PT], tx2[PT];       // 2*x for the in-slab levels 1, 2
        double txo[CP_MAXR][PT];       // 2*x for the levels across slabs (local memory)
        double ob[CP_MAXR][PT], ob1[CP_MAXR][PT];
        int64_t pidx[PT];
        {
            double x0[PT];
#pragma unroll
            for (int p = 0; p < PT; ++p) {
                pidx[p] = tile * TP + (int64_t)p * (NW * 32) + tid;
                const bool ok = pidx[p] < a.npts;
                const double *xp = a.x + pidx[p] * rank;
                if (kPrefetch && pre) {
                    x0[p] = ok ? premap_coord(a.pm, 0, xn[kPrefetch ? p : 0][0]) : 0.0;
                    tx1[p] = (ok && rank > 1) ? 2.0 * premap_coord(a.pm, 1, xn[kPrefetch ? p : 0][1]) : 0.0;
                    tx2[p] = 0.0;
                    continue;
                }
                x0[p] = ok ? premap_coord(a.pm, 0, xp[0]) : 0.0;
                tx1[p] = (ok && rank > 1) ? 2.0 * premap_coord(a.pm, 1, xp[1]) : 0.0;
                tx2[p] = (ok && rank > 2) ? 2.0 * premap_coord(a.pm, 2, xp[2]) : 0.0;
                for (int l = depth; l < rank; ++l) {
                    txo[l][p] = ok ? 2.0 * premap_coord(a.pm, l, xp[l]) : 0.0;
                    ob[l][p] = 0.0;
                    ob1[l][p] = 0.0;
                }
            }
            re.init(x0, a.n0);
            if (pre) fetch_x(tile + gridDim.x);
        }

        double res[PT];
#pragma unroll
        for (int p = 0; p < PT; ++p) res[p] = 0.0;
        for (int s = a.nslabs - 1; s >= 0; --s) {
            const int st = a.resident ? s : stage;
            mbar_wait(&full_bar[st], a.resident ? 0u : (round & 1));
            const double *slab = reinterpret_cast<const double *>(ring + (size_t)st * a.stage_stride);

            double v[PT];
            if (depth == 1) {
                re.dot(slab, v);
            } else if (depth == 2) {
                level1_eval<N0P, PT, RB>(re, slab + (size_t)(n1 - 1) * N0P, n1, tx1, v);
            } else {
                double c[PT], c1[PT];
#pragma unroll
                for (int p = 0; p < PT; ++p) c[p] = c1[p] = 0.0;
                const double *top = slab + ((size_t)n2 * n1 - 1) * N0P;
                for (int i2 = n2 - 1; i2 >= 0; --i2, top -= (size_t)n1 * N0P) {
                    double w[PT];
                    level1_eval<N0P, PT, RB>(re, top, n1, tx1, w);
                    clenshaw_step<PT>(c, c1, tx2, w);
                }
                clenshaw_finish<PT>(c, c1, tx2, v);
            }

            if (!a.resident) {
                // release the stage; the last warp to finish it issues the
                // refill (fill + ns), so no warp ever waits on a producer
                __syncwarp();
                if (lane == 0) {
                    __threadfence_block();
                    const int old = atomicAdd(&done_cnt[stage], 1);
                    if (old == NW - 1) {
                        done_cnt[stage] = 0;
                        const int64_t nf = fill + ns;
                        if (nf < total_fills) {
                            const int ss = a.nslabs - 1 - (int)(nf % a.nslabs);
                            __threadfence_block();
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                            mbar_expect_tx(&full_bar[stage], slab_bytes);
                            tma_bulk_g2s(ring + (size_t)stage * a.stage_stride,
                                         a.cf + (size_t)ss * a.slab_elems, slab_bytes, &full_bar[stage]);
                        }
                    }
                }
                ++fill;
                if (++stage == ns) { stage = 0; ++round; }
            }

            // levels across slabs: feed v into level `depth`, carry upwards
            // whenever a level has consumed its index 0
            int l = depth;
            while (l < rank) {
#pragma unroll
                for (int p = 0; p < PT; ++p) {
                    const double nb = fma(txo[l][p], ob[l][p], v[p] - ob1[l][p]);
                    ob1[l][p] = ob[l][p];
                    ob[l][p] = nb;
                }
                if (((unsigned)s % a.cum[l]) != 0u) break;
#pragma unroll
                for (int p = 0; p < PT; ++p) {
                    v[p] = fma(-0.5 * txo[l][p], ob1[l][p], ob[l][p]);
                    ob[l][p] = 0.0;
                    ob1[l][p] = 0.0;
                }
                ++l;
            }
            if (l == rank) {
#pragma unroll
                for (int p = 0; p < PT; ++p) res[p] = v[p];
            }
        }
#pragma unroll
        for (int p = 0; p < PT; ++p)
            if (pidx[p] < a.npts) a.out[pidx[p]] = res[p];
    }
}

template <int N0P, int PT, int NW, int MINB, int RB = 1>
cudaError_t launch_one(const SlabArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    auto kern = cheb_slab_kernel<N0P, PT, NW, MINB, RB>;
    const int smem = kBarBytes + a.ns * a.stage_stride;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    constexpr int TP = NW * 32 * PT;
    int64_t ntiles = (a.npts + TP - 1) / TP;
    int64_t cap = (int64_t)num_sms * MINB;
    int grid = (int)(ntiles < cap ? ntiles : cap);
    if (grid < 1) grid = 1;
    kern<<<grid, NW * 32, smem, s>>>(a);
    if (li) { li->grid = grid; li->block = NW * 32; li->smem = smem; li->kernel = K_CHEB_SLAB; }
    return cudaGetLastError();
}

}  // namespace

// padded innermost length for which an instantiation exists (0: none)
int cheb_slab_pick_n0p(int n0)
{
    static const int avail[] = {2, 4, 6, 8, 10, 12, 14, 16, 20, 24, 28, 32};
    for (int v : avail)
        if (n0 <= v) return v;
    return 0;
}

// variant: 0 = default (PT, NW) for the shape; otherwise PT*100 + NW selects
// one of the tuning instantiations compiled for n0p = 10 and 12 (NW = 4 runs
// two CTAs per SM).
cudaError_t launch_cheb_slab(const SlabArgs &a, int variant, int num_sms, cudaStream_t s, LaunchInfo *li)
{
#define CASE(N0P, PT, NW, MINB) return launch_one<N0P, PT, NW, MINB>(a, num_sms, s, li)
    if (variant == 1) variant = 0;  // "the DFMA slab kernel, default tuning"
    if (variant != 0 && (a.n0p == 10 || a.n0p == 12)) {
        const int rb = variant / 10000 ? variant / 10000 : 1;
        const int pt = (variant % 10000) / 100, nw = variant % 100;
#define TUNE(N0P)                                        \
    if (a.n0p == N0P && rb == 2) {                       \
        if (pt == 3 && nw == 8) return launch_one<N0P, 3, 8, 1, 2>(a, num_sms, s, li); \
        if (pt == 4 && nw == 8) return launch_one<N0P, 4, 8, 1, 2>(a, num_sms, s, li); \
        if (pt == 5 && nw == 8) return launch_one<N0P, 5, 8, 1, 2>(a, num_sms, s, li); \
        if (pt == 6 && nw == 8) return launch_one<N0P, 6, 8, 1, 2>(a, num_sms, s, li); \
        if (pt == 4 && nw == 4) return launch_one<N0P, 4, 4, 2, 2>(a, num_sms, s, li); \
        if (pt == 6 && nw == 4) return launch_one<N0P, 6, 4, 1, 2>(a, num_sms, s, li); \
        if (pt == 3 && nw == 12) return launch_one<N0P, 3, 12, 1, 2>(a, num_sms, s, li); \
    }                                                    \
    if (a.n0p == N0P && rb == 1) {                       \
        if (pt == 2 && nw == 8) CASE(N0P, 2, 8, 1);      \
        if (pt == 3 && nw == 8) CASE(N0P, 3, 8, 1);      \
        if (pt == 4 && nw == 8) CASE(N0P, 4, 8, 1);      \
        if (pt == 5 && nw == 8) CASE(N0P, 5, 8, 1);      \
        if (pt == 6 && nw == 8) CASE(N0P, 6, 8, 1);      \
        if (pt == 4 && nw == 4) CASE(N0P, 4, 4, 2);      \
        if (pt == 6 && nw == 4) CASE(N0P, 6, 4, 2);      \
        if (pt == 2 && nw == 16) CASE(N0P, 2, 16, 1);    \
        if (pt == 3 && nw == 12) CASE(N0P, 3, 12, 1);    \
    }
        TUNE(10)
        TUNE(12)
#undef TUNE
        return cudaErrorNotSupported;
    }
    switch (a.n0p) {
    case 2: CASE(2, 6, 8, 1);
    case 4: CASE(4, 6, 8, 1);
    case 6: CASE(6, 6, 8, 1);
    case 8: CASE(8, 6, 8, 1);
    case 10: CASE(10, 4, 8, 1);
    case 12: CASE(12, 4, 8, 1);
    case 14: CASE(14, 4, 8, 1);
    case 16: CASE(16, 4, 8, 1);
    case 20: CASE(20, 3, 8, 1);
    case 24: CASE(24, 3, 8, 1);
    case 28: CASE(28, 2, 8, 1);
    case 32: CASE(32, 2, 8, 1);
    default: return cudaErrorNotSupported;
    }
#undef CASE
}

