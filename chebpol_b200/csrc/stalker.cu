// stalker.cu -- stalker / hyperbolic stalker spline evaluation for sm_100a.
//
// Replaces the point loops of R_evalstalk / R_evalhyp (src/stalker.c:765-832) over evalstalk
// (:366-442) and evalhyp (:537-613).  Same shape as the multilinear evaluator -- locate the
// cell, visit its 2^rank corners, blend -- but every corner carries a one-dimensional basis per
// dimension with three tabulated parameters, so a point touches 2^rank*(3*rank+1) doubles.
//  * The four (five) reference tables are packed at plan time into ONE record per grid point,
//    [val (+a), b_0.., c_0.., r_0.. or d_0.., pad] (16-byte aligned), so a corner is a single
//    contiguous run fetched with 128-bit read-only loads instead of four scattered segments;
//    the two corners along dimension 0 are adjacent records.
//  * The blend weights and the corner-relative coordinates are computed once per dimension
//    (the reference calls blendfun 2^rank*rank times per point); the corner weight is the same
//    product in the same order, so the `cw == 0` skip (:411,582) fires for the same corners.
//  * hstalker: d/(1 + c z) by the hardware reciprocal seed, two Newton steps and one residual
//    correction instead of the IEEE division slow path; stalker: |z|^r short-cuts r == 2 and
//    r == 1 (what the fit produces on most grid points) and calls pow only otherwise.
// One thread per point, corners fully unrolled up to rank 3 so all record loads are in flight
// together.  Differences from the reference: FMA contraction and the division: <= 1e-12*max|f|.
#include "common.cuh"
#include "ptx.cuh"
#include <cfloat>

namespace {

__host__ __device__ constexpr int record_stride_of(int rank) { return (3 * rank + 2) & ~1; }

__device__ __forceinline__ double blend_of(double w, int blend)
{
    switch (blend) {
    case 1: return w < 0.5 ? 0.5 * exp(2 - 1 / w) : 1 - 0.5 * exp(2 - 1 / (1 - w));
    case 2: return w < 0.5 ? 0.5 * exp(4 - 1 / (w * w)) : 1 - 0.5 * exp(4 - 1 / ((1 - w) * (1 - w)));
    case 3: return (-2 * w + 3) * w * w;
    case 4: return (w < 0.5) ? 0.0 : 1.0;
    case 5: return 0.5;
    default: return w;
    }
}

// knots <= x by bisection, clamped: R's findInterval(rightmost_closed = all_inside = TRUE) - 1
__device__ __forceinline__ int find_interval0(const double *__restrict__ g, int n, double x)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(g + mid) <= x) lo = mid + 1; else hi = mid;
    }
    lo = lo < 1 ? 1 : lo;
    lo = lo > n - 1 ? n - 1 : lo;
    return lo - 1;
}

// n/t to within an ulp without the IEEE slow path.  A zero or infinite t (the hyperbola's pole,
// which exact knot hits do reach: 1 + c*z == 0) takes n * seed, which is what IEEE division
// returns there (+-inf, 0 or NaN), so poles show up exactly where the reference's do.
__device__ __forceinline__ double fast_div(double n, double t)
{
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(t));
    double e = fma(-t, r0, 1.0);
    double r = fma(r0, e, r0);
    e = fma(-t, r, 1.0);
    r = fma(r, e, r);
    const double q = n * r;
    const double refined = fma(fma(-t, q, n), r, q);
    return (r0 == 0.0 || fabs(r0) > DBL_MAX) ? n * r0 : refined;
}

template <int RANK, int KIND>
__global__ void __launch_bounds__(256) stalk_cell_kernel(const __grid_constant__ StalkArgs a)
{
    constexpr int RS = record_stride_of(RANK);
    constexpr int NC = 1 << RANK;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < a.npts;
         p += (int64_t)gridDim.x * blockDim.x) {
        double wl[RANK], wh[RANK], zl[RANK], zh[RANK];
        int64_t ll = 0;
#pragma unroll
        for (int g = 0; g < RANK; ++g) {
            const double xg = ldg_stream_f64(a.x + p * RANK + g);
            const double *gr = a.grid + a.goff[g];
            const int gp = find_interval0(gr, a.dims[g], xg);
            const double g0 = __ldg(gr + gp), g1 = __ldg(gr + gp + 1);
            double w = 1 - (xg - g0) / (g1 - g0);
            w = w < 0 ? 0.0 : w;
            w = w > 1 ? 1.0 : w;
            wl[g] = blend_of(w, a.blend);
            wh[g] = blend_of(1 - w, a.blend);
            zl[g] = xg - g0;
            zh[g] = xg - g1;
            ll += a.stride[g] * gp;
        }
        double V = 0.0, sumw = 0.0;
#pragma unroll(RANK <= 3 ? NC : 4)
        for (int i = 0; i < NC; ++i) {
            int64_t corner = ll;
            double cw = 1.0;
#pragma unroll
            for (int g = 0; g < RANK; ++g) {
                const bool up = (i >> g) & 1;
                cw *= up ? wh[g] : wl[g];
                corner += up ? a.stride[g] : 0;
            }
            if (cw == 0.0) continue;
            double rec[RS];
            const double2 *rp = reinterpret_cast<const double2 *>(a.records + corner * RS);
#pragma unroll
            for (int q = 0; q < RS / 2; ++q) {
                const double2 v = __ldg(rp + q);
                rec[2 * q] = v.x;
                rec[2 * q + 1] = v.y;
            }
            double fval = rec[0];
#pragma unroll
            for (int j = 0; j < RANK; ++j) {
                const double z = ((i >> j) & 1) ? zh[j] : zl[j];
                const double bj = rec[1 + j], cj = rec[1 + RANK + j], tj = rec[1 + 2 * RANK + j];
                if (KIND == 0) {
                    double term = bj * z;
                    if (cj != 0) {
                        const double az = fabs(z);
                        const double pw = (tj == 2.0) ? z * z : ((tj == 1.0) ? az : pow(az, tj));
                        term = fma(cj, pw, term);
                    }
                    fval += term;
                } else {
                    // the denominator rounded as the reference rounds it (product, then sum), so that it
                    // vanishes for the same z
                    const double hyp = fma(bj, z, fast_div(tj, __dadd_rn(1.0, __dmul_rn(cj, z))));
                    fval += (cj == cj) ? hyp : bj * z * z;
                }
            }
            sumw += cw;
            V = fma(cw, fval, V);
        }
        a.out[p] = V / sumw;
    }
}

template <int RANK>
cudaError_t launch_rank(const StalkArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    const int block = 256;
    int64_t nb = (a.npts + block - 1) / block;
    const int64_t cap = (int64_t)num_sms * 16;
    if (nb > cap) nb = cap;
    if (nb < 1) nb = 1;
    if (a.kind == 0) stalk_cell_kernel<RANK, 0><<<(int)nb, block, 0, s>>>(a);
    else stalk_cell_kernel<RANK, 1><<<(int)nb, block, 0, s>>>(a);
    if (li) { li->grid = (int)nb; li->block = block; li->smem = 0; li->kernel = K_STALK_CELL; }
    return cudaGetLastError();
}

// [val (+a), b, t1, t2, pad] per grid point
__global__ void stalk_pack_kernel(const double *__restrict__ val, const double *__restrict__ aa,
                                  const double *__restrict__ b, const double *__restrict__ t1,
                                  const double *__restrict__ t2, int rank, int rs, int64_t n, double *__restrict__ rec)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double *r = rec + i * rs;
        r[0] = aa ? val[i] + aa[i] : val[i];  // the reference's val[corner] + a[corner] (:589), rounded once
        for (int j = 0; j < rank; ++j) {
            r[1 + j] = b[rank * i + j];
            r[1 + rank + j] = t1[rank * i + j];
            r[1 + 2 * rank + j] = t2[rank * i + j];
        }
        for (int q = 1 + 3 * rank; q < rs; ++q) r[q] = 0.0;
    }
}

}  // namespace

int stalk_record_stride(int rank, int kind)
{
    (void)kind;
    return (rank >= 1 && rank <= 6) ? record_stride_of(rank) : 0;
}

cudaError_t stalk_pack_records(const double *val, const double *aa, const double *b, const double *t1,
                               const double *t2, int rank, int64_t n, double *rec, cudaStream_t s)
{
    const int rs = stalk_record_stride(rank, 0);
    if (!rs) return cudaErrorNotSupported;
    int64_t nb = (n + 255) / 256;
    if (nb > 148 * 32) nb = 148 * 32;
    if (nb < 1) nb = 1;
    stalk_pack_kernel<<<(int)nb, 256, 0, s>>>(val, aa, b, t1, t2, rank, rs, n, rec);
    return cudaGetLastError();
}

cudaError_t launch_stalk_cell(const StalkArgs &a, int variant, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    (void)variant;
    if (!a.records) return cudaErrorNotSupported;
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
