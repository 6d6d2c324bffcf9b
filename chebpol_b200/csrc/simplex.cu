// simplex.cu -- simplex-linear interpolation on a Delaunay triangulation (slappx.real,
// R/simplexlinear.R:27-40): R_evalsl / findsimplex (src/chebpol.c:738-828) and the fit step
// R_analyzesimplex (:830-875).  Compiled with -fmad=false: every floating-point operation is
// the reference's, in the reference's order, so BOTH kernels are bit-identical to the reference C
// (x86-64 gcc -O2, no FMA, netlib LAPACK semantics for dgetrf / dgetrs) -- what differs between
// them is only how the candidate simplices are found.
//
//   sl_strict  one thread per point, the reference's linear scan over all simplices (bounding
//              boxes read at warp-uniform addresses), any dimension <= 31.
//   sl_cell    the same acceptance test on far fewer candidates.  The reference returns the FIRST
//              simplex, in index order, whose bounding box holds the point and whose barycentric
//              coordinates are all >= -1e-10.  A uniform grid over the knots' bounding box is built
//              at plan time; every cell lists, in ascending order, the simplices whose bounding box
//              overlaps the cell.  A point's cell is found with the same monotone, FMA-free
//              expression the host used to bin the boxes, so every simplex whose box holds the point
//              is in its cell's list, in index order: the first accepted candidate is the
//              reference's answer.  Per simplex one packed record (box, LU factors, row permutation,
//              vertex values) is read; the table lives in L2.  dim <= 6.
//
// Outside the convex hull: R's NA_real_ (bit pattern included), or with `epol` the first simplex
// holding the nearest knot (tabulated per knot at plan time), plain barycentric combination.
// NaN coordinates pass every box test in the reference (comparisons are false), so simplex 0
// decides; both kernels reproduce that.
#include "common.cuh"
#include <float.h>
#include <math.h>
#include <string.h>
#include <algorithm>
#include <vector>

#define SL_MAXCELLDIM 6

struct SlDevice {
    int dim = 0, nknots = 0, ns = 0;
    double *knots = nullptr, *lumats = nullptr, *bbox = nullptr, *val = nullptr, *records = nullptr;
    int *dtri = nullptr, *pivots = nullptr, *knot_simplex = nullptr, *cell_start = nullptr, *cell_items = nullptr;
    int rs = 0;  // record stride (doubles); 0: no cell structure (dim too large / no room)
    int G[SL_MAXCELLDIM] = {0};
    double lo[SL_MAXCELLDIM] = {0}, inv[SL_MAXCELLDIM] = {0};
    double glo[CP_MAXR] = {0}, ghi[CP_MAXR] = {0};  // bounding box of all simplices
    int64_t bytes = 0, nitems = 0, ncells = 0, cell_bytes = 0;
    int forced_G = 0;
    int64_t points_seen = 0;  // evaluated by this plan so far: sizes the grid (sl_wanted_G)
    // host copies kept for (re)building the cell lists
    std::vector<double> h_bbox, h_aff, h_tol;
};

struct SlArgs {
    int dim, nknots, ns, rs, epol, blend;
    const double *knots, *lumats, *bbox, *val, *records;
    const int *dtri, *pivots, *knot_simplex, *cell_start, *cell_items;
    int G[SL_MAXCELLDIM];
    double lo[SL_MAXCELLDIM], inv[SL_MAXCELLDIM];
    const double *x;
    int64_t npts;
    double *out;
};

__host__ __device__ __forceinline__ double sl_na_real()
{
#ifdef __CUDA_ARCH__
    return __longlong_as_double(0x7FF00000000007A2LL);  // R's NA_real_ (low word 1954)
#else
    union { double d; unsigned long long u; } v;
    v.u = 0x7FF00000000007A2ULL;
    return v.d;
#endif
}

// blendfun, src/chebpol.h:24-45
__device__ __forceinline__ double sl_blend(double w, int blend)
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

// cell coordinate of x along one dimension: the SAME expression on host (binning the boxes) and
// device (binning the points); monotone in x, so box_lo <= x <= box_hi implies
// cell(box_lo) <= cell(x) <= cell(box_hi)
__host__ __device__ __forceinline__ int sl_cell_of(double x, double lo, double inv, int G)
{
#ifdef __CUDA_ARCH__
    const double t = __dmul_rn(__dsub_rn(x, lo), inv);
#else
    const volatile double d = x - lo;  // keep the host compiler from contracting
    const double t = d * inv;
#endif
    if (!(t > 0.0)) return 0;
    if (t >= (double)G) return G - 1;
    return (int)t;
}

// dgetrs 'N' on one right-hand side, reference BLAS order (see oracle/ref_wrap.c): the row
// interchanges are applied by the caller (they only move data); unit-lower then upper
// substitution, each skipping zero entries
template <int N>
__device__ __forceinline__ void sl_substitute(const double *__restrict__ lu, double (&b)[N])
{
#pragma unroll
    for (int k = 0; k < N; ++k)
        if (b[k] != 0.0) {
#pragma unroll
            for (int i = k + 1; i < N; ++i) b[i] = b[i] - b[k] * lu[i + k * N];
        }
#pragma unroll
    for (int k = N - 1; k >= 0; --k)
        if (b[k] != 0.0) {
            b[k] = b[k] / lu[k + k * N];
#pragma unroll
            for (int i = 0; i < k; ++i) b[i] = b[i] - b[k] * lu[i + k * N];
        }
}

__device__ void sl_solve_generic(int N, const double *__restrict__ lu, const int *__restrict__ piv, double *b)
{
    for (int i = 0; i < N; ++i) {
        const int ip = piv[i] - 1;
        if (ip != i) { const double t = b[i]; b[i] = b[ip]; b[ip] = t; }
    }
    for (int k = 0; k < N; ++k)
        if (b[k] != 0.0)
            for (int i = k + 1; i < N; ++i) b[i] = b[i] - b[k] * lu[i + k * N];
    for (int k = N - 1; k >= 0; --k)
        if (b[k] != 0.0) {
            b[k] = b[k] / lu[k + k * N];
            for (int i = 0; i < k; ++i) b[i] = b[i] - b[k] * lu[i + k * N];
        }
}

// row permutation packed 4 bits per row: after dlaswp, b[i] = [x; 1][perm_i]
template <int DIM>
__device__ __forceinline__ void sl_permuted_rhs(const double (&x)[DIM], unsigned perm, double (&b)[DIM + 1])
{
#pragma unroll
    for (int i = 0; i <= DIM; ++i) {
        const int pi = (perm >> (4 * i)) & 15;
        double v = 1.0;
#pragma unroll
        for (int j = 0; j < DIM; ++j) v = (pi == j) ? x[j] : v;
        b[i] = v;
    }
}

// nearest knot, then the first simplex holding it (findsimplex :775-799)
__device__ double sl_extrapolate(const SlArgs &a, const double *x)
{
    const int dim = a.dim, N = dim + 1;
    double mindist = 1e99;
    int nearknot = 0;
    for (int k = 0; k < a.nknots; ++k) {
        double dist = 0.0;
        for (int i = 0; i < dim; ++i) {
            const double kv = __ldg(a.knots + (int64_t)k * dim + i);
            dist += (x[i] - kv) * (x[i] - kv);
        }
        if (dist < mindist) { mindist = dist; nearknot = k; }
    }
    const int s = __ldg(a.knot_simplex + nearknot);
    if (s < 0) return sl_na_real();
    double vec[CP_MAXR];
    for (int d = 0; d < dim; ++d) vec[d] = x[d];
    vec[dim] = 1.0;
    sl_solve_generic(N, a.lumats + (int64_t)s * N * N, a.pivots + (int64_t)s * N, vec);
    const int *tri = a.dtri + (int64_t)s * N;
    double sum = 0.0;
    for (int d = 0; d <= dim; ++d) sum += vec[d] * __ldg(a.val + tri[d] - 1);
    return sum;
}

// ------------------------------------------------------------------ strict: the linear scan
__global__ void __launch_bounds__(256) sl_strict_kernel(const __grid_constant__ SlArgs a)
{
    const int dim = a.dim, N = dim + 1;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < a.npts; p += (int64_t)gridDim.x * blockDim.x) {
        double x[CP_MAXR], vec[CP_MAXR];
        for (int d = 0; d < dim; ++d) x[d] = a.x[p * dim + d];
        bool found = false;
        double res = 0.0;
        for (int s = 0; s < a.ns && !found; ++s) {
            const double *box = a.bbox + (int64_t)s * 2 * dim;
            bool bad = false;
            for (int i = 0; i < dim; ++i)
                if (x[i] < __ldg(box + 2 * i) || x[i] > __ldg(box + 2 * i + 1)) { bad = true; break; }
            if (bad) continue;
            for (int d = 0; d < dim; ++d) vec[d] = x[d];
            vec[dim] = 1.0;
            sl_solve_generic(N, a.lumats + (int64_t)s * N * N, a.pivots + (int64_t)s * N, vec);
            for (int d = 0; d < N; ++d)
                if (vec[d] < -1e-10) { bad = true; break; }
            if (bad) continue;
            const int *tri = a.dtri + (int64_t)s * N;
            double sum = 0.0, wsum = 0.0;
            for (int d = 0; d < N; ++d) {
                const double w = sl_blend(vec[d], a.blend);
                wsum += w;
                sum += w * __ldg(a.val + tri[d] - 1);
            }
            res = sum / wsum;
            found = true;
        }
        if (!found) res = a.epol ? sl_extrapolate(a, x) : sl_na_real();
        a.out[p] = res;
    }
}

// ------------------------------------------------------------------ cell lists
// record of simplex s at records + s*rs: [box lo/hi interleaved (2 DIM) | LU (N*N, column-major) |
// vertex values (N) | row permutation (low 32 bits of one slot)]
template <int DIM>
__device__ __forceinline__ bool sl_box(const double *__restrict__ rec, const double (&x)[DIM])
{
    bool in = true;
#pragma unroll
    for (int i = 0; i < DIM; ++i) {
        const double2 bx = __ldg(reinterpret_cast<const double2 *>(rec) + i);
        in &= !(x[i] < bx.x || x[i] > bx.y);
    }
    return in;
}

template <int DIM>
__device__ __forceinline__ bool sl_accept(const double *__restrict__ rec, const double (&x)[DIM], int blend, double *res)
{
    constexpr int N = DIM + 1;
    double lu[N * N];
    {   // the LU block starts 16-byte aligned (2 DIM doubles into a 32-byte aligned record)
        const double2 *q = reinterpret_cast<const double2 *>(rec + 2 * DIM);
#pragma unroll
        for (int i = 0; i + 1 < N * N; i += 2) {
            const double2 v = __ldg(q + i / 2);
            lu[i] = v.x;
            lu[i + 1] = v.y;
        }
        if ((N * N) & 1) lu[N * N - 1] = __ldg(rec + 2 * DIM + N * N - 1);
    }
    const unsigned perm = (unsigned)__double_as_longlong(__ldg(rec + 2 * DIM + N * N + N));
    double b[N];
    sl_permuted_rhs<DIM>(x, perm, b);
    sl_substitute<N>(lu, b);
#pragma unroll
    for (int d = 0; d < N; ++d)
        if (b[d] < -1e-10) return false;
    double sum = 0.0, wsum = 0.0;
#pragma unroll
    for (int d = 0; d < N; ++d) {
        const double w = sl_blend(b[d], blend);
        wsum += w;
        sum += w * __ldg(rec + 2 * DIM + N * N + d);
    }
    *res = sum / wsum;
    return true;
}

// SL_BATCH candidates are box-tested together, so their loads are in flight at the same time; the
// ones that pass are then solved in index order, which keeps "first accepted" the reference's
template <int DIM, int SL_BATCH, int MINB>
__global__ void __launch_bounds__(128, MINB) sl_cell_kernel(const __grid_constant__ SlArgs a)
{
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < a.npts; p += (int64_t)gridDim.x * blockDim.x) {
        double x[DIM];
        bool nan = false;
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
            x[d] = a.x[p * DIM + d];
            nan |= (x[d] != x[d]);
        }
        double res = 0.0;
        bool found = false;
        if (nan) {
            // a NaN coordinate passes every box test of the reference (comparisons are false), so its
            // scan is over all simplices, restricted by the finite coordinates: do exactly that
            for (int s = 0; s < a.ns && !found; ++s) {
                const double *rec = a.records + (int64_t)s * a.rs;
                if (sl_box<DIM>(rec, x)) found = sl_accept<DIM>(rec, x, a.blend, &res);
            }
        } else {
            int cell = 0;
#pragma unroll
            for (int d = DIM - 1; d >= 0; --d) cell = cell * a.G[d] + sl_cell_of(x[d], a.lo[d], a.inv[d], a.G[d]);
            const int beg = __ldg(a.cell_start + cell), end = __ldg(a.cell_start + cell + 1);
            for (int it = beg; it < end && !found; it += SL_BATCH) {
                int sid[SL_BATCH];
                unsigned mask = 0;
#pragma unroll
                for (int j = 0; j < SL_BATCH; ++j) sid[j] = (it + j < end) ? __ldg(a.cell_items + it + j) : -1;
#pragma unroll
                for (int j = 0; j < SL_BATCH; ++j) {
                    const double *rec = a.records + (int64_t)(sid[j] < 0 ? 0 : sid[j]) * a.rs;
                    if (sl_box<DIM>(rec, x) && sid[j] >= 0) mask |= 1u << j;
                }
                while (mask && !found) {  // ascending j = ascending simplex index
                    const int j = __ffs(mask) - 1;
                    mask &= mask - 1;
                    int sj = sid[0];
#pragma unroll
                    for (int q = 1; q < SL_BATCH; ++q) sj = (j == q) ? sid[q] : sj;
                    found = sl_accept<DIM>(a.records + (int64_t)sj * a.rs, x, a.blend, &res);
                }
            }
        }
        if (!found) {
            if (a.epol) {
                double xl[DIM];  // a copy whose address may be taken: x itself stays in registers
#pragma unroll
                for (int d = 0; d < DIM; ++d) xl[d] = x[d];
                res = sl_extrapolate(a, xl);
            } else
                res = sl_na_real();
        }
        a.out[p] = res;
    }
}

// The same search as a per-lane state machine.  In sl_cell_kernel a warp stays on its 32 points until
// the slowest lane has found its simplex (1 to 6+ LU solves per point, so half the issue slots go to
// masked lanes).  Here every trip of the loop does, for each lane independently, "advance to the next
// candidate whose box holds my point, solve it, and on acceptance (or an exhausted list) store the
// result and move on to my next point": lanes progress through their own points at their own pace and
// the expensive solve is executed by all lanes together.  Candidate order per point is unchanged, so
// the result is the same bits.
template <int DIM, int MINB>
__global__ void __launch_bounds__(128, MINB) sl_cell_sm_kernel(const __grid_constant__ SlArgs a)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double x[DIM];
    int it = 0, end = 0;
    bool ident = false;  // NaN coordinate: the candidates are all simplices in order (see sl_cell_kernel)
    auto open_point = [&]() {
        bool nan = false;
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
            x[d] = a.x[p * DIM + d];
            nan |= (x[d] != x[d]);
        }
        ident = nan;
        if (nan) { it = 0; end = a.ns; return; }
        int cell = 0;
#pragma unroll
        for (int d = DIM - 1; d >= 0; --d) cell = cell * a.G[d] + sl_cell_of(x[d], a.lo[d], a.inv[d], a.G[d]);
        it = __ldg(a.cell_start + cell);
        end = __ldg(a.cell_start + cell + 1);
    };
    if (p < a.npts) open_point();
    while (p < a.npts) {
        // next candidate whose box holds the point
        const double *rec = nullptr;
        while (it < end) {
            const int s = ident ? it : __ldg(a.cell_items + it);
            ++it;
            const double *r = a.records + (int64_t)s * a.rs;
            if (sl_box<DIM>(r, x)) { rec = r; break; }
        }
        double res = 0.0;
        bool done = rec == nullptr;
        if (rec) done = sl_accept<DIM>(rec, x, a.blend, &res);
        else if (a.epol) {
            double xl[DIM];
#pragma unroll
            for (int d = 0; d < DIM; ++d) xl[d] = x[d];
            res = sl_extrapolate(a, xl);
        } else
            res = sl_na_real();
        if (done) {
            a.out[p] = res;
            p += stride;
            if (p < a.npts) open_point();
        }
    }
}

// ------------------------------------------------------------------ fit: R_analyzesimplex
// one thread per simplex; the matrix is factorised in place in the output array (dgetf2 order)
__global__ void sl_analyze_kernel(const int *__restrict__ dtri, int ns, const double *__restrict__ knots, int dim,
                                  double *__restrict__ lumats, int *__restrict__ pivots, double *__restrict__ bbox)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= ns) return;
    const int N = dim + 1;
    const int *tri = dtri + (int64_t)s * N;
    double *a = lumats + (int64_t)s * N * N, *m = bbox + (int64_t)s * 2 * dim;
    int *ipiv = pivots + (int64_t)s * N;
    for (int j = 0; j < dim; ++j) { m[2 * j] = 1e100; m[2 * j + 1] = -1e100; }
    for (int v = 0; v < N; ++v) {
        const double *knot = knots + (int64_t)(tri[v] - 1) * dim;
        for (int j = 0; j < dim; ++j) {
            a[v * N + j] = knot[j];
            if (knot[j] < m[2 * j]) m[2 * j] = knot[j];
            if (knot[j] > m[2 * j + 1]) m[2 * j + 1] = knot[j];
        }
        a[v * N + dim] = 1.0;
    }
    for (int j = 0; j < N; ++j) {
        int jp = j;
        double mx = fabs(a[j + j * N]);
        for (int i = j + 1; i < N; ++i)
            if (fabs(a[i + j * N]) > mx) { mx = fabs(a[i + j * N]); jp = i; }
        ipiv[j] = jp + 1;
        if (a[jp + j * N] != 0.0) {
            if (jp != j)
                for (int c = 0; c < N; ++c) { const double t = a[j + c * N]; a[j + c * N] = a[jp + c * N]; a[jp + c * N] = t; }
            if (j < N - 1) {
                if (fabs(a[j + j * N]) >= DBL_MIN) {
                    const double r = 1.0 / a[j + j * N];
                    for (int i = j + 1; i < N; ++i) a[i + j * N] *= r;
                } else
                    for (int i = j + 1; i < N; ++i) a[i + j * N] /= a[j + j * N];
            }
        }
        if (j < N - 1)
            for (int c = j + 1; c < N; ++c) {
                const double y = a[j + c * N];
                if (y != 0.0) {
                    const double t = -y;
                    for (int i = j + 1; i < N; ++i) a[i + c * N] += a[i + j * N] * t;
                }
            }
    }
}

cudaError_t sl_analyze_device(const int *d_dtri, int ns, const double *d_knots, int dim, double *d_lumats,
                              int *d_pivots, double *d_bbox, cudaStream_t s)
{
    if (ns <= 0) return cudaSuccess;
    sl_analyze_kernel<<<(ns + 127) / 128, 128, 0, s>>>(d_dtri, ns, d_knots, dim, d_lumats, d_pivots, d_bbox);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ plan-time tables
void sl_destroy(SlDevice *d)
{
    if (!d) return;
    cudaFree(d->knots);
    cudaFree(d->lumats);
    cudaFree(d->bbox);
    cudaFree(d->val);
    cudaFree(d->records);
    cudaFree(d->dtri);
    cudaFree(d->pivots);
    cudaFree(d->knot_simplex);
    cudaFree(d->cell_start);
    cudaFree(d->cell_items);
    delete d;
}

template <class T>
static cudaError_t sl_upload(T **dst, const T *src, size_t n, int64_t *bytes)
{
    cudaError_t e = cudaMalloc(dst, sizeof(T) * (n ? n : 1));
    if (e != cudaSuccess) return e;
    *bytes += sizeof(T) * n;
    return n ? cudaMemcpy(*dst, src, sizeof(T) * n, cudaMemcpyHostToDevice) : cudaSuccess;
}

// Exact barycentric maps lambda_d(x) = a_d . x + c_d of every simplex (host, long double Gauss-Jordan
// on [vertices; 1]), used ONLY to prune the cell lists: a simplex is left out of a cell's list when
// some lambda_d stays below -tol on the whole (padded) cell, tol covering the -1e-10 acceptance slack
// plus a bound on what the kernel's LU solve can deviate from the exact value
// (1e-12 (1 + |M|)(1 + |M^-1|) max|lambda| -- ten times the classical GEPP bound).  Such a simplex cannot
// be accepted for any point of the cell, so "first accepted candidate" is unchanged.  Degenerate
// simplices (no inverse) are never pruned.
static void sl_affine_maps(SlDevice *d, const double *knots, const int *dtri)
{
    const int dim = d->dim, N = dim + 1, ns = d->ns;
    d->h_aff.assign((size_t)ns * N * N, 0.0);
    d->h_tol.assign((size_t)ns, -1.0);  // < 0: do not prune
#pragma omp parallel for schedule(static)
    for (int s = 0; s < ns; ++s) {
        long double m[SL_MAXCELLDIM + 1][2 * (SL_MAXCELLDIM + 1)];
        double mnorm = (double)N;
        for (int r = 0; r < N; ++r)
            for (int c = 0; c < N; ++c) {
                const int k = dtri[(size_t)s * N + c] - 1;
                m[r][c] = (r < dim) ? (long double)knots[(size_t)k * dim + r] : 1.0L;
                m[r][N + c] = (r == c) ? 1.0L : 0.0L;
            }
        for (int r = 0; r < dim; ++r) {
            double rs = 0.0;
            for (int c = 0; c < N; ++c) rs += fabs((double)m[r][c]);
            mnorm = std::max(mnorm, rs);
        }
        bool ok = true;
        for (int c = 0; c < N && ok; ++c) {
            int pr = c;
            for (int r = c + 1; r < N; ++r)
                if (fabsl(m[r][c]) > fabsl(m[pr][c])) pr = r;
            if (!(fabsl(m[pr][c]) > 0.0L) || !std::isfinite((double)m[pr][c])) { ok = false; break; }
            if (pr != c)
                for (int q = 0; q < 2 * N; ++q) std::swap(m[c][q], m[pr][q]);
            const long double inv = 1.0L / m[c][c];
            for (int q = 0; q < 2 * N; ++q) m[c][q] *= inv;
            for (int r = 0; r < N; ++r)
                if (r != c && m[r][c] != 0.0L) {
                    const long double f = m[r][c];
                    for (int q = 0; q < 2 * N; ++q) m[r][q] -= f * m[c][q];
                }
        }
        if (!ok) continue;
        // row v of the inverse: lambda_v = sum_j inv[v][j] x_j + inv[v][dim]
        double inorm = 0.0, lam = 0.0;
        const double *bx = &d->h_bbox[(size_t)s * 2 * dim];
        for (int v = 0; v < N; ++v) {
            double rs = 0.0, mag = 0.0;
            for (int j = 0; j < N; ++j) {
                const double a = (double)m[v][N + j];
                d->h_aff[((size_t)s * N + v) * N + j] = a;
                rs += fabs(a);
                mag += (j < dim) ? fabs(a) * std::max(fabs(bx[2 * j]), fabs(bx[2 * j + 1])) : fabs(a);
            }
            inorm = std::max(inorm, rs);
            lam = std::max(lam, mag);
        }
        const double tol = 1e-10 + 1e-12 * (1.0 + mnorm) * (1.0 + inorm) * lam + 1e-14 * lam;
        if (std::isfinite(tol)) d->h_tol[s] = tol;
    }
}

// (Re)build the cell lists with G cells per dimension.  Lists hold, in ascending order, the simplices
// whose bounding box overlaps the cell and which the pruning test above cannot rule out.
static void sl_host_lists(SlDevice *d, int G, std::vector<int> &start, std::vector<int> &items)
{
    const int dim = d->dim, N = dim + 1, ns = d->ns;
    if (G < 1) G = 1;
    while (G > 1 && pow((double)G, dim) > (double)(1 << 24)) G = (G + 1) / 2;
    const double *bbox = d->h_bbox.data();
    for (;; G = (G + 1) / 2) {
        double w[SL_MAXCELLDIM], pad[SL_MAXCELLDIM];
        int64_t nc = 1;
        for (int j = 0; j < dim; ++j) {
            d->G[j] = G;
            d->lo[j] = d->glo[j];
            const double span = d->ghi[j] - d->glo[j];
            const bool good = span > 0.0 && std::isfinite(span);
            d->inv[j] = good ? (double)G / span : 0.0;
            w[j] = good ? span / (double)G : 0.0;
            pad[j] = 1e-6 * w[j] + 1e-12 * (fabs(d->glo[j]) + fabs(d->ghi[j]));
            nc *= G;
        }
        std::vector<int> count((size_t)nc + 1, 0);
        // a (simplex, cell) pair is kept unless some barycentric coordinate is surely below -tol on the cell
        auto keep = [&](int s, const int *c) -> bool {
            const double tol = d->h_tol[s];
            if (tol < 0.0) return true;
            const double *aff = &d->h_aff[(size_t)s * N * N];
            for (int v = 0; v < N; ++v) {
                double ub = aff[v * N + dim];
                for (int j = 0; j < dim; ++j) {
                    const double a = aff[v * N + j];
                    const double lo = d->glo[j] + c[j] * w[j] - pad[j], hi = d->glo[j] + (c[j] + 1) * w[j] + pad[j];
                    ub += std::max(a * lo, a * hi);
                }
                if (ub < -tol) return false;
            }
            return true;
        };
        auto visit = [&](int s, auto &&fn) {
            int c0[SL_MAXCELLDIM], c1[SL_MAXCELLDIM], c[SL_MAXCELLDIM];
            for (int j = 0; j < dim; ++j) {
                c0[j] = sl_cell_of(bbox[(size_t)s * 2 * dim + 2 * j], d->lo[j], d->inv[j], G);
                c1[j] = sl_cell_of(bbox[(size_t)s * 2 * dim + 2 * j + 1], d->lo[j], d->inv[j], G);
                if (c1[j] < c0[j]) c1[j] = c0[j];  // NaN box: bin it somewhere, the box test rejects it anyway
                c[j] = c0[j];
            }
            for (;;) {
                if (keep(s, c)) {
                    int64_t id = 0;
                    for (int j = dim - 1; j >= 0; --j) id = id * G + c[j];
                    fn(id);
                }
                int j = 0;
                for (; j < dim; ++j) {
                    if (++c[j] <= c1[j]) break;
                    c[j] = c0[j];
                }
                if (j == dim) break;
            }
        };
#pragma omp parallel for schedule(dynamic, 256)
        for (int s = 0; s < ns; ++s)
            visit(s, [&](int64_t id) {
#pragma omp atomic
                ++count[(size_t)id + 1];
            });
        int64_t total = 0;
        for (int64_t i = 0; i < nc; ++i) total += count[(size_t)i + 1];
        if (total > (int64_t(1) << 28) && G > 1) continue;  // too fine for the memory it takes: coarsen
        for (int64_t i = 0; i < nc; ++i) count[(size_t)i + 1] += count[(size_t)i];
        start.assign(count.begin(), count.end());
        items.assign((size_t)total, 0);
        std::vector<int> fill(count.begin(), count.end() - 1);
#pragma omp parallel for schedule(dynamic, 256)
        for (int s = 0; s < ns; ++s)
            visit(s, [&](int64_t id) {
                int pos;
#pragma omp atomic capture
                pos = fill[(size_t)id]++;
                items[(size_t)pos] = s;
            });
        // the parallel fill loses the order inside a list: restore ascending simplex index
#pragma omp parallel for schedule(dynamic, 4096)
        for (int64_t i = 0; i < nc; ++i)
            if (start[(size_t)i + 1] - start[(size_t)i] > 1) std::sort(items.begin() + start[(size_t)i], items.begin() + start[(size_t)i + 1]);
        d->nitems = total;
        d->ncells = nc;
        break;
    }
}

static cudaError_t sl_build_cells(SlDevice *d, int G)
{
    NvtxRange nvtx("chebpol_cuda:sl_candidate_lists");
    std::vector<int> start, items;
    sl_host_lists(d, G, start, items);
    cudaError_t e = cudaDeviceSynchronize();  // a previous launch may still read the old lists
    if (e != cudaSuccess) return e;
    cudaFree(d->cell_start);
    cudaFree(d->cell_items);
    d->cell_start = d->cell_items = nullptr;
    d->bytes -= d->cell_bytes;
    int64_t nb = 0;
    e = sl_upload(&d->cell_start, start.data(), start.size(), &nb);
    if (e == cudaSuccess) e = sl_upload(&d->cell_items, items.data(), items.size(), &nb);
    if (e != cudaSuccess) {
        cudaFree(d->cell_start);
        cudaFree(d->cell_items);
        d->cell_start = d->cell_items = nullptr;
        d->cell_bytes = 0;
        d->ncells = 0;
        return e;
    }
    d->cell_bytes = nb;
    d->bytes += nb;
    // a pageable cudaMemcpy may return while its DMA is still in flight on the legacy stream; the
    // kernel that follows runs on a non-blocking stream, so wait for the lists to land
    return cudaDeviceSynchronize();
}

// bounding box of all simplices and the barycentric maps: everything the list builder needs
static void sl_host_setup(SlDevice *d, const double *knots, const int *dtri, const double *bbox)
{
    const int dim = d->dim, ns = d->ns;
    d->h_bbox.assign(bbox, bbox + (size_t)2 * dim * ns);
    for (int j = 0; j < dim; ++j) { d->glo[j] = 1e300; d->ghi[j] = -1e300; }
    for (int s = 0; s < ns; ++s)
        for (int j = 0; j < dim; ++j) {
            d->glo[j] = std::min(d->glo[j], bbox[(size_t)s * 2 * dim + 2 * j]);
            d->ghi[j] = std::max(d->ghi[j], bbox[(size_t)s * 2 * dim + 2 * j + 1]);
        }
    sl_affine_maps(d, knots, dtri);
}

// Host-only view of the candidate lists (no GPU needed): tests check, against the oracle, that the
// simplex the reference accepts for a point is always in the list of the point's cell.
int sl_debug_lists(const double *knots, int dim, int nknots, const int *dtri, int ns, const double *bbox, int G,
                   int *G_used, double *lo, double *inv, int *start, int64_t start_cap, int *items, int64_t items_cap,
                   int64_t *nitems)
{
    (void)nknots;
    if (dim < 1 || dim > SL_MAXCELLDIM || ns < 1) return 1;
    SlDevice d;
    d.dim = dim;
    d.ns = ns;
    sl_host_setup(&d, knots, dtri, bbox);
    std::vector<int> st, it;
    sl_host_lists(&d, G, st, it);
    if ((int64_t)st.size() > start_cap || (int64_t)it.size() > items_cap) return 2;
    memcpy(start, st.data(), sizeof(int) * st.size());
    memcpy(items, it.data(), sizeof(int) * it.size());
    *nitems = (int64_t)it.size();
    *G_used = d.G[0];
    for (int j = 0; j < dim; ++j) { lo[j] = d.lo[j]; inv[j] = d.inv[j]; }
    return 0;
}

// cells_per_dim: 0 = the grid is built on first use, sized for the batch (launch_sl); > 0 forces it
cudaError_t sl_create(const double *knots, int dim, int nknots, const int *dtri, int ns, const double *lumats,
                      const int *pivots, const double *bbox, const double *val, int cells_per_dim, SlDevice **out)
{
    SlDevice *d = new SlDevice();
    d->dim = dim;
    d->nknots = nknots;
    d->ns = ns;
    d->forced_G = cells_per_dim > 0 ? cells_per_dim : 0;
    const int N = dim + 1;
    cudaError_t e;
#define SL_TRY(call) do { e = (call); if (e != cudaSuccess) { sl_destroy(d); return e; } } while (0)
    SL_TRY(sl_upload(&d->knots, knots, (size_t)dim * nknots, &d->bytes));
    SL_TRY(sl_upload(&d->dtri, dtri, (size_t)N * ns, &d->bytes));
    SL_TRY(sl_upload(&d->lumats, lumats, (size_t)N * N * ns, &d->bytes));
    SL_TRY(sl_upload(&d->pivots, pivots, (size_t)N * ns, &d->bytes));
    SL_TRY(sl_upload(&d->bbox, bbox, (size_t)2 * dim * ns, &d->bytes));
    SL_TRY(sl_upload(&d->val, val, (size_t)nknots, &d->bytes));
    // first simplex (index order) holding each knot: the epol search of findsimplex :786-790
    std::vector<int> ks(nknots > 0 ? nknots : 1, -1);
    for (int s = ns - 1; s >= 0; --s)
        for (int j = 0; j < N; ++j) {
            const int k = dtri[(size_t)s * N + j] - 1;
            if (k >= 0 && k < nknots) ks[k] = s;
        }
    SL_TRY(sl_upload(&d->knot_simplex, ks.data(), (size_t)nknots, &d->bytes));
    if (dim > SL_MAXCELLDIM || ns == 0) { *out = d; return cudaSuccess; }

    sl_host_setup(d, knots, dtri, bbox);
    // packed records
    const int rs = ((2 * dim + N * N + N + 1) + 3) & ~3;
    std::vector<double> rec((size_t)ns * rs, 0.0);
    for (int s = 0; s < ns; ++s) {
        double *r = &rec[(size_t)s * rs];
        for (int i = 0; i < 2 * dim; ++i) r[i] = bbox[(size_t)s * 2 * dim + i];
        for (int i = 0; i < N * N; ++i) r[2 * dim + i] = lumats[(size_t)s * N * N + i];
        for (int v = 0; v < N; ++v) {
            const int k = dtri[(size_t)s * N + v] - 1;
            r[2 * dim + N * N + v] = (k >= 0 && k < nknots) ? val[k] : sl_na_real();
        }
        int perm[SL_MAXCELLDIM + 1];
        for (int i = 0; i < N; ++i) perm[i] = i;
        for (int i = 0; i < N; ++i) {
            int ip = pivots[(size_t)s * N + i] - 1;
            if (ip < 0 || ip >= N) ip = i;
            std::swap(perm[i], perm[ip]);
        }
        unsigned long long code = 0;
        for (int i = 0; i < N; ++i) code |= (unsigned long long)perm[i] << (4 * i);
        memcpy(&r[2 * dim + N * N + N], &code, sizeof code);
    }
    SL_TRY(sl_upload(&d->records, rec.data(), rec.size(), &d->bytes));
    d->rs = rs;
    if (d->forced_G) SL_TRY(sl_build_cells(d, d->forced_G));
#undef SL_TRY
    *out = d;
    return cudaSuccess;
}

// Cells wanted once a plan has evaluated `seen` points in total.  Building visits every (simplex,
// overlapped cell) pair on the host (133k simplices in 3-D, 8 cores: 0.07 / 0.22 / 0.97 / 2.7 s for
// 41 / 82 / 163 / 256 cells per dimension) while a point costs 0.46 / 0.3 / 0.18 / 0.15 ns to
// evaluate, so a fine grid only pays for a plan that lives long: about one cell per 64 points seen,
// between ns/4 and 128 ns (8 ns in 1-D and 2-D).  launch_sl rebuilds when this has grown 8-fold, so the total build time
// stays a fraction of the evaluation time; CHEBPOL_CUDA_OPT_SL_CELLS / CHEBPOL_CUDA_SL_CELLS pin it.
static int sl_wanted_G(const SlDevice *d, int64_t seen)
{
    double cells = (double)seen / 64.0;
    cells = std::max(cells, (double)d->ns / 4.0);
    // measured upper sizes: 3-D (133k simplices) is fastest at the finest grid tried, 126 cells per simplex;
    // 2-D (200k triangles) peaks at ~5 cells per triangle (1.62e10 evals/s; 1.45e10 at 84 per triangle)
    cells = std::min(cells, (d->dim <= 2 ? 8.0 : 128.0) * (double)d->ns);
    cells = std::min(cells, (double)(1 << 24));
    return std::max(1, (int)ceil(pow(cells, 1.0 / d->dim)));
}

cudaError_t sl_force_cells(SlDevice *d, int cells_per_dim)
{
    if (!d || d->rs <= 0) return cudaErrorNotSupported;
    d->forced_G = cells_per_dim > 0 ? cells_per_dim : 0;
    if (!d->forced_G) return cudaSuccess;
    return sl_build_cells(d, d->forced_G);
}

int64_t sl_bytes(const SlDevice *d) { return d ? d->bytes : 0; }

// kernel: 0 auto, 1 strict, 2 fast (cudaErrorNotSupported if there is no cell structure)
#define SL_DEFAULT_VARIANT 900  // per-lane state machine at 64 registers (tools/sl_sweep.sh: 8.6e9 vs 7.4e9 batched)
cudaError_t launch_sl(SlDevice *d, int kernel, int variant, int epol, int blend, const double *x, int64_t npts,
                      double *out, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    SlArgs a;
    memset(&a, 0, sizeof a);
    a.dim = d->dim; a.nknots = d->nknots; a.ns = d->ns; a.rs = d->rs; a.epol = epol; a.blend = blend;
    a.knots = d->knots; a.lumats = d->lumats; a.bbox = d->bbox; a.val = d->val; a.records = d->records;
    a.dtri = d->dtri; a.pivots = d->pivots; a.knot_simplex = d->knot_simplex;
    a.cell_start = d->cell_start; a.cell_items = d->cell_items;
    for (int j = 0; j < SL_MAXCELLDIM; ++j) { a.G[j] = d->G[j]; a.lo[j] = d->lo[j]; a.inv[j] = d->inv[j]; }
    a.x = x; a.npts = npts; a.out = out;
    const bool can_cell = d->rs > 0;
    if (kernel == CHEBPOL_CUDA_KERNEL_FAST && !can_cell) return cudaErrorNotSupported;
    bool use_cell = can_cell && kernel != CHEBPOL_CUDA_KERNEL_STRICT;
    if (use_cell && !d->forced_G) {
        // a small first batch is cheaper through the plain scan than through building the lists
        if (!d->cell_start && kernel == CHEBPOL_CUDA_KERNEL_AUTO && (d->points_seen + npts) * (int64_t)d->ns < (int64_t(1) << 32))
            use_cell = false;  // scan: ~2.5 ps per point and simplex; counted over the plan's life, not per call
        else {
            const int G = sl_wanted_G(d, d->points_seen + npts);
            if (!d->cell_start || pow((double)G, d->dim) >= 8.0 * (double)d->ncells) {
                cudaError_t be = sl_build_cells(d, G);
                if (be != cudaSuccess) {
                    (void)cudaGetLastError();
                    if (kernel == CHEBPOL_CUDA_KERNEL_FAST || !d->cell_start) {
                        if (kernel == CHEBPOL_CUDA_KERNEL_FAST) return be;
                        use_cell = false;
                    }
                }
            }
        }
    }
    d->points_seen += npts;
    a.cell_start = d->cell_start; a.cell_items = d->cell_items;
    for (int j = 0; j < SL_MAXCELLDIM; ++j) { a.G[j] = d->G[j]; a.lo[j] = d->lo[j]; a.inv[j] = d->inv[j]; }
    if (use_cell) {
        const int block = 128;
        int64_t nb = (npts + block - 1) / block;
        const int64_t cap = (int64_t)num_sms * 16;
        if (nb > cap) nb = cap;
        if (nb < 1) nb = 1;
        // variant (tuning / ncu evidence; 0: the measured default): 900 / 901 per-lane state machine at
        // 64 registers / unconstrained; batch + 100*minblocks for the point-per-thread kernel
        const int code = variant > 0 ? variant : SL_DEFAULT_VARIANT;
#define SL_LAUNCH(D, B, M) sl_cell_kernel<D, B, M><<<(int)nb, block, 0, s>>>(a)
#define SL_DIMS(B, M)                                                                                          \
    switch (d->dim) {                                                                                          \
    case 1: SL_LAUNCH(1, B, M); break;                                                                         \
    case 2: SL_LAUNCH(2, B, M); break;                                                                         \
    case 3: SL_LAUNCH(3, B, M); break;                                                                         \
    case 4: SL_LAUNCH(4, B, M); break;                                                                         \
    case 5: SL_LAUNCH(5, B, M); break;                                                                         \
    default: SL_LAUNCH(6, B, M); break;                                                                        \
    }
#define SL_SM(M)                                                                                               \
    switch (d->dim) {                                                                                          \
    case 1: sl_cell_sm_kernel<1, M><<<(int)nb, block, 0, s>>>(a); break;                                       \
    case 2: sl_cell_sm_kernel<2, M><<<(int)nb, block, 0, s>>>(a); break;                                       \
    case 3: sl_cell_sm_kernel<3, M><<<(int)nb, block, 0, s>>>(a); break;                                       \
    case 4: sl_cell_sm_kernel<4, M><<<(int)nb, block, 0, s>>>(a); break;                                       \
    case 5: sl_cell_sm_kernel<5, M><<<(int)nb, block, 0, s>>>(a); break;                                       \
    default: sl_cell_sm_kernel<6, M><<<(int)nb, block, 0, s>>>(a); break;                                      \
    }
        switch (code) {
        case 900: SL_SM(8); break;
        case 901: SL_SM(1); break;
        case 1: SL_DIMS(1, 1); break;
        case 2: SL_DIMS(2, 1); break;
        case 8: SL_DIMS(8, 1); break;
        case 802: SL_DIMS(2, 8); break;
        case 804: SL_DIMS(4, 8); break;
        case 4: SL_DIMS(4, 1); break;
        default: SL_DIMS(4, 8); break;
        }
#undef SL_DIMS
#undef SL_SM
#undef SL_LAUNCH
        if (li) { li->grid = (int)nb; li->block = block; li->smem = 0; li->kernel = K_SL_CELL; }
        return cudaGetLastError();
    }
    const int block = 256;
    int64_t nb = (npts + block - 1) / block;
    if (nb > (int64_t)num_sms * 8) nb = (int64_t)num_sms * 8;
    if (nb < 1) nb = 1;
    sl_strict_kernel<<<(int)nb, block, 0, s>>>(a);
    if (li) { li->grid = (int)nb; li->block = block; li->smem = 0; li->kernel = K_SL_STRICT; }
    return cudaGetLastError();
}
