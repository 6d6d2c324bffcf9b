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
    int64_t bytes = 0, nitems = 0;
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
__device__ __forceinline__ bool sl_try(const double *__restrict__ rec, const double (&x)[DIM], int blend, double *res)
{
    constexpr int N = DIM + 1;
#pragma unroll
    for (int i = 0; i < DIM; ++i) {
        const double2 bx = __ldg(reinterpret_cast<const double2 *>(rec) + i);
        if (x[i] < bx.x || x[i] > bx.y) return false;
    }
    double lu[N * N];
#pragma unroll
    for (int i = 0; i < N * N; ++i) lu[i] = __ldg(rec + 2 * DIM + i);
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

template <int DIM>
__global__ void __launch_bounds__(128) sl_cell_kernel(const __grid_constant__ SlArgs a)
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
            // a NaN coordinate passes every box test of the reference; the scan is the reference's
            // own, restricted to the finite coordinates, so fall back to it for this rare point
            for (int s = 0; s < a.ns && !found; ++s) found = sl_try<DIM>(a.records + (int64_t)s * a.rs, x, a.blend, &res);
        } else {
            int cell = 0;
#pragma unroll
            for (int d = DIM - 1; d >= 0; --d) cell = cell * a.G[d] + sl_cell_of(x[d], a.lo[d], a.inv[d], a.G[d]);
            const int beg = __ldg(a.cell_start + cell), end = __ldg(a.cell_start + cell + 1);
            for (int it = beg; it < end && !found; ++it) {
                const int s = __ldg(a.cell_items + it);
                found = sl_try<DIM>(a.records + (int64_t)s * a.rs, x, a.blend, &res);
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

// cells_per_dim: 0 = automatic (about 4 cells per simplex), > 0 forces the resolution
cudaError_t sl_create(const double *knots, int dim, int nknots, const int *dtri, int ns, const double *lumats,
                      const int *pivots, const double *bbox, const double *val, int cells_per_dim, SlDevice **out)
{
    SlDevice *d = new SlDevice();
    d->dim = dim;
    d->nknots = nknots;
    d->ns = ns;
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

    // bounding box of all simplices and the cell grid over it
    for (int j = 0; j < dim; ++j) { d->glo[j] = 1e300; d->ghi[j] = -1e300; }
    for (int s = 0; s < ns; ++s)
        for (int j = 0; j < dim; ++j) {
            d->glo[j] = std::min(d->glo[j], bbox[(size_t)s * 2 * dim + 2 * j]);
            d->ghi[j] = std::max(d->ghi[j], bbox[(size_t)s * 2 * dim + 2 * j + 1]);
        }
    int G = cells_per_dim > 0 ? cells_per_dim : (int)ceil(pow(4.0 * ns, 1.0 / dim));
    std::vector<int> start, items;
    for (;; G = (G + 1) / 2) {
        if (G < 1) G = 1;
        double ncell = pow((double)G, dim);
        if (ncell > (double)(1 << 24) && G > 1) continue;
        for (int j = 0; j < dim; ++j) {
            d->G[j] = G;
            d->lo[j] = d->glo[j];
            const double span = d->ghi[j] - d->glo[j];
            d->inv[j] = (span > 0.0 && std::isfinite(span)) ? (double)G / span : 0.0;
        }
        const int64_t nc = (int64_t)ncell;
        std::vector<int64_t> count(nc + 1, 0);
        int c0[SL_MAXCELLDIM], c1[SL_MAXCELLDIM], c[SL_MAXCELLDIM];
        auto range = [&](int s) {
            for (int j = 0; j < dim; ++j) {
                c0[j] = sl_cell_of(bbox[(size_t)s * 2 * dim + 2 * j], d->lo[j], d->inv[j], G);
                c1[j] = sl_cell_of(bbox[(size_t)s * 2 * dim + 2 * j + 1], d->lo[j], d->inv[j], G);
                if (c1[j] < c0[j]) c1[j] = c0[j];  // NaN box: bin it somewhere, the box test rejects it anyway
            }
        };
        auto for_cells = [&](auto &&fn) {
            for (int j = 0; j < dim; ++j) c[j] = c0[j];
            for (;;) {
                int64_t id = 0;
                for (int j = dim - 1; j >= 0; --j) id = id * G + c[j];
                fn(id);
                int j = 0;
                for (; j < dim; ++j) {
                    if (++c[j] <= c1[j]) break;
                    c[j] = c0[j];
                }
                if (j == dim) break;
            }
        };
        int64_t total = 0;
        for (int s = 0; s < ns; ++s) {
            range(s);
            for_cells([&](int64_t id) { ++count[id + 1]; ++total; });
        }
        if (total > (int64_t(1) << 27) && G > 1 && cells_per_dim <= 0) continue;  // too fine: coarsen
        for (int64_t i = 0; i < nc; ++i) count[i + 1] += count[i];
        start.assign(count.begin(), count.end());
        items.assign((size_t)total, 0);
        std::vector<int64_t> fill(count.begin(), count.end() - 1);
        for (int s = 0; s < ns; ++s) {  // ascending s: every list comes out in index order
            range(s);
            for_cells([&](int64_t id) { items[(size_t)fill[id]++] = s; });
        }
        d->nitems = total;
        break;
    }
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
    SL_TRY(sl_upload(&d->cell_start, start.data(), start.size(), &d->bytes));
    SL_TRY(sl_upload(&d->cell_items, items.data(), items.size(), &d->bytes));
    d->rs = rs;
#undef SL_TRY
    *out = d;
    return cudaSuccess;
}

int64_t sl_bytes(const SlDevice *d) { return d ? d->bytes : 0; }

// kernel: 0 auto, 1 strict, 2 fast (cudaErrorNotSupported if there is no cell structure)
cudaError_t launch_sl(const SlDevice *d, int kernel, int epol, int blend, const double *x, int64_t npts, double *out,
                      int num_sms, cudaStream_t s, LaunchInfo *li)
{
    SlArgs a;
    memset(&a, 0, sizeof a);
    a.dim = d->dim; a.nknots = d->nknots; a.ns = d->ns; a.rs = d->rs; a.epol = epol; a.blend = blend;
    a.knots = d->knots; a.lumats = d->lumats; a.bbox = d->bbox; a.val = d->val; a.records = d->records;
    a.dtri = d->dtri; a.pivots = d->pivots; a.knot_simplex = d->knot_simplex;
    a.cell_start = d->cell_start; a.cell_items = d->cell_items;
    for (int j = 0; j < SL_MAXCELLDIM; ++j) { a.G[j] = d->G[j]; a.lo[j] = d->lo[j]; a.inv[j] = d->inv[j]; }
    a.x = x; a.npts = npts; a.out = out;
    const bool cell_ok = d->rs > 0;
    if (kernel == CHEBPOL_CUDA_KERNEL_FAST && !cell_ok) return cudaErrorNotSupported;
    if (kernel != CHEBPOL_CUDA_KERNEL_STRICT && cell_ok) {
        const int block = 128;
        int64_t nb = (npts + block - 1) / block;
        const int64_t cap = (int64_t)num_sms * 16;
        if (nb > cap) nb = cap;
        if (nb < 1) nb = 1;
        switch (d->dim) {
        case 1: sl_cell_kernel<1><<<(int)nb, block, 0, s>>>(a); break;
        case 2: sl_cell_kernel<2><<<(int)nb, block, 0, s>>>(a); break;
        case 3: sl_cell_kernel<3><<<(int)nb, block, 0, s>>>(a); break;
        case 4: sl_cell_kernel<4><<<(int)nb, block, 0, s>>>(a); break;
        case 5: sl_cell_kernel<5><<<(int)nb, block, 0, s>>>(a); break;
        default: sl_cell_kernel<6><<<(int)nb, block, 0, s>>>(a); break;
        }
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
