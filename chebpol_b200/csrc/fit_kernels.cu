// fit_kernels.cu -- the fit step next to the evaluation path:
//   chebcoef   values on a Chebyshev grid -> Chebyshev coefficients   (src/chebpol.c:8-141)
//   fhweights  Floater-Hormann weights of a knot vector, eq. (18)      (src/chebpol.c:267-282)
// chebcoef is a separable DCT-II: per dimension d, with the tensor viewed as
// (inner, N = n_d, outer),  out[a,k,b] = alpha * sum_i cos(pi*k*(i+1/2)/N) * in[a,i,b],
// alpha = 2/N (2 when dct), then every index-0 hyperplane is halved (unless dct).  This is
// the reference's matrix path (:65-140; its FFTW path computes the same numbers).  A CTA
// owns one (b, k-tile) and stages the needed rows of the cosine matrix in shared memory; the
// matrix is generated on the device with cospi (exact at multiples of 1/2, like R's cospi).
#include "common.cuh"

namespace {

constexpr int kTileK = 8;

// Mode product along one dimension of a tensor viewed as (inner, N, outer):
//   out[a, k, b] = alpha * sum_i Mx[k][i] * in[a, i, b],   k < M
// One thread per output element.  Mx: [M][N] row-major, or null for the DCT-II cosine
// matrix generated on the fly (very long 1-D transforms).
__global__ void __launch_bounds__(256) mode_product_kernel(const double *__restrict__ in, double *__restrict__ out,
                                                           const double *__restrict__ Mx, int64_t inner, int N, int M,
                                                           int64_t outer, double alpha)
{
    const int64_t total = inner * (int64_t)M * outer;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t a = idx % inner;
        const int k = (int)((idx / inner) % M);
        const int64_t b = idx / (inner * M);
        const double *src = in + a + b * inner * N;
        double s = 0.0;
        if (Mx) {
            const double *c = Mx + (size_t)k * N;
            for (int i = 0; i < N; ++i) s = fma(__ldg(c + i), __ldg(src + (int64_t)i * inner), s);
        } else {
            for (int i = 0; i < N; ++i) s = fma(cospi((double)k * ((double)i + 0.5) / (double)N), __ldg(src + (int64_t)i * inner), s);
        }
        out[idx] = alpha * s;
    }
}

// ---- basis matrices B[m][n] of one dimension at m grid coordinates (evalongridV) -------
// kind 0: Chebyshev T_i(x) after the pre-map; 1: Floater-Hormann barycentric basis (one-hot on
// a knot hit, src/chebpol.c:195-197); 2: multilinear hat weights with blending, normalised
// per dimension (the reference's sum of corner weights factorises over the dimensions).
__device__ __forceinline__ double grid_blend(double w, int blend)
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

__global__ void basis_matrix_kernel(int kind, double *__restrict__ B, const double *__restrict__ pts, int m, int n,
                                    const double *__restrict__ knots, const double *__restrict__ weights, int blend,
                                    int pm_mode, double mid, double ispan, double nd)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    double *row = B + (size_t)j * n;
    double x = pts[j];
    if (kind == 0) {
        if (pm_mode & 1) x = (x - mid) * ispan;
        if (pm_mode & 2) x = sin(0.5 * 3.141592653589793 * x * (1.0 - nd) / nd);
        const double tx = 2.0 * x;
        double tkm1 = 1.0, tk = x;
        row[0] = 1.0;
        if (n > 1) row[1] = x;
        for (int i = 2; i < n; ++i) {
            const double t = fma(tx, tk, -tkm1);
            tkm1 = tk;
            tk = t;
            row[i] = t;
        }
    } else if (kind == 1) {
        int hit = -1;
        for (int i = 0; i < n; ++i)
            if (fabs(x - knots[i]) < 10.0 * 2.220446049250313e-16) { hit = i; break; }
        if (hit >= 0 || x != x) {
            for (int i = 0; i < n; ++i) row[i] = (i == hit) ? 1.0 : ((x != x) ? x : 0.0);
            return;
        }
        double den = 0.0;
        for (int i = 0; i < n; ++i) {
            const double p = weights[i] / (x - knots[i]);
            row[i] = p;
            den += p;
        }
        for (int i = 0; i < n; ++i) row[i] /= den;
    } else {
        for (int i = 0; i < n; ++i) row[i] = 0.0;
        int lo = 0, hi = n - 1;  // src/chebpol.c:543-556, NaN-safe
        while (lo + 1 < hi) {
            const int mdl = (lo + hi) >> 1;
            if (knots[mdl] >= x) hi = mdl;
            else lo = mdl;
        }
        const double w = (x - knots[hi - 1]) / (knots[hi] - knots[hi - 1]);
        const double bh = grid_blend(w, blend), bl = grid_blend(1 - w, blend);
        const double sum = bh + bl;
        row[hi] = bh / sum;
        row[hi - 1] = bl / sum;
        if (x != x) row[hi] = row[hi - 1] = x;
    }
}

__global__ void cos_matrix_kernel(double *cosm, int N)
{
    const int64_t total = (int64_t)N * N;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(idx / N), i = (int)(idx % N);
        cosm[idx] = cospi((double)k * ((double)i + 0.5) / (double)N);
    }
}

// halve the index-0 hyperplane of dimension d
__global__ void halve_plane_kernel(double *F, int64_t inner, int N, int64_t outer)
{
    const int64_t total = inner * outer;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t a = idx % inner, b = idx / inner;
        F[a + b * inner * N] *= 0.5;
    }
}

__global__ void fhweights_kernel(const double *__restrict__ grid, int n /* knots - 1 */, int d, double *__restrict__ w)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > n) return;
    const int start = (k < d) ? 0 : k - d, end = (k < n - d) ? k : n - d;
    double sum = 0.0;
    for (int i = start; i <= end; ++i) {
        double prod = 1.0;
        for (int j = i; j <= i + d; ++j)
            if (j != k) prod *= grid[k] - grid[j];
        sum += 1.0 / fabs(prod);
    }
    w[k] = sum * ((k % 2 != d % 2) ? -1.0 : 1.0);
}

int blocks_for(int64_t total)
{
    int64_t nb = (total + 255) / 256;
    if (nb > 148 * 16) nb = 148 * 16;
    return (int)(nb < 1 ? 1 : nb);
}

}  // namespace

// d_a holds the values (overwritten), d_b is scratch of the same size; returns the buffer
// that holds the result in *result.  launches counts kernels issued.
cudaError_t fit_chebcoef_device(double *d_a, double *d_b, const int *dims, int rank, int dct, cudaStream_t s,
                                double **result, int *launches)
{
    int64_t siz = 1;
    for (int d = 0; d < rank; ++d) siz *= dims[d];
    double *src = d_a, *dst = d_b;
    int64_t inner = 1;
    *launches = 0;
    for (int d = 0; d < rank; ++d) {
        const int N = dims[d];
        const int64_t outer = siz / (inner * N);
        double *cosm = nullptr;
        if (N <= 1024) {
            cudaError_t e = cudaMalloc(&cosm, sizeof(double) * N * N);
            if (e != cudaSuccess) return e;
            cos_matrix_kernel<<<blocks_for((int64_t)N * N), 256, 0, s>>>(cosm, N);
            ++*launches;
        }
        mode_product_kernel<<<blocks_for(siz), 256, 0, s>>>(src, dst, cosm, inner, N, N, outer, dct ? 2.0 : 2.0 / N);
        ++*launches;
        if (!dct) {
            halve_plane_kernel<<<blocks_for(inner * outer), 256, 0, s>>>(dst, inner, N, outer);
            ++*launches;
        }
        if (cosm) {
            cudaError_t e = cudaStreamSynchronize(s);
            cudaFree(cosm);
            if (e != cudaSuccess) return e;
        }
        double *t = src; src = dst; dst = t;
        inner *= N;
    }
    *result = src;
    return cudaGetLastError();
}

cudaError_t fit_fhweights_device(const double *d_grid, int nknots, int d, double *d_w, cudaStream_t s)
{
    fhweights_kernel<<<(nknots + 127) / 128, 128, 0, s>>>(d_grid, nknots - 1, d, d_w);
    return cudaGetLastError();
}

// Evaluate a tensor-product interpolant on the tensor grid pts[0] x ... x pts[rank-1]:
// out = T x_0 B_0 x_1 B_1 ... with the per-dimension basis matrices above.  d_tensor is the
// interpolant's tensor (R layout, left untouched); the result (R array of dims gdims) is
// returned in *result, which is one of the two scratch buffers.
cudaError_t fit_evalgrid_device(int kind, const double *d_tensor, const int *dims, int rank, const double *const *d_pts,
                                const int *gdims, const double *const *d_knots, const double *const *d_weights, int blend,
                                const PreMap *pm, double *d_s0, double *d_s1, cudaStream_t s, double **result, int *launches)
{
    *launches = 0;
    const double *src = d_tensor;
    double *bufs[2] = {d_s0, d_s1};
    int which = 0;
    int64_t inner = 1;  // product of the grid sizes of the dims already done
    int64_t outer = 1;
    for (int d = 0; d < rank; ++d) outer *= dims[d];
    for (int d = 0; d < rank; ++d) {
        const int N = dims[d], M = gdims[d];
        outer /= N;
        double *B = nullptr;
        cudaError_t e = cudaMalloc(&B, sizeof(double) * (size_t)M * N);
        if (e != cudaSuccess) return e;
        basis_matrix_kernel<<<(M + 127) / 128, 128, 0, s>>>(kind, B, d_pts[d], M, N, d_knots ? d_knots[d] : nullptr,
                                                           d_weights ? d_weights[d] : nullptr, blend, pm ? pm->mode : 0,
                                                           pm ? pm->mid[d] : 0.0, pm ? pm->ispan[d] : 1.0, pm ? pm->nd[d] : 1.0);
        double *dst = bufs[which];
        mode_product_kernel<<<blocks_for(inner * M * outer), 256, 0, s>>>(src, dst, B, inner, N, M, outer, 1.0);
        *launches += 2;
        e = cudaStreamSynchronize(s);
        cudaFree(B);
        if (e != cudaSuccess) return e;
        src = dst;
        which ^= 1;
        inner *= M;
    }
    *result = const_cast<double *>(src);
    return cudaGetLastError();
}
