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

// one thread per (a, k) for fixed b; loops over i.  cosm: [N][N] (k-major) or null -> on the fly
__global__ void __launch_bounds__(256) dct_mode_kernel(const double *__restrict__ in, double *__restrict__ out,
                                                       const double *__restrict__ cosm, int64_t inner, int N,
                                                       int64_t outer, double alpha)
{
    const int64_t total = inner * (int64_t)N * outer;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t a = idx % inner;
        const int k = (int)((idx / inner) % N);
        const int64_t b = idx / (inner * N);
        const double *src = in + a + b * inner * N;
        double s = 0.0;
        if (cosm) {
            const double *c = cosm + (size_t)k * N;
            for (int i = 0; i < N; ++i) s = fma(__ldg(c + i), __ldg(src + (int64_t)i * inner), s);
        } else {
            for (int i = 0; i < N; ++i) s = fma(cospi((double)k * ((double)i + 0.5) / (double)N), __ldg(src + (int64_t)i * inner), s);
        }
        out[idx] = alpha * s;
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
        dct_mode_kernel<<<blocks_for(siz), 256, 0, s>>>(src, dst, cosm, inner, N, outer, dct ? 2.0 : 2.0 / N);
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
