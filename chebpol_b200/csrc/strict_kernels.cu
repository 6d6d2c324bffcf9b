// strict_kernels.cu -- the "reference order" kernels: one thread per point,
// exactly the floating-point operation sequence of the reference C, and this
// TU is compiled with -fmad=false so no multiply-add is fused.  Results are
// bit-identical to /root/reference/src/chebpol.c built by gcc -O2 (x86-64, no
// FMA), except where libm is involved (exp in blends 1/2, sin in the uniform
// map: <= 1 ulp apart).  They accept any rank <= 31 and any dims, and are the
// parity anchor for, and the fallback of, the fast kernels.
#include "common.cuh"
#include <float.h>

// ---------------------------------------------------------------- Chebyshev
// C_evalcheb, src/chebpol.c:314-338 (see oracle/chebpol_oracle.c for the
// iterative form used here).
__device__ __forceinline__ double strict_cheb_row(const double *__restrict__ c, int n, double x0)
{
    const double two_x = 2.0 * x0;
    double b = 0.0, b1 = 0.0, b2;
    for (int i = n - 1; i >= 1; --i) {
        b2 = b1;
        b1 = b;
        b = two_x * b1 - b2 + __ldg(c + i);
    }
    return x0 * b - b1 + __ldg(c);
}

__global__ void __launch_bounds__(256) cheb_strict_kernel(const __grid_constant__ ChebStrictArgs a)
{
    const int rank = a.rank;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < a.npts;
         p += (int64_t)gridDim.x * blockDim.x) {
        double x[CP_MAXR];
        for (int d = 0; d < rank; ++d) x[d] = premap_coord(a.pm, d, a.x[p * rank + d]);
        double res;
        if (rank == 1) {
            res = strict_cheb_row(a.cf, a.dims[0], x[0]);
        } else {
            int idx[CP_MAXR];
            double b[CP_MAXR], b1[CP_MAXR];
            int64_t off = 0;
            for (int l = 1; l < rank; ++l) {
                idx[l] = a.dims[l] - 1;
                b[l] = b1[l] = 0.0;
                off += (int64_t)idx[l] * a.stride[l];
            }
            for (;;) {
                double v = strict_cheb_row(a.cf + off, a.dims[0], x[0]);
                int l = 1;
                bool done = false;
                for (;;) {
                    const double b2 = b1[l];
                    b1[l] = b[l];
                    b[l] = 2.0 * x[l] * b1[l] - b2 + v;
                    if (idx[l] > 0) {
                        --idx[l];
                        off -= a.stride[l];
                        break;
                    }
                    v = b[l] - x[l] * b1[l];
                    b[l] = b1[l] = 0.0;
                    idx[l] = a.dims[l] - 1;
                    off += (int64_t)idx[l] * a.stride[l];
                    if (++l == rank) { done = true; break; }
                }
                if (done) { res = v; break; }
            }
        }
        a.out[p] = res;
    }
}

cudaError_t launch_cheb_strict(const ChebStrictArgs &a, cudaStream_t s, LaunchInfo *li)
{
    const int block = 256;
    int64_t nb = (a.npts + block - 1) / block;
    if (nb > 148 * 64) nb = 148 * 64;
    if (nb < 1) nb = 1;
    cheb_strict_kernel<<<(int)nb, block, 0, s>>>(a);
    if (li) { li->grid = (int)nb; li->block = block; li->smem = 0; li->kernel = K_CHEB_STRICT; }
    return cudaGetLastError();
}

// -------------------------------------------------------------- multilinear
// binsearch src/chebpol.c:543-556 (NaN-terminating), blendfun
// src/chebpol.h:24-45, C_evalmlip src/chebpol.c:617-667.
__device__ __forceinline__ int strict_cell_upper(int n, const double *__restrict__ arr, double val)
{
    int lo = 0, hi = n - 1;
    while (lo + 1 < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(arr + mid) >= val) hi = mid;
        else lo = mid;
    }
    return hi;
}

__device__ __forceinline__ double strict_blend(double w, int blend)
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

__global__ void __launch_bounds__(256) mlip_strict_kernel(const __grid_constant__ MlipArgs a)
{
    const int rank = a.rank;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < a.npts;
         p += (int64_t)gridDim.x * blockDim.x) {
        double w[CP_MAXR];
        int64_t top = 0;
        bool nan_in = false;
        for (int d = 0; d < rank; ++d) {
            const double xd = a.x[p * rank + d];
            nan_in |= (xd != xd);
            const double *g = a.grid + a.goff[d];
            const int gp = strict_cell_upper(a.dims[d], g, xd);
            w[d] = (xd - __ldg(g + gp - 1)) / (__ldg(g + gp) - __ldg(g + gp - 1));
            top += a.stride[d] * gp;
        }
        double wsum = 0.0, acc = 0.0;
        const unsigned long ncorner = 1ul << rank;
        for (unsigned long corner = 0; corner < ncorner; ++corner) {
            int64_t pos = top;
            double cw = 1;
            for (int d = 0; d < rank; ++d) {
                if (corner & (1ul << d)) {
                    cw *= strict_blend(w[d], a.blend);
                } else {
                    cw *= strict_blend(1 - w[d], a.blend);
                    pos -= a.stride[d];
                }
            }
            wsum += cw;
            acc += cw * __ldg(a.values + pos);
        }
        double r = acc / wsum;
        // blends 4/5 ignore the weight, so a NaN coordinate would not propagate
        if (nan_in) r = __longlong_as_double(0x7ff8000000000000LL);
        a.out[p] = r;
    }
}

cudaError_t launch_mlip_strict(const MlipArgs &a, cudaStream_t s, LaunchInfo *li)
{
    const int block = 256;
    int64_t nb = (a.npts + block - 1) / block;
    if (nb > 148 * 64) nb = 148 * 64;
    if (nb < 1) nb = 1;
    mlip_strict_kernel<<<(int)nb, block, 0, s>>>(a);
    if (li) { li->grid = (int)nb; li->block = block; li->smem = 0; li->kernel = K_MLIP_STRICT; }
    return cudaGetLastError();
}

// ---------------------------------------------------------- Floater-Hormann
// FH, src/chebpol.c:182-218, as the explicit-frame iteration of
// oracle/chebpol_oracle.c (fh_point).
__global__ void __launch_bounds__(256) fh_strict_kernel(const __grid_constant__ FhArgs a)
{
    const int rank = a.rank;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < a.npts;
         p += (int64_t)gridDim.x * blockDim.x) {
        double x[CP_MAXR];
        for (int d = 0; d < rank; ++d) x[d] = a.x[p * rank + d];
        int i[CP_MAXR], hit[CP_MAXR];
        double num[CP_MAXR], den[CP_MAXR];
        int64_t base[CP_MAXR];
        int l = rank - 1;
        base[l] = 0;
        double ret = 0.0;
        bool entering = true;
        for (;;) {
            if (entering) {
                const double xx = x[l];
                const double *kn = a.grid + a.goff[l];
                hit[l] = -1;
                for (int k = 0; k < a.dims[l]; ++k)
                    if (fabs(xx - __ldg(kn + k)) < 10.0 * DBL_EPSILON) { hit[l] = k; break; }
                num[l] = den[l] = 0.0;
                i[l] = (hit[l] >= 0) ? hit[l] : 0;
                if (l == 0) {
                    const double *fv = a.vals + base[0];
                    if (hit[0] >= 0) {
                        ret = __ldg(fv + hit[0]);
                    } else {
                        double nu = 0, de = 0;
                        const double *w = a.weights + a.goff[0];
                        for (int k = 0; k < a.dims[0]; ++k) {
                            const double val = __ldg(fv + k);
                            const double pole = __ldg(w + k) / (xx - __ldg(kn + k));
                            nu += pole * val;
                            de += pole;
                        }
                        ret = nu / de;
                    }
                    entering = false;
                    ++l;
                    if (l == rank) break;
                    continue;
                }
                base[l - 1] = base[l] + (int64_t)i[l] * a.stride[l];
                --l;
                continue;
            }
            if (hit[l] < 0) {
                const double pole =
                    __ldg(a.weights + a.goff[l] + i[l]) / (x[l] - __ldg(a.grid + a.goff[l] + i[l]));
                num[l] += pole * ret;
                den[l] += pole;
                if (++i[l] < a.dims[l]) {
                    base[l - 1] = base[l] + (int64_t)i[l] * a.stride[l];
                    --l;
                    entering = true;
                    continue;
                }
                ret = num[l] / den[l];
            }
            ++l;
            if (l == rank) break;
        }
        a.out[p] = ret;
    }
}

cudaError_t launch_fh_strict(const FhArgs &a, cudaStream_t s, LaunchInfo *li)
{
    const int block = 256;
    int64_t nb = (a.npts + block - 1) / block;
    if (nb > 148 * 64) nb = 148 * 64;
    if (nb < 1) nb = 1;
    fh_strict_kernel<<<(int)nb, block, 0, s>>>(a);
    if (li) { li->grid = (int)nb; li->block = block; li->smem = 0; li->kernel = K_FH_STRICT; }
    return cudaGetLastError();
}

// ---------------------------------------------------------- polyharmonic spline
// C_evalpolyh, src/chebpol.c:383-415, operation for operation (this TU is compiled without
// FMA contraction): squared distance summed over the coordinates in order, knots in order,
// knots at distance zero skipped for k >= 0, then the affine part.  r^k is R's R_pow_di
// (repeated squaring).  sqrt is IEEE; exp and log are CUDA's (<= 1 ulp), so k odd is
// bit-identical to the reference and k even / k < 0 agree to ~1e-15 relative per term.
__device__ __forceinline__ double pow_di_strict(double x, int n)
{
    double xn = 1.0;
    if (x != x) return x;
    if (n != 0) {
        if (isinf(x)) return pow(x, (double)n);
        const bool neg = n < 0;
        if (neg) n = -n;
        for (;;) {
            if (n & 1) xn *= x;
            n >>= 1;
            if (!n) break;
            x *= x;
        }
        if (neg) xn = 1.0 / xn;
    }
    return xn;
}

__global__ void __launch_bounds__(256) polyh_strict_kernel(const __grid_constant__ PolyhArgs a)
{
    const int rank = a.rank;
    const int ki = (int)a.k;
    const int mode = (a.k < 0) ? 0 : ((ki % 2 == 1) ? 1 : 2);
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < a.npts;
         p += (int64_t)gridDim.x * blockDim.x) {
        double x[CP_MAXR];
        for (int d = 0; d < rank; ++d) x[d] = premap_coord(a.pm, d, a.x[p * rank + d]);
        double wsum = 0.0;
        for (int i = 0; i < a.nknots; ++i) {
            const double *kn = a.knots + (size_t)i * rank;
            double sqdist = 0.0;
            for (int j = 0; j < rank; ++j) {
                const double t = x[j] - __ldg(kn + j);
                sqdist += t * t;
            }
            if (mode == 0) {
                wsum += __ldg(a.weights + i) * exp(a.k * sqdist);
            } else if (sqdist != 0) {
                const double r = sqrt(sqdist);
                if (mode == 1) wsum += __ldg(a.weights + i) * pow_di_strict(r, ki);
                else wsum += __ldg(a.weights + i) * log(r) * pow_di_strict(r, ki);
            }
        }
        wsum += __ldg(a.lweights);
        for (int j = 0; j < rank; ++j) wsum += __ldg(a.lweights + j + 1) * x[j];
        a.out[p] = wsum;
    }
}

cudaError_t launch_polyh_strict(const PolyhArgs &a, cudaStream_t s, LaunchInfo *li)
{
    const int block = 256;
    int64_t nb = (a.npts + block - 1) / block;
    if (nb > 148 * 64) nb = 148 * 64;
    if (nb < 1) nb = 1;
    polyh_strict_kernel<<<(int)nb, block, 0, s>>>(a);
    if (li) { li->grid = (int)nb; li->block = block; li->smem = 0; li->kernel = K_POLYH_STRICT; }
    return cudaGetLastError();
}

// ---------------------------------------------------------- stalker splines
// evalstalk (src/stalker.c:366-442) and evalhyp (:537-613), operation for operation.  The cell
// is R's findInterval(rightmost_closed = all_inside = TRUE): knots <= x counted by bisection,
// clamped, so g[gp] <= x < g[gp+1] inside the grid.  evalhyp uses only + - * /: bit-identical
// to the reference; evalstalk calls pow (device pow within 2 ulp of libm).
__device__ __forceinline__ int strict_find_interval0(const double *__restrict__ g, int n, double x)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) / 2;
        if (__ldg(g + mid) <= x) lo = mid + 1; else hi = mid;
    }
    if (lo < 1) lo = 1;
    if (lo > n - 1) lo = n - 1;
    return lo - 1;
}

template <int KIND>
__global__ void __launch_bounds__(256) stalk_strict_kernel(const __grid_constant__ StalkArgs a)
{
    const int rank = a.rank;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < a.npts;
         p += (int64_t)gridDim.x * blockDim.x) {
        double weight[CP_MAXR], x[CP_MAXR], zx[CP_MAXR];
        int gbase[CP_MAXR];
        int64_t llcorner = 0;
        for (int g = 0; g < rank; ++g) {
            x[g] = a.x[p * rank + g];
            const double *gr = a.grid + a.goff[g];
            const int gp = strict_find_interval0(gr, a.dims[g], x[g]);
            gbase[g] = gp;
            double w = 1 - (x[g] - __ldg(gr + gp)) / (__ldg(gr + gp + 1) - __ldg(gr + gp));
            if (w < 0) w = 0;
            if (w > 1) w = 1;
            weight[g] = w;
            llcorner += a.stride[g] * gp;
        }
        double V = 0.0, sumw = 0.0;
        const unsigned long ncorner = 1ul << rank;
        for (unsigned long i = 0; i < ncorner; ++i) {
            int64_t corner = llcorner;
            double cw = 1.0;
            for (int g = 0; g < rank; ++g) {
                const double *gr = a.grid + a.goff[g];
                if (i & (1ul << g)) {
                    cw *= strict_blend(1 - weight[g], a.blend);
                    corner += a.stride[g];
                    zx[g] = x[g] - __ldg(gr + gbase[g] + 1);
                } else {
                    zx[g] = x[g] - __ldg(gr + gbase[g]);
                    cw *= strict_blend(weight[g], a.blend);
                }
            }
            if (cw == 0.0) continue;
            const double *bb = a.b + rank * corner, *p1 = a.t1 + rank * corner, *p2 = a.t2 + rank * corner;
            double fval;
            if (KIND == 0) {
                fval = __ldg(a.val + corner);
                for (int j = 0; j < rank; ++j) {
                    const double cj = __ldg(p1 + j);
                    if (cj == 0) fval += __ldg(bb + j) * zx[j];
                    else fval += __ldg(bb + j) * zx[j] + cj * pow(fabs(zx[j]), __ldg(p2 + j));
                }
            } else {
                fval = __ldg(a.val + corner) + __ldg(a.a + corner);
                for (int j = 0; j < rank; ++j) {
                    const double cj = __ldg(p1 + j);
                    if (cj == cj) fval += __ldg(bb + j) * zx[j] + __ldg(p2 + j) / (1.0 + cj * zx[j]);
                    else fval += __ldg(bb + j) * zx[j] * zx[j];
                }
            }
            sumw += cw;
            V += cw * fval;
        }
        a.out[p] = V / sumw;
    }
}

cudaError_t launch_stalk_strict(const StalkArgs &a, cudaStream_t s, LaunchInfo *li)
{
    const int block = 256;
    int64_t nb = (a.npts + block - 1) / block;
    if (nb > 148 * 64) nb = 148 * 64;
    if (nb < 1) nb = 1;
    if (a.kind == 0) stalk_strict_kernel<0><<<(int)nb, block, 0, s>>>(a);
    else stalk_strict_kernel<1><<<(int)nb, block, 0, s>>>(a);
    if (li) { li->grid = (int)nb; li->block = block; li->smem = 0; li->kernel = K_STALK_STRICT; }
    return cudaGetLastError();
}
