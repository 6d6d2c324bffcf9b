// mlip_quad.cu -- multilinear evaluation from a row-pair-interleaved table (sm_100a).
//
// C_evalmlip (src/chebpol.c:617-667) gathers the 2^D corners of a cell from the value array.  In
// the R layout they are 2^(D-1) pairs adjacent along dim 0, every pair in a different 128-byte
// line: the gather costs 2^(D-1) L1 wavefronts and, when the table is larger than L2, as many
// missed sectors per point (mlip_pair: 426 B/point of DRAM traffic on the 64^4 table against
// 168 algorithmic, profiles/ncu_mlip_pair_cfg3_r01.txt).  The cell-major table (mlip_cell) fixes
// that with 2^D times the memory.  This kernel is the one in between, for tables whose 2^D-fold
// expansion does not fit and for batches too small to pay for it: TWICE the memory.
//
// Layout Q ("quads"): for every index r of the dims >= 2 and every PAIR of neighbouring dim-1 rows
// (i1, i1+1), the two rows interleaved element by element,
//     Q[((r*(n1-1) + i1)*n0 + i0)*2 + j] = V[i0, i1+j, r],          2*(n1-1)/n1 of the table,
// so the four corners of a cell in dims (0,1) -- V[c0,c1], V[c0,c1+1], V[c0+1,c1], V[c0+1,c1+1] --
// are 32 contiguous, 16-byte aligned bytes: a point reads 2^(D-2) quads instead of 2^(D-1) pairs,
// half the L1 wavefronts and 1.25 sixty-four-byte DRAM bursts per 32 useful bytes instead of 2.25
// per two pairs.  (B200's L2 fetches 64 bytes from HBM whatever cudaLimitMaxL2FetchGranularity says:
// 32 / 64 / 128 measured identical, as did a persisting-L2 window on the table and CTAs per SM;
// profiles/mlip_variants_r02.jsonl.)  Measured on the 64^4 table: 1.51e10 evaluations/s against
// 1.14e10 for the pair gather, 5.2e9 against 2.7e9 on a 48^5 table (2 GB, far beyond L2 and TLB reach).
//
// Mapping: two adjacent lanes share a point; lane h loads the 16 bytes of dim-0 corner h of every
// quad (one LDG.128 per quad and lane, both lanes of a point in the same line), computes the cell
// search for half of the dimensions and exchanges by shuffle; partial sums are combined with one
// shuffle.  Table loads carry an L2 evict_last policy and the point stream / results evict_first,
// so that what L2 there is keeps table sectors rather than data that is never read again.
// Rounding: two interleaved partial sums (few-ulp differences from the reference's corner order);
// weights, blends and the final division are the reference's.
#include "mlip_common.cuh"

namespace {

using mlip::blend_w;
using mlip::cell_search;

constexpr int kBlock = 256;

__device__ __forceinline__ uint64_t policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ double2 ldg_f64x2_hint(const double *p, uint64_t pol)
{
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(r.x), "=d"(r.y) : "l"(p), "l"(pol));
    return r;
}
// (the .L2::64B fill qualifier was measured here: DRAM traffic 422 -> 289 B/point, 1.51e10 -> 1.42e10 evals/s --
// the gather is bound by the number of DRAM accesses, not their bytes; profiles/mlip_variants_r02o.jsonl)
__device__ __forceinline__ double ldg_f64_hint(const double *p, uint64_t pol)
{
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void stg_f64_hint(double *p, double v, uint64_t pol)
{
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}

// ---- expansion: values (R layout) -> interleaved row pairs ------------------------------
__global__ void __launch_bounds__(256) mlip_quad_expand_kernel(const double *__restrict__ values, double *__restrict__ q, int n0,
                                                               int n1, int64_t nrest)
{
    // one thread per output element pair (i0; j = 0,1): coalesced 16-byte stores
    const int64_t total = nrest * (n1 - 1) * n0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int i0 = (int)(i % n0);
        const int64_t t = i / n0;
        const int i1 = (int)(t % (n1 - 1));
        const int64_t r = t / (n1 - 1);
        const double *src = values + (r * n1 + i1) * (int64_t)n0 + i0;
        double2 v;
        v.x = __ldg(src);
        v.y = __ldg(src + n0);
        reinterpret_cast<double2 *>(q)[i] = v;
    }
}

template <int RANK, bool LINEAR>
__global__ void __launch_bounds__(kBlock) mlip_quad_kernel(const __grid_constant__ MlipArgs a, const double *__restrict__ quads,
                                                           int grid_elems, int lut_elems)
{
    static_assert(RANK >= 2, "the quad layout needs two dimensions");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sgrid = reinterpret_cast<double *>(smem_raw);
    int *slut = reinterpret_cast<int *>(smem_raw + (size_t)grid_elems * 8);
    for (int i = threadIdx.x; i < grid_elems; i += kBlock) sgrid[i] = a.grid[i];
    for (int i = threadIdx.x; i < lut_elems; i += kBlock) slut[i] = a.lut[i];
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int h = lane & 1;
    const unsigned full = 0xffffffffu;
    const uint64_t pol_tab = policy_evict_last(), pol_str = policy_evict_first();
    const int64_t pts_per_grid = (int64_t)gridDim.x * (kBlock / 2);
    // quad strides: dim 1 advances one interleaved row (2*n0), dim d >= 2 advances (n1-1)*2*n0 * prod_{2<=j<d} n_j
    int64_t qstride[RANK];
    qstride[0] = 2;
    qstride[1] = 2 * (int64_t)a.dims[0];
    {
        int64_t s = 2 * (int64_t)a.dims[0] * (a.dims[1] - 1);
#pragma unroll
        for (int d = 2; d < RANK; ++d) {
            qstride[d] = s;
            s *= a.dims[d];
        }
    }

    // which dimensions this lane searches: halves of the coordinate vector, 16-byte aligned when RANK is even
    auto mine = [&](int d) -> bool {
        if constexpr ((RANK & 1) == 0) return (((d >> 1) & 1) == h) || RANK == 2;
        else return (d & 1) == h;
    };

    auto load_x = [&](int64_t p, double (&xd)[RANK]) {
#pragma unroll
        for (int d = 0; d < RANK; ++d) xd[d] = 0.0;
        if (p < a.npts) {
            const double *xp = a.x + p * RANK;
            if constexpr ((RANK & 1) == 0) {
#pragma unroll
                for (int d = 0; d < RANK; d += 2)
                    if (mine(d)) {
                        const double2 v = ldg_f64x2_hint(xp + d, pol_str);
                        xd[d] = v.x;
                        xd[d + 1] = v.y;
                    }
            } else {
#pragma unroll
                for (int d = 0; d < RANK; ++d)
                    if (mine(d)) xd[d] = ldg_f64_hint(xp + d, pol_str);
            }
        }
    };

    int64_t p = (int64_t)blockIdx.x * (kBlock / 2) + (threadIdx.x >> 1);
    double xcur[RANK];
    load_x(p, xcur);
    for (; (p - (threadIdx.x >> 1)) < a.npts; p += pts_per_grid) {  // block-uniform trip count
        const bool ok = p < a.npts;
        double xnext[RANK];
        load_x(p + pts_per_grid, xnext);  // next point's coordinates while this one's quads are in flight
        int cl[RANK];
        double w[RANK];
        bool nn = false;
#pragma unroll
        for (int d = 0; d < RANK; ++d) {
            cl[d] = 0;
            w[d] = 0.0;
            if (mine(d)) {
                const double v = xcur[d];
                nn |= (v != v);
                double g0, g1;
                const int hi = cell_search(sgrid + a.goff[d], a.dims[d], slut + a.loff[d], a.nb[d], a.lscale[d], v, g0, g1);
                cl[d] = hi - 1;
                w[d] = (v - g0) / (g1 - g0);
            }
        }
        if constexpr (RANK != 2) {
#pragma unroll
            for (int d = 0; d < RANK; ++d) {
                int owner;
                if constexpr ((RANK & 1) == 0) owner = (d >> 1) & 1;
                else owner = d & 1;
                const int src = (lane & ~1) | owner;
                cl[d] = __shfl_sync(full, cl[d], src);
                w[d] = __shfl_sync(full, w[d], src);
            }
        }
        nn = (__shfl_xor_sync(full, (int)nn, 1) | (int)nn) != 0;

        // this lane's corners: dim 0 fixed to h; one 16-byte load per quad = (dim-1 corner 0, 1)
        int64_t base = (int64_t)(cl[0] + h) * 2;
#pragma unroll
        for (int d = 1; d < RANK; ++d) base += qstride[d] * cl[d];
        constexpr int NQ = 1 << (RANK - 2);
        double2 vals[NQ];
#pragma unroll
        for (int c = 0; c < NQ; ++c) {
            int64_t pos = base;
#pragma unroll
            for (int d = 2; d < RANK; ++d)
                if ((c >> (d - 2)) & 1) pos += qstride[d];
            vals[c] = ok ? ldg_f64x2_hint(quads + pos, pol_tab) : make_double2(0.0, 0.0);
        }
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
        const double w0 = h ? whi[0] : wlo[0];
        const double w00 = w0 * wlo[1], w01 = w0 * whi[1];  // the reference's running product, g = 0, 1
        double wsum = 0.0, acc = 0.0;
#pragma unroll
        for (int c = 0; c < NQ; ++c) {
            double c0 = w00, c1 = w01;
#pragma unroll
            for (int d = 2; d < RANK; ++d) {
                const double f = ((c >> (d - 2)) & 1) ? whi[d] : wlo[d];
                c0 *= f;
                c1 *= f;
            }
            wsum += c0;
            acc = fma(c0, vals[c].x, acc);
            wsum += c1;
            acc = fma(c1, vals[c].y, acc);
        }
        wsum += __shfl_xor_sync(full, wsum, 1);
        acc += __shfl_xor_sync(full, acc, 1);
        if (ok && h == 0) {
            double r = acc / wsum;
            if (nn) r = __longlong_as_double(0x7ff8000000000000LL);
            stg_f64_hint(a.out + p, r, pol_str);
        }
#pragma unroll
        for (int d = 0; d < RANK; ++d) xcur[d] = xnext[d];
    }
}

template <int RANK>
cudaError_t launch_rank(const MlipArgs &a, const double *quads, int variant, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    int grid_elems = 0, lut_elems = 0;
    for (int d = 0; d < RANK; ++d) { grid_elems += a.dims[d]; lut_elems += a.nb[d]; }
    const int smem = (grid_elems * 8 + lut_elems * 4 + 15) & ~15;
    if (smem > 160 * 1024 || !a.lut) return cudaErrorNotSupported;
    if ((RANK & 1) == 0 && (((uintptr_t)a.x) & 15) != 0) return cudaErrorNotSupported;
    // CTAs per SM: tuning code variant % 100 (0: default)
    int per_sm = variant % 100;
    if (per_sm <= 0) per_sm = 8;
    const int64_t need = (a.npts + kBlock / 2 - 1) / (kBlock / 2);
    int grid = num_sms * per_sm;
    if (need < grid) grid = (int)(need < 1 ? 1 : need);
    cudaError_t e;
    if (a.blend == 0) {
        auto k = mlip_quad_kernel<RANK, true>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        k<<<grid, kBlock, smem, s>>>(a, quads, grid_elems, lut_elems);
    } else {
        auto k = mlip_quad_kernel<RANK, false>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        k<<<grid, kBlock, smem, s>>>(a, quads, grid_elems, lut_elems);
    }
    if (li) { li->grid = grid; li->block = kBlock; li->smem = smem; li->kernel = K_MLIP_QUAD; }
    return cudaGetLastError();
}

}  // namespace

// bytes of the interleaved row-pair table, 0 if the shape is not supported
int64_t mlip_quad_bytes(int rank, const int *dims)
{
    if (rank < 2 || rank > 6) return 0;
    int64_t n = 2 * (int64_t)dims[0] * (dims[1] - 1);
    for (int d = 2; d < rank; ++d) n *= dims[d];
    return n * 8;
}

cudaError_t mlip_quad_expand(const MlipArgs &a, const double *values, double *quads, cudaStream_t s)
{
    int64_t nrest = 1;
    for (int d = 2; d < a.rank; ++d) nrest *= a.dims[d];
    const int64_t total = nrest * (a.dims[1] - 1) * a.dims[0];
    int64_t nb = (total + 255) / 256;
    if (nb > 148 * 32) nb = 148 * 32;
    mlip_quad_expand_kernel<<<(int)nb, 256, 0, s>>>(values, quads, a.dims[0], a.dims[1], nrest);
    return cudaGetLastError();
}

cudaError_t launch_mlip_quad(const MlipArgs &a, const double *quads, int variant, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    switch (a.rank) {
    case 2: return launch_rank<2>(a, quads, variant, num_sms, s, li);
    case 3: return launch_rank<3>(a, quads, variant, num_sms, s, li);
    case 4: return launch_rank<4>(a, quads, variant, num_sms, s, li);
    case 5: return launch_rank<5>(a, quads, variant, num_sms, s, li);
    case 6: return launch_rank<6>(a, quads, variant, num_sms, s, li);
    default: return cudaErrorNotSupported;
    }
}
