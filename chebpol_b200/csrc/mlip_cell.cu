// mlip_cell.cu -- multilinear evaluation from a cell-major table (sm_100a).
//
// The reference's C_evalmlip (src/chebpol.c:617-667) gathers the 2^D corner values of a
// cell from the D-dimensional value array.  On a random point batch those gathers miss:
// the 64^4 table of BASELINE config 3 (134 MB) does not survive in B200's two-partition L2
// (ncu on the pair kernel, profiles/: L2 hit rate 28%, 426 B/point of DRAM traffic against
// 168 B/point of algorithmic bytes, every missed 16-byte pair costing a 64-byte burst).
//
// B200 has 180 GB of HBM3e, so the plan trades capacity for bandwidth: the table is
// expanded once into CELL-MAJOR form -- for every cell (i_0..i_{D-1}), i_d in [0, n_d-2],
// its 2^D corner values stored contiguously (corner c = sum_d bit_d 2^d, bit_d = 1 the
// upper knot).  For D = 4 that is one aligned 128-byte line per cell (2.0 GB for 64^4), so
// a point reads exactly its algorithmic bytes: D coordinates, one cell block, one result.
//
// When 2^D times the table does not fit (64^5: 275 GB), the expansion stops after the first m
// dimensions: BRICKS of 2^m corner values, one per cell of dims < m and per GRID INDEX of dims >= m
// (2^m x the memory; m = 3: 64-byte bricks, the DRAM access granularity; m = D: the cell-major table
// above).  A point then reads 2^(D-m) bricks -- still only bytes it uses -- and because the corner
// index has the upper dimensions in its high bits, the bricks land in the staging row in exactly the
// cell-major order, so only the address of each 16-byte piece changes (a per-lane constant offset).
//
// Mapping: one thread per point for D <= 4 (a lane holds up to 16 corner values; for D = 5, 6
// two or four adjacent lanes share a point and reduce by shuffle).  The kernel is bounded by
// DRAM latency x points in flight, so per-point state is kept private to one thread (most
// points in flight per register) and the next point's coordinates are prefetched while the
// current cell block is in flight.  Weights, blend functions and the final division follow
// the reference (blend weights once per dimension, running products in the order
// g = 0..D-1 as at :654-662); the 2^D corner terms are summed in corner order like the
// reference for D <= 4.
#include "mlip_common.cuh"

namespace {

using mlip::blend_w;
using mlip::cell_search;

// ---- expansion: values (R layout) -> bricks of 2^m corners -------------------------------
// brick index = upper * nlow + lowcell:  lowcell over the cells of dims < m (dim 0 fastest), upper
// over the GRID indices of dims >= m (dim m fastest)
__global__ void __launch_bounds__(256) mlip_expand_kernel(const double *__restrict__ values, double *__restrict__ cells,
                                                          const __grid_constant__ MlipArgs a, int64_t nbricks, int m)
{
    const int D = a.rank;
    const int64_t total = nbricks << m;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int corner = (int)(i & ((1 << m) - 1));
        int64_t rem = i >> m;
        int64_t pos = 0;
        for (int d = 0; d < D; ++d) {
            const int n = d < m ? a.dims[d] - 1 : a.dims[d];
            const int id = (int)(rem % n);
            rem /= n;
            pos += (int64_t)(id + (d < m ? ((corner >> d) & 1) : 0)) * a.stride[d];
        }
        cells[i] = __ldg(values + pos);
    }
}


// NaN-safe version of the reference's bisection (src/chebpol.c:543-556) on a knot vector
// in shared memory; returns the upper knot index in [1, n-1]
__device__ __forceinline__ int cell_upper(const double *g, int n, double v)
{
    int lo = 0, hi = n - 1;
    while (lo + 1 < hi) {
        const int mid = (lo + hi) >> 1;
        if (g[mid] >= v) hi = mid;
        else lo = mid;
    }
    return hi;
}

__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
// bricks of at most 64 bytes: ask for the 64-byte L2 fill (the default fills the 128-byte line from DRAM)
__device__ __forceinline__ void cp_async16_fill64(void *dst_smem, const void *src)
{
    asm volatile("cp.async.cg.shared.global.L2::64B [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One thread per point.  The cell blocks travel global -> shared memory with cp.async
// (LDGSTS): in round k the 32 lanes of a warp copy the blocks of 32/CH points, CH = 2^(D-1)
// consecutive lanes covering one block in 16-byte pieces, so every request is a whole
// contiguous block (one 128-byte line for D = 4) and the data in flight occupies shared
// memory, not registers.  Each warp owns S stages of 32 rows; a row holds the point's cell
// block followed by its D weights, so nothing but the point index is carried in registers
// between the issue of a stage and its consumption S-1 iterations later.  The next
// point's coordinates are prefetched in registers.
template <int RANK, bool LINEAR>
__global__ void __launch_bounds__(512) mlip_cell_kernel(const __grid_constant__ MlipArgs a, const double *__restrict__ cells,
                                                        int grid_elems, int lut_elems, int S, int rowb, int m)
{
    constexpr int NCORN = 1 << RANK;
    constexpr int CH = NCORN / 2;                  // 16-byte pieces per block (RANK >= 1)
    constexpr int PPI = 32 / CH;                   // points served per copy instruction
    constexpr int LOWD = RANK < 4 ? RANK : 4;
    constexpr int NLOW = 1 << LOWD;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sgrid = reinterpret_cast<double *>(smem_raw);
    // 1 / (g[i] - g[i-1]) beside the knots: with linear blending the weight (x - g0) / (g1 - g0) becomes one
    // multiply (<= 1.5 ulp from the quotient, the blend is continuous: far inside the fast kernels' 1e-12 bar);
    // the other blends keep the division -- blend 4 is a step at w = 1/2, where one ulp decides the corner
    double *srinv = sgrid + grid_elems;
    int *slut = reinterpret_cast<int *>(smem_raw + (size_t)grid_elems * 16);
    const int head_bytes = (grid_elems * 16 + lut_elems * 4 + 15) & ~15;
    for (int i = threadIdx.x; i < grid_elems; i += blockDim.x) {
        const double gi = a.grid[i];
        sgrid[i] = gi;
        srinv[i] = i > 0 ? 1.0 / (gi - a.grid[i - 1]) : 0.0;  // entries at the first knot of a dimension are never read
    }
    for (int i = threadIdx.x; i < lut_elems; i += blockDim.x) slut[i] = a.lut[i];
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *wstage = smem_raw + head_bytes + (size_t)warp * S * 32 * rowb;
    const unsigned full = 0xffffffffu;
    const int64_t stride_pts = (int64_t)gridDim.x * blockDim.x;
    const int64_t niter = (a.npts + stride_pts - 1) / stride_pts;  // iterations that issue
    const int64_t total = niter + (S - 1);
    const int64_t pbase = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // brick strides: dims < m count cells, dims >= m count grid indices; this lane's 16-byte piece
    // (lane % CH) lies in brick (piece >> (m-1)) of the point, i.e. at upper-dimension corner bits b
    int64_t bstr[RANK];
    unsigned long long lane_off;
    {
        int64_t sacc = 1;
#pragma unroll
        for (int d = 0; d < RANK; ++d) {
            bstr[d] = sacc;
            sacc *= d < m ? a.dims[d] - 1 : a.dims[d];
        }
        const int piece = lane % CH;
        const int b = m >= 1 ? piece >> (m - 1) : 0;
        int64_t boff = 0;
#pragma unroll
        for (int d = 0; d < RANK; ++d)
            if (d >= m && ((b >> (d - m)) & 1)) boff += bstr[d];
        lane_off = ((unsigned long long)boff << (m + 3)) + 16ull * (unsigned)(piece & ((1 << (m - 1)) - 1));
    }

    auto load_x = [&](int64_t p, double (&xv)[RANK]) {
        const bool ok = p < a.npts;
        if constexpr ((RANK % 2) == 0) {
#pragma unroll
            for (int d = 0; d < RANK; d += 2) {
                double2 t = make_double2(0.0, 0.0);
                if (ok) t = ldg_stream_f64x2(a.x + p * RANK + d);
                xv[d] = t.x;
                xv[d + 1] = t.y;
            }
        } else {
#pragma unroll
            for (int d = 0; d < RANK; ++d) xv[d] = ok ? ldg_stream_f64(a.x + p * RANK + d) : 0.0;
        }
    };

    double xcur[RANK];
    load_x(pbase, xcur);
    int si = 0;  // stage to issue into

    for (int64_t it = 0; it < total; ++it) {
        // ---- issue: stage si receives the blocks (and weights) of this iteration's 32 points
        if (it < niter) {  // warp-uniform
            const int64_t p = pbase + it * stride_pts;
            double xnext[RANK];
            load_x(p + stride_pts, xnext);
            unsigned char *st = wstage + (size_t)si * 32 * rowb;
            double *wrow = reinterpret_cast<double *>(st + (size_t)lane * rowb + NCORN * 8);
            int64_t cell = 0;
            bool nn = false;
#pragma unroll
            for (int d = 0; d < RANK; ++d) {
                const double v = xcur[d];
                nn |= (v != v);
                double g0, g1;
                const int hi = cell_search(sgrid + a.goff[d], a.dims[d], slut + a.loff[d], a.nb[d], a.lscale[d], v, g0, g1);
                if constexpr (LINEAR) wrow[d] = v == g1 ? 1.0 : (v - g0) * srinv[a.goff[d] + hi];  // knots reproduce exactly
                else wrow[d] = (v - g0) / (g1 - g0);
                cell += bstr[d] * (hi - 1);
            }
            if (nn) wrow[0] = __longlong_as_double(0x7ff8000000000000LL);  // NaN in -> NaN out, whatever the blend
            if (p >= a.npts) cell = 0;
            const unsigned long long base = reinterpret_cast<unsigned long long>(cells) + ((unsigned long long)cell << (m + 3));
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                const int pt = lane / CH + PPI * k;  // the lane whose point this lane serves in round k
                const unsigned long long b = __shfl_sync(full, base, pt);
                void *dst = st + (size_t)pt * rowb + 16 * (lane % CH);
                if (m <= 3) cp_async16_fill64(dst, reinterpret_cast<const void *>(b + lane_off));  // warp-uniform
                else cp_async16(dst, reinterpret_cast<const void *>(b + lane_off));
            }
#pragma unroll
            for (int d = 0; d < RANK; ++d) xcur[d] = xnext[d];
        }
        cp_async_commit();
        // ---- consume the oldest stage (issued S-1 iterations ago)
        if (S == 2) cp_async_wait<1>();
        else if (S == 3) cp_async_wait<2>();
        else cp_async_wait<3>();
        __syncwarp();
        const int sc = (si + 1 == S) ? 0 : si + 1;
        const int64_t pc = pbase + (it - (S - 1)) * stride_pts;
        if (it >= S - 1 && pc < a.npts) {
            const unsigned char *row = wstage + (size_t)sc * 32 * rowb + (size_t)lane * rowb;
            const double *wrow = reinterpret_cast<const double *>(row + NCORN * 8);
            double whi[RANK], wlo[RANK];
            const double w0 = wrow[0];
            const bool isnan_ = (w0 != w0);
#pragma unroll
            for (int d = 0; d < RANK; ++d) {
                const double w = wrow[d];
                if constexpr (LINEAR) {
                    whi[d] = w;
                    wlo[d] = 1 - w;
                } else {
                    whi[d] = blend_w(w, a.blend);
                    wlo[d] = blend_w(1 - w, a.blend);
                }
            }
            // corner weights of the low dims as running products, reference order
            double cw[NLOW];
            cw[0] = wlo[0];
            if constexpr (NLOW > 1) cw[1] = whi[0];
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
            for (int h = 0; h < NCORN / NLOW; ++h) {  // high-dim corner bits (RANK > 4)
                double f = 1.0;
#pragma unroll
                for (int d = LOWD; d < RANK; ++d) f *= ((h >> (d - LOWD)) & 1) ? whi[d] : wlo[d];
                if constexpr (NLOW == 1) {
                    const double v = *reinterpret_cast<const double *>(row);
                    wsum += cw[0];
                    acc = fma(cw[0], v, acc);
                } else {
#pragma unroll
                    for (int c = 0; c < NLOW; c += 2) {
                        const double2 v = *reinterpret_cast<const double2 *>(row + (size_t)(h * NLOW + c) * 8);
                        double c0 = cw[c], c1 = cw[c + 1];
                        if constexpr (RANK > LOWD) { c0 *= f; c1 *= f; }
                        wsum += c0;
                        acc = fma(c0, v.x, acc);
                        wsum += c1;
                        acc = fma(c1, v.y, acc);
                    }
                }
            }
            double r = acc / wsum;
            if (isnan_) r = __longlong_as_double(0x7ff8000000000000LL);
            stg_stream_f64(a.out + pc, r);
        }
        __syncwarp();  // stage sc is overwritten by the next issue
        si = sc;
    }
}

// row stride of a stage: block, D weights; a multiple of 16 B whose 16-byte count is odd
// (conflict-free LDS.128 across 8 consecutive rows)
int stage_row_bytes(int rank)
{
    int b = ((1 << rank) * 8 + rank * 8 + 15) & ~15;
    if (((b / 16) & 1) == 0) b += 16;
    return b;
}

template <int RANK>
cudaError_t launch_rank(const MlipArgs &a, const double *cells, int m, int variant, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    int grid_elems = 0, lut_elems = 0;
    for (int d = 0; d < RANK; ++d) { grid_elems += a.dims[d]; lut_elems += a.nb[d]; }
    const int head_bytes = (grid_elems * 16 + lut_elems * 4 + 15) & ~15;  // knots, reciprocal spacings, bucket tables
    const int rowb = stage_row_bytes(RANK);
    if ((((uintptr_t)a.x) & 15) != 0 || !a.lut || m < 1 || m > RANK) return cudaErrorNotSupported;
    // layout: `warps` warps per CTA, S stages each; tuning code = warps*10 + S (0: default)
    int warps = 16, S = 2;
    if (variant >= 20) { warps = variant / 10; S = variant % 10; }
    if (S < 2 || S > 4 || warps < 1 || warps > 16) return cudaErrorNotSupported;
    const int budget = 220 * 1024;
    while (warps > 1 && head_bytes + warps * S * 32 * rowb > budget) warps /= 2;
    const int smem = head_bytes + warps * S * 32 * rowb;
    if (smem > budget) return cudaErrorNotSupported;
    const int block = warps * 32;
    int per_sm = budget / (smem + 1024);
    if (per_sm < 1) per_sm = 1;
    if (per_sm * block > 2048) per_sm = 2048 / block;
    const int64_t need = (a.npts + block - 1) / block;
    int grid = num_sms * per_sm;
    if (need < grid) grid = (int)(need < 1 ? 1 : need);
    cudaError_t e;
    if (a.blend == 0) {
        auto k = mlip_cell_kernel<RANK, true>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        k<<<grid, block, smem, s>>>(a, cells, grid_elems, lut_elems, S, rowb, m);
    } else {
        auto k = mlip_cell_kernel<RANK, false>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        k<<<grid, block, smem, s>>>(a, cells, grid_elems, lut_elems, S, rowb, m);
    }
    if (li) { li->grid = grid; li->block = block; li->smem = smem; li->kernel = K_MLIP_CELL; }
    return cudaGetLastError();
}

}  // namespace

// bytes of the brick table with the first m dimensions expanded (m = rank: cell-major), 0 if the
// shape is not supported
int64_t mlip_cell_bytes(int rank, const int *dims, int m)
{
    if (rank < 1 || rank > 6 || m < 1 || m > rank) return 0;
    int64_t bricks = 1;
    for (int d = 0; d < rank; ++d) {
        if (dims[d] < 2) return 0;
        bricks *= d < m ? dims[d] - 1 : dims[d];
        if (bricks > (int64_t(1) << 40)) return 0;
    }
    return (bricks << m) * 8;
}

cudaError_t mlip_expand(const MlipArgs &a, const double *values, double *cells, int m, cudaStream_t s)
{
    int64_t nbricks = 1;
    for (int d = 0; d < a.rank; ++d) nbricks *= d < m ? a.dims[d] - 1 : a.dims[d];
    const int64_t total = nbricks << m;
    int64_t nb = (total + 255) / 256;
    if (nb > 148 * 32) nb = 148 * 32;
    mlip_expand_kernel<<<(int)nb, 256, 0, s>>>(values, cells, a, nbricks, m);
    return cudaGetLastError();
}

cudaError_t launch_mlip_cell(const MlipArgs &a, const double *cells, int m, int variant, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    switch (a.rank) {
    case 1: return launch_rank<1>(a, cells, m, variant, num_sms, s, li);
    case 2: return launch_rank<2>(a, cells, m, variant, num_sms, s, li);
    case 3: return launch_rank<3>(a, cells, m, variant, num_sms, s, li);
    case 4: return launch_rank<4>(a, cells, m, variant, num_sms, s, li);
    case 5: return launch_rank<5>(a, cells, m, variant, num_sms, s, li);
    case 6: return launch_rank<6>(a, cells, m, variant, num_sms, s, li);
    default: return cudaErrorNotSupported;
    }
}
