// polyh.cu -- polyharmonic spline evaluation for sm_100a.
//
// Replaces the point loop of R_evalpolyh over C_evalpolyh (src/chebpol.c:383-451):
//   f(x) = sum_i w_i phi(|x - k_i|) + v_0 + sum_j v_{j+1} x_j
//   phi(r) = exp(k r^2) (k < 0), r^k (k odd), log(r) r^k (k even).
// Every point meets every knot: N*(3*rank + phi) flops per point against 8*(rank+1) bytes,
// so the kernel is FP64-pipe bound and organised like the DFMA Chebyshev kernel:
//  * a thread owns PT points for a whole pass over the knots, their (pre-mapped)
//    coordinates live in registers;
//  * the knot table is packed host-side into rows [w_i, k_i0 .. k_i(R-1), 0 pad] (R = rank
//    padded to the instantiated size, zero pad on both sides, so the padded terms add
//    exact zeros) and streamed L2 -> shared memory in slabs through a TMA bulk-copy ring
//    (cp.async.bulk + mbarrier complete_tx; the last warp to finish a stage re-arms it;
//    a table that fits stays resident);
//  * a knot row is read as warp-broadcast LDS.128s and feeds PT independent distance
//    chains (one DADD + one DFMA per coordinate), then phi and one DFMA into the sum.
// Summation runs over the knots in the reference's order, one accumulator per point;
// differences from the reference are FMA contraction and the phi formulation
// (d^((k+1)/2) * rsqrt(d) for odd k with a Newton rsqrt, 0.5*log(d) * d^(k/2) for even k:
// a few ulp per term).
#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int kMaxStages = 4;
constexpr int kBarBytes = 128;
constexpr int kNW = 8;  // warps per CTA

__host__ __device__ constexpr int row_stride_of(int R) { return (R + 2) & ~1; }

// d^M for the small compile-time powers (k <= 5 or 4); M == 3 means "run-time m >= 3"
template <int M>
__device__ __forceinline__ double ipow(double d, int m)
{
    if (M == 0) return 1.0;
    if (M == 1) return d;
    if (M == 2) return d * d;
    double pw = d * d * d;
    for (int q = 3; q < m; ++q) pw *= d;
    return pw;
}

// 1/sqrt(d) for normal positive d without the IEEE slow path: the hardware seed (~2^-22) and two
// Newton steps y <- y*(1.5 - 0.5*d*y^2), good to ~2 ulp.  d == 0 / subnormal give inf/NaN; the
// caller selects those away.
__device__ __forceinline__ double fast_rsqrt(double d)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double hd = 0.5 * d;
    double e = fma(-hd, y * y, 1.5);
    y *= e;
    e = fma(-hd, y * y, 1.5);
    y *= e;
    return y;
}

// phi(r) on d = r^2.  MODE 0: exp(k d).  MODE 1 (k odd, M = (k-1)/2): r^k = d^(M+1)/sqrt(d);
// MODE 2 (k even, M = k/2): log(r) r^k = 0.5 log(d) d^M.  d == 0 contributes nothing, as the
// reference's `continue` (src/chebpol.c:400,408); below 1e-290 r^k is far under one ulp of any sum.
template <int MODE, int M>
__device__ __forceinline__ double phi_of(double d, double k, int m)
{
    if (MODE == 0) return exp(k * d);
    if (MODE == 1) {
        const double v = ipow<M>(d, m) * d * fast_rsqrt(d);
        return d < 1e-290 ? 0.0 : v;
    }
    const double v = 0.5 * log(d) * ipow<M>(d, m);
    return d == 0.0 ? 0.0 : v;
}

// the closure's normfun x <- a*(x - b) (R/polyharmonic.R:77,82); PreMap bit 0 only
__device__ __forceinline__ double polyh_premap(const PreMap &pm, int d, double v)
{
    return (pm.mode & 1) ? (v - pm.mid[d]) * pm.ispan[d] : v;
}

struct TileArgs {
    PolyhArgs p;
    int rows_per_slab, nslabs, ns, resident, stage_bytes;
    int m;  // integer power applied to d: (k-1)/2 for odd k, k/2 for even k
};

template <int R, int PT, int MODE, int M>
__global__ void __launch_bounds__(kNW * 32) polyh_tile_kernel(const __grid_constant__ TileArgs ta)
{
    constexpr int RS = row_stride_of(R);
    constexpr int TP = kNW * 32 * PT;
    const PolyhArgs &a = ta.p;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem);
    int *done_cnt = reinterpret_cast<int *>(smem + kMaxStages * 8);
    unsigned char *ring = smem + kBarBytes;

    const int tid = threadIdx.x, lane = tid & 31;
    const int rank = a.rank, ns = ta.ns, rps = ta.rows_per_slab;
    const int64_t ntiles = (a.npts + TP - 1) / TP;
    const int64_t my_tiles = (ntiles > (int64_t)blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t total_fills = ta.resident ? (int64_t)ta.nslabs : my_tiles * ta.nslabs;

    auto slab_rows = [&](int s) { const int left = a.nknots - s * rps; return left < rps ? left : rps; };
    auto issue_fill = [&](int64_t f, int stage) {
        const int s = (int)(f % ta.nslabs);
        const uint32_t bytes = (uint32_t)slab_rows(s) * RS * 8u;
        mbar_expect_tx(&full_bar[stage], bytes);
        tma_bulk_g2s(ring + (size_t)stage * ta.stage_bytes, a.packed + (size_t)s * rps * RS, bytes, &full_bar[stage]);
    };

    if (tid == 0) {
        for (int s = 0; s < ns; ++s) {
            mbar_init(&full_bar[s], 1);
            done_cnt[s] = 0;
        }
        mbar_fence_init();
        if (my_tiles > 0)
            for (int f = 0; f < ns && f < total_fills; ++f) issue_fill(f, f);
    }
    __syncthreads();

    int stage = 0;
    uint32_t round = 0;
    int64_t fill = 0;

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        double x[PT][R], acc[PT];
        int64_t pidx[PT];
#pragma unroll
        for (int p = 0; p < PT; ++p) {
            pidx[p] = tile * TP + (int64_t)p * (kNW * 32) + tid;
            const bool ok = pidx[p] < a.npts;
            const double *xp = a.x + pidx[p] * rank;
#pragma unroll
            for (int j = 0; j < R; ++j) x[p][j] = (ok && j < rank) ? polyh_premap(a.pm, j, xp[j]) : 0.0;
            acc[p] = 0.0;
        }

        for (int s = 0; s < ta.nslabs; ++s) {
            const int st = ta.resident ? s : stage;
            mbar_wait(&full_bar[st], ta.resident ? 0u : (round & 1));
            const double *slab = reinterpret_cast<const double *>(ring + (size_t)st * ta.stage_bytes);
            const int rows = slab_rows(s);
#pragma unroll 2
            for (int i = 0; i < rows; ++i) {
                double kr[RS];
                const double2 *row = reinterpret_cast<const double2 *>(slab + (size_t)i * RS);
#pragma unroll
                for (int q = 0; q < RS / 2; ++q) {
                    const double2 v = row[q];
                    kr[2 * q] = v.x;
                    kr[2 * q + 1] = v.y;
                }
#pragma unroll
                for (int p = 0; p < PT; ++p) {
                    double d = 0.0;
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        const double t = x[p][j] - kr[1 + j];
                        d = fma(t, t, d);
                    }
                    acc[p] = fma(kr[0], phi_of<MODE, M>(d, a.k, ta.m), acc[p]);
                }
            }
            if (!ta.resident) {
                __syncwarp();
                if (lane == 0) {
                    __threadfence_block();
                    const int old = atomicAdd(&done_cnt[stage], 1);
                    if (old == kNW - 1) {
                        done_cnt[stage] = 0;
                        const int64_t nf = fill + ns;
                        if (nf < total_fills) {
                            __threadfence_block();
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                            issue_fill(nf, stage);
                        }
                    }
                }
                ++fill;
                if (++stage == ns) { stage = 0; ++round; }
            }
        }

        // the affine part, in the reference's order (src/chebpol.c:412-414)
#pragma unroll
        for (int p = 0; p < PT; ++p) {
            double v = acc[p] + __ldg(a.lweights);
#pragma unroll
            for (int j = 0; j < R; ++j)
                if (j < rank) v = fma(__ldg(a.lweights + j + 1), x[p][j], v);
            if (pidx[p] < a.npts) a.out[pidx[p]] = v;
        }
    }
}

template <int R, int PT, int MODE, int M>
cudaError_t launch_inst(const PolyhArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    constexpr int RS = row_stride_of(R);
    auto kern = polyh_tile_kernel<R, PT, MODE, M>;
    TileArgs ta;
    ta.p = a;
    const int ki = (int)a.k;
    ta.m = MODE == 1 ? (ki - 1) / 2 : (MODE == 2 ? ki / 2 : 0);
    (void)M;
    // two CTAs per SM: up to ~100 KB of shared memory each
    const int budget = 100 * 1024 - kBarBytes;
    const int row_bytes = RS * 8;
    const int64_t table = (int64_t)a.nknots * row_bytes;
    if (table <= budget) {
        ta.resident = 1;
        ta.rows_per_slab = a.nknots < 1 ? 1 : a.nknots;
        ta.nslabs = 1;
        ta.ns = 1;
    } else {
        ta.resident = 0;
        ta.ns = 3;
        ta.rows_per_slab = budget / ta.ns / row_bytes;
        ta.nslabs = (a.nknots + ta.rows_per_slab - 1) / ta.rows_per_slab;
    }
    ta.stage_bytes = ((ta.rows_per_slab * row_bytes + 127) / 128) * 128;
    const int smem = kBarBytes + ta.ns * ta.stage_bytes;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    constexpr int TP = kNW * 32 * PT;
    const int64_t ntiles = (a.npts + TP - 1) / TP;
    int grid = (int)(ntiles < 2 * num_sms ? ntiles : 2 * num_sms);
    if (grid < 1) grid = 1;
    kern<<<grid, kNW * 32, smem, s>>>(ta);
    if (li) { li->grid = grid; li->block = kNW * 32; li->smem = smem; li->kernel = K_POLYH_TILE; }
    return cudaGetLastError();
}

template <int R, int PT, int MODE>
cudaError_t launch_pow(const PolyhArgs &a, int m, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    switch (m < 3 ? m : 3) {
    case 0: return launch_inst<R, PT, MODE, 0>(a, num_sms, s, li);
    case 1: return launch_inst<R, PT, MODE, 1>(a, num_sms, s, li);
    case 2: return launch_inst<R, PT, MODE, 2>(a, num_sms, s, li);
    default: return launch_inst<R, PT, MODE, 3>(a, num_sms, s, li);
    }
}

template <int R, int PT>
cudaError_t launch_mode(const PolyhArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    const int ki = (int)a.k;
    if (a.k < 0) return launch_inst<R, PT, 0, 0>(a, num_sms, s, li);
    if (ki % 2 == 1) return launch_pow<R, PT, 1>(a, (ki - 1) / 2, num_sms, s, li);
    return launch_pow<R, PT, 2>(a, ki / 2, num_sms, s, li);
}

constexpr int kRanks[] = {1, 2, 3, 4, 6, 8, 12, 16, 24, 32};

}  // namespace

static int padded_rank(int rank)
{
    for (int r : kRanks)
        if (rank <= r) return r;
    return 0;
}

int polyh_row_stride(int rank)
{
    const int R = padded_rank(rank);
    return R ? row_stride_of(R) : 0;
}

cudaError_t launch_polyh_tile(const PolyhArgs &a, int variant, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    (void)variant;
    if (!a.packed || a.nknots < 1) return cudaErrorNotSupported;
    switch (padded_rank(a.rank)) {
    case 1: return launch_mode<1, 4>(a, num_sms, s, li);
    case 2: return launch_mode<2, 4>(a, num_sms, s, li);
    case 3: return launch_mode<3, 4>(a, num_sms, s, li);
    case 4: return launch_mode<4, 4>(a, num_sms, s, li);
    case 6: return launch_mode<6, 4>(a, num_sms, s, li);
    case 8: return launch_mode<8, 4>(a, num_sms, s, li);
    case 12: return launch_mode<12, 2>(a, num_sms, s, li);
    case 16: return launch_mode<16, 2>(a, num_sms, s, li);
    case 24: return launch_mode<24, 1>(a, num_sms, s, li);
    case 32: return launch_mode<32, 1>(a, num_sms, s, li);
    default: return cudaErrorNotSupported;
    }
}
