#pragma once
// contract_mma.cuh -- FP64 tensor-core (DMMA) contraction engine for sm_100a.
//
// Serves the two dense evaluators of the path:
//   Chebyshev  f(x) = sum_i c[i_0..i_{D-1}] prod_d T_{i_d}(x_d)     (C_evalcheb, src/chebpol.c:314-338)
//   Floater-Hormann  nested barycentric rational, whose value is the same
//              contraction with b_d[i] = (w_i/(x_d-k_i)) / sum_j w_j/(x_d-k_j)
//              and a one-hot b_d on a knot hit                      (FH, src/chebpol.c:182-218)
// Both are  f(x) = sum_r' W_r'(x) * sum_iA b_A[iA](x) * ( sum_k A_k(x) * C[(iA,r')][k] )  with k
// running over the first `nfold` dimensions (folded, product basis A_k), iA over the next
// one and r' over the rest (weights W_r').
//
// Why tensor cores: on B200 the DMMA pipe has the same peak as the DFMA pipe (36.9 vs
// 36.8 TFLOP/s measured, tools/dmma_peak.cu) but one DMMA.8x8x4 retires 256 FMAs with
// four 64-bit register operands per thread, whereas the DFMA loop of cheb_slab.cu is
// limited by register-file operand bandwidth (ncu: ~2 extra issue cycles per 6 DFMAs,
// FP64 pipe 80% active; profiles/).  The inner sum over k is a GEMM:
//   D[point][r] = sum_k A[point][k] * C[r][k]    (M = 8 points, N = 8 rows r, K = 4 per DMMA)
//
// Mapping:
//  * a CTA owns NPT = NW*MT*8 points per pass; a warp owns MT tiles of 8 points;
//  * per point and dimension the basis vector is computed once per pass into shared
//    memory (Tb); each thread then builds its A fragments (MT*KC doubles, registers):
//    A[point g][k = 4*kc + t] = prod over folded dims of Tb;
//  * the tensor, stored as rows C[r][0..K) padded to `rs` doubles, is streamed through a
//    shared-memory ring by the TMA bulk-copy engine; the last warp to finish a stage
//    issues its refill (no producer warp);
//  * the rows are ordered so that consecutive 8-row N-tiles walk the first unfolded dim
//    ("level A") for the same eight columns r'; per pair of N-tiles: KC x (2 B-fragment
//    LDS.64 + 2*MT DMMA), and the epilogue is one DFMA per D element (acc2 += b_A[iA] * D);
//    the weights of the other unfolded dims (product of Tb entries) are applied once per
//    group of nA N-tiles (acc += W_r' * acc2); the four lanes of a quad are summed by two
//    shuffles at the end of the pass.
//
// Rounding: explicit-basis sums (not Clenshaw / not the reference's num/den recursion):
// differences are O(n*eps*sum|c*basis|), inside 1e-12*max|f| (tests/test_parity_gpu.py).
#include "common.cuh"
#include "ptx.cuh"
#include <float.h>
#include <cstdio>

#ifdef MMA_TIMING
// debug build only (-DMMA_TIMING): clock64() of CTA 0's warps at the phases of their first passes,
// [pass < 16][warp < 16][stamp < 8]: 0 pass begin (before barrier 1), 1 after barrier 1, 2 basis done (before
// barrier 2), 3 after barrier 2, 4 A fragments built, 5 first slab arrived, 6 last step issued, 7 pass stored.
// (ptxas hoists the clock reads that follow a barrier above it: "after barrier" stamps read as "arrived at".)
static __device__ long long g_mma_timing[16 * 16 * 8];
#define MMA_STAMP(k) do { if (blockIdx.x == 0 && lane == 0 && tpass < 16 && warp < 16) g_mma_timing[(tpass * 16 + warp) * 8 + (k)] = clock64(); } while (0)
#else
#define MMA_STAMP(k) do { } while (0)
#endif

namespace mma_detail {

constexpr int kMaxStages = 8;
constexpr int kBarBytes = 128;
constexpr int kSmemBudget = 220 * 1024;
constexpr int kLevMax = 4;  // rest levels (rank - nfold) the engine supports

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// ---- per-point, per-dimension basis vectors into shared memory -------------------
// column layout: col[i*NPT] for i = 0..n-1
template <int NPT>
__device__ __forceinline__ void cheb_basis(double *col, int n, double x)
{
    const double tx = 2.0 * x;
    double tkm1 = 1.0, tk = x;
    col[0] = 1.0;
    if (n > 1) col[NPT] = x;
    for (int i = 2; i < n; ++i) {
        const double t = fma(tx, tk, -tkm1);
        tkm1 = tk;
        tk = t;
        col[(size_t)i * NPT] = t;
    }
}

// 1/d to full double precision without the IEEE slow path: the 20-bit hardware seed and two
// Newton steps.  d is never subnormal here (knot hits are peeled off first).
__device__ __forceinline__ double fast_rcp(double d)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// Floater-Hormann basis b_i = (w_i/(x-k_i)) / sum_j (w_j/(x-k_j)), the reference's own form
// (src/chebpol.c:199-215).  The n reciprocals are independent, so the loop pipelines; the only
// serial chain is the running sum.  Knot hit (first i with |x-k_i| < 10*DBL_EPSILON,
// src/chebpol.c:195-197): one-hot.
template <int NPT>
__device__ __forceinline__ void fh_basis(double *col, int n, double x, const double *__restrict__ kn,
                                         const double *__restrict__ w)
{
    int hit = n;
    double den0 = 0.0, den1 = 0.0;
    int i = 0;
#pragma unroll 2
    for (; i + 1 < n; i += 2) {
        const double d0 = x - kn[i], d1 = x - kn[i + 1];
        const double p0 = w[i] * fast_rcp(d0), p1 = w[i + 1] * fast_rcp(d1);
        col[(size_t)i * NPT] = p0;
        col[(size_t)(i + 1) * NPT] = p1;
        den0 += p0;
        den1 += p1;
        hit = (fabs(d1) < 10.0 * DBL_EPSILON && i + 1 < hit) ? i + 1 : hit;
        hit = (fabs(d0) < 10.0 * DBL_EPSILON && i < hit) ? i : hit;
    }
    if (i < n) {
        const double d0 = x - kn[i];
        const double p0 = w[i] * fast_rcp(d0);
        col[(size_t)i * NPT] = p0;
        den0 += p0;
        hit = (fabs(d0) < 10.0 * DBL_EPSILON && i < hit) ? i : hit;
    }
    if (hit < n || x != x) {
        for (int j = 0; j < n; ++j) col[(size_t)j * NPT] = (j == hit) ? 1.0 : ((x != x) ? x : 0.0);
        return;
    }
    const double den = den0 + den1;
    const double r = 1.0 / den;
    if (isfinite(r)) {
#pragma unroll 4
        for (int j = 0; j < n; ++j) col[(size_t)j * NPT] *= r;
    } else {
        // the sum cancelled to zero or left the range (extreme extrapolation): redo it with IEEE
        // divisions in the reference's order, so the result is finite exactly when the reference's is
        double dsum = 0.0;
        for (int j = 0; j < n; ++j) {
            const double p = w[j] / (x - kn[j]);
            col[(size_t)j * NPT] = p;
            dsum += p;
        }
        for (int j = 0; j < n; ++j) col[(size_t)j * NPT] /= dsum;
    }
}

// KC: K/4 (padded); MT: 8-point tiles per warp; NW: warps; NTB: N-tiles (level-A indices) per step.
template <int KC, int MT, int NW, int NTB>
__global__ void __launch_bounds__(NW * 32, 1) contract_mma_kernel(const __grid_constant__ MmaArgs a)
{
    constexpr int NPT = NW * MT * 8;  // points per pass
    constexpr int NT = NW * 32;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem);
    int *done_cnt = reinterpret_cast<int *>(smem + kMaxStages * 8);
    unsigned char *ring = smem + kBarBytes;
    double *Tb = reinterpret_cast<double *>(ring + (size_t)a.ns * a.stage_bytes);
    // Small tables copied to shared memory once (read from global memory they cost an L2 round trip each at
    // every pass, all warps at the same time): the digit rows of the folded dims for each k; the
    // Floater-Hormann knots and weights and the group-level table when the launcher found room for them.
    int *sktab = reinterpret_cast<int *>(Tb + (size_t)(a.sumn + 1) * NPT);
    double *sknots = reinterpret_cast<double *>(sktab + 4 * KC * 3);
    const bool fh_local = a.fh_local != 0;
    const double *fh_grid = fh_local ? sknots : a.grid, *fh_w = fh_local ? sknots + a.sumn : a.weights;
    int *sginfo = reinterpret_cast<int *>(sknots + (fh_local ? 2 * a.sumn : 0));
    const int ngi = (a.nrp + 7) / 8 * 8 * 3;
    const int *ginfo = a.gi_local ? sginfo : a.ginfo;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int rank = a.rank, ns = a.ns;
    const int64_t ntiles = (a.npts + NPT - 1) / NPT;
    const int64_t my_tiles = (ntiles > (int64_t)blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t total_fills = a.resident ? (int64_t)a.nslabs : my_tiles * a.nslabs;
    const int rps = a.rows_per_slab;
    const uint32_t row_bytes = (uint32_t)a.rs * 8u;
    const int spg = a.nAp / NTB;  // steps per group (a group = one tile of 8 r' values, all of level A)
    constexpr int ROWS_STEP = 8 * NTB;

    auto slab_rows = [&](int s) { const int left = a.nrows - s * rps; return left < rps ? left : rps; };
    auto issue_fill = [&](int64_t f, int stage) {
        const int s = (int)(f % a.nslabs);
        const uint32_t bytes = (uint32_t)slab_rows(s) * row_bytes;
        mbar_expect_tx(&full_bar[stage], bytes);
        tma_bulk_g2s(ring + (size_t)stage * a.stage_bytes, a.tensor + (size_t)s * rps * a.rs, bytes, &full_bar[stage]);
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
    // the ones row of the basis table (absent levels)
    for (int i = tid; i < NPT; i += NT) Tb[(size_t)a.sumn * NPT + i] = 1.0;
    for (int i = tid; i < 4 * KC * 3; i += NT) sktab[i] = a.ktab[i];
    if (fh_local)
        for (int i = tid; i < a.sumn; i += NT) { sknots[i] = a.grid[i]; sknots[a.sumn + i] = a.weights[i]; }
    if (a.gi_local)
        for (int i = tid; i < ngi; i += NT) sginfo[i] = a.ginfo[i];

    int stage = 0;
    uint32_t round = 0;
    int64_t fill = 0;

    // coordinates of the next pass are fetched (into registers) while the current pass
    // computes, so the basis prologue never waits on DRAM
    constexpr int kMaxTasks = (NPT * (3 + kLevMax) + NT - 1) / NT;  // rank <= nfold(3) + kLevMax
    double xpre[kMaxTasks];
    auto prefetch_x = [&](int64_t tile) {
#pragma unroll
        for (int q = 0; q < kMaxTasks; ++q) {
            const int task = tid + q * NT;
            const int pt = task % NPT, L = task / NPT;
            const int64_t p = tile * NPT + pt;
            xpre[q] = (task < NPT * rank && tile < ntiles && p < a.npts) ? ldg_stream_f64(a.x + p * rank + L) : 0.0;
        }
    };
    prefetch_x(blockIdx.x);

#ifdef MMA_TIMING
    int tpass = -1;
#endif
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
#ifdef MMA_TIMING
        ++tpass;
#endif
        MMA_STAMP(0);
        __syncthreads();  // previous pass has finished reading Tb (and, first time, barrier init is visible)
        MMA_STAMP(1);
        // ---- 1. basis vectors of this pass's points
#pragma unroll
        for (int q = 0; q < kMaxTasks; ++q) {
            const int task = tid + q * NT;
            if (task >= NPT * rank) break;
            const int pt = task % NPT, L = task / NPT;
            const int64_t p = tile * NPT + pt;
            double *col = Tb + (size_t)a.boff[L] * NPT + pt;
            const int n = a.dims[L];
#if defined(MMA_ABL) && (MMA_ABL & 8)
            if (p < 0) {
#else
            if (p < a.npts) {
#endif
                const double xv = xpre[q];
                if (a.kind == 0) cheb_basis<NPT>(col, n, premap_coord(a.pm, L, xv));
                else fh_basis<NPT>(col, n, xv, fh_grid + a.goff[L], fh_w + a.goff[L]);
            } else {
                for (int i = 0; i < n; ++i) col[(size_t)i * NPT] = 0.0;
            }
        }
        prefetch_x(tile + gridDim.x);
        MMA_STAMP(2);
        __syncthreads();
        MMA_STAMP(3);

        // ---- 2. A fragments: A[mt][kc] = prod_{l < nfold} Tb_l[digit_l(k)][point], k = 4*kc + t;
        //         the digit rows come from a host-built table (no integer division here)
        double A[MT][KC];
        // (straight-line per number of folded dims, padding k and absent dims naming the ones row: every load of
        //  the pass's A fragments is in flight at once; the branchy per-k form with the digit rows in global memory
        //  took 6.3 k cycles of a 238 k-cycle pass on cfg2, all warps at the same time)
        if constexpr (KC * MT >= 60) {
            // instantiations that hold 60 or more A fragments per lane (cfg5: <36, 2, 8>, 243 registers) keep the
            // per-k form: the straight-line build costs them 2.3 % in the step loop's register allocation
            // (5.98e6 vs 5.85e6 evals/s on 12^6), and their passes are hundreds of slabs long
#pragma unroll
            for (int kc = 0; kc < KC; ++kc) {
                const int *kt = a.ktab + (4 * kc + t) * 3;
                const int r0 = __ldg(kt), r1 = __ldg(kt + 1), r2 = __ldg(kt + 2);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    const int pt = (warp * MT + mt) * 8 + g;
                    double v = Tb[(size_t)r0 * NPT + pt];
                    if (r1 != a.sumn) v *= Tb[(size_t)r1 * NPT + pt];
                    if (r2 != a.sumn) v *= Tb[(size_t)r2 * NPT + pt];
                    A[mt][kc] = v;
                }
            }
        } else {
            const double *tb = Tb + warp * MT * 8 + g;
            if (a.nfold == 1) {
#pragma unroll
                for (int kc = 0; kc < KC; ++kc) {
                    const int r0 = sktab[(4 * kc + t) * 3];
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) A[mt][kc] = tb[(size_t)r0 * NPT + mt * 8];
                }
            } else if (a.nfold == 2) {
#pragma unroll
                for (int kc = 0; kc < KC; ++kc) {
                    const int *kt = sktab + (4 * kc + t) * 3;
                    const int r0 = kt[0], r1 = kt[1];
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) A[mt][kc] = tb[(size_t)r0 * NPT + mt * 8] * tb[(size_t)r1 * NPT + mt * 8];
                }
            } else {
#pragma unroll
                for (int kc = 0; kc < KC; ++kc) {
                    const int *kt = sktab + (4 * kc + t) * 3;
                    const int r0 = kt[0], r1 = kt[1], r2 = kt[2];
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt)
                        A[mt][kc] = tb[(size_t)r0 * NPT + mt * 8] * tb[(size_t)r1 * NPT + mt * 8] * tb[(size_t)r2 * NPT + mt * 8];
                }
            }
        }

        // ---- 3. stream the tensor rows.  A step is 8*NTB rows: NTB 8-row N-tiles that hold the SAME
        // eight r' columns at NTB consecutive level-A indices; a group is the nAp/NTB steps that
        // walk level A for those columns.  Per step the epilogue is therefore one
        // basis value per N-tile and point tile and one DFMA per D element (acc2 += b_A * D);
        // the weights of the remaining levels are applied once per group (acc += W * acc2).
        // Software pipeline: the D tiles of step q-1 are folded in while the DMMAs of step q are
        // in flight (the per-step part is branch-free so ptxas interleaves the two).
        double acc[MT], acc2[2][MT];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) acc[mt] = acc2[0][mt] = acc2[1][mt] = 0.0;
        // a step's D tiles and what the epilogue needs to know about them
        struct Step {
            double D[NTB][MT][2];
            int iA;    // level-A index of the first N-tile
            bool end;  // the step closes its group ...
            int g8;    // ... which is this one (eight r' columns)
        };
        Step P0, P1;  // the step loop is unrolled by two: the roles (in flight / pending) alternate
#pragma unroll
        for (int j = 0; j < NTB; ++j)
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) P1.D[j][mt][0] = P1.D[j][mt][1] = 0.0;
        P1.iA = 0;
        P1.end = false;
        P1.g8 = 0;
        const double *tcol = Tb + warp * MT * 8 + g;
        const int nA1 = a.nA - 1;

        auto fold_pending = [&](const Step &P) {
            double bA[NTB][MT];
#pragma unroll
            for (int j = 0; j < NTB; ++j) {
                const int ia = P.iA + j < nA1 ? P.iA + j : nA1;  // padded indices: their tensor rows are zero
                const double *tj = tcol + (size_t)(a.boffA + ia) * NPT;
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) bA[j][mt] = tj[mt * 8];
            }
#pragma unroll
            for (int j = 0; j < NTB; ++j)
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) acc2[e][mt] = fma(bA[j][mt], P.D[j][mt][e], acc2[e][mt]);
        };
        // Runs on group boundaries only (a warp-uniform branch).  Straight-line: the loads of all
        // weights go out together (absent levels read the ones row), so the exposed latency is one
        // shared-memory round trip and three dependent multiplies, not a chain per level.
        auto close_group = [&](const Step &P) {
            // basis-table rows of the group levels of this lane's two columns (padding columns and absent
            // levels name the ones row; the accumulated value of a padding column is zero: its tensor rows are)
            const int *gi = ginfo + (size_t)(8 * P.g8 + 2 * t) * 3;
            int info[2][3];
#pragma unroll
            for (int e = 0; e < 2; ++e)
#pragma unroll
                for (int l = 0; l < 3; ++l) info[e][l] = gi[3 * e + l];
            double w[2][MT][3];
#pragma unroll
            for (int e = 0; e < 2; ++e)
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int l = 0; l < 3; ++l) w[e][mt][l] = tcol[(size_t)info[e][l] * NPT + mt * 8];
#pragma unroll
            for (int e = 0; e < 2; ++e)
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    const double ww = w[e][mt][0] * w[e][mt][1] * w[e][mt][2];
                    acc[mt] = fma(ww, acc2[e][mt], acc[mt]);
                    acc2[e][mt] = 0.0;
                }
        };

        int sg = 0, g8 = 0;  // step within the group, group index
        // one step: issue the DMMAs of rows [r16, r16 + ROWS_STEP) into C while P is folded in
        auto step = [&](Step &C, const Step &P, const double *slab, int r16) {
#pragma unroll
            for (int j = 0; j < NTB; ++j)
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) C.D[j][mt][0] = C.D[j][mt][1] = 0.0;
            const double *bp = slab + (size_t)(r16 + g) * a.rs + t;
            C.end = (sg == spg - 1);
            C.iA = NTB * sg;
            C.g8 = g8;
            if constexpr (mma_pair_rows(KC)) {
                // pair-interleaved rows: one LDS.128 per N-tile holds this lane's B fragments of two k-chunks
                const double *bq = slab + (size_t)(r16 + g) * a.rs + 2 * t;
#pragma unroll
                for (int kc = 0; kc + 1 < KC; kc += 2) {
                    double2 b[NTB];
#pragma unroll
                    for (int j = 0; j < NTB; ++j) b[j] = *reinterpret_cast<const double2 *>(bq + (size_t)j * 8 * a.rs + 4 * kc);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                        for (int j = 0; j < NTB; ++j) dmma884(C.D[j][mt][0], C.D[j][mt][1], A[mt][kc], b[j].x);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                        for (int j = 0; j < NTB; ++j) dmma884(C.D[j][mt][0], C.D[j][mt][1], A[mt][kc + 1], b[j].y);
                }
                if constexpr (KC & 1) {
#pragma unroll
                    for (int j = 0; j < NTB; ++j) {
                        const double b = bp[(size_t)j * 8 * a.rs + 4 * (KC - 1)];
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) dmma884(C.D[j][mt][0], C.D[j][mt][1], A[mt][KC - 1], b);
                    }
                }
            } else {
#pragma unroll
                for (int kc = 0; kc < KC; ++kc) {
                    double b[NTB];
#pragma unroll
#if defined(MMA_ABL) && (MMA_ABL & 4)
                    for (int j = 0; j < NTB; ++j) b[j] = bp[(size_t)j * 8 * a.rs];
#else
                    for (int j = 0; j < NTB; ++j) b[j] = bp[(size_t)j * 8 * a.rs + 4 * kc];
#endif
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                        for (int j = 0; j < NTB; ++j) dmma884(C.D[j][mt][0], C.D[j][mt][1], A[mt][kc], b[j]);
                }
            }
#if defined(MMA_ABL) && (MMA_ABL & 1)
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) acc[mt] += P.D[0][mt][0] + P.D[1][mt][1];
#else
            fold_pending(P);
            if (P.end) close_group(P);
#endif
            if (++sg == spg) { sg = 0; ++g8; }
        };

        MMA_STAMP(4);
        for (int s = 0; s < a.nslabs; ++s) {
            const int st = a.resident ? s : stage;
            mbar_wait(&full_bar[st], a.resident ? 0u : (round & 1));
            if (s == 0) MMA_STAMP(5);
            const double *slab = reinterpret_cast<const double *>(ring + (size_t)st * a.stage_bytes);
            const int rows = slab_rows(s);

            // pending is P1 on entry and on exit
            int r16 = 0;
            for (; r16 + 2 * ROWS_STEP <= rows; r16 += 2 * ROWS_STEP) {
                step(P0, P1, slab, r16);
                step(P1, P0, slab, r16 + ROWS_STEP);
            }
            if (r16 < rows) {
                step(P0, P1, slab, r16);
                P1 = P0;
            }

            if (!a.resident) {
                __syncwarp();
                if (lane == 0) {
                    __threadfence_block();
                    const int old = atomicAdd(&done_cnt[stage], 1);
                    if (old == NW - 1) {
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
        MMA_STAMP(6);
        fold_pending(P1);  // drain: the last step always closes its group
        close_group(P1);

        // ---- 4. sum the four lanes of each quad, store
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            double v = acc[mt];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            const int64_t p = tile * NPT + (warp * MT + mt) * 8 + g;
            if (t == 0 && p < a.npts) a.out[p] = v;
        }
        MMA_STAMP(7);
    }
}

}  // namespace mma_detail
bool mma_fit_ring(MmaArgs &a, int npt);
namespace mma_detail {

template <int KC, int MT, int NW, int NTB>
cudaError_t launch_one(const MmaArgs &a_in, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    auto kern = contract_mma_kernel<KC, MT, NW, NTB>;
    constexpr int NPT = NW * MT * 8;
    MmaArgs a = a_in;
    if (a.ntb != NTB || !mma_fit_ring(a, NPT)) return cudaErrorNotSupported;
    // beside the ring budget (220 KB of the 227 KB a CTA may have): the k digit rows, and when they fit the
    // remaining 4.5 KB the FH knots + weights and the group-level table
    int extra = 4 * KC * 3 * 4, room = 4608;
    a.fh_local = (a.kind == 1 && 2 * a.sumn * 8 <= room) ? 1 : 0;
    if (a.fh_local) { extra += 2 * a.sumn * 8; room -= 2 * a.sumn * 8; }
    const int gi_bytes = (a.nrp + 7) / 8 * 8 * 3 * 4;
    a.gi_local = gi_bytes <= room ? 1 : 0;
    if (a.gi_local) extra += gi_bytes;
    const int smem = kBarBytes + a.ns * a.stage_bytes + (a.sumn + 1) * NPT * 8 + extra;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    const int64_t ntiles = (a.npts + NPT - 1) / NPT;
    const int ctas = num_sms * (a.occ == 2 ? 2 : 1);
    int grid = (int)(ntiles < ctas ? ntiles : ctas);
    if (grid < 1) grid = 1;
    kern<<<grid, NW * 32, smem, s>>>(a);
#ifdef MMA_TIMING
    {
        static long long h[16 * 16 * 8];
        cudaStreamSynchronize(s);
        cudaMemcpyFromSymbol(h, g_mma_timing, sizeof h);
        const int np = (int)((ntiles + grid - 1) / grid) < 16 ? (int)((ntiles + grid - 1) / grid) : 16;
        static const char *names[7] = {"wait at barrier 1", "basis", "wait at barrier 2", "A fragments", "first slab wait", "step loop", "drain + store"};
        fprintf(stderr, "mma timing <KC %d, MT %d, NW %d>: %d slabs of %d rows, %d stages; cycles per phase, CTA 0, passes 2..%d: mean over warps [min, max]\n",
                KC, MT, NW, a.nslabs, a.rows_per_slab, a.ns, np - 1);
        for (int ph = 0; ph < 7; ++ph) {
            double sum = 0; long long mn = 1ll << 60, mx = 0; int cnt = 0;
            for (int p = 2; p < np; ++p)
                for (int w = 0; w < NW && w < 16; ++w) {
                    const long long d = h[(p * 16 + w) * 8 + ph + 1] - h[(p * 16 + w) * 8 + ph];
                    sum += d; mn = d < mn ? d : mn; mx = d > mx ? d : mx; ++cnt;
                }
            if (cnt) fprintf(stderr, "  %-18s %9.0f [%lld, %lld]\n", names[ph], sum / cnt, mn, mx);
        }
        for (int p = 2; p < np && p < 6; ++p) {
            long long a0 = 1ll << 60, e0 = 1ll << 60, e1 = 0, b0 = 1ll << 60, pe0 = 1ll << 60;
            for (int w = 0; w < NW && w < 16; ++w) {
                const long long *q = h + (p * 16 + w) * 8, v = h[((p - 1) * 16 + w) * 8 + 6];
                a0 = q[0] < a0 ? q[0] : a0; e0 = q[6] < e0 ? q[6] : e0; e1 = q[6] > e1 ? q[6] : e1; b0 = q[4] < b0 ? q[4] : b0; pe0 = v < pe0 ? v : pe0;
            }
            fprintf(stderr, "  pass %d: warps reach the pass end over %lld cycles; from the first warp leaving the previous step loop to the first warp entering this one: %lld; pass length %lld\n",
                    p, e1 - e0, b0 - pe0, e1 - a0);
        }
    }
#endif
    if (li) { li->grid = grid; li->block = NW * 32; li->smem = smem; li->kernel = a.kind == 0 ? K_CHEB_MMA : K_FH_MMA; }
    return cudaGetLastError();
}

}  // namespace mma_detail
