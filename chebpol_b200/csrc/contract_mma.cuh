#pragma once
// contract_mma.cuh -- FP64 tensor-core (DMMA) contraction engine for sm_100a.
//
// Serves the two dense evaluators of the path:
//   Chebyshev  f(x) = sum_i c[i_0..i_{D-1}] prod_d T_{i_d}(x_d)     (C_evalcheb, src/chebpol.c:314-338)
//   Floater-Hormann  nested barycentric rational, whose value is the same
//              contraction with b_d[i] = (w_i/(x_d-k_i)) / sum_j w_j/(x_d-k_j)
//              and a one-hot b_d on a knot hit                      (FH, src/chebpol.c:182-218)
// Both are  f(x) = sum_r W_r(x) * ( sum_k A_k(x) * C[r][k] )  with k running over the
// first `nfold` dimensions (folded, product basis A_k) and r over the rest (weights W_r).
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
//  * per pair of 8-row N-tiles: KC x (2 B-fragment LDS.64 + 2*MT DMMA), then the epilogue
//    multiplies the 2x2 D elements of each point tile by W_r (product of Tb entries of
//    the remaining dims) into a per-lane accumulator; the four lanes of a quad are summed
//    by two shuffles at the end of the pass.
//
// Rounding: explicit-basis sums (not Clenshaw / not the reference's num/den recursion):
// differences are O(n*eps*sum|c*basis|), inside 1e-12*max|f| (tests/test_parity_gpu.py).
#include "common.cuh"
#include "ptx.cuh"
#include <float.h>

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

// Floater-Hormann basis b_i = q_i / sum q, q_i = w_i * prod_{j != i} (x - k_j) * s^(n-1)
// (prefix/suffix products, one reciprocal), which equals (w_i/(x-k_i)) / sum_j (w_j/(x-k_j)).
// Knot hit (first i with |x-k_i| < 10*DBL_EPSILON, src/chebpol.c:195-197): one-hot.
template <int NPT>
__device__ __forceinline__ void fh_basis(double *col, int n, double x, const double *__restrict__ kn,
                                         const double *__restrict__ w)
{
    int hit = -1;
    for (int i = 0; i < n; ++i)
        if (fabs(x - __ldg(kn + i)) < 10.0 * DBL_EPSILON) { hit = i; break; }
    if (hit >= 0 || x != x) {
        for (int i = 0; i < n; ++i) col[(size_t)i * NPT] = (i == hit) ? 1.0 : ((x != x) ? x : 0.0);
        return;
    }
    // distances scaled so that the products stay near 1 inside the grid
    const double k0 = __ldg(kn), k1 = __ldg(kn + n - 1);
    const double span = fabs(k1 - k0);
    const double inv_s = (n > 1 && span > 0.0) ? 4.0 / span : 1.0;
    double pre = 1.0;
    for (int i = 0; i < n; ++i) {
        col[(size_t)i * NPT] = pre;
        pre *= (x - __ldg(kn + i)) * inv_s;
    }
    double suf = 1.0, sum = 0.0;
    for (int i = n - 1; i >= 0; --i) {
        const double q = __ldg(w + i) * col[(size_t)i * NPT] * suf;
        col[(size_t)i * NPT] = q;
        sum += q;
        suf *= (x - __ldg(kn + i)) * inv_s;
    }
    const double r = 1.0 / sum;
    if (isfinite(r) && isfinite(sum) && r != 0.0) {
        for (int i = 0; i < n; ++i) col[(size_t)i * NPT] *= r;
    } else {
        // products over/underflowed (extreme extrapolation or very long grids): the
        // reference's own form, one division per knot
        double den = 0.0;
        for (int i = 0; i < n; ++i) {
            const double p = __ldg(w + i) / (x - __ldg(kn + i));
            col[(size_t)i * NPT] = p;
            den += p;
        }
        for (int i = 0; i < n; ++i) col[(size_t)i * NPT] /= den;
    }
}

// KC: K/4 (padded); MT: 8-point tiles per warp; NW: warps; LEVU: compile-time bound (2 or 4) on
// the number of unfolded ("rest") dims -- absent levels multiply by 1.0 so the epilogue
// stays branch-free.
template <int KC, int MT, int NW, int LEVU>
__global__ void __launch_bounds__(NW * 32, 1) contract_mma_kernel(const __grid_constant__ MmaArgs a)
{
    constexpr int NPT = NW * MT * 8;  // points per pass
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem);
    int *done_cnt = reinterpret_cast<int *>(smem + kMaxStages * 8);
    unsigned char *ring = smem + kBarBytes;
    double *Tb = reinterpret_cast<double *>(ring + (size_t)a.ns * a.stage_bytes);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int rank = a.rank, nlev = a.nlev, ns = a.ns;
    const int64_t ntiles = (a.npts + NPT - 1) / NPT;
    const int64_t my_tiles = (ntiles > (int64_t)blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t total_fills = a.resident ? (int64_t)a.nslabs : my_tiles * a.nslabs;
    const int rps = a.rows_per_slab;
    const uint32_t row_bytes = (uint32_t)a.rs * 8u;

    auto slab_rows = [&](int s) { const int left = a.nrest - s * rps; return left < rps ? left : rps; };
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

    int stage = 0;
    uint32_t round = 0;
    int64_t fill = 0;

    // coordinates of the next pass are fetched (into registers) while the current pass
    // computes, so the basis prologue never waits on DRAM.  Task q of a thread is
    // (point pt, dim L) = ((tid + q*NT) % NPT, (tid + q*NT) / NPT); the coordinates of a
    // pass are the contiguous range x[tile*NPT*rank ...), which task index `pt*rank + L`
    // addresses directly.
    constexpr int NT = NW * 32;
    constexpr int kMaxTasks = (NPT * (3 + LEVU) + NT - 1) / NT;  // rank <= nfold(3) + LEVU
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

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        __syncthreads();  // previous pass has finished reading Tb (and, first time, barrier init is visible)
        // ---- 1. basis vectors of this pass's points
#pragma unroll
        for (int q = 0; q < kMaxTasks; ++q) {
            const int task = tid + q * NT;
            if (task >= NPT * rank) break;
            const int pt = task % NPT, L = task / NPT;
            const int64_t p = tile * NPT + pt;
            double *col = Tb + (size_t)a.boff[L] * NPT + pt;
            const int n = a.dims[L];
            if (p < a.npts) {
                const double xv = xpre[q];
                if (a.kind == 0) cheb_basis<NPT>(col, n, premap_coord(a.pm, L, xv));
                else fh_basis<NPT>(col, n, xv, a.grid + a.goff[L], a.weights + a.goff[L]);
            } else {
                for (int i = 0; i < n; ++i) col[(size_t)i * NPT] = 0.0;
            }
        }
        prefetch_x(tile + gridDim.x);
        __syncthreads();

        // ---- 2. A fragments: A[mt][kc] = prod_{l < nfold} Tb_l[digit_l(k)][point], k = 4*kc + t;
        //         the digit rows come from a host-built table (no integer division here)
        double A[MT][KC];
#pragma unroll
        for (int kc = 0; kc < KC; ++kc) {
            const int *kt = a.ktab + (4 * kc + t) * 3;
            const int r0 = __ldg(kt), r1 = __ldg(kt + 1), r2 = __ldg(kt + 2);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                const int pt = (warp * MT + mt) * 8 + g;
                double v = 0.0;
                if (r0 >= 0) {
                    v = Tb[(size_t)r0 * NPT + pt];
                    if (r1 >= 0) v *= Tb[(size_t)r1 * NPT + pt];
                    if (r2 >= 0) v *= Tb[(size_t)r2 * NPT + pt];
                }
                A[mt][kc] = v;
            }
        }

        double acc[MT];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) acc[mt] = 0.0;

        // ---- 3. stream the tensor rows.  Software pipeline: while the DMMAs of step q are
        // in flight, the D tiles of step q-1 are weighted into `acc` (branch-free, so ptxas
        // can interleave the two).  A step = 16 tensor rows (two 8-row N-tiles).
        double Dp[2][MT][2];          // D tiles of the previous step
        int ip[4][LEVU];             // basis-table rows of its 4 columns' rest-level digits
        bool vp[4];                   // column valid (row < nrest)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) Dp[j][mt][0] = Dp[j][mt][1] = 0.0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            vp[c] = false;
#pragma unroll
            for (int l = 0; l < LEVU; ++l) ip[c][l] = 0;
        }
        const double *tcol = Tb + warp * MT * 8 + g;

        auto weigh_pending = [&]() {
            // per column: its table loads are independent of each other (and of the DMMAs)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                double tv[LEVU][MT];
#pragma unroll
                for (int l = 0; l < LEVU; ++l)
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt)
                        tv[l][mt] = (l < nlev) ? tcol[(size_t)ip[c][l] * NPT + mt * 8] : 1.0;
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    double w = tv[0][mt];
#pragma unroll
                    for (int l = 1; l < LEVU; ++l) w *= tv[l][mt];  // absent levels are 1.0: keeps the block branch-free
                    const double nv = fma(w, Dp[c >> 1][mt][c & 1], acc[mt]);
                    acc[mt] = vp[c] ? nv : acc[mt];
                }
            }
        };

        for (int s = 0; s < a.nslabs; ++s) {
            const int st = a.resident ? s : stage;
            mbar_wait(&full_bar[st], a.resident ? 0u : (round & 1));
            const double *slab = reinterpret_cast<const double *>(ring + (size_t)st * a.stage_bytes);
            const int rows = slab_rows(s);

            for (int r16 = 0; r16 < rows; r16 += 16) {
                double D[2][MT][2];
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) D[j][mt][0] = D[j][mt][1] = 0.0;
                const double *b0p = slab + (size_t)(r16 + g) * a.rs + t;
                const double *b1p = b0p + (size_t)8 * a.rs;
                // this step's row info (column c = 2*j + e -> local row r16 + 8j + 2t + e),
                // clamped to a valid row so the table index is always in range
                int ic[4][LEVU];
                bool vc[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int lr = r16 + (c >> 1) * 8 + 2 * t + (c & 1);
                    vc[c] = lr < rows;
                    const int lrc = vc[c] ? lr : rows - 1;
                    const int *info = reinterpret_cast<const int *>(slab + (size_t)lrc * a.rs + a.kp);
#pragma unroll
                    for (int l = 0; l < LEVU; ++l) ic[c][l] = (l < nlev) ? info[l] : 0;
                }
#pragma unroll
                for (int kc = 0; kc < KC; ++kc) {
                    const double b0 = b0p[4 * kc];
                    const double b1 = b1p[4 * kc];
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        dmma884(D[0][mt][0], D[0][mt][1], A[mt][kc], b0);
                        dmma884(D[1][mt][0], D[1][mt][1], A[mt][kc], b1);
                    }
                }
                weigh_pending();
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        Dp[j][mt][0] = D[j][mt][0];
                        Dp[j][mt][1] = D[j][mt][1];
                    }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    vp[c] = vc[c];
#pragma unroll
                    for (int l = 0; l < LEVU; ++l) ip[c][l] = ic[c][l];
                }
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
        weigh_pending();  // drain

        // ---- 4. sum the four lanes of each quad, store
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            double v = acc[mt];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            const int64_t p = tile * NPT + (warp * MT + mt) * 8 + g;
            if (t == 0 && p < a.npts) a.out[p] = v;
        }
    }
}

}  // namespace mma_detail
bool mma_fit_ring(MmaArgs &a, int npt);
namespace mma_detail {

template <int KC, int MT, int NW, int LEVU>
cudaError_t launch_one(const MmaArgs &a_in, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    auto kern = contract_mma_kernel<KC, MT, NW, LEVU>;
    constexpr int NPT = NW * MT * 8;
    MmaArgs a = a_in;
    if (a.nlev > LEVU || !mma_fit_ring(a, NPT)) return cudaErrorNotSupported;
    const int smem = kBarBytes + a.ns * a.stage_bytes + a.sumn * NPT * 8;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    const int64_t ntiles = (a.npts + NPT - 1) / NPT;
    int grid = (int)(ntiles < num_sms ? ntiles : num_sms);
    if (grid < 1) grid = 1;
    kern<<<grid, NW * 32, smem, s>>>(a);
    if (li) { li->grid = grid; li->block = NW * 32; li->smem = smem; li->kernel = a.kind == 0 ? K_CHEB_MMA : K_FH_MMA; }
    return cudaGetLastError();
}

}  // namespace mma_detail
