// common.cuh -- shared declarations of libchebpol_cuda (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/chebpol_cuda.h"

#define CP_MAXR 32  // CHEBPOL_CUDA_MAXRANK + 1

// NVTX ranges around the host-side phases (plan build, staged evaluation, list building) so a timeline
// tool shows where a call spends its time; header-only NVTX3 costs nothing when no tool is attached.
#include <nvtx3/nvToolsExt.h>
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

// Pre-map applied to every coordinate before evaluation (Chebyshev plans):
//   bit 0: x <- (x - mid)*ispan                 R/chebyshev.R:223-227 (imap)
//   bit 1: x <- sin(0.5*pi*x*(1-n)/n)           R/uniform.R:6 (ugm)
//   bit 2: x <- gridmaps[[d]](x)                R/chebyshev.R:349-354 (chebappxg.real): the monotone
//          cubic spline stats::splinefun(grid, chebknots, method='hyman') of dimension d, evaluated
//          as R's spline_eval does for a scalar call.  Table of dimension d at spl + soff[d]:
//          sn knots (padded to a multiple of 4 doubles), then sn records [y, b, c, d].
struct PreMap {
    int mode;
    double mid[CP_MAXR];
    double ispan[CP_MAXR];
    double nd[CP_MAXR];
    const double *spl;
    int soff[CP_MAXR];
    int sn[CP_MAXR];
};

__device__ __forceinline__ double premap_spline(const PreMap &pm, int d, double u)
{
    const double *__restrict__ sx = pm.spl + pm.soff[d];
    const int n = pm.sn[d];
    // a scalar call of the closure starts its search at i = 0 and keeps it when x[0] <= u <= x[1]
    // (NaN keeps it too, and propagates through dx); else bisection for the last knot <= u
    int i = 0;
    if (u < __ldg(sx) || __ldg(sx + 1) < u) {
        int j = n;
        do {
            const int k = (i + j) >> 1;
            if (u < __ldg(sx + k)) j = k; else i = k;
        } while (j > i + 1);
    }
    const double2 *rec = reinterpret_cast<const double2 *>(sx + ((n + 3) & ~3) + 4 * i);
    const double2 yb = __ldg(rec), cd = __ldg(rec + 1);
    const double dx = __dsub_rn(u, __ldg(sx + i));
    // Horner in R's order, no fused multiply-add: the mapped coordinate is bit-identical in every kernel
    double t = __dadd_rn(cd.x, __dmul_rn(dx, cd.y));
    t = __dadd_rn(yb.y, __dmul_rn(dx, t));
    return __dadd_rn(yb.x, __dmul_rn(dx, t));
}

__device__ __forceinline__ double premap_coord(const PreMap &pm, int d, double v)
{
    if (pm.mode & 1) v = (v - pm.mid[d]) * pm.ispan[d];
    if (pm.mode & 2) {
        // R evaluates 0.5*pi*x*(1-n)/n left to right
        const double n = pm.nd[d];
        v = sin(0.5 * 3.141592653589793 * v * (1.0 - n) / n);
    }
#ifndef CP_NO_SPLINE  // A/B switch for measuring what the extra prologue branch costs the small kernels
    if (pm.mode & 4) v = premap_spline(pm, d, v);
#endif
    return v;
}

// ---- kernel ids reported through CHEBPOL_CUDA_GET_KERNEL_USED ------------
enum KernelId {
    K_NONE = 0,
    K_CHEB_STRICT = 10,  // thread-per-point Clenshaw, reference op order, no FMA
    K_CHEB_SLAB = 11,    // basis-in-registers DFMA contraction, TMA-streamed slabs
    K_CHEB_MMA = 12,     // DMMA (FP64 tensor core) contraction, TMA-streamed rows
    K_MLIP_STRICT = 20,  // thread-per-point, reference op order
    K_MLIP_PAIR = 21,    // lane-pair gather kernel
    K_MLIP_CELL = 22,    // cell-major table: one contiguous 2^D block per point
    K_MLIP_BIN = 24,     // points binned by upper-dimension cell, sub-table staged in shared memory (no extra table memory)
    K_MLIP_QUAD = 23,    // row-pair-interleaved table (2x): one 32-byte quad per (dim 0, dim 1) corner set
    K_FH_STRICT = 30,    // thread-per-point nested barycentric, reference op order
    K_FH_MMA = 32,       // barycentric-basis DMMA contraction, TMA-streamed rows
    K_STALK_STRICT = 50, // stalker / hstalker, thread per point, reference op order
    K_STALK_CELL = 51,   // packed per-grid-point records, blend weights hoisted, fast reciprocal
    K_POLYH_STRICT = 40, // thread-per-point radial sum, reference op order
    K_POLYH_TILE = 41,   // points in registers, knot rows broadcast from a TMA-filled ring
    K_OLDSTALK = 52,     // deprecated recursive stalker evaluator, thread per point, reference op order
    K_SL_STRICT = 60,    // simplex-linear: the reference's linear scan over all simplices
    K_SL_CELL = 61,      // simplex-linear: candidates from a uniform cell grid, same acceptance test
};

// ---- launch descriptors filled by api.cu, consumed by the kernel TUs -------
struct ChebStrictArgs {
    const double *cf;  // original layout (R array)
    int rank;
    int dims[CP_MAXR];
    int64_t stride[CP_MAXR];
    const double *x;
    int64_t npts;
    double *out;
    PreMap pm;
};

struct MlipArgs {
    const double *values;
    const double *grid;  // all knot vectors, concatenated
    int goff[CP_MAXR];   // offset of grid d in `grid`
    int rank;
    int dims[CP_MAXR];
    int64_t stride[CP_MAXR];
    int blend;
    const double *x;
    int64_t npts;
    double *out;
    // bucket table that accelerates the cell search (mlip_cell.cu): per dim, lut[b] is the
    // reference bisection's answer for the left edge of bucket b of [g_0, g_{n-1}]
    const int *lut;
    int loff[CP_MAXR];     // offset of dim d's table in `lut`
    int nb[CP_MAXR];       // buckets of dim d
    double lscale[CP_MAXR];  // buckets per unit length: nb / (g_{n-1} - g_0)
};

// Stalker splines (evalstalk src/stalker.c:366-442, evalhyp :537-613): cell lookup, 2^rank corners,
// per corner and dimension a one-dimensional basis with tabulated parameters.
struct StalkArgs {
    int kind;               // 0 stalker: b z + c |z|^r;  1 hstalker: b z + d/(1 + c z), NaN c: b z^2
    const double *val;      // prod(dims), R layout
    const double *a;        // hstalker only: prod(dims)
    const double *b, *t1, *t2;  // rank x prod(dims) each (entry j of grid point i at [rank*i + j]); t1 = c, t2 = r or d
    const double *records;  // packed per grid point: [val (+a), b_0.., t1_0.., t2_0.., pad], see stalker.cu
    const double *grid;     // all knot vectors, concatenated
    int goff[CP_MAXR];
    int rank;
    int dims[CP_MAXR];
    int64_t stride[CP_MAXR];
    int blend;
    const double *x;
    int64_t npts;
    double *out;
};

// Polyharmonic spline evaluation (C_evalpolyh, src/chebpol.c:383-415)
struct PolyhArgs {
    const double *knots;    // rank x nknots, knot i at knots + i*rank (R matrix)
    const double *weights;  // nknots
    const double *lweights; // rank + 1: constant, then one per coordinate
    const double *packed;   // nknots rows of polyh_row_stride(rank) doubles: w, coordinates, zero pad
    int rank, nknots;
    double k;               // < 0: exp(k r^2); odd (int)k: r^k; even: log(r) r^k
    const double *x;
    int64_t npts;
    double *out;
    PreMap pm;              // bit 0: x <- (x - b)*a, the closure's normfun (R/polyharmonic.R:77,82)
};

struct FhArgs {
    const double *vals;
    const double *grid;     // concatenated knots
    const double *weights;  // concatenated weights, same offsets
    int goff[CP_MAXR];
    int rank;
    int dims[CP_MAXR];
    int64_t stride[CP_MAXR];
    const double *x;
    int64_t npts;
    double *out;
};

// Slab-streamed contraction kernels (Chebyshev fast path, FH fast path).
// The tensor is stored padded: dim 0 rounded up to an even length n0p so
// every row is 16-byte aligned; a slab is the sub-tensor spanned by the first
// `depth` dims (slab_elems doubles); slabs are consumed last-to-first.
struct SlabArgs {
    const double *cf;  // padded tensor
    int rank;
    int depth;         // dims inside a slab (1..3)
    int n0, n0p;
    int dims[CP_MAXR];
    int slab_elems;
    int nslabs;        // < 2^31 (R arrays have < 2^31 elements)
    int ns;            // ring stages in shared memory
    int stage_stride;  // bytes between stages (slab bytes rounded up to 128)
    int resident;      // nslabs <= ns: every slab loaded once per CTA
    unsigned cum[CP_MAXR];  // cum[L] = prod_{j=depth..L} dims[j], L >= depth
    const double *x;
    int64_t npts;
    double *out;
    PreMap pm;
    // FH only
    const double *grid;
    const double *weights;
    int goff[CP_MAXR];
};

struct LaunchInfo {
    int grid, block, smem, kernel;
};

// DMMA contraction engine (contract_mma.cuh):
//   f(x) = sum_{r'} W_{r'}(x) * sum_{iA} b_A[iA](x) * sum_k A_k(x) * C[(iA, r')][k]
// k runs over the first `nfold` dims (folded into the MMA K dimension, product basis A_k),
// iA over the next dim ("level A"), r' over the remaining <= 3 dims (weights W_{r'}); b, A, W
// are (products of) per-dimension basis values: Chebyshev T_i(x_d) or Floater-Hormann
// barycentric weights.
struct MmaArgs {
    const double *tensor;  // [nrows][rs] doubles, rows ordered (r'-tile of 8, iA, r' within the tile)
    int kind;              // 0 Chebyshev basis, 1 Floater-Hormann basis
    int rank;
    int dims[CP_MAXR];
    int boff[CP_MAXR];     // row offset of dim d in the shared-memory basis table (prefix sum of dims)
    int sumn;              // sum of dims; table row `sumn` holds ones (used when a level is absent)
    int nfold;             // leading dims folded into K
    int K;                 // prod of folded dims
    int rs;                // row stride in doubles (== 4 or 12 mod 16: conflict-free B loads)
    int kp;                // 4*KC: padded K
    int nlev;              // rank - nfold: unfolded dims
    int nA, nAp;           // size of level A (1 if absent) and its padding to a multiple of ntb
    int occ;               // CTAs per SM the ring is sized for (1 or 2)
    int ntb;               // N-tiles (consecutive level-A indices) per step: 2, or 4 when K is small
    int boffA;             // basis-table row of level A (the ones row if absent)
    int nrp;               // number of r' values (prod of the dims after level A; 1 if none)
    int nlg;               // group levels: max(nlev - 1, 0) <= 3
    int nrows;             // tensor rows: ceil(nrp/8) * nAp * 8 (a multiple of 8*ntb)
    const int *ktab;       // [kp][3] basis-table rows of the folded dims for each k (the ones row: padding k, absent dims)
    const int *ginfo;      // [8*ceil(nrp/8)][3] basis-table rows of the group levels of each r' (ones row: absent / padding)
    int fh_local, gi_local; // set by the launcher: FH knots + weights / the ginfo table are copied to shared memory (they fit
                           // the few KB beside the ring budget); otherwise they are read from global memory
    int rows_per_slab;     // multiple of 8*ntb
    int nslabs, ns, stage_bytes, resident;
    const double *x;
    int64_t npts;
    double *out;
    PreMap pm;
    const double *grid;    // FH: concatenated knots / weights
    const double *weights;
    int goff[CP_MAXR];
};
// Rows of long K store the coefficients of two consecutive k-chunks pair-interleaved, so that one LDS.128
// fetches a lane's B fragments of both (position of k = 4*kc + t: 8*(kc/2) + 2*t + kc%2; an odd last chunk
// stays in place); the row stride is then 8 mod 16 doubles (conflict-free 128-bit loads).  Measured on the
// step loop in isolation (tools/mma_lab.cu): KC = 25 0.948 -> 0.965 of the FP64 peak, KC = 10 0.928 -> 0.923.
__host__ __device__ constexpr bool mma_pair_rows(int kc) { return kc >= 20; }
__host__ __device__ constexpr int mma_row_pos(int kc_total, int k)
{
    const int kc = k / 4, t = k % 4;
    return (mma_pair_rows(kc_total) && kc < (kc_total & ~1)) ? 8 * (kc / 2) + 2 * t + (kc & 1) : k;
}
bool mma_configure(MmaArgs &a, int kind, int rank, const int *dims, int *kc_out);
cudaError_t launch_contract_mma(const MmaArgs &a, int variant, int num_sms, cudaStream_t s, LaunchInfo *li);

// Launchers implemented in the kernel TUs.  Each returns cudaGetLastError().
cudaError_t launch_cheb_strict(const ChebStrictArgs &a, cudaStream_t s, LaunchInfo *li);
cudaError_t launch_mlip_strict(const MlipArgs &a, cudaStream_t s, LaunchInfo *li);
cudaError_t launch_fh_strict(const FhArgs &a, cudaStream_t s, LaunchInfo *li);
cudaError_t launch_polyh_strict(const PolyhArgs &a, cudaStream_t s, LaunchInfo *li);
cudaError_t launch_stalk_strict(const StalkArgs &a, cudaStream_t s, LaunchInfo *li);
// fast stalker kernel (stalker.cu): rank <= 6; record stride in doubles for a rank and kind
cudaError_t launch_stalk_cell(const StalkArgs &a, int variant, int num_sms, cudaStream_t s, LaunchInfo *li);
int stalk_record_stride(int rank, int kind);
// fast polyharmonic kernel (polyh.cu); cudaErrorNotSupported for rank > 32
cudaError_t launch_polyh_tile(const PolyhArgs &a, int variant, int num_sms, cudaStream_t s, LaunchInfo *li);
// row stride (doubles) of the packed knot table the tile kernel reads: [w, k_0..k_{rank-1}, 0...]
int polyh_row_stride(int rank);
// returns cudaErrorNotSupported when the shape has no fast instantiation
cudaError_t launch_cheb_slab(const SlabArgs &a, int variant, int num_sms, cudaStream_t s, LaunchInfo *li);
int cheb_slab_pick_n0p(int n0);
cudaError_t launch_mlip_pair(const MlipArgs &a, int variant, int num_sms, cudaStream_t s, LaunchInfo *li);
bool mlip_pair_supported(int rank, int blend);
int64_t mlip_cell_bytes(int rank, const int *dims, int m);
cudaError_t mlip_expand(const MlipArgs &a, const double *values, double *cells, int m, cudaStream_t s);
cudaError_t launch_mlip_cell(const MlipArgs &a, const double *cells, int m, int variant, int num_sms, cudaStream_t s, LaunchInfo *li);
struct MlipBinState;
bool mlip_bin_supported(int rank, const int *dims, int64_t npts, int blend, bool force);
cudaError_t mlip_bin_prepare(MlipBinState **st, int rank, const int *dims, int64_t npts);
void mlip_bin_destroy(MlipBinState *st);
int64_t mlip_bin_bytes(const MlipBinState *st);
cudaError_t launch_mlip_bin(const MlipArgs &a, MlipBinState *st, int num_sms, cudaStream_t s, LaunchInfo *li, int *launches);
int64_t mlip_quad_bytes(int rank, const int *dims);
cudaError_t mlip_quad_expand(const MlipArgs &a, const double *values, double *quads, cudaStream_t s);
cudaError_t launch_mlip_quad(const MlipArgs &a, const double *quads, int variant, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t fit_chebcoef_device(double *d_a, double *d_b, const int *dims, int rank, int dct, cudaStream_t s,
                                double **result, int *launches);
cudaError_t fit_evalgrid_device(int kind, const double *d_tensor, const int *dims, int rank, const double *const *d_pts,
                                const int *gdims, const double *const *d_knots, const double *const *d_weights, int blend,
                                const PreMap *pm, double *d_s0, double *d_s1, cudaStream_t s, double **result, int *launches);
cudaError_t fit_fhweights_device(const double *d_grid, int nknots, int d, double *d_w, cudaStream_t s);
// counter-based synthetic points (synth.cu)
cudaError_t synth_fill_points(uint64_t seed, int64_t first_point, int64_t npts, int rank, double lo, double hi, double *d_x,
                              int num_sms, cudaStream_t s);
cudaError_t run_fp64_peak(int num_sms, int repeats, cudaStream_t s, double *tflops);
// deprecated recursive stalker evaluator (oldstalk.cu)
cudaError_t launch_oldstalk(const double *val, const double *grid, const double *det, const double *pmin,
                            const double *pplus, const int *goff, int rank, const int *dims, const int64_t *stride,
                            const double *degree, const int *uniform, int blend, const double *x, int64_t npts,
                            double *out, int num_sms, cudaStream_t s, LaunchInfo *li);
// simplex-linear interpolation (simplex.cu): device tables of one triangulation
struct SlDevice;
cudaError_t sl_create(const double *knots, int dim, int nknots, const int *dtri, int ns, const double *lumats,
                      const int *pivots, const double *bbox, const double *val, int cells_per_dim, SlDevice **out);
void sl_destroy(SlDevice *d);
int64_t sl_bytes(const SlDevice *d);
cudaError_t sl_force_cells(SlDevice *d, int cells_per_dim);
cudaError_t launch_sl(SlDevice *d, int kernel, int variant, int epol, int blend, const double *x, int64_t npts,
                      double *out, int num_sms, cudaStream_t s, LaunchInfo *li);
int sl_debug_lists(const double *knots, int dim, int nknots, const int *dtri, int ns, const double *bbox, int G,
                   int *G_used, double *lo, double *inv, int *start, int64_t start_cap, int *items, int64_t items_cap,
                   int64_t *nitems);
cudaError_t sl_analyze_device(const int *d_dtri, int ns, const double *d_knots, int dim, double *d_lumats,
                              int *d_pivots, double *d_bbox, cudaStream_t s);
