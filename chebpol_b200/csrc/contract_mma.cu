// contract_mma.cu -- configuration and dispatch of the DMMA contraction engine
// (kernel: contract_mma.cuh; instantiations: contract_mma_inst*.cu).
#include "common.cuh"
#include <cstdlib>

namespace {
constexpr int kMaxStages = 8;
constexpr int kBarBytes = 128;
constexpr int kSmemBudget = 220 * 1024;
constexpr int kSmemBudget2 = 111 * 1024;  // two CTAs per SM (228 KB per SM, 1 KB reserved per CTA)
constexpr int kLevMax = 4;
}  // namespace

cudaError_t mma_launch_2_1_4_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_2_1_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_2_1_16_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_2_2_4_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_2_2_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_2_2_12_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_2_4_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_3_1_4_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_3_1_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_3_1_16_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_3_2_4_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_3_2_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_3_2_12_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_3_4_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_4_1_4_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_4_1_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_4_1_16_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_4_2_4_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_4_2_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_4_2_12_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_4_4_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_6_1_4_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_6_1_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_6_1_16_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_6_2_4_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_6_2_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_6_2_12_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_6_4_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_8_1_4_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_8_1_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_8_1_16_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_8_2_4_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_8_2_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_8_2_12_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_8_4_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_10_1_4_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_10_1_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_10_1_16_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_10_2_4_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_10_2_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_10_2_12_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_10_4_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_12_1_4_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_12_1_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_12_1_16_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_12_2_4_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_12_2_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_12_2_12_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_12_4_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_16_1_4_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_16_1_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_16_1_16_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_16_2_4_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_16_2_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_16_3_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_20_1_4_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_20_1_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_20_1_16_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_20_2_4_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_20_2_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_25_1_4_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_25_1_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_25_1_16_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_25_2_4_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_25_2_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_30_1_4_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_30_1_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_30_2_4_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_30_2_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_36_1_4_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_36_1_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_36_1_16_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_36_2_4_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_36_2_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_40_1_4_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_40_1_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_48_1_4_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);
cudaError_t mma_launch_48_1_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li);

namespace {
typedef cudaError_t (*LaunchFn)(const MmaArgs &, int, cudaStream_t, LaunchInfo *);
struct Inst { int kc, mt, nw, ntb; LaunchFn fn; };
const Inst kInsts[] = {
    {2, 1, 4, 2, mma_launch_2_1_4_2},
    {2, 1, 8, 2, mma_launch_2_1_8_2},
    {2, 1, 16, 2, mma_launch_2_1_16_2},
    {2, 2, 4, 2, mma_launch_2_2_4_2},
    {2, 2, 8, 2, mma_launch_2_2_8_2},
    {2, 2, 12, 2, mma_launch_2_2_12_2},
    {2, 4, 8, 2, mma_launch_2_4_8_2},
    {3, 1, 4, 2, mma_launch_3_1_4_2},
    {3, 1, 8, 2, mma_launch_3_1_8_2},
    {3, 1, 16, 2, mma_launch_3_1_16_2},
    {3, 2, 4, 2, mma_launch_3_2_4_2},
    {3, 2, 8, 2, mma_launch_3_2_8_2},
    {3, 2, 12, 2, mma_launch_3_2_12_2},
    {3, 4, 8, 2, mma_launch_3_4_8_2},
    {4, 1, 4, 2, mma_launch_4_1_4_2},
    {4, 1, 8, 2, mma_launch_4_1_8_2},
    {4, 1, 16, 2, mma_launch_4_1_16_2},
    {4, 2, 4, 2, mma_launch_4_2_4_2},
    {4, 2, 8, 2, mma_launch_4_2_8_2},
    {4, 2, 12, 2, mma_launch_4_2_12_2},
    {4, 4, 8, 2, mma_launch_4_4_8_2},
    {6, 1, 4, 2, mma_launch_6_1_4_2},
    {6, 1, 8, 2, mma_launch_6_1_8_2},
    {6, 1, 16, 2, mma_launch_6_1_16_2},
    {6, 2, 4, 2, mma_launch_6_2_4_2},
    {6, 2, 8, 2, mma_launch_6_2_8_2},
    {6, 2, 12, 2, mma_launch_6_2_12_2},
    {6, 4, 8, 2, mma_launch_6_4_8_2},
    {8, 1, 4, 2, mma_launch_8_1_4_2},
    {8, 1, 8, 2, mma_launch_8_1_8_2},
    {8, 1, 16, 2, mma_launch_8_1_16_2},
    {8, 2, 4, 2, mma_launch_8_2_4_2},
    {8, 2, 8, 2, mma_launch_8_2_8_2},
    {8, 2, 12, 2, mma_launch_8_2_12_2},
    {8, 4, 8, 2, mma_launch_8_4_8_2},
    {10, 1, 4, 2, mma_launch_10_1_4_2},
    {10, 1, 8, 2, mma_launch_10_1_8_2},
    {10, 1, 16, 2, mma_launch_10_1_16_2},
    {10, 2, 4, 2, mma_launch_10_2_4_2},
    {10, 2, 8, 2, mma_launch_10_2_8_2},
    {10, 2, 12, 2, mma_launch_10_2_12_2},
    {10, 4, 8, 2, mma_launch_10_4_8_2},
    {12, 1, 4, 2, mma_launch_12_1_4_2},
    {12, 1, 8, 2, mma_launch_12_1_8_2},
    {12, 1, 16, 2, mma_launch_12_1_16_2},
    {12, 2, 4, 2, mma_launch_12_2_4_2},
    {12, 2, 8, 2, mma_launch_12_2_8_2},
    {12, 2, 12, 2, mma_launch_12_2_12_2},
    {12, 4, 8, 2, mma_launch_12_4_8_2},
    {16, 1, 4, 2, mma_launch_16_1_4_2},
    {16, 1, 8, 2, mma_launch_16_1_8_2},
    {16, 1, 16, 2, mma_launch_16_1_16_2},
    {16, 2, 4, 2, mma_launch_16_2_4_2},
    {16, 2, 8, 2, mma_launch_16_2_8_2},
    {16, 3, 8, 2, mma_launch_16_3_8_2},
    {20, 1, 4, 2, mma_launch_20_1_4_2},
    {20, 1, 8, 2, mma_launch_20_1_8_2},
    {20, 1, 16, 2, mma_launch_20_1_16_2},
    {20, 2, 4, 2, mma_launch_20_2_4_2},
    {20, 2, 8, 2, mma_launch_20_2_8_2},
    {25, 1, 4, 2, mma_launch_25_1_4_2},
    {25, 1, 8, 2, mma_launch_25_1_8_2},
    {25, 1, 16, 2, mma_launch_25_1_16_2},
    {25, 2, 4, 2, mma_launch_25_2_4_2},
    {25, 2, 8, 2, mma_launch_25_2_8_2},
    {30, 1, 4, 2, mma_launch_30_1_4_2},
    {30, 1, 8, 2, mma_launch_30_1_8_2},
    {30, 2, 4, 2, mma_launch_30_2_4_2},
    {30, 2, 8, 2, mma_launch_30_2_8_2},
    {36, 1, 4, 2, mma_launch_36_1_4_2},
    {36, 1, 8, 2, mma_launch_36_1_8_2},
    {36, 1, 16, 2, mma_launch_36_1_16_2},
    {36, 2, 4, 2, mma_launch_36_2_4_2},
    {36, 2, 8, 2, mma_launch_36_2_8_2},
    {40, 1, 4, 2, mma_launch_40_1_4_2},
    {40, 1, 8, 2, mma_launch_40_1_8_2},
    {48, 1, 4, 2, mma_launch_48_1_4_2},
    {48, 1, 8, 2, mma_launch_48_1_8_2}
};

// smallest instantiated KC with 4*KC >= K (0: none)
int pick_kc(int K)
{
    int best = 0;
    for (const auto &i : kInsts)
        if (4 * i.kc >= K && (best == 0 || i.kc < best)) best = i.kc;
    return best;
}
}  // namespace

// Ring geometry for `npt` points per pass (steps of 16 tensor rows): slab size, number of
// slabs, stages.  False if it does not fit in shared memory.
bool mma_fit_ring(MmaArgs &a, int npt)
{
    const int budget = a.occ == 2 ? kSmemBudget2 : kSmemBudget;
    const int left = budget - kBarBytes - (a.sumn + 1) * npt * 8;
    if (left <= 0) return false;
    const int row_bytes = a.rs * 8;
    // Two large slabs (the ring space halved, at most 128 KB each) beat three smaller ones on every shape of
    // tools/ring_sweep.py (profiles/ring_sweep_r02.jsonl: cfg2 0.954 -> 0.958, FH 40^3 0.840 -> 0.854,
    // 12^6 0.998 -> 1.047 of the FP64 peak): fewer hand-shakes per pass, and a warp can run ahead of the
    // slowest one by one slab only, which shortens the drain at the end of a pass.
    // Tuning aids: CHEBPOL_CUDA_MMA_RINGDIV (2..8 parts), CHEBPOL_CUDA_MMA_SLAB_KB (8..200).
    static const int ringdiv = [] { const char *e = getenv("CHEBPOL_CUDA_MMA_RINGDIV"); const int v = e ? atoi(e) : 2; return v >= 2 && v <= 8 ? v : 2; }();
    int slab_cap = left / ringdiv;
    static const int slab_max = [] { const char *e = getenv("CHEBPOL_CUDA_MMA_SLAB_KB"); const int v = e ? atoi(e) : 128; return (v >= 8 && v <= 200 ? v : 128) * 1024; }();
    if (slab_cap > slab_max) slab_cap = slab_max;
    const int step_rows = 8 * a.ntb;
    int64_t rps = slab_cap / row_bytes / step_rows * step_rows;
    if (rps < step_rows) rps = step_rows;
    if (rps > a.nrows) rps = a.nrows;
    a.rows_per_slab = (int)rps;
    a.nslabs = (int)((a.nrows + rps - 1) / rps);
    a.stage_bytes = (int)rps * row_bytes;
    int ns = left / a.stage_bytes;
    if (ns > kMaxStages) ns = kMaxStages;
    if (ns >= a.nslabs && ns >= 1) { a.resident = 1; a.ns = a.nslabs; return true; }
    if (ns < 2) return false;
    a.resident = 0;
    a.ns = ns;
    return true;
}

// Choose the folding and the padded row layout for a tensor.  Returns false when the
// shape is outside what the engine supports (the caller then uses another kernel).
bool mma_configure(MmaArgs &a, int kind, int rank, const int *dims, int *kc_out)
{
    a.kind = kind;
    a.rank = rank;
    int sumn = 0;
    for (int d = 0; d < rank; ++d) {
        a.dims[d] = dims[d];
        a.boff[d] = sumn;
        sumn += dims[d];
        if (kind == 1 && dims[d] > 256) return false;  // FH product form: keep products in range
    }
    a.sumn = sumn;
    // folding: maximise K / (4*KCp + R + 1)  (DMMA cycles vs epilogue cycles)
    double best = -1.0;
    int best_nf = 0;
    int64_t K = 1;
    for (int nf = 1; nf <= rank && nf <= 3; ++nf) {
        K *= dims[nf - 1];
        if (K > 192) break;
        const int kc = pick_kc((int)K);
        if (!kc) break;
        if (rank - nf > kLevMax) continue;
        const double score = (double)K / (4.0 * kc + (rank - nf) + 1.0);
        if (score > best) { best = score; best_nf = nf; }
    }
    if (best_nf == 0) return false;
    a.nfold = best_nf;
    K = 1;
    for (int d = 0; d < best_nf; ++d) K *= dims[d];
    a.K = (int)K;
    const int kc = pick_kc(a.K);
    *kc_out = kc;
    a.kp = 4 * kc;
    a.nlev = rank - best_nf;
    a.nlg = a.nlev > 1 ? a.nlev - 1 : 0;
    // level A: the first unfolded dim (absent: one entry weighted by the table's ones row)
    a.nA = a.nlev >= 1 ? dims[best_nf] : 1;
    // 4 N-tiles per step measured slower, twice: FH 40^3 0.725 vs 0.72 on the first engine, and with the
    // current one (gpurun_out/ntb_sweep.log) FH 0.794 vs 0.828, cfg2 0.79 vs 0.945 of the FP64 peak
    a.ntb = 2;
    a.nAp = (a.nA + a.ntb - 1) / a.ntb * a.ntb;
    a.boffA = a.nlev >= 1 ? a.boff[best_nf] : a.sumn;
    int64_t nrp = 1;
    for (int d = best_nf + 1; d < rank; ++d) nrp *= dims[d];
    a.nrp = (int)nrp;
    const int64_t nrows = (nrp + 7) / 8 * a.nAp * 8;
    if (nrows >= (int64_t(1) << 31)) return false;
    a.nrows = (int)nrows;
    int rs = a.kp;  // coefficients only (the group-level table rows live in MmaArgs::ginfo)
    if (mma_pair_rows(kc)) while (rs % 16 != 8) ++rs;   // 128-bit B loads: two rows fill the 32 banks
    else while (rs % 16 != 4 && rs % 16 != 12) ++rs;    // 64-bit B loads: eight rows x four k
    a.rs = rs;
    // some layout must fit
    for (const auto &i : kInsts) {
        if (i.kc != kc || i.ntb != a.ntb) continue;
        MmaArgs probe = a;
        if (mma_fit_ring(probe, i.nw * i.mt * 8)) {
            a.rows_per_slab = probe.rows_per_slab;
            a.nslabs = probe.nslabs;
            a.stage_bytes = probe.stage_bytes;
            a.ns = probe.ns;
            a.resident = probe.resident;
            return true;
        }
    }
    return false;
}

// variant 0: default layout -- 8 warps with the most point tiles per warp (MT) whose basis
// table leaves room for the ring, else 4 warps.  variant OCC*10000 + MT*100 + NW forces a layout
// (OCC = 2: ring sized for two co-resident CTAs per SM, grid of 2 x SMs).
cudaError_t launch_contract_mma(const MmaArgs &a, int variant, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    const int kc = pick_kc(a.K);
    if (!kc) return cudaErrorNotSupported;
    if (variant >= 100) {
        const int occ = variant / 10000, mt = (variant / 100) % 100, nw = variant % 100;
        MmaArgs b = a;
        b.occ = occ == 2 ? 2 : 1;
        for (const auto &i : kInsts)
            if (i.kc == kc && i.mt == mt && i.nw == nw && i.ntb == a.ntb) return i.fn(b, num_sms, s, li);
        return cudaErrorNotSupported;
    }
    // Measured preference (tools/layout_sweep.py): up to K = 100 sixteen one-tile warps hide the
    // per-step epilogue best, then twelve two-tile warps; longer K wants the most point tiles
    // per warp that still leave room for the ring.
    const Inst *best = nullptr;
    auto fits = [&](const Inst &i) {
        MmaArgs probe = a;
        return mma_fit_ring(probe, i.nw * i.mt * 8);
    };
    if (kc <= 25) {
        static const int pref[2][2] = {{1, 16}, {2, 12}};
        for (int q = 0; q < 2 && !best; ++q)
            for (const auto &i : kInsts)
                if (i.kc == kc && i.ntb == a.ntb && i.mt == pref[q][0] && i.nw == pref[q][1] && fits(i)) { best = &i; break; }
    }
    for (int pass = 0; pass < 2 && !best; ++pass) {
        const int nw = pass == 0 ? 8 : 4;
        for (const auto &i : kInsts) {
            if (i.kc != kc || i.nw != nw || i.ntb != a.ntb || (kc <= 12 && i.mt > 2)) continue;
            if (!fits(i)) continue;
            if (!best || i.mt > best->mt) best = &i;
        }
    }
    return best ? best->fn(a, num_sms, s, li) : cudaErrorNotSupported;
}
