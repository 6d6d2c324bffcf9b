// api.cu -- the C ABI of libchebpol_cuda (include/chebpol_cuda.h): argument
// validation mirroring the reference's entry points, plan management, kernel
// selection, the chunked host<->device pipeline and the multi-GPU point
// partitioner.  No CPU evaluation path exists in this library.
#include "common.cuh"

#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <exception>

// --------------------------------------------------------------- error state
static thread_local std::string g_last_error;
static std::atomic<int64_t> g_launches{0};

static int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

static int cuda_fail(cudaError_t e, const char *what)
{
    const int code = (e == cudaErrorMemoryAllocation) ? CHEBPOL_CUDA_ENOMEM
                     : (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ||
                        e == cudaErrorInvalidDevice)
                         ? CHEBPOL_CUDA_ENODEV
                         : CHEBPOL_CUDA_ECUDA;
    return fail(code, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
}

#define CU(call)                                             \
    do {                                                     \
        cudaError_t e_ = (call);                             \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call);  \
    } while (0)

extern "C" const char *chebpol_cuda_strerror(int code)
{
    switch (code) {
    case CHEBPOL_CUDA_OK: return "success";
    case CHEBPOL_CUDA_EINVAL: return "invalid argument";
    case CHEBPOL_CUDA_ERANK: return "rank must be positive (and at most 31)";
    case CHEBPOL_CUDA_ESIZE: return "dimensions must be positive and their product below 2^31";
    case CHEBPOL_CUDA_EBLEND: return "unknown blend";
    case CHEBPOL_CUDA_ENODEV: return "no usable CUDA device";
    case CHEBPOL_CUDA_ENOMEM: return "out of device or pinned host memory";
    case CHEBPOL_CUDA_ECUDA: return "CUDA runtime error";
    case CHEBPOL_CUDA_EARCH: return "device is not an sm_100 (Blackwell B200) GPU";
    default: return "unknown error code";
    }
}

extern "C" const char *chebpol_cuda_last_error(void) { return g_last_error.c_str(); }
extern "C" int chebpol_cuda_version(void) { return 100; }
extern "C" int64_t chebpol_cuda_launch_count(void) { return g_launches.load(); }

extern "C" int chebpol_cuda_device_count(int *count)
{
    if (!count) return fail(CHEBPOL_CUDA_EINVAL, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        (void)cudaGetLastError();
        return cuda_fail(e, "cudaGetDeviceCount");
    }
    *count = n;
    return CHEBPOL_CUDA_OK;
}

// ---------------------------------------------------------------------- plan
enum PlanKind { PLAN_CHEB = 1, PLAN_MLIP = 2, PLAN_FH = 3, PLAN_POLYH = 4, PLAN_STALK = 5, PLAN_SL = 6, PLAN_OLDSTALK = 7 };

cudaError_t stalk_pack_records(const double *val, const double *aa, const double *b, const double *t1,
                               const double *t2, int rank, int64_t n, double *rec, cudaStream_t s);

struct chebpol_cuda_plan {
    int kind = 0;
    int device = 0;
    int num_sms = 0;
    int rank = 0;
    int dims[CP_MAXR] = {0};
    int64_t stride[CP_MAXR] = {0};
    int64_t nelem = 0;
    double *d_tensor = nullptr;  // R layout
    double *d_padded = nullptr;  // dim 0 padded to n0p (slab kernels)
    double *d_grid = nullptr;    // concatenated knots
    double *d_weights = nullptr; // concatenated FH weights
    int goff[CP_MAXR] = {0};
    int grid_elems = 0;
    PreMap pm;
    double *d_spline = nullptr;  // 'general' grid maps: per dimension knots + [y, b, c, d] records (PreMap bit 2)
    int blend = 0;
    // options
    int kernel_opt = CHEBPOL_CUDA_KERNEL_AUTO;
    int variant = 0;
    int64_t chunk_points = 0;
    int64_t call_points = 0;  // points of the host call in progress (a chunk is launched on fewer): what the lazily built tables amortise over
    // slab configuration (Chebyshev DFMA fast path)
    bool slab_ok = false;
    SlabArgs slab;
    // DMMA contraction engine (Chebyshev and Floater-Hormann)
    bool mma_ok = false;
    MmaArgs mma;
    double *d_mma = nullptr;  // [nrest][rs] padded rows
    int *d_ktab = nullptr;    // [kp][3] folded-dim digit rows
    // multilinear: cell-major expansion of the value table, built on first use
    double *d_cells = nullptr;
    int *d_lut = nullptr;     // bucket tables of the cell search
    int loff[CP_MAXR] = {0}, nbk[CP_MAXR] = {0};
    double lscale[CP_MAXR] = {0};
    int64_t cell_bytes = 0;   // 0: shape not supported / disabled
    int cell_m = 0;           // leading dimensions expanded into bricks of 2^m corners (rank: cell-major)
    int64_t ncells = 0;
    bool cells_failed = false;
    // multilinear: row-pair-interleaved table (mlip_quad.cu, 2x the values), built on first use
    double *d_quads = nullptr;
    int64_t quad_bytes = 0;   // 0: shape not supported
    bool quads_failed = false;
    // multilinear: workspace of the binned (cache-blocked) path, mlip_bin.cu
    MlipBinState *bin = nullptr;
    bool bin_failed = false;
    // polyharmonic spline: knots in d_tensor, weights in d_weights, affine part in d_grid,
    // packed [w, knot, pad] rows in d_padded
    int nknots = 0;
    double polyh_k = 0.0;
    // stalker splines: values in d_tensor, knots in d_grid, packed records in d_padded
    int stalk_kind = 0;  // 0 stalker, 1 hstalker
    double *d_stalk[4] = {nullptr, nullptr, nullptr, nullptr};  // a, b, t1, t2
    // simplex-linear: every device table of the triangulation (simplex.cu)
    SlDevice *sl = nullptr;
    int epol = 0;
    // deprecated recursive stalker: values in d_tensor, knots in d_grid, det / pmin / pplus in d_stalk[0..2]
    double os_degree[CP_MAXR] = {0};
    int os_uniform[CP_MAXR] = {0};
    // statistics
    LaunchInfo last = {0, 0, 0, 0};
    int64_t launches = 0;
    int64_t tensor_bytes = 0;
    // host pipeline resources, created on first eval_host
    cudaStream_t st_in = nullptr, st_k = nullptr, st_out = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_k[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    double *d_x[2] = {nullptr, nullptr}, *d_o[2] = {nullptr, nullptr};
    int64_t buf_points = 0;
    int host_threads = 0;  // threads that fill / drain the pinned staging buffers (0: staging_threads())
    // derived tables built lazily on the stream of the evaluation that first needs them (padded
    // slab tensor, multilinear cell-major table): later evaluations on another stream wait for this
    cudaEvent_t ev_build = nullptr;
    cudaStream_t build_stream = nullptr;
    bool built_lazily = false;
};

// Plan constructors upload with synchronous cudaMemcpy from pageable memory, which may return while
// the last DMA is still in flight on the legacy stream; evaluations run on non-blocking streams, so
// a constructor only returns once everything it uploaded has landed.
extern "C" void chebpol_cuda_plan_destroy(chebpol_cuda_plan *p);
static int plan_ready(chebpol_cuda_plan **plan)
{
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) return CHEBPOL_CUDA_OK;
    chebpol_cuda_plan_destroy(*plan);
    *plan = nullptr;
    return cuda_fail(e, "plan upload");
}

static int check_dims(const int *dims, int rank, int64_t *nelem)
{
    if (rank <= 0) return fail(CHEBPOL_CUDA_ERANK, "rank must be positive");
    if (rank > CHEBPOL_CUDA_MAXRANK) return fail(CHEBPOL_CUDA_ERANK, "rank %d exceeds %d", rank, CHEBPOL_CUDA_MAXRANK);
    if (!dims) return fail(CHEBPOL_CUDA_EINVAL, "dims is NULL");
    int64_t n = 1;
    for (int i = 0; i < rank; ++i) {
        if (dims[i] <= 0) return fail(CHEBPOL_CUDA_ESIZE, "dims[%d] = %d must be positive", i, dims[i]);
        n *= dims[i];
        if (n >= (int64_t(1) << 31)) return fail(CHEBPOL_CUDA_ESIZE, "tensor has 2^31 or more elements");
    }
    *nelem = n;
    return CHEBPOL_CUDA_OK;
}

static int select_device(int device, int *num_sms)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return cuda_fail(e, "cudaGetDeviceCount"); }
    if (n <= 0) return fail(CHEBPOL_CUDA_ENODEV, "no CUDA device visible");
    if (device < 0 || device >= n) return fail(CHEBPOL_CUDA_ENODEV, "device %d not in 0..%d", device, n - 1);
    CU(cudaSetDevice(device));
    int major = 0, sms = 0;
    CU(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    if (major != 10) return fail(CHEBPOL_CUDA_EARCH, "device %d has compute capability %d.x; this library is sm_100a only", device, major);
    *num_sms = sms;
    return CHEBPOL_CUDA_OK;
}

static chebpol_cuda_plan *new_plan(int kind, int device, int num_sms, const int *dims, int rank, int64_t nelem)
{
    nvtxMarkA("chebpol_cuda:plan_build");
    chebpol_cuda_plan *p = new chebpol_cuda_plan();
    p->kind = kind;
    p->device = device;
    p->num_sms = num_sms;
    p->rank = rank;
    p->nelem = nelem;
    int64_t s = 1;
    for (int i = 0; i < rank; ++i) {
        p->dims[i] = dims[i];
        p->stride[i] = s;
        s *= dims[i];
    }
    memset(&p->pm, 0, sizeof p->pm);
    memset(&p->slab, 0, sizeof p->slab);
    memset(&p->mma, 0, sizeof p->mma);
    return p;
}

extern "C" void chebpol_cuda_plan_destroy(chebpol_cuda_plan *p)
{
    if (!p) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(p->device);
    if (p->st_in) {
        cudaStreamSynchronize(p->st_in);
        cudaStreamSynchronize(p->st_k);
        cudaStreamSynchronize(p->st_out);
        cudaStreamDestroy(p->st_in);
        cudaStreamDestroy(p->st_k);
        cudaStreamDestroy(p->st_out);
        for (int i = 0; i < 2; ++i) {
            cudaEventDestroy(p->ev_in[i]);
            cudaEventDestroy(p->ev_k[i]);
            cudaEventDestroy(p->ev_out[i]);
        }
    }
    for (int i = 0; i < 2; ++i) {
        cudaFree(p->d_x[i]);
        cudaFree(p->d_o[i]);
    }
    if (p->ev_build) cudaEventDestroy(p->ev_build);
    cudaFree(p->d_tensor);
    cudaFree(p->d_padded);
    cudaFree(p->d_mma);
    cudaFree(p->d_ktab);
    cudaFree(p->d_cells);
    cudaFree(p->d_quads);
    mlip_bin_destroy(p->bin);
    cudaFree(p->d_lut);
    cudaFree(p->d_grid);
    cudaFree(p->d_weights);
    cudaFree(p->d_spline);
    sl_destroy(p->sl);
    for (int i = 0; i < 4; ++i) cudaFree(p->d_stalk[i]);
    delete p;
    if (prev >= 0) cudaSetDevice(prev);
}

// upload concatenated per-dimension vectors
static int upload_lists(chebpol_cuda_plan *p, const double *const *grid, const double *const *weights)
{
    int tot = 0;
    for (int d = 0; d < p->rank; ++d) {
        if (!grid[d] || (weights && !weights[d])) return fail(CHEBPOL_CUDA_EINVAL, "grid/weights[%d] is NULL", d);
        p->goff[d] = tot;
        tot += p->dims[d];
    }
    p->grid_elems = tot;
    std::vector<double> g(tot), w(weights ? tot : 0);
    for (int d = 0; d < p->rank; ++d) {
        memcpy(&g[p->goff[d]], grid[d], sizeof(double) * p->dims[d]);
        if (weights) memcpy(&w[p->goff[d]], weights[d], sizeof(double) * p->dims[d]);
    }
    CU(cudaMalloc(&p->d_grid, sizeof(double) * tot));
    CU(cudaMemcpy(p->d_grid, g.data(), sizeof(double) * tot, cudaMemcpyHostToDevice));
    if (weights) {
        CU(cudaMalloc(&p->d_weights, sizeof(double) * tot));
        CU(cudaMemcpy(p->d_weights, w.data(), sizeof(double) * tot, cudaMemcpyHostToDevice));
    }
    return CHEBPOL_CUDA_OK;
}

// Slab configuration of the Chebyshev fast path (see cheb_slab.cu).
static void configure_slab(chebpol_cuda_plan *p)
{
    p->slab_ok = false;
    const int n0 = p->dims[0];
    const int n0p = cheb_slab_pick_n0p(n0);
    if (n0p == 0) return;
    SlabArgs &s = p->slab;
    s.rank = p->rank;
    s.n0 = n0;
    s.n0p = n0p;
    for (int i = 0; i < p->rank; ++i) s.dims[i] = p->dims[i];
    const int64_t cap = 48 * 1024;
    int depth = 1;
    int64_t elems = n0p;
    while (depth < p->rank && depth < 3 && elems * p->dims[depth] * 8 <= cap) {
        elems *= p->dims[depth];
        ++depth;
    }
    s.depth = depth;
    s.slab_elems = (int)elems;
    int64_t nsl = 1;
    for (int l = depth; l < p->rank; ++l) {
        nsl *= p->dims[l];
        s.cum[l] = (unsigned)nsl;
    }
    s.nslabs = (int)nsl;
    s.stage_stride = (int)((elems * 8 + 127) / 128 * 128);
    if (nsl <= 8 && nsl * s.stage_stride <= 200 * 1024) {
        s.resident = 1;
        s.ns = (int)nsl;
    } else {
        s.resident = 0;
        int ns = (int)((96 * 1024) / s.stage_stride);
        if (ns < 2) ns = 2;
        if (ns > 8) ns = 8;
        s.ns = ns;
    }
    s.pm = p->pm;
    p->slab_ok = true;
}

// Row-padded copy of the tensor for the DMMA engine: row r holds the K entries of the
// folded leading dims (contiguous in the R layout) followed by zeros up to rs.
static int configure_mma(chebpol_cuda_plan *p, int kind, const double *host_tensor)
{
    p->mma_ok = false;
    int kc = 0;
    if (!mma_configure(p->mma, kind, p->rank, p->dims, &kc)) return CHEBPOL_CUDA_OK;
    MmaArgs &m = p->mma;
    m.pm = p->pm;
    // rows ordered (r'-tile of 8, level-A index, r' within the tile); zero rows pad level A to an
    // even count and r' to a multiple of 8
    const size_t rows = (size_t)m.nrows, rs = (size_t)m.rs, K = (size_t)m.K;
    std::vector<double> padded(rows * rs, 0.0);
    const int ng = m.nrows / (m.nAp * 8);
    for (int g = 0; g < ng; ++g)
        for (int ia = 0; ia < m.nAp; ++ia)
            for (int u = 0; u < 8; ++u) {
                const size_t dst = ((size_t)g * m.nAp + ia) * 8 + u;
                const int rp = 8 * g + u;
                if (rp < m.nrp && ia < m.nA) {
                    const double *src = host_tensor + ((size_t)ia + (size_t)m.nA * rp) * K;
                    if (!mma_pair_rows(kc)) memcpy(&padded[dst * rs], src, sizeof(double) * K);
                    else
                        for (size_t k = 0; k < K; ++k) padded[dst * rs + (size_t)mma_row_pos(kc, (int)k)] = src[k];
                }
            }
    std::vector<int> ktab((size_t)m.kp * 3, m.sumn);  // padding k and absent folded dims name the ones row (the padded coefficients are zero)
    for (int k = 0; k < m.K; ++k) {
        int rem = k;
        for (int l = 0; l < m.nfold; ++l) {
            ktab[(size_t)k * 3 + l] = m.boff[l] + rem % m.dims[l];
            rem /= m.dims[l];
        }
    }
    // basis-table rows of the group-level digits of every r' (the ones row for absent levels and for the
    // r' that pad the last tile): read once per group by close_group
    const size_t ktab_ints = ktab.size();
    ktab.resize(ktab_ints + (size_t)ng * 8 * 3, m.sumn);
    for (int rp = 0; rp < m.nrp; ++rp) {
        size_t rem = (size_t)rp;
        for (int l = 0; l < m.nlg; ++l) {
            const int d = m.nfold + 1 + l;
            ktab[ktab_ints + (size_t)rp * 3 + l] = m.boff[d] + (int)(rem % (size_t)m.dims[d]);
            rem /= (size_t)m.dims[d];
        }
    }
    cudaError_t ek = cudaMalloc(&p->d_ktab, sizeof(int) * ktab.size());
    if (ek != cudaSuccess) return cuda_fail(ek, "cudaMalloc(ktab)");
    ek = cudaMemcpy(p->d_ktab, ktab.data(), sizeof(int) * ktab.size(), cudaMemcpyHostToDevice);
    if (ek != cudaSuccess) return cuda_fail(ek, "cudaMemcpy(ktab)");
    m.ktab = p->d_ktab;
    m.ginfo = p->d_ktab + ktab_ints;
    cudaError_t e = cudaMalloc(&p->d_mma, sizeof(double) * rows * rs);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(mma tensor)");
    e = cudaMemcpy(p->d_mma, padded.data(), sizeof(double) * rows * rs, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy(mma tensor)");
    m.tensor = p->d_mma;
    m.grid = p->d_grid;
    m.weights = p->d_weights;
    for (int d = 0; d < p->rank; ++d) m.goff[d] = p->goff[d];
    p->tensor_bytes += sizeof(double) * rows * rs;
    p->mma_ok = true;
    return CHEBPOL_CUDA_OK;
}

// spline tables of the 'general' grid maps as the caller holds them (one pointer per dimension)
struct SplineTabs {
    const int *sn;
    const double *const *x, *const *y, *const *b, *const *c, *const *d;
};

static int check_spline_tabs(const SplineTabs &t, int rank)
{
    if (!t.sn || !t.x || !t.y || !t.b || !t.c || !t.d) return fail(CHEBPOL_CUDA_EINVAL, "spline table list is NULL");
    for (int k = 0; k < rank; ++k) {
        if (t.sn[k] < 2) return fail(CHEBPOL_CUDA_ESIZE, "grid map %d needs at least 2 knots", k);
        if (!t.x[k] || !t.y[k] || !t.b[k] || !t.c[k] || !t.d[k])
            return fail(CHEBPOL_CUDA_EINVAL, "spline table of dimension %d is NULL", k);
        for (int i = 1; i < t.sn[k]; ++i)
            if (!(t.x[k][i] > t.x[k][i - 1]))
                return fail(CHEBPOL_CUDA_EINVAL, "grid map %d: knots must be strictly increasing (as splinefun stores them)", k);
    }
    return CHEBPOL_CUDA_OK;
}

// device layout read by premap_spline (common.cuh): per dimension the knots, padded to a multiple
// of 4 doubles, then one 32-byte record [y, b, c, d] per knot
static int upload_spline_tabs(chebpol_cuda_plan *p, const SplineTabs &t)
{
    std::vector<double> h;
    for (int k = 0; k < p->rank; ++k) {
        const int n = t.sn[k], npad = (n + 3) & ~3;
        p->pm.soff[k] = (int)h.size();
        p->pm.sn[k] = n;
        h.insert(h.end(), t.x[k], t.x[k] + n);
        h.resize(h.size() + (npad - n), 0.0);
        for (int i = 0; i < n; ++i) {
            h.push_back(t.y[k][i]);
            h.push_back(t.b[k][i]);
            h.push_back(t.c[k][i]);
            h.push_back(t.d[k][i]);
        }
    }
    CU(cudaMalloc(&p->d_spline, sizeof(double) * h.size()));
    CU(cudaMemcpy(p->d_spline, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice));
    p->pm.spl = p->d_spline;
    p->pm.mode |= 4;
    return CHEBPOL_CUDA_OK;
}

static int plan_cheb_impl(const double *coef, const int *dims, int rank, const double *mid, const double *ispan,
                          const int *ugm_n, const SplineTabs *spl, int device, chebpol_cuda_plan **plan);

extern "C" int chebpol_cuda_plan_cheb(const double *coef, const int *dims, int rank, const double *mid,
                                      const double *ispan, const int *ugm_n, int device,
                                      chebpol_cuda_plan **plan)
{
    return plan_cheb_impl(coef, dims, rank, mid, ispan, ugm_n, nullptr, device, plan);
}

extern "C" int chebpol_cuda_plan_chebg(const double *coef, const int *dims, int rank, const int *sn,
                                       const double *const *sx, const double *const *sy, const double *const *sb,
                                       const double *const *sc, const double *const *sd, int device,
                                       chebpol_cuda_plan **plan)
{
    SplineTabs t = {sn, sx, sy, sb, sc, sd};
    return plan_cheb_impl(coef, dims, rank, nullptr, nullptr, nullptr, &t, device, plan);
}

static int plan_cheb_impl(const double *coef, const int *dims, int rank, const double *mid, const double *ispan,
                          const int *ugm_n, const SplineTabs *spl, int device, chebpol_cuda_plan **plan)
{
    if (!plan) return fail(CHEBPOL_CUDA_EINVAL, "plan is NULL");
    *plan = nullptr;
    if (!coef) return fail(CHEBPOL_CUDA_EINVAL, "coef is NULL");
    if ((mid == nullptr) != (ispan == nullptr)) return fail(CHEBPOL_CUDA_EINVAL, "mid and ispan must be given together");
    int64_t nelem = 0;
    int rc = check_dims(dims, rank, &nelem);
    if (rc) return rc;
    if (spl && (rc = check_spline_tabs(*spl, rank))) return rc;
    int sms = 0;
    rc = select_device(device, &sms);
    if (rc) return rc;
    chebpol_cuda_plan *p = new_plan(PLAN_CHEB, device, sms, dims, rank, nelem);
    if (mid) {
        p->pm.mode |= 1;
        for (int d = 0; d < rank; ++d) { p->pm.mid[d] = mid[d]; p->pm.ispan[d] = ispan[d]; }
    }
    if (ugm_n) {
        p->pm.mode |= 2;
        for (int d = 0; d < rank; ++d) {
            if (ugm_n[d] <= 0) { delete p; return fail(CHEBPOL_CUDA_ESIZE, "ugm_n[%d] must be positive", d); }
            p->pm.nd[d] = (double)ugm_n[d];
        }
    }
    auto bail = [&](int code) { chebpol_cuda_plan_destroy(p); return code; };
    if (spl && (rc = upload_spline_tabs(p, *spl))) return bail(rc);
    cudaError_t e = cudaMalloc(&p->d_tensor, sizeof(double) * nelem);
    if (e != cudaSuccess) return bail(cuda_fail(e, "cudaMalloc(tensor)"));
    e = cudaMemcpy(p->d_tensor, coef, sizeof(double) * nelem, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return bail(cuda_fail(e, "cudaMemcpy(tensor)"));
    p->tensor_bytes = sizeof(double) * nelem;
    configure_slab(p);  // the padded copy for the DFMA slab kernel is made on first use (ensure_slab_tensor)
    rc = configure_mma(p, 0, coef);
    if (rc) return bail(rc);
    *plan = p;
    return plan_ready(plan);
}

extern "C" int chebpol_cuda_plan_mlip(const double *const *grid, const int *dims, int rank,
                                      const double *values, int blend, int device, chebpol_cuda_plan **plan)
{
    if (!plan) return fail(CHEBPOL_CUDA_EINVAL, "plan is NULL");
    *plan = nullptr;
    if (!grid || !values) return fail(CHEBPOL_CUDA_EINVAL, "grid or values is NULL");
    if (blend < 0 || blend > 5) return fail(CHEBPOL_CUDA_EBLEND, "blend %d not in 0..5", blend);
    int64_t nelem = 0;
    int rc = check_dims(dims, rank, &nelem);
    if (rc) return rc;
    for (int d = 0; d < rank; ++d)
        if (dims[d] < 2) return fail(CHEBPOL_CUDA_ESIZE, "multilinear grid %d needs at least 2 knots", d);
    int sms = 0;
    rc = select_device(device, &sms);
    if (rc) return rc;
    chebpol_cuda_plan *p = new_plan(PLAN_MLIP, device, sms, dims, rank, nelem);
    p->blend = blend;
    auto bail = [&](int code) { chebpol_cuda_plan_destroy(p); return code; };
    rc = upload_lists(p, grid, nullptr);
    if (rc) return bail(rc);
    cudaError_t e = cudaMalloc(&p->d_tensor, sizeof(double) * nelem);
    if (e != cudaSuccess) return bail(cuda_fail(e, "cudaMalloc(values)"));
    e = cudaMemcpy(p->d_tensor, values, sizeof(double) * nelem, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return bail(cuda_fail(e, "cudaMemcpy(values)"));
    p->tensor_bytes = sizeof(double) * nelem;
    // bucket tables for the cell search: 4 buckets per knot interval on average; entry b is
    // the reference bisection's answer for the bucket's left edge
    {
        std::vector<int> lut;
        for (int d = 0; d < rank; ++d) {
            const int n = dims[d];
            const double *g = grid[d];
            int nb = 1;
            while (nb < 4 * n && nb < 4096) nb <<= 1;
            const double span = g[n - 1] - g[0];
            const double scale = span > 0 ? nb / span : 0.0;
            p->loff[d] = (int)lut.size();
            p->nbk[d] = nb;
            p->lscale[d] = scale;
            for (int b = 0; b < nb; ++b) {
                const double edge = g[0] + (scale > 0 ? b / scale : 0.0);
                int lo = 0, hi = n - 1;  // src/chebpol.c:543-556
                while (lo + 1 < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (g[mid] >= edge) hi = mid;
                    else lo = mid;
                }
                lut.push_back(hi);
            }
        }
        e = cudaMalloc(&p->d_lut, sizeof(int) * lut.size());
        if (e != cudaSuccess) return bail(cuda_fail(e, "cudaMalloc(lut)"));
        e = cudaMemcpy(p->d_lut, lut.data(), sizeof(int) * lut.size(), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return bail(cuda_fail(e, "cudaMemcpy(lut)"));
    }
    // cell-major expansion (mlip_cell.cu) is built lazily, on the first evaluation large
    // enough to amortise it; capped by CHEBPOL_CUDA_CELL_MAX_GB (default: a quarter of the device) and free memory
    // The largest expansion that fits the cap: all rank dimensions (cell-major, 2^rank x the table),
    // else bricks over the first m >= 3 of them (2^m x; 64-byte bricks are the DRAM access
    // granularity, smaller ones waste half of every burst -- the 2 x quad table covers that end).
    // CHEBPOL_CUDA_CELL_M pins m (tests, tuning).
    p->ncells = 1;
    for (int d = 0; d < rank; ++d) p->ncells *= dims[d] - 1;
    {
        double cap_gb = 16.0;
        {
            size_t free_b = 0, total_b = 0;  // default cap: a quarter of the device (45 GB on a B200)
            if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) cap_gb = 0.25 * (double)total_b / 1e9;
            (void)cudaGetLastError();
        }
        if (const char *env = getenv("CHEBPOL_CUDA_CELL_MAX_GB")) cap_gb = atof(env);
        int m_lo = rank < 3 ? rank : 3, m_hi = rank;
        if (const char *env = getenv("CHEBPOL_CUDA_CELL_M")) {
            const int mm = atoi(env);
            if (mm >= 1 && mm <= rank) m_lo = m_hi = mm;
        }
        p->cell_bytes = 0;
        p->cell_m = 0;
        for (int m = m_hi; m >= m_lo; --m) {
            const int64_t b = mlip_cell_bytes(rank, dims, m);
            if (b > 0 && (double)b <= cap_gb * 1e9) {
                p->cell_bytes = b;
                p->cell_m = m;
                break;
            }
        }
    }
    p->quad_bytes = mlip_quad_bytes(rank, dims);
    *plan = p;
    return plan_ready(plan);
}

extern "C" int chebpol_cuda_plan_fh(const double *vals, const double *const *grid,
                                    const double *const *weights, const int *dims, int rank, int device,
                                    chebpol_cuda_plan **plan)
{
    if (!plan) return fail(CHEBPOL_CUDA_EINVAL, "plan is NULL");
    *plan = nullptr;
    if (!vals || !grid || !weights) return fail(CHEBPOL_CUDA_EINVAL, "vals, grid or weights is NULL");
    int64_t nelem = 0;
    int rc = check_dims(dims, rank, &nelem);
    if (rc) return rc;
    int sms = 0;
    rc = select_device(device, &sms);
    if (rc) return rc;
    chebpol_cuda_plan *p = new_plan(PLAN_FH, device, sms, dims, rank, nelem);
    auto bail = [&](int code) { chebpol_cuda_plan_destroy(p); return code; };
    rc = upload_lists(p, grid, weights);
    if (rc) return bail(rc);
    cudaError_t e = cudaMalloc(&p->d_tensor, sizeof(double) * nelem);
    if (e != cudaSuccess) return bail(cuda_fail(e, "cudaMalloc(vals)"));
    e = cudaMemcpy(p->d_tensor, vals, sizeof(double) * nelem, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return bail(cuda_fail(e, "cudaMemcpy(vals)"));
    p->tensor_bytes = sizeof(double) * nelem;
    rc = configure_mma(p, 1, vals);
    if (rc) return bail(rc);
    *plan = p;
    return plan_ready(plan);
}

extern "C" int chebpol_cuda_plan_polyh(const double *knots, int rank, int nknots, const double *weights,
                                       const double *lweights, double k, const double *norm_a,
                                       const double *norm_b, int device, chebpol_cuda_plan **plan)
{
    if (!plan) return fail(CHEBPOL_CUDA_EINVAL, "plan is NULL");
    *plan = nullptr;
    if (rank <= 0) return fail(CHEBPOL_CUDA_ERANK, "rank must be positive");
    if (rank > CHEBPOL_CUDA_MAXRANK) return fail(CHEBPOL_CUDA_ERANK, "rank %d exceeds %d", rank, CHEBPOL_CUDA_MAXRANK);
    if (nknots < 0) return fail(CHEBPOL_CUDA_ESIZE, "negative number of knots");
    if ((int64_t)nknots * rank >= (int64_t(1) << 31)) return fail(CHEBPOL_CUDA_ESIZE, "knot matrix has 2^31 or more elements");
    if (!lweights || (nknots > 0 && (!knots || !weights))) return fail(CHEBPOL_CUDA_EINVAL, "knots, weights or lweights is NULL");
    if ((norm_a == nullptr) != (norm_b == nullptr)) return fail(CHEBPOL_CUDA_EINVAL, "norm_a and norm_b go together");
    if (k != k) return fail(CHEBPOL_CUDA_EINVAL, "k is NaN");
    int sms = 0;
    int rc = select_device(device, &sms);
    if (rc) return rc;
    int dims[CP_MAXR];
    for (int i = 0; i < rank; ++i) dims[i] = nknots > 0 ? nknots : 1;
    chebpol_cuda_plan *p = new_plan(PLAN_POLYH, device, sms, dims, rank, (int64_t)nknots * rank);
    auto bail = [&](int code) { chebpol_cuda_plan_destroy(p); return code; };
    p->nknots = nknots;
    p->polyh_k = k;
    if (norm_a) {
        p->pm.mode = 1;
        for (int i = 0; i < rank; ++i) { p->pm.mid[i] = norm_b[i]; p->pm.ispan[i] = norm_a[i]; }
    }
    cudaError_t e = cudaMalloc(&p->d_grid, sizeof(double) * (rank + 1));
    if (e != cudaSuccess) return bail(cuda_fail(e, "cudaMalloc(lweights)"));
    e = cudaMemcpy(p->d_grid, lweights, sizeof(double) * (rank + 1), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return bail(cuda_fail(e, "cudaMemcpy(lweights)"));
    if (nknots > 0) {
        const size_t nk = (size_t)nknots;
        e = cudaMalloc(&p->d_tensor, sizeof(double) * nk * rank);
        if (e != cudaSuccess) return bail(cuda_fail(e, "cudaMalloc(knots)"));
        e = cudaMemcpy(p->d_tensor, knots, sizeof(double) * nk * rank, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return bail(cuda_fail(e, "cudaMemcpy(knots)"));
        e = cudaMalloc(&p->d_weights, sizeof(double) * nk);
        if (e != cudaSuccess) return bail(cuda_fail(e, "cudaMalloc(weights)"));
        e = cudaMemcpy(p->d_weights, weights, sizeof(double) * nk, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return bail(cuda_fail(e, "cudaMemcpy(weights)"));
        const int rs = polyh_row_stride(rank);
        if (rs > 0) {
            std::vector<double> rows(nk * rs, 0.0);
            for (size_t i = 0; i < nk; ++i) {
                rows[i * rs] = weights[i];
                for (int j = 0; j < rank; ++j) rows[i * rs + 1 + j] = knots[i * rank + j];
            }
            e = cudaMalloc(&p->d_padded, sizeof(double) * rows.size());
            if (e != cudaSuccess) return bail(cuda_fail(e, "cudaMalloc(packed knots)"));
            e = cudaMemcpy(p->d_padded, rows.data(), sizeof(double) * rows.size(), cudaMemcpyHostToDevice);
            if (e != cudaSuccess) return bail(cuda_fail(e, "cudaMemcpy(packed knots)"));
        }
    }
    p->tensor_bytes = sizeof(double) * ((int64_t)nknots * (rank + 1) + rank + 1);
    *plan = p;
    return plan_ready(plan);
}

extern "C" int chebpol_cuda_plan_stalk(int kind, const double *val, const double *const *grid, const int *dims,
                                       int rank, const double *a, const double *b, const double *t1,
                                       const double *t2, int blend, int device, chebpol_cuda_plan **plan)
{
    if (!plan) return fail(CHEBPOL_CUDA_EINVAL, "plan is NULL");
    *plan = nullptr;
    if (kind != 0 && kind != 1) return fail(CHEBPOL_CUDA_EINVAL, "kind %d: 0 stalker, 1 hstalker", kind);
    if (!val || !grid || !b || !t1 || !t2 || (kind == 1 && !a)) return fail(CHEBPOL_CUDA_EINVAL, "a table pointer is NULL");
    if (blend < 0 || blend > 5) return fail(CHEBPOL_CUDA_EBLEND, "blend %d not in 0..5", blend);
    int64_t nelem = 0;
    int rc = check_dims(dims, rank, &nelem);
    if (rc) return rc;
    if (rank > 24) return fail(CHEBPOL_CUDA_ERANK, "rank %d: the corner loop supports at most 24 dimensions", rank);
    for (int d = 0; d < rank; ++d)
        if (dims[d] < 2) return fail(CHEBPOL_CUDA_ESIZE, "dims[%d] = %d: a stalker grid needs two knots per dimension", d, dims[d]);
    if (nelem * rank >= (int64_t(1) << 31)) return fail(CHEBPOL_CUDA_ESIZE, "parameter tables have 2^31 or more elements");
    int sms = 0;
    rc = select_device(device, &sms);
    if (rc) return rc;
    chebpol_cuda_plan *p = new_plan(PLAN_STALK, device, sms, dims, rank, nelem);
    auto bail = [&](int code) { chebpol_cuda_plan_destroy(p); return code; };
    p->stalk_kind = kind;
    p->blend = blend;
    rc = upload_lists(p, grid, nullptr);
    if (rc) return bail(rc);
    auto up = [&](double **dst, const double *src, size_t n, const char *what) -> int {
        cudaError_t e = cudaMalloc(dst, sizeof(double) * n);
        if (e != cudaSuccess) return cuda_fail(e, what);
        e = cudaMemcpy(*dst, src, sizeof(double) * n, cudaMemcpyHostToDevice);
        return e == cudaSuccess ? CHEBPOL_CUDA_OK : cuda_fail(e, what);
    };
    const size_t n = (size_t)nelem, nr = n * rank;
    if ((rc = up(&p->d_tensor, val, n, "upload(val)"))) return bail(rc);
    if (kind == 1 && (rc = up(&p->d_stalk[0], a, n, "upload(a)"))) return bail(rc);
    if ((rc = up(&p->d_stalk[1], b, nr, "upload(b)"))) return bail(rc);
    if ((rc = up(&p->d_stalk[2], t1, nr, "upload(c)"))) return bail(rc);
    if ((rc = up(&p->d_stalk[3], t2, nr, "upload(r/d)"))) return bail(rc);
    p->tensor_bytes = sizeof(double) * (n * (kind == 1 ? 2 : 1) + 3 * nr);
    const int rs = stalk_record_stride(rank, kind);
    if (rs > 0) {
        cudaError_t e = cudaMalloc(&p->d_padded, sizeof(double) * n * rs);
        if (e == cudaSuccess) {
            e = stalk_pack_records(p->d_tensor, p->d_stalk[0], p->d_stalk[1], p->d_stalk[2], p->d_stalk[3], rank,
                                   nelem, p->d_padded, 0);
            if (e == cudaSuccess) e = cudaDeviceSynchronize();
            if (e != cudaSuccess) return bail(cuda_fail(e, "pack stalker records"));
        } else {
            (void)cudaGetLastError();  // no room for the packed copy: the strict kernel serves the plan
            p->d_padded = nullptr;
        }
    }
    *plan = p;
    return plan_ready(plan);
}

// Deprecated recursive stalker evaluator (oldstalkerappx, R/stalker.R:11-49; R_evalstalker src/stalker.c:834-891)
extern "C" int chebpol_cuda_plan_oldstalk(const double *val, const double *const *grid, const int *dims, int rank,
                                          const double *degree, const double *const *det, const double *const *pmin,
                                          const double *const *pplus, const int *uniform, int blend, int device,
                                          chebpol_cuda_plan **plan)
{
    if (!plan) return fail(CHEBPOL_CUDA_EINVAL, "plan is NULL");
    *plan = nullptr;
    if (!val || !grid || !degree || !det || !pmin || !pplus || !uniform) return fail(CHEBPOL_CUDA_EINVAL, "a table pointer is NULL");
    if (blend < 0 || blend > 5) return fail(CHEBPOL_CUDA_EBLEND, "blend %d not in 0..5", blend);
    int64_t nelem = 0;
    int rc = check_dims(dims, rank, &nelem);
    if (rc) return rc;
    for (int d = 0; d < rank; ++d) {
        if (dims[d] < 2) return fail(CHEBPOL_CUDA_ESIZE, "stalker grid %d needs at least 2 knots", d);
        if (!det[d] || !pmin[d] || !pplus[d]) return fail(CHEBPOL_CUDA_EINVAL, "context vector of dimension %d is NULL", d);
        // the reference's message when it is built without GSL (src/stalker.c:871-874)
        if (!uniform[d] && degree[d] != degree[d])
            return fail(CHEBPOL_CUDA_EINVAL, "Non-uniform grid in dimension %d not supported without GSL", d + 1);
    }
    int sms = 0;
    rc = select_device(device, &sms);
    if (rc) return rc;
    chebpol_cuda_plan *p = new_plan(PLAN_OLDSTALK, device, sms, dims, rank, nelem);
    p->blend = blend;
    for (int d = 0; d < rank; ++d) { p->os_degree[d] = degree[d]; p->os_uniform[d] = uniform[d] != 0; }
    auto bail = [&](int code) { chebpol_cuda_plan_destroy(p); return code; };
    rc = upload_lists(p, grid, nullptr);
    if (rc) return bail(rc);
    const double *const *ctx[3] = {det, pmin, pplus};
    for (int j = 0; j < 3; ++j) {
        std::vector<double> h(p->grid_elems);
        for (int d = 0; d < rank; ++d) memcpy(&h[p->goff[d]], ctx[j][d], sizeof(double) * dims[d]);
        cudaError_t e = cudaMalloc(&p->d_stalk[j], sizeof(double) * h.size());
        if (e == cudaSuccess) e = cudaMemcpy(p->d_stalk[j], h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return bail(cuda_fail(e, "upload(stalker context)"));
    }
    cudaError_t e = cudaMalloc(&p->d_tensor, sizeof(double) * nelem);
    if (e == cudaSuccess) e = cudaMemcpy(p->d_tensor, val, sizeof(double) * nelem, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return bail(cuda_fail(e, "upload(values)"));
    p->tensor_bytes = sizeof(double) * (nelem + 4 * (int64_t)p->grid_elems);
    *plan = p;
    return plan_ready(plan);
}

static int check_sl(const double *knots, int dim, int nknots, const int *dtri, int ns, const double *lumats,
                    const int *pivots, const double *bbox, const double *val)
{
    if (dim <= 0) return fail(CHEBPOL_CUDA_ERANK, "rank must be positive");
    if (dim > CHEBPOL_CUDA_MAXRANK - 1) return fail(CHEBPOL_CUDA_ERANK, "dimension %d exceeds %d", dim, CHEBPOL_CUDA_MAXRANK - 1);
    if (nknots < 0 || ns < 0) return fail(CHEBPOL_CUDA_ESIZE, "negative number of knots or simplices");
    if ((int64_t)ns * (dim + 1) * (dim + 1) >= (int64_t(1) << 31) || (int64_t)nknots * dim >= (int64_t(1) << 31))
        return fail(CHEBPOL_CUDA_ESIZE, "a simplex table has 2^31 or more elements");
    if ((nknots > 0 && (!knots || !val)) || (ns > 0 && (!dtri || !lumats || !pivots || !bbox)))
        return fail(CHEBPOL_CUDA_EINVAL, "a simplex table pointer is NULL");
    for (int64_t i = 0; i < (int64_t)ns * (dim + 1); ++i)
        if (dtri[i] < 1 || dtri[i] > nknots)
            return fail(CHEBPOL_CUDA_EINVAL, "dtri[%lld] = %d is not a knot index in 1..%d", (long long)i, dtri[i], nknots);
    // LAPACK's dgetrf records, for row r (1-based), the row it was interchanged with: r <= piv <= dim+1.
    // The kernels apply the interchanges without bounds checks.
    const int N = dim + 1;
    for (int64_t j = 0; j < (int64_t)ns; ++j)
        for (int r = 0; r < N; ++r) {
            const int pv = pivots[j * N + r];
            if (pv < r + 1 || pv > N)
                return fail(CHEBPOL_CUDA_EINVAL, "pivots[%lld] = %d is not a dgetrf pivot (%d..%d)", (long long)(j * N + r), pv, r + 1, N);
        }
    return CHEBPOL_CUDA_OK;
}

extern "C" int chebpol_cuda_plan_sl(const double *knots, int dim, int nknots, const int *dtri, int numsimplex,
                                    const double *lumats, const int *pivots, const double *bbox, const double *val,
                                    int device, chebpol_cuda_plan **plan)
{
    if (!plan) return fail(CHEBPOL_CUDA_EINVAL, "plan is NULL");
    *plan = nullptr;
    int rc = check_sl(knots, dim, nknots, dtri, numsimplex, lumats, pivots, bbox, val);
    if (rc) return rc;
    int sms = 0;
    rc = select_device(device, &sms);
    if (rc) return rc;
    int dims[CP_MAXR];
    for (int i = 0; i < dim; ++i) dims[i] = nknots > 0 ? nknots : 1;
    chebpol_cuda_plan *p = new_plan(PLAN_SL, device, sms, dims, dim, (int64_t)nknots * dim);
    p->blend = 3;  // the closure's default, 'cubic' (R/simplexlinear.R:36)
    p->nknots = nknots;
    const char *env = getenv("CHEBPOL_CUDA_SL_CELLS");
    cudaError_t e = cudaSuccess;
    try {  // host-side tables are std::vectors: an allocation failure must not cross the C boundary
        e = sl_create(knots, dim, nknots, dtri, numsimplex, lumats, pivots, bbox, val, env ? atoi(env) : 0, &p->sl);
    } catch (const std::exception &ex) {
        chebpol_cuda_plan_destroy(p);
        return fail(CHEBPOL_CUDA_ENOMEM, "simplex tables: %s", ex.what());
    }
    if (e != cudaSuccess) {
        chebpol_cuda_plan_destroy(p);
        return cuda_fail(e, "simplex tables");
    }
    p->tensor_bytes = sl_bytes(p->sl);
    *plan = p;
    return plan_ready(plan);
}

extern "C" int chebpol_cuda_analyzesimplex(const int *dtri, int numsimplex, const double *knots, int dim, int nknots,
                                           int device, double *lumats, int *pivots, double *bbox)
{
    if (dim <= 0 || dim > CHEBPOL_CUDA_MAXRANK - 1) return fail(CHEBPOL_CUDA_ERANK, "dimension %d not in 1..%d", dim, CHEBPOL_CUDA_MAXRANK - 1);
    if (numsimplex < 0 || nknots < 0) return fail(CHEBPOL_CUDA_ESIZE, "negative number of knots or simplices");
    if (numsimplex == 0) return CHEBPOL_CUDA_OK;
    if (!dtri || !knots || !lumats || !pivots || !bbox) return fail(CHEBPOL_CUDA_EINVAL, "a pointer is NULL");
    const int N = dim + 1;
    if ((int64_t)numsimplex * N * N >= (int64_t(1) << 31)) return fail(CHEBPOL_CUDA_ESIZE, "too many simplices");
    for (int64_t i = 0; i < (int64_t)numsimplex * N; ++i)
        if (dtri[i] < 1 || dtri[i] > nknots)
            return fail(CHEBPOL_CUDA_EINVAL, "dtri[%lld] = %d is not a knot index in 1..%d", (long long)i, dtri[i], nknots);
    int sms = 0;
    int rc = select_device(device, &sms);
    if (rc) return rc;
    int *d_tri = nullptr, *d_piv = nullptr;
    double *d_kn = nullptr, *d_lu = nullptr, *d_bb = nullptr;
    const size_t ns = (size_t)numsimplex;
    cudaError_t e = cudaMalloc(&d_tri, sizeof(int) * ns * N);
    if (e == cudaSuccess) e = cudaMalloc(&d_piv, sizeof(int) * ns * N);
    if (e == cudaSuccess) e = cudaMalloc(&d_kn, sizeof(double) * (size_t)nknots * dim);
    if (e == cudaSuccess) e = cudaMalloc(&d_lu, sizeof(double) * ns * N * N);
    if (e == cudaSuccess) e = cudaMalloc(&d_bb, sizeof(double) * ns * 2 * dim);
    if (e == cudaSuccess) e = cudaMemcpy(d_tri, dtri, sizeof(int) * ns * N, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_kn, knots, sizeof(double) * (size_t)nknots * dim, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = sl_analyze_device(d_tri, numsimplex, d_kn, dim, d_lu, d_piv, d_bb, 0);
    if (e == cudaSuccess) { g_launches.fetch_add(1); e = cudaMemcpy(lumats, d_lu, sizeof(double) * ns * N * N, cudaMemcpyDeviceToHost); }
    if (e == cudaSuccess) e = cudaMemcpy(pivots, d_piv, sizeof(int) * ns * N, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(bbox, d_bb, sizeof(double) * ns * 2 * dim, cudaMemcpyDeviceToHost);
    cudaFree(d_tri); cudaFree(d_piv); cudaFree(d_kn); cudaFree(d_lu); cudaFree(d_bb);
    if (e != cudaSuccess) return cuda_fail(e, "analyzesimplex");
    return CHEBPOL_CUDA_OK;
}

extern "C" int chebpol_cuda_debug_sl_lists(const double *knots, int dim, int nknots, const int *dtri, int numsimplex,
                                           const double *bbox, int cells_per_dim, int *cells_used, double *lo,
                                           double *inv, int *start, int64_t start_cap, int *items,
                                           int64_t items_cap, int64_t *nitems)
{
    if (!knots || !dtri || !bbox || !cells_used || !lo || !inv || !start || !items || !nitems)
        return fail(CHEBPOL_CUDA_EINVAL, "a pointer is NULL");
    int rc = 0;
    try {
        rc = sl_debug_lists(knots, dim, nknots, dtri, numsimplex, bbox, cells_per_dim, cells_used, lo, inv, start,
                            start_cap, items, items_cap, nitems);
    } catch (const std::exception &ex) {
        return fail(CHEBPOL_CUDA_ENOMEM, "candidate lists: %s", ex.what());
    }
    if (rc == 1) return fail(CHEBPOL_CUDA_ERANK, "cell lists exist for 1..6 dimensions and at least one simplex");
    if (rc == 2) return fail(CHEBPOL_CUDA_ESIZE, "start / items buffers too small");
    return CHEBPOL_CUDA_OK;
}

extern "C" int chebpol_cuda_plan_set(chebpol_cuda_plan *p, int option, int64_t value)
{
    if (!p) return fail(CHEBPOL_CUDA_EINVAL, "plan is NULL");
    switch (option) {
    case CHEBPOL_CUDA_OPT_KERNEL:
        if (value < 0 || value > 2) return fail(CHEBPOL_CUDA_EINVAL, "kernel option %lld", (long long)value);
        p->kernel_opt = (int)value;
        return CHEBPOL_CUDA_OK;
    case CHEBPOL_CUDA_OPT_BLEND:
        if (p->kind != PLAN_MLIP && p->kind != PLAN_STALK && p->kind != PLAN_SL && p->kind != PLAN_OLDSTALK)
            return fail(CHEBPOL_CUDA_EINVAL, "blend applies to multilinear, stalker and simplex-linear plans");
        if (value < 0 || value > 5) return fail(CHEBPOL_CUDA_EBLEND, "blend %lld not in 0..5", (long long)value);
        p->blend = (int)value;
        return CHEBPOL_CUDA_OK;
    case CHEBPOL_CUDA_OPT_CHUNK_POINTS:
        if (value < 0) return fail(CHEBPOL_CUDA_EINVAL, "negative chunk size");
        p->chunk_points = value;
        return CHEBPOL_CUDA_OK;
    case CHEBPOL_CUDA_OPT_VARIANT:
        p->variant = (int)value;
        return CHEBPOL_CUDA_OK;
    case CHEBPOL_CUDA_OPT_SL_CELLS: {
        if (p->kind != PLAN_SL) return fail(CHEBPOL_CUDA_EINVAL, "cell grid applies to simplex-linear plans");
        if (value < 0 || value > 4096) return fail(CHEBPOL_CUDA_EINVAL, "cells per dimension %lld", (long long)value);
        int cur = -1;
        CU(cudaGetDevice(&cur));
        if (cur != p->device) CU(cudaSetDevice(p->device));
        cudaError_t e = cudaSuccess;
        try {
            e = sl_force_cells(p->sl, (int)value);
        } catch (const std::exception &ex) {
            if (cur != p->device) cudaSetDevice(cur);
            return fail(CHEBPOL_CUDA_ENOMEM, "candidate lists: %s", ex.what());
        }
        if (cur != p->device) cudaSetDevice(cur);
        if (e == cudaErrorNotSupported) return fail(CHEBPOL_CUDA_EINVAL, "no cell-list kernel for %d-dimensional simplices", p->rank);
        if (e != cudaSuccess) return cuda_fail(e, "cell lists");
        p->tensor_bytes = sl_bytes(p->sl);
        return CHEBPOL_CUDA_OK;
    }
    case CHEBPOL_CUDA_OPT_EPOL:
        if (p->kind != PLAN_SL) return fail(CHEBPOL_CUDA_EINVAL, "epol applies to simplex-linear plans");
        p->epol = value != 0;
        return CHEBPOL_CUDA_OK;
    default: return fail(CHEBPOL_CUDA_EINVAL, "unknown option %d", option);
    }
}

extern "C" int chebpol_cuda_plan_get(chebpol_cuda_plan *p, int what, int64_t *value)
{
    if (!p || !value) return fail(CHEBPOL_CUDA_EINVAL, "plan or value is NULL");
    switch (what) {
    case CHEBPOL_CUDA_GET_KERNEL_USED: *value = p->last.kernel; break;
    case CHEBPOL_CUDA_GET_LAUNCHES: *value = p->launches; break;
    case CHEBPOL_CUDA_GET_TENSOR_BYTES: *value = p->tensor_bytes; break;
    case CHEBPOL_CUDA_GET_DEVICE: *value = p->device; break;
    case CHEBPOL_CUDA_GET_SMEM_BYTES: *value = p->last.smem; break;
    case CHEBPOL_CUDA_GET_GRID: *value = p->last.grid; break;
    case CHEBPOL_CUDA_GET_RANK: *value = p->rank; break;
    case CHEBPOL_CUDA_GET_MMA_MACS:
        *value = p->mma_ok ? (int64_t)p->mma.nrows * p->mma.kp + p->mma.nrows + (int64_t)(p->mma.nrows / (p->mma.nAp > 0 ? p->mma.nAp : 1)) : 0;
        break;
    default: return fail(CHEBPOL_CUDA_EINVAL, "unknown query %d", what);
    }
    return CHEBPOL_CUDA_OK;
}

// A derived table has just been enqueued on stream s: remember where, so that an evaluation on
// any other stream (chebpol_cuda_plan_eval_device takes the caller's) is ordered after the build.
static int mark_lazy_build(chebpol_cuda_plan *p, cudaStream_t s)
{
    if (!p->ev_build) CU(cudaEventCreateWithFlags(&p->ev_build, cudaEventDisableTiming));
    CU(cudaEventRecord(p->ev_build, s));
    p->build_stream = s;
    p->built_lazily = true;
    return CHEBPOL_CUDA_OK;
}

// Row-padded copy of the tensor for the DFMA slab kernel, made on the device from the
// resident R-layout tensor the first time that kernel is asked for.
static int ensure_slab_tensor(chebpol_cuda_plan *p, cudaStream_t s)
{
    if (p->d_padded || !p->slab_ok) return CHEBPOL_CUDA_OK;
    const size_t n0 = (size_t)p->dims[0], n0p = (size_t)p->slab.n0p, rows = (size_t)(p->nelem / p->dims[0]);
    CU(cudaMalloc(&p->d_padded, sizeof(double) * rows * n0p));
    CU(cudaMemsetAsync(p->d_padded, 0, sizeof(double) * rows * n0p, s));
    CU(cudaMemcpy2DAsync(p->d_padded, n0p * 8, p->d_tensor, n0 * 8, n0 * 8, rows, cudaMemcpyDeviceToDevice, s));
    p->slab.cf = p->d_padded;
    p->tensor_bytes += sizeof(double) * rows * n0p;
    return mark_lazy_build(p, s);
}

// ------------------------------------------------------------------- launch
// The current device must be p->device.
static int launch_plan(chebpol_cuda_plan *p, const double *d_x, int64_t npts, double *d_out, cudaStream_t s)
{
    if (npts == 0) return CHEBPOL_CUDA_OK;
    if (p->built_lazily && s != p->build_stream) CU(cudaStreamWaitEvent(s, p->ev_build, 0));
    cudaError_t e = cudaErrorNotSupported;
    LaunchInfo li = {0, 0, 0, 0};
    const bool want_fast = p->kernel_opt != CHEBPOL_CUDA_KERNEL_STRICT;
    const bool aligned16 = (((uintptr_t)d_x) & 15) == 0;
    // variant: 0 auto (DMMA engine, else DFMA slab kernel, else strict); 1..99999 tuning
    // code of the DFMA slab kernel; >= 100000 DMMA engine with tuning code variant-100000
    const bool force_slab = p->variant > 0 && p->variant < 100000;
    const int mma_variant = p->variant >= 100000 ? p->variant - 100000 : 0;
    if (p->kind == PLAN_CHEB) {
        // AUTO between the fast kernels (measured crossover, tools/crossover.py; S = Clenshaw
        // multiply-adds per point): below S ~ 25000 the DMMA engine's per-pass basis prologue
        // dominates and the DFMA slab kernel wins (up to 6x on 2-D tensors); above it the DMMA
        // engine wins by 1.1-1.7x.  Small tensors the slab kernel cannot take (first dimension
        // too long) go to the strict kernel, which beats the DMMA engine there (64x64: 1.5x).
        bool prefer_slab = false, prefer_strict = false;
        if (p->variant == 0 && p->mma_ok) {
            double S = 0, prod = 1;
            for (int d = p->rank - 1; d >= 0; --d) { prod *= p->dims[d]; S += prod; }
            const bool slab_healthy = p->slab_ok && (p->slab.slab_elems >= 256 || p->slab.nslabs == 1);
            prefer_slab = slab_healthy && S < 25000.0;
            prefer_strict = !slab_healthy && S < 8000.0 && p->kernel_opt == CHEBPOL_CUDA_KERNEL_AUTO;
        }
        if (want_fast && p->mma_ok && !force_slab && !prefer_slab && !prefer_strict) {
            MmaArgs a = p->mma;
            a.x = d_x;
            a.npts = npts;
            a.out = d_out;
            e = launch_contract_mma(a, mma_variant, p->num_sms, s, &li);
        }
        if (e == cudaErrorNotSupported && want_fast && p->slab_ok && p->variant < 100000) {
            const int rcs = ensure_slab_tensor(p, s);
            if (rcs) return rcs;
            SlabArgs a = p->slab;
            a.x = d_x;
            a.npts = npts;
            a.out = d_out;
            e = launch_cheb_slab(a, p->variant, p->num_sms, s, &li);
        }
        if (e == cudaErrorNotSupported) {
            if (p->kernel_opt == CHEBPOL_CUDA_KERNEL_FAST)
                return fail(CHEBPOL_CUDA_EINVAL, "no fast Chebyshev kernel for dims[0]=%d (variant %d)", p->dims[0], p->variant);
            ChebStrictArgs a;
            a.cf = p->d_tensor;
            a.rank = p->rank;
            for (int i = 0; i < p->rank; ++i) { a.dims[i] = p->dims[i]; a.stride[i] = p->stride[i]; }
            a.x = d_x;
            a.npts = npts;
            a.out = d_out;
            a.pm = p->pm;
            e = launch_cheb_strict(a, s, &li);
        }
    } else if (p->kind == PLAN_MLIP) {
        MlipArgs a;
        a.values = p->d_tensor;
        a.grid = p->d_grid;
        a.rank = p->rank;
        for (int i = 0; i < p->rank; ++i) { a.dims[i] = p->dims[i]; a.stride[i] = p->stride[i]; a.goff[i] = p->goff[i]; }
        a.blend = p->blend;
        a.x = d_x;
        a.npts = npts;
        a.out = d_out;
        a.lut = p->d_lut;
        for (int i = 0; i < p->rank; ++i) { a.loff[i] = p->loff[i]; a.nb[i] = p->nbk[i]; a.lscale[i] = p->lscale[i]; }
        // variant: 0 auto, 1 pair-gather kernel, 2 brick / cell-major kernel (always expand), 3 quad kernel
        // (always expand), 4 binned path, 20..199 brick kernel with layout code warps*10 + stages,
        // 300+c quad kernel with c CTAs per SM.
        // AUTO climbs a ladder of table memory against speed (what decides is how many
        // DRAM pages a point opens -- B200 serves about 4e10 random accesses a second whatever their
        // size -- 1 for the cell-major table, 2^(D-m) for bricks, 2^(D-2) for quads, 2^(D-1) for pairs):
        //   bricks over the first m dims, the largest 3 <= m <= D that fits the cap (m = D: cell-major),
        //     once a batch has at least cells/8 points;
        //   else the row-pair-interleaved quad table (2 x) once a batch has at least half as many
        //     points as the table has values;
        //   else, if even that cannot be allocated, the binned path (no extra table memory);
        //   else the pair gather on the values as given.
        const bool want_quad = p->variant == 3 || (p->variant >= 300 && p->variant < 400);
        const bool want_bin = p->variant == 4;
        const bool allow_cell = p->variant == 0 || p->variant == 2 || (p->variant >= 20 && p->variant < 200);
        if (want_fast && allow_cell && p->cell_bytes > 0 && !p->d_cells && !p->cells_failed &&
            (p->variant >= 2 || (npts > p->call_points ? npts : p->call_points) >= p->ncells / 8)) {
            size_t free_b = 0, total_b = 0;
            cudaError_t ce = cudaMemGetInfo(&free_b, &total_b);
            if (ce == cudaSuccess && (size_t)p->cell_bytes < free_b / 2) ce = cudaMalloc(&p->d_cells, (size_t)p->cell_bytes);
            else if (ce == cudaSuccess) ce = cudaErrorMemoryAllocation;
            if (ce == cudaSuccess) ce = mlip_expand(a, p->d_tensor, p->d_cells, p->cell_m, s);
            if (ce != cudaSuccess) {
                (void)cudaGetLastError();
                cudaFree(p->d_cells);
                p->d_cells = nullptr;
                p->cells_failed = true;  // fall back to the pair kernel for the plan's lifetime
            } else {
                p->tensor_bytes += p->cell_bytes;
                p->launches += 1;
                g_launches.fetch_add(1);
                const int rcb = mark_lazy_build(p, s);
                if (rcb) return rcb;
            }
        }
        if (want_fast && allow_cell && p->d_cells) e = launch_mlip_cell(a, p->d_cells, p->cell_m, p->variant, p->num_sms, s, &li);
        if (e == cudaErrorNotSupported && want_fast && (want_quad || p->variant == 0) && p->quad_bytes > 0 && !p->quads_failed) {
            if (!p->d_quads && (want_quad || (npts > p->call_points ? npts : p->call_points) >= p->nelem / 2)) {
                size_t free_b = 0, total_b = 0;
                cudaError_t ce = cudaMemGetInfo(&free_b, &total_b);
                if (ce == cudaSuccess && (size_t)p->quad_bytes < free_b / 2) ce = cudaMalloc(&p->d_quads, (size_t)p->quad_bytes);
                else if (ce == cudaSuccess) ce = cudaErrorMemoryAllocation;
                if (ce == cudaSuccess) ce = mlip_quad_expand(a, p->d_tensor, p->d_quads, s);
                if (ce != cudaSuccess) {
                    (void)cudaGetLastError();
                    cudaFree(p->d_quads);
                    p->d_quads = nullptr;
                    p->quads_failed = true;  // the pair kernel serves the plan
                } else {
                    p->tensor_bytes += p->quad_bytes;
                    p->launches += 1;
                    g_launches.fetch_add(1);
                    const int rcb = mark_lazy_build(p, s);
                    if (rcb) return rcb;
                }
            }
            if (p->d_quads) e = launch_mlip_quad(a, p->d_quads, want_quad && p->variant >= 300 ? p->variant - 300 : 0, p->num_sms, s, &li);
        }
        if (e == cudaErrorNotSupported && want_fast && (want_bin || (p->variant == 0 && p->quads_failed)) && !p->bin_failed && aligned16 &&
            (want_bin || p->nelem * 8 > (48ll << 20)) && mlip_bin_supported(p->rank, p->dims, npts, p->blend, want_bin)) {
            cudaError_t ce = mlip_bin_prepare(&p->bin, p->rank, p->dims, npts);
            if (ce == cudaSuccess) {
                int nl = 0;
                e = launch_mlip_bin(a, p->bin, p->num_sms, s, &li, &nl);
                if (e == cudaSuccess && nl > 1) {
                    p->launches += nl - 1;
                    g_launches.fetch_add(nl - 1);
                }
            } else {
                (void)cudaGetLastError();
                if (ce != cudaErrorNotSupported) p->bin_failed = true;  // no room for the workspace
            }
        }
        if (e == cudaErrorNotSupported && want_fast && p->variant < 2 && aligned16 && mlip_pair_supported(p->rank, p->blend))
            e = launch_mlip_pair(a, p->variant, p->num_sms, s, &li);
        if (e == cudaErrorNotSupported) {
            if (p->kernel_opt == CHEBPOL_CUDA_KERNEL_FAST)
                return fail(CHEBPOL_CUDA_EINVAL, "no fast multilinear kernel for rank %d", p->rank);
            e = launch_mlip_strict(a, s, &li);
        }
    } else if (p->kind == PLAN_FH) {
        FhArgs a;
        a.vals = p->d_tensor;
        a.grid = p->d_grid;
        a.weights = p->d_weights;
        a.rank = p->rank;
        for (int i = 0; i < p->rank; ++i) { a.dims[i] = p->dims[i]; a.stride[i] = p->stride[i]; a.goff[i] = p->goff[i]; }
        a.x = d_x;
        a.npts = npts;
        a.out = d_out;
        if (want_fast && p->mma_ok) {
            MmaArgs m = p->mma;
            m.x = d_x;
            m.npts = npts;
            m.out = d_out;
            e = launch_contract_mma(m, mma_variant, p->num_sms, s, &li);
        }
        if (e == cudaErrorNotSupported) {
            if (p->kernel_opt == CHEBPOL_CUDA_KERNEL_FAST)
                return fail(CHEBPOL_CUDA_EINVAL, "no fast Floater-Hormann kernel for this shape");
            e = launch_fh_strict(a, s, &li);
        }
    } else if (p->kind == PLAN_POLYH) {
        PolyhArgs a;
        a.knots = p->d_tensor;
        a.weights = p->d_weights;
        a.lweights = p->d_grid;
        a.packed = p->d_padded;
        a.rank = p->rank;
        a.nknots = p->nknots;
        a.k = p->polyh_k;
        a.x = d_x;
        a.npts = npts;
        a.out = d_out;
        a.pm = p->pm;
        if (want_fast) e = launch_polyh_tile(a, p->variant, p->num_sms, s, &li);
        if (e == cudaErrorNotSupported) {
            if (p->kernel_opt == CHEBPOL_CUDA_KERNEL_FAST)
                return fail(CHEBPOL_CUDA_EINVAL, "no fast polyharmonic kernel for rank %d with %d knots", p->rank, p->nknots);
            e = launch_polyh_strict(a, s, &li);
        }
    } else if (p->kind == PLAN_STALK) {
        StalkArgs a;
        a.kind = p->stalk_kind;
        a.val = p->d_tensor;
        a.a = p->d_stalk[0];
        a.b = p->d_stalk[1];
        a.t1 = p->d_stalk[2];
        a.t2 = p->d_stalk[3];
        a.records = p->d_padded;
        a.grid = p->d_grid;
        a.rank = p->rank;
        for (int i = 0; i < p->rank; ++i) { a.dims[i] = p->dims[i]; a.stride[i] = p->stride[i]; a.goff[i] = p->goff[i]; }
        a.blend = p->blend;
        a.x = d_x;
        a.npts = npts;
        a.out = d_out;
        if (want_fast) e = launch_stalk_cell(a, p->variant, p->num_sms, s, &li);
        if (e == cudaErrorNotSupported) {
            if (p->kernel_opt == CHEBPOL_CUDA_KERNEL_FAST)
                return fail(CHEBPOL_CUDA_EINVAL, "no fast stalker kernel for rank %d", p->rank);
            e = launch_stalk_strict(a, s, &li);
        }
    } else if (p->kind == PLAN_OLDSTALK) {
        e = launch_oldstalk(p->d_tensor, p->d_grid, p->d_stalk[0], p->d_stalk[1], p->d_stalk[2], p->goff, p->rank, p->dims,
                            p->stride, p->os_degree, p->os_uniform, p->blend, d_x, npts, d_out, p->num_sms, s, &li);
    } else if (p->kind == PLAN_SL) {
        try {  // may (re)build the candidate lists on the host
            e = launch_sl(p->sl, p->kernel_opt, p->variant, p->epol, p->blend, d_x, npts, d_out, p->num_sms, s, &li);
        } catch (const std::exception &ex) {
            return fail(CHEBPOL_CUDA_ENOMEM, "candidate lists: %s", ex.what());
        }
        if (e == cudaErrorNotSupported)
            return fail(CHEBPOL_CUDA_EINVAL, "no cell-list kernel for %d-dimensional simplices", p->rank);
        p->tensor_bytes = sl_bytes(p->sl);  // the candidate lists may just have been (re)built
    } else {
        return fail(CHEBPOL_CUDA_EINVAL, "corrupt plan");
    }
    if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
    p->last = li;
    p->launches += 1;
    g_launches.fetch_add(1);
    return CHEBPOL_CUDA_OK;
}

extern "C" int chebpol_cuda_plan_eval_device(chebpol_cuda_plan *p, const double *d_x, int64_t npts,
                                             double *d_out, void *stream)
{
    if (!p) return fail(CHEBPOL_CUDA_EINVAL, "plan is NULL");
    if (npts < 0) return fail(CHEBPOL_CUDA_EINVAL, "negative point count");
    if (npts > 0 && (!d_x || !d_out)) return fail(CHEBPOL_CUDA_EINVAL, "x or out is NULL");
    int cur = -1;
    CU(cudaGetDevice(&cur));
    if (cur != p->device) CU(cudaSetDevice(p->device));
    int rc = launch_plan(p, d_x, npts, d_out, (cudaStream_t)stream);
    if (cur != p->device) cudaSetDevice(cur);
    return rc;
}

// ------------------------------------------------------------ pageable staging
// R hands over pageable memory.  cudaMemcpyAsync from pageable memory is staged by the driver
// on one thread (measured 12 GB/s against 66 GB/s from pinned memory), so large pageable
// batches go through our own pinned staging buffers, filled and drained by several host
// threads (OpenMP) while the GPU works on the neighbouring chunks.
#include <omp.h>

struct Staging {
    double *in[2] = {nullptr, nullptr};
    double *out[2] = {nullptr, nullptr};
    size_t in_bytes = 0, out_bytes = 0;
};
static std::vector<Staging *> g_staging_free;
static std::mutex g_staging_mu;
static const size_t kStageInBytes = 32u << 20;   // per input staging buffer
static const size_t kStageOutBytes = 16u << 20;  // per result staging buffer (rank 1 is limited by this one)

static Staging *staging_get(size_t in_bytes, size_t out_bytes)
{
    {
        std::lock_guard<std::mutex> lk(g_staging_mu);
        for (size_t i = 0; i < g_staging_free.size(); ++i)
            if (g_staging_free[i]->in_bytes >= in_bytes && g_staging_free[i]->out_bytes >= out_bytes) {
                Staging *s = g_staging_free[i];
                g_staging_free.erase(g_staging_free.begin() + i);
                return s;
            }
    }
    Staging *s = new Staging();
    bool ok = true;
    for (int i = 0; i < 2 && ok; ++i) {
        ok = cudaHostAlloc(&s->in[i], in_bytes, cudaHostAllocPortable) == cudaSuccess &&
             cudaHostAlloc(&s->out[i], out_bytes, cudaHostAllocPortable) == cudaSuccess;
    }
    if (!ok) {
        (void)cudaGetLastError();
        for (int i = 0; i < 2; ++i) { cudaFreeHost(s->in[i]); cudaFreeHost(s->out[i]); }
        delete s;
        return nullptr;
    }
    s->in_bytes = in_bytes;
    s->out_bytes = out_bytes;
    return s;
}

static void staging_put(Staging *s)
{
    Staging *victim = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_staging_mu);
        if (g_staging_free.size() < 8) { g_staging_free.push_back(s); return; }
        // full: keep the larger buffers (a pool of sets that are too small for the next caller would make
        // every call pin and unpin its own -- cudaHostAlloc costs tens of milliseconds)
        size_t small = 0;
        for (size_t i = 1; i < g_staging_free.size(); ++i)
            if (g_staging_free[i]->in_bytes + g_staging_free[i]->out_bytes < g_staging_free[small]->in_bytes + g_staging_free[small]->out_bytes) small = i;
        if (g_staging_free[small]->in_bytes + g_staging_free[small]->out_bytes < s->in_bytes + s->out_bytes) {
            victim = g_staging_free[small];
            g_staging_free[small] = s;
        } else {
            victim = s;
        }
    }
    for (int i = 0; i < 2; ++i) { cudaFreeHost(victim->in[i]); cudaFreeHost(victim->out[i]); }
    delete victim;
}

static int staging_threads()
{
    static int n = 0;
    if (n == 0) {
        const char *e = getenv("CHEBPOL_CUDA_STAGING_THREADS");
        n = e ? atoi(e) : 0;
        if (n <= 0) {
            n = (int)std::thread::hardware_concurrency() / 2;
            if (n > 8) n = 8;
        }
        if (n < 1) n = 1;
    }
    return n;
}

// Copy with non-temporal stores: the destination (a pinned staging buffer about to be read by the
// DMA engine, or the caller's result vector) is not read again by this core, so the write-allocate
// read of every destination line is pure waste on a path that is bound by host memory bandwidth.
#include <immintrin.h>
__attribute__((target("avx2"))) static void copy_stream_avx2(void *dst, const void *src, size_t bytes)
{
    char *d = static_cast<char *>(dst);
    const char *s = static_cast<const char *>(src);
    size_t head = (32 - ((uintptr_t)d & 31)) & 31;
    if (head > bytes) head = bytes;
    memcpy(d, s, head);
    d += head; s += head; bytes -= head;
    size_t i = 0;
    for (; i + 128 <= bytes; i += 128) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + i));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + i + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + i + 64));
        const __m256i e = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + i + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i *>(d + i), a);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(d + i + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(d + i + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(d + i + 96), e);
    }
    _mm_sfence();
    memcpy(d + i, s + i, bytes - i);
}

static void copy_stream(void *dst, const void *src, size_t bytes)
{
    static const int mode = [] {
        const char *e = getenv("CHEBPOL_CUDA_STAGING_NT");
        if (e && atoi(e) == 0) return 0;
        return __builtin_cpu_supports("avx2") ? 1 : 0;
    }();
    if (mode == 1 && bytes >= 4096) copy_stream_avx2(dst, src, bytes);
    else memcpy(dst, src, bytes);
}

static void par_memcpy(void *dst, const void *src, size_t bytes, int T)
{
    if (T == 1 || bytes < (4u << 20)) { copy_stream(dst, src, bytes); return; }
    const size_t slice = ((bytes + T - 1) / T + 4095) & ~(size_t)4095;
#pragma omp parallel for num_threads(T) schedule(static)
    for (int i = 0; i < T; ++i) {
        const size_t lo = (size_t)i * slice;
        if (lo < bytes) copy_stream((char *)dst + lo, (const char *)src + lo, (lo + slice <= bytes) ? slice : bytes - lo);
    }
}

static bool is_pageable(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return true;
    }
    return at.type == cudaMemoryTypeUnregistered;
}

// ------------------------------------------------------------ host pipeline
static int ensure_pipeline(chebpol_cuda_plan *p, int64_t chunk)
{
    if (!p->st_in) {
        CU(cudaStreamCreateWithFlags(&p->st_in, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&p->st_k, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&p->st_out, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CU(cudaEventCreateWithFlags(&p->ev_in[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&p->ev_k[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&p->ev_out[i], cudaEventDisableTiming));
        }
    }
    if (p->buf_points < chunk) {
        for (int i = 0; i < 2; ++i) {
            cudaFree(p->d_x[i]);
            cudaFree(p->d_o[i]);
            p->d_x[i] = p->d_o[i] = nullptr;
        }
        p->buf_points = 0;
        for (int i = 0; i < 2; ++i) {
            CU(cudaMalloc(&p->d_x[i], sizeof(double) * chunk * p->rank));
            CU(cudaMalloc(&p->d_o[i], sizeof(double) * chunk));
        }
        p->buf_points = chunk;
    }
    return CHEBPOL_CUDA_OK;
}

extern "C" int chebpol_cuda_plan_eval_host(chebpol_cuda_plan *p, const double *x, int64_t npts, double *out)
{
    if (!p) return fail(CHEBPOL_CUDA_EINVAL, "plan is NULL");
    if (npts < 0) return fail(CHEBPOL_CUDA_EINVAL, "negative point count");
    if (npts == 0) return CHEBPOL_CUDA_OK;
    if (!x || !out) return fail(CHEBPOL_CUDA_EINVAL, "x or out is NULL");
    NvtxRange nvtx("chebpol_cuda:eval_host");
    int cur = -1;
    CU(cudaGetDevice(&cur));
    if (cur != p->device) CU(cudaSetDevice(p->device));

    struct CallPoints {  // the whole call's size, visible to launch_plan while the chunks run
        chebpol_cuda_plan *p;
        CallPoints(chebpol_cuda_plan *q, int64_t n) : p(q) { p->call_points = n; }
        ~CallPoints() { p->call_points = 0; }
    } call_points(p, npts);
    int64_t chunk = p->chunk_points;
    // large pageable batches: our own multi-threaded pinned staging (smaller chunks)
    Staging *stg = nullptr;
    if (chunk <= 0 && (size_t)npts * p->rank * 8 >= (16u << 20) && (is_pageable(x) || is_pageable(out))) {
        // every staging set has the same size, whatever the rank, so that any set serves any plan
        int64_t sc = (int64_t)(kStageInBytes / (8 * (size_t)p->rank));
        if (sc > (int64_t)(kStageOutBytes / 8)) sc = (int64_t)(kStageOutBytes / 8);
        stg = staging_get(kStageInBytes, kStageOutBytes);
        if (stg) chunk = sc;
    }
    if (chunk <= 0) {
        chunk = (int64_t(128) << 20) / (8 * p->rank);
        if (chunk < (1 << 20)) chunk = 1 << 20;
        // the first chunk's upload and the last chunk's download are not overlapped with anything: at least
        // eight chunks (of at least 2^18 points) keep that below an eighth of the copy time
        int64_t c8 = (npts + 7) / 8;
        if (c8 < (1 << 18)) c8 = 1 << 18;
        if (chunk > c8) chunk = c8;
    }
    if (chunk > npts) chunk = npts;
    int rc = ensure_pipeline(p, chunk);
    if (rc) { if (stg) staging_put(stg); return rc; }
    const int nthreads = p->host_threads > 0 ? p->host_threads : staging_threads();
    // CHEBPOL_CUDA_TRACE=1: where the host thread of a staged call spends its time (stderr, one line per call)
    static const bool trace = getenv("CHEBPOL_CUDA_TRACE") && atoi(getenv("CHEBPOL_CUDA_TRACE")) != 0;
    double t_in = 0, t_out = 0, t_wait_buf = 0, t_wait_res = 0;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = trace ? now() : 0.0;
    if (stg) {
        // chunk c: threads fill in[b]; H2D -> kernel -> D2H into out[b]; chunk c-1 is drained to
        // the caller's buffer while chunk c runs on the GPU
        const int64_t nch = (npts + chunk - 1) / chunk;
        auto bail = [&](int code) {
            cudaStreamSynchronize(p->st_in);
            cudaStreamSynchronize(p->st_k);
            cudaStreamSynchronize(p->st_out);
            staging_put(stg);
            if (cur != p->device) cudaSetDevice(cur);
            return code;
        };
        auto drain = [&](int64_t c) -> int {
            const int b = (int)(c & 1);
            const int64_t off = c * chunk, n = (off + chunk <= npts) ? chunk : npts - off;
            const double t0 = trace ? now() : 0.0;
            cudaError_t e = cudaEventSynchronize(p->ev_out[b]);
            if (e != cudaSuccess) return cuda_fail(e, "D2H (staged)");
            const double t1 = trace ? now() : 0.0;
            par_memcpy(out + off, stg->out[b], sizeof(double) * n, nthreads);
            if (trace) { t_wait_res += t1 - t0; t_out += now() - t1; }
            return CHEBPOL_CUDA_OK;
        };
        for (int64_t c = 0; c < nch; ++c) {
            const int b = (int)(c & 1);
            const int64_t off = c * chunk, n = (off + chunk <= npts) ? chunk : npts - off;
            const double t0 = trace ? now() : 0.0;
            if (c >= 2) {
                cudaError_t e = cudaEventSynchronize(p->ev_in[b]);  // staging buffer b has been read
                if (e != cudaSuccess) return bail(cuda_fail(e, "H2D (staged)"));
            }
            const double t1 = trace ? now() : 0.0;
            par_memcpy(stg->in[b], x + off * p->rank, sizeof(double) * n * p->rank, nthreads);
            if (trace) { t_wait_buf += t1 - t0; t_in += now() - t1; }
            cudaError_t e = cudaSuccess;
            if (c >= 2) e = cudaStreamWaitEvent(p->st_in, p->ev_k[b], 0);
            if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_x[b], stg->in[b], sizeof(double) * n * p->rank, cudaMemcpyHostToDevice, p->st_in);
            if (e == cudaSuccess) e = cudaEventRecord(p->ev_in[b], p->st_in);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(p->st_k, p->ev_in[b], 0);
            if (e == cudaSuccess && c >= 2) e = cudaStreamWaitEvent(p->st_k, p->ev_out[b], 0);
            if (e != cudaSuccess) return bail(cuda_fail(e, "staged pipeline"));
            rc = launch_plan(p, p->d_x[b], n, p->d_o[b], p->st_k);
            if (rc) return bail(rc);
            e = cudaEventRecord(p->ev_k[b], p->st_k);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(p->st_out, p->ev_k[b], 0);
            if (e == cudaSuccess) e = cudaMemcpyAsync(stg->out[b], p->d_o[b], sizeof(double) * n, cudaMemcpyDeviceToHost, p->st_out);
            if (e == cudaSuccess) e = cudaEventRecord(p->ev_out[b], p->st_out);
            if (e != cudaSuccess) return bail(cuda_fail(e, "staged pipeline"));
            if (c >= 1) {
                rc = drain(c - 1);
                if (rc) return bail(rc);
            }
        }
        rc = drain(nch - 1);
        if (trace)
            fprintf(stderr, "chebpol_cuda trace: device %d, %lld points in %lld chunks, %d host threads: total %.1f ms = copy-in %.1f (%.1f GB/s) + "
                            "copy-out %.1f + wait for a free staging buffer %.1f + wait for results %.1f + enqueue/other %.1f\n",
                    p->device, (long long)npts, (long long)nch, nthreads, 1e3 * (now() - t_begin), 1e3 * t_in,
                    t_in > 0 ? npts * p->rank * 8.0 / t_in / 1e9 : 0.0, 1e3 * t_out, 1e3 * t_wait_buf, 1e3 * t_wait_res,
                    1e3 * (now() - t_begin - t_in - t_out - t_wait_buf - t_wait_res));
        return bail(rc);
    }
    const int64_t nchunks = (npts + chunk - 1) / chunk;
    if (nchunks == 1) {
        // a small call (the closure on one or a few points): one stream, one synchronisation
        cudaError_t e = cudaMemcpyAsync(p->d_x[0], x, sizeof(double) * npts * p->rank, cudaMemcpyHostToDevice, p->st_k);
        if (e == cudaSuccess) {
            rc = launch_plan(p, p->d_x[0], npts, p->d_o[0], p->st_k);
            if (!rc) e = cudaMemcpyAsync(out, p->d_o[0], sizeof(double) * npts, cudaMemcpyDeviceToHost, p->st_k);
        }
        const cudaError_t es = cudaStreamSynchronize(p->st_k);
        if (cur != p->device) cudaSetDevice(cur);
        if (rc) return rc;
        if (e != cudaSuccess) return cuda_fail(e, "small-call pipeline");
        if (es != cudaSuccess) return cuda_fail(es, "kernel stream");
        return CHEBPOL_CUDA_OK;
    }
    // chunk c uses buffer c&1:  H2D(c) -> kernel(c) -> D2H(c); D2H(c-1) is
    // enqueued after kernel(c) so that a blocking (pageable) copy-out overlaps
    // the next kernel.
    // Nothing in the loop returns early: asynchronous copies that touch the caller's x / out may be
    // in flight, so every exit goes through the stream synchronisation below.
    auto copy_out = [&](int64_t c) -> int {
        const int b = (int)(c & 1);
        const int64_t off = c * chunk, n = (off + chunk <= npts) ? chunk : npts - off;
        cudaError_t e = cudaStreamWaitEvent(p->st_out, p->ev_k[b], 0);
        if (e == cudaSuccess) e = cudaMemcpyAsync(out + off, p->d_o[b], sizeof(double) * n, cudaMemcpyDeviceToHost, p->st_out);
        if (e == cudaSuccess) e = cudaEventRecord(p->ev_out[b], p->st_out);
        return e == cudaSuccess ? CHEBPOL_CUDA_OK : cuda_fail(e, "D2H pipeline");
    };
    for (int64_t c = 0; c < nchunks && !rc; ++c) {
        const int b = (int)(c & 1);
        const int64_t off = c * chunk, n = (off + chunk <= npts) ? chunk : npts - off;
        cudaError_t e = cudaSuccess;
        if (c >= 2) e = cudaStreamWaitEvent(p->st_in, p->ev_k[b], 0);  // x buffer free
        if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_x[b], x + off * p->rank, sizeof(double) * n * p->rank, cudaMemcpyHostToDevice, p->st_in);
        if (e == cudaSuccess) e = cudaEventRecord(p->ev_in[b], p->st_in);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(p->st_k, p->ev_in[b], 0);
        if (e == cudaSuccess && c >= 2) e = cudaStreamWaitEvent(p->st_k, p->ev_out[b], 0);  // out buffer free
        if (e != cudaSuccess) { rc = cuda_fail(e, "H2D pipeline"); break; }
        rc = launch_plan(p, p->d_x[b], n, p->d_o[b], p->st_k);
        if (rc) break;
        e = cudaEventRecord(p->ev_k[b], p->st_k);
        if (e != cudaSuccess) { rc = cuda_fail(e, "kernel event"); break; }
        if (c >= 1) rc = copy_out(c - 1);
    }
    if (!rc) rc = copy_out(nchunks - 1);
    cudaError_t e1 = cudaStreamSynchronize(p->st_in);
    cudaError_t e2 = cudaStreamSynchronize(p->st_k);
    cudaError_t e3 = cudaStreamSynchronize(p->st_out);
    if (cur != p->device) cudaSetDevice(cur);
    if (rc) return rc;
    if (e1 != cudaSuccess) return cuda_fail(e1, "H2D stream");
    if (e2 != cudaSuccess) return cuda_fail(e2, "kernel stream");
    if (e3 != cudaSuccess) return cuda_fail(e3, "D2H stream");
    return CHEBPOL_CUDA_OK;
}

// ---------------------------------------------------------------- plan cache
// The R closures re-pass the tensor on every call (no C-side state in the reference), so a
// stateless call would rebuild and re-upload its plan each time (4 ms for a small tensor; 134 MB
// of upload plus a 2 GB cell-table expansion for the 64^4 multilinear table).  Interpolants are
// therefore kept in a process-wide cache keyed by the CONTENT of everything that defines the plan
// (kind, device, dims, byte count and a 128-bit fingerprint of tensor, knots, weights, pre-map:
// hash128.h), so a closure restored by load() or a tensor re-allocated at another address still
// hits, and a modified tensor never does.  A hit is verified: entries up to 32 MB keep a host copy
// of their defining arrays and the bytes are compared; larger ones rely on the 128-bit fingerprint
// (collision probability ~2^-128 per pair, see hash128.h).  Per device at most 8 entries and
// CHEBPOL_CUDA_CACHE_MAX_GB (default 64) of device memory, least recently used out first.
// CHEBPOL_CUDA_PLAN_CACHE=0 disables the cache.
#include "hash128.h"
#include <memory>

typedef std::shared_ptr<const std::vector<unsigned char>> KeptBytes;

struct CacheKey {
    int kind, device, rank;
    int dims[CP_MAXR];
    uint64_t bytes;
    cphash::H128 hash;
    bool operator==(const CacheKey &o) const
    {
        if (kind != o.kind || device != o.device || rank != o.rank || bytes != o.bytes || !(hash == o.hash)) return false;
        for (int i = 0; i < rank; ++i)
            if (dims[i] != o.dims[i]) return false;
        return true;
    }
};
struct CacheEntry {
    CacheKey key;
    chebpol_cuda_plan *plan;
    uint64_t stamp;
    bool busy, doomed;
    KeptBytes kept;  // null: fingerprint only (> kRetainMaxBytes)
};
static std::mutex g_cache_mu;
static std::vector<CacheEntry> g_cache;
static uint64_t g_cache_clock = 0;
static std::atomic<int64_t> g_cache_hits{0}, g_cache_misses{0}, g_cache_rejected{0};
static std::atomic<int> g_weak_hash{0};
static const size_t kCacheSlotsPerDevice = 8;
static const uint64_t kRetainMaxBytes = 32ull << 20;

static bool cache_enabled()
{
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("CHEBPOL_CUDA_PLAN_CACHE");
        on = (e && atoi(e) == 0) ? 0 : 1;
    }
    return on == 1;
}

static int64_t cache_budget_bytes()
{
    static int64_t b = -1;
    if (b < 0) {
        double gb = 64.0;  // per device; a multilinear plan with its expanded table can hold a quarter of a B200
        if (const char *e = getenv("CHEBPOL_CUDA_CACHE_MAX_GB")) gb = atof(e);
        b = gb > 0 ? (int64_t)(gb * 1e9) : 0;
    }
    return b;
}

static int host_threads_total()
{
    static int n = 0;
    if (n == 0) {
        const char *e = getenv("CHEBPOL_CUDA_HOST_THREADS");
        n = e ? atoi(e) : 0;
        if (n <= 0) {
            n = (int)std::thread::hardware_concurrency() / 2;
            if (n > 16) n = 16;
        }
        if (n < 1) n = 1;
    }
    return n;
}

// Opt-in (CHEBPOL_CUDA_CACHE_TRUST_ADDRESS=1), for interpolants of 8 MB and more: identify them by
// the ADDRESSES and lengths of their arrays plus a sparse sample of the content (both ends and 512
// strided words of every array) instead of reading every byte -- the exact fingerprint of the
// 134 MB 64^4 table costs milliseconds of host memory bandwidth per call.  This is NOT exact: an
// array modified in place between two calls, outside the sampled words, is served from the stale
// device copy.  R only modifies a vector in place when nothing else references it, which is not
// how the closures use their tensors, but the default stays the exact content key.
static bool cache_trust_address()
{
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("CHEBPOL_CUDA_CACHE_TRUST_ADDRESS");
        on = (e && atoi(e) != 0) ? 1 : 0;
    }
    return on == 1;
}

// everything that identifies the interpolant of one stateless call
struct Identity {
    CacheKey key;
    std::vector<cphash::Blob> blobs;
    bool cacheable = false;
    bool by_address = false;
    KeptBytes kept;          // made on the first miss of this call, shared by the per-device entries
    const void *verified = nullptr;  // a retained copy this call has already compared equal
    std::mutex mu;
};

static void make_identity(Identity &id, int kind, const int *dims, int rank, uint64_t seed, bool as_large = false)
{
    memset(&id.key, 0, sizeof id.key);
    id.key.kind = kind;
    id.key.rank = rank;
    for (int i = 0; i < rank && i < CP_MAXR; ++i) id.key.dims[i] = dims[i];
    id.key.bytes = cphash::blobs_bytes(id.blobs);
    id.cacheable = cache_enabled() && (int64_t)id.key.bytes <= cache_budget_bytes();
    if (!id.cacheable) return;
    if (cache_trust_address() && id.key.bytes >= (8ull << 20)) {
        std::vector<uint64_t> sample;
        for (const cphash::Blob &b : id.blobs) {
            sample.push_back((uint64_t)(uintptr_t)b.p);
            sample.push_back((uint64_t)b.n);
            const unsigned char *p = static_cast<const unsigned char *>(b.p);
            const size_t nw = b.n / 8;
            if (nw <= 2048) {
                for (size_t i = 0; i < nw; ++i) sample.push_back(cphash::rd64(p + 8 * i));
            } else {
                for (size_t i = 0; i < 512; ++i) sample.push_back(cphash::rd64(p + 8 * i));
                for (size_t i = nw - 512; i < nw; ++i) sample.push_back(cphash::rd64(p + 8 * i));
                const size_t step = nw / 512;
                for (size_t i = 0; i < 512; ++i) sample.push_back(cphash::rd64(p + 8 * (i * step + step / 2)));
            }
        }
        std::vector<cphash::Blob> sb{{sample.data(), sample.size() * 8}};
        id.key.hash = cphash::hash_blobs(sb, seed ^ 0xADD2E55ull, 1);
        id.by_address = true;
        return;
    }
    // Interpolants small enough to be retained are told apart by the byte comparison alone (one pass
    // over the arrays on a hit instead of two); the fingerprint is for the large ones.
    if ((id.key.bytes > kRetainMaxBytes || as_large) && !g_weak_hash.load())
        id.key.hash = cphash::hash_blobs(id.blobs, seed, host_threads_total());
}

// does a cached entry hold exactly the interpolant of this call?
static bool identity_matches(Identity &id, const CacheKey &ekey_in, const KeptBytes &ekept)
{
    CacheKey ekey = ekey_in;
    ekey.device = id.key.device;  // the identity describes the interpolant; entries of every device are candidates
    if (!(ekey == id.key)) return false;
    if (!ekept) return true;  // fingerprint-only entry
    {
        std::lock_guard<std::mutex> lk(id.mu);
        if (id.verified == ekept.get() || id.kept.get() == ekept.get()) return true;
    }
    if (!cphash::blobs_equal(id.blobs, *ekept, host_threads_total())) return false;
    std::lock_guard<std::mutex> lk(id.mu);
    id.verified = ekept.get();
    return true;
}

static int64_t plan_device_bytes(const chebpol_cuda_plan *p)
{
    return p->tensor_bytes + 2 * p->buf_points * (int64_t)(p->rank + 1) * 8;
}

static chebpol_cuda_plan *cache_take(Identity &id, int device)
{
    CacheKey k = id.key;
    k.device = device;
    std::vector<chebpol_cuda_plan *> skipped;
    for (;;) {
        chebpol_cuda_plan *cand = nullptr;
        KeptBytes kept;
        {
            std::lock_guard<std::mutex> lk(g_cache_mu);
            for (auto &e : g_cache) {
                if (e.busy || e.doomed || !(e.key == k)) continue;
                bool seen = false;
                for (auto *q : skipped) seen |= (q == e.plan);
                if (seen) continue;
                e.busy = true;  // held while the bytes are compared outside the lock
                cand = e.plan;
                kept = e.kept;
                break;
            }
        }
        if (!cand) {
            g_cache_misses.fetch_add(1);
            return nullptr;
        }
        Identity &idr = id;
        CacheKey kk = k;
        if (identity_matches(idr, kk, kept)) {
            std::lock_guard<std::mutex> lk(g_cache_mu);
            for (auto &e : g_cache)
                if (e.plan == cand) e.stamp = ++g_cache_clock;
            g_cache_hits.fetch_add(1);
            return cand;
        }
        // same fingerprint, different bytes: not this interpolant
        g_cache_rejected.fetch_add(1);
        skipped.push_back(cand);
        chebpol_cuda_plan *doomed = nullptr;
        {
            std::lock_guard<std::mutex> lk(g_cache_mu);
            for (size_t i = 0; i < g_cache.size(); ++i)
                if (g_cache[i].plan == cand) {
                    g_cache[i].busy = false;
                    if (g_cache[i].doomed) { doomed = cand; g_cache.erase(g_cache.begin() + i); }
                    break;
                }
        }
        if (doomed) chebpol_cuda_plan_destroy(doomed);
    }
}

// give a plan (back) to the cache; takes ownership
static void cache_put(Identity &id, int device, chebpol_cuda_plan *p)
{
    CacheKey k = id.key;
    k.device = device;
    std::vector<chebpol_cuda_plan *> evict;
    bool known = false;
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        for (size_t i = 0; i < g_cache.size(); ++i)
            if (g_cache[i].plan == p) {
                known = true;
                g_cache[i].busy = false;
                if (g_cache[i].doomed) { evict.push_back(p); g_cache.erase(g_cache.begin() + i); }
                break;
            }
    }
    if (!known) {
        KeptBytes kept;
        if (id.key.bytes <= kRetainMaxBytes && !id.by_address) {
            std::lock_guard<std::mutex> lk(id.mu);
            if (!id.kept) {
                auto v = std::make_shared<std::vector<unsigned char>>();
                cphash::blobs_copy(id.blobs, *v);
                id.kept = v;
            }
            kept = id.kept;
        }
        std::lock_guard<std::mutex> lk(g_cache_mu);
        g_cache.push_back({k, p, ++g_cache_clock, false, false, kept});
    }
    {
        // per-device limits: slots and device bytes; least recently used idle entries go first
        std::lock_guard<std::mutex> lk(g_cache_mu);
        for (;;) {
            size_t count = 0, victim = g_cache.size();
            int64_t bytes = 0;
            for (size_t i = 0; i < g_cache.size(); ++i) {
                if (g_cache[i].key.device != device) continue;
                ++count;
                bytes += plan_device_bytes(g_cache[i].plan);
                if (!g_cache[i].busy && (victim == g_cache.size() || g_cache[i].stamp < g_cache[victim].stamp)) victim = i;
            }
            if ((count <= kCacheSlotsPerDevice && bytes <= cache_budget_bytes()) || victim == g_cache.size()) break;
            evict.push_back(g_cache[victim].plan);
            g_cache.erase(g_cache.begin() + victim);
        }
    }
    for (auto *q : evict) chebpol_cuda_plan_destroy(q);
}

// a plan that failed: never goes back
static void cache_drop(chebpol_cuda_plan *p)
{
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        for (size_t i = 0; i < g_cache.size(); ++i)
            if (g_cache[i].plan == p) { g_cache.erase(g_cache.begin() + i); break; }
    }
    chebpol_cuda_plan_destroy(p);
}

extern "C" void chebpol_cuda_cache_clear(void)
{
    std::vector<chebpol_cuda_plan *> old;
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        for (size_t i = 0; i < g_cache.size();) {
            if (g_cache[i].busy) {  // another thread is evaluating it: destroyed when it comes back
                g_cache[i].doomed = true;
                ++i;
            } else {
                old.push_back(g_cache[i].plan);
                g_cache.erase(g_cache.begin() + i);
            }
        }
    }
    for (auto *p : old) chebpol_cuda_plan_destroy(p);
}

extern "C" int chebpol_cuda_debug_cache_stats(int64_t *entries, int64_t *hits, int64_t *misses, int64_t *rejected,
                                              int64_t *device_bytes)
{
    std::lock_guard<std::mutex> lk(g_cache_mu);
    int64_t bytes = 0;
    for (auto &e : g_cache) bytes += plan_device_bytes(e.plan);
    if (entries) *entries = (int64_t)g_cache.size();
    if (hits) *hits = g_cache_hits.load();
    if (misses) *misses = g_cache_misses.load();
    if (rejected) *rejected = g_cache_rejected.load();
    if (device_bytes) *device_bytes = bytes;
    return CHEBPOL_CUDA_OK;
}

extern "C" void chebpol_cuda_debug_cache_weak_hash(int on) { g_weak_hash.store(on ? 1 : 0); }

bool mma_fit_ring(MmaArgs &a, int npt);
extern "C" int chebpol_cuda_debug_mma_geometry(int kind, const int *dims, int rank, int npt, int64_t *out)
{
    if (!dims || !out || rank < 1 || rank > CP_MAXR || npt < 8 || npt % 8) return fail(CHEBPOL_CUDA_EINVAL, "mma geometry: bad arguments");
    for (int i = 0; i < 16; ++i) out[i] = 0;
    MmaArgs a = {};
    int kc = 0;
    if (!mma_configure(a, kind, rank, dims, &kc)) return CHEBPOL_CUDA_OK;
    a.occ = 1;
    if (!mma_fit_ring(a, npt)) return CHEBPOL_CUDA_OK;
    out[0] = 1; out[1] = a.nfold; out[2] = a.K; out[3] = kc; out[4] = a.rs; out[5] = a.nrows;
    out[6] = a.rows_per_slab; out[7] = a.nslabs; out[8] = a.ns; out[9] = a.stage_bytes; out[10] = a.resident;
    out[11] = mma_pair_rows(kc) ? 1 : 0;
    out[12] = 128 + (int64_t)a.ns * a.stage_bytes + (int64_t)(a.sumn + 1) * npt * 8;
    out[13] = a.nA; out[14] = a.nAp; out[15] = a.nrp;
    return CHEBPOL_CUDA_OK;
}
extern "C" int chebpol_cuda_debug_mma_row_pos(int kc, int k) { return mma_row_pos(kc, k); }

extern "C" int chebpol_cuda_debug_fingerprint(const void *data, int64_t nbytes, int threads, uint64_t *out2)
{
    if (!out2 || nbytes < 0 || (nbytes > 0 && !data)) return fail(CHEBPOL_CUDA_EINVAL, "NULL argument");
    std::vector<cphash::Blob> b{{data, (size_t)nbytes}};
    const cphash::H128 h = cphash::hash_blobs(b, 1, threads);
    out2[0] = h.a;
    out2[1] = h.b;
    return CHEBPOL_CUDA_OK;
}

// Host-only model of one cache decision (no device needed): an entry is made from array a as a
// stateless call would make it, then a call with array b looks it up.  *hit = 1 when b would be
// served by a's entry.  retain = 0 models an entry above the 32 MB retention limit.
extern "C" int chebpol_cuda_debug_cache_would_hit(const void *a, int64_t na, const void *b, int64_t nb, int retain, int *hit)
{
    if (!hit || na < 0 || nb < 0 || (na > 0 && !a) || (nb > 0 && !b)) return fail(CHEBPOL_CUDA_EINVAL, "NULL argument");
    const int dims[1] = {1};
    Identity ia, ib;
    ia.blobs.push_back({a, (size_t)na});
    ib.blobs.push_back({b, (size_t)nb});
    make_identity(ia, PLAN_CHEB, dims, 1, 1, !retain);
    make_identity(ib, PLAN_CHEB, dims, 1, 1, !retain);
    KeptBytes kept;
    if (retain) {
        auto v = std::make_shared<std::vector<unsigned char>>();
        cphash::blobs_copy(ia.blobs, *v);
        kept = v;
    }
    *hit = identity_matches(ib, ia.key, kept) ? 1 : 0;
    return CHEBPOL_CUDA_OK;
}

// -------------------------------------------------- stateless entry points
// Split [0, npts) into contiguous ranges over `ngpus` devices; one host thread
// per device builds a plan there (replicating the tensor) and runs the host
// pipeline on its slice.  No collective: every GPU writes its own slice.
// The host threads that copy pageable batches into the pinned staging buffers are a
// per-process budget (CHEBPOL_CUDA_HOST_THREADS, default min(16, cores/2)) divided among the
// devices, not a team per device.
template <class MakePlan>
static int run_sharded(int rank, const double *x, int64_t npts, int ngpus, double *out, Identity &id,
                       MakePlan make_plan, int blend = -1, int epol = -1)
{
    if (npts < 0) return fail(CHEBPOL_CUDA_EINVAL, "negative point count");
    if (npts > 0 && (!x || !out)) return fail(CHEBPOL_CUDA_EINVAL, "x or out is NULL");
    int ndev = 0;
    int rc = chebpol_cuda_device_count(&ndev);
    if (rc) return rc;
    if (ndev <= 0) return fail(CHEBPOL_CUDA_ENODEV, "no CUDA device visible");
    if (ngpus <= 0 || ngpus > ndev) ngpus = ndev;
    // do not spread tiny batches
    const int64_t min_per_gpu = 1 << 16;
    while (ngpus > 1 && npts / ngpus < min_per_gpu) --ngpus;
    const bool cacheable = id.cacheable;
    int per_dev_threads = host_threads_total() / ngpus;
    if (per_dev_threads < 1) per_dev_threads = 1;
    if (per_dev_threads > 8) per_dev_threads = 8;

    static const bool trace = getenv("CHEBPOL_CUDA_TRACE") && atoi(getenv("CHEBPOL_CUDA_TRACE")) != 0;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_call = trace ? now() : 0.0;
    auto on_device = [&](int dev, const double *xs, int64_t n, double *os) -> int {
        const double t0 = trace ? now() : 0.0;
        chebpol_cuda_plan *p = cacheable ? cache_take(id, dev) : nullptr;
        int r = CHEBPOL_CUDA_OK;
        if (!p) r = make_plan(dev, &p);
        if (!r && blend >= 0) p->blend = blend;  // per-call argument of the multilinear closure, not part of the key
        if (!r && epol >= 0) p->epol = epol;
        if (!r) p->host_threads = per_dev_threads;
        const double t1 = trace ? now() : 0.0;
        if (!r && n > 0) r = chebpol_cuda_plan_eval_host(p, xs, n, os);
        const double t2 = trace ? now() : 0.0;
        if (p) {
            if (cacheable && !r) cache_put(id, dev, p);
            else if (cacheable) cache_drop(p);  // failed: remove it from the cache if it came from there
            else chebpol_cuda_plan_destroy(p);
        }
        if (trace)
            fprintf(stderr, "chebpol_cuda trace: sharded call, device %d: thread started %.1f ms into the call, plan %.1f ms, evaluation %.1f ms, "
                            "hand-back %.1f ms\n", dev, 1e3 * (t0 - t_call), 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (now() - t2));
        return r;
    };

    if (ngpus == 1) return on_device(0, x, npts, out);
    std::vector<int> rcs(ngpus, 0);
    std::vector<std::string> msgs(ngpus);
    std::vector<std::thread> th;
    const int64_t per = (npts + ngpus - 1) / ngpus;
    for (int g = 0; g < ngpus; ++g) {
        th.emplace_back([&, g]() {
            const int64_t lo = (int64_t)g * per;
            const int64_t n = (lo >= npts) ? 0 : ((lo + per <= npts) ? per : npts - lo);
            rcs[g] = on_device(g, x + lo * rank, n, out + lo);
            if (rcs[g]) msgs[g] = g_last_error;
        });
    }
    for (auto &t : th) t.join();
    if (trace) fprintf(stderr, "chebpol_cuda trace: sharded call over %d devices took %.1f ms\n", ngpus, 1e3 * (now() - t_call));
    for (int g = 0; g < ngpus; ++g)
        if (rcs[g]) {
            g_last_error = "GPU " + std::to_string(g) + ": " + msgs[g];
            return rcs[g];
        }
    return CHEBPOL_CUDA_OK;
}

extern "C" int chebpol_cuda_evalcheb(const double *coef, const int *dims, int rank, const double *x,
                                     int64_t npts, const double *mid, const double *ispan,
                                     const int *ugm_n, int ngpus, double *out)
{
    int64_t nelem = 0;
    if (!coef) return fail(CHEBPOL_CUDA_EINVAL, "coef is NULL");
    int rc = check_dims(dims, rank, &nelem);
    if (rc) return rc;
    if ((mid == nullptr) != (ispan == nullptr)) return fail(CHEBPOL_CUDA_EINVAL, "mid and ispan must be given together");
    Identity id;
    const int32_t have[2] = {mid ? 1 : 0, ugm_n ? 1 : 0};
    id.blobs = {{have, sizeof have}, {coef, (size_t)nelem * 8}};
    if (mid) { id.blobs.push_back({mid, (size_t)rank * 8}); id.blobs.push_back({ispan, (size_t)rank * 8}); }
    if (ugm_n) id.blobs.push_back({ugm_n, (size_t)rank * 4});
    make_identity(id, PLAN_CHEB, dims, rank, 1);
    return run_sharded(rank, x, npts, ngpus, out, id, [&](int dev, chebpol_cuda_plan **p) {
        return chebpol_cuda_plan_cheb(coef, dims, rank, mid, ispan, ugm_n, dev, p);
    });
}

extern "C" int chebpol_cuda_evalchebg(const double *coef, const int *dims, int rank, const double *x,
                                      int64_t npts, const int *sn, const double *const *sx,
                                      const double *const *sy, const double *const *sb, const double *const *sc,
                                      const double *const *sd, int ngpus, double *out)
{
    int64_t nelem = 0;
    if (!coef) return fail(CHEBPOL_CUDA_EINVAL, "coef is NULL");
    int rc = check_dims(dims, rank, &nelem);
    if (rc) return rc;
    SplineTabs t = {sn, sx, sy, sb, sc, sd};
    if ((rc = check_spline_tabs(t, rank))) return rc;
    Identity id;
    id.blobs = {{coef, (size_t)nelem * 8}, {sn, (size_t)rank * 4}};
    {
        const double *const *tabs[5] = {sx, sy, sb, sc, sd};
        for (int k = 0; k < rank; ++k)
            for (int j = 0; j < 5; ++j) id.blobs.push_back({tabs[j][k], (size_t)sn[k] * 8});
    }
    make_identity(id, PLAN_CHEB, dims, rank, 7);
    return run_sharded(rank, x, npts, ngpus, out, id, [&](int dev, chebpol_cuda_plan **p) {
        return plan_cheb_impl(coef, dims, rank, nullptr, nullptr, nullptr, &t, dev, p);
    });
}

extern "C" int chebpol_cuda_evalmlip(const double *const *grid, const int *dims, int rank,
                                     const double *values, const double *x, int64_t npts, int blend,
                                     int ngpus, double *out)
{
    int64_t nelem = 0;
    if (!grid || !values) return fail(CHEBPOL_CUDA_EINVAL, "grid or values is NULL");
    if (blend < 0 || blend > 5) return fail(CHEBPOL_CUDA_EBLEND, "blend %d not in 0..5", blend);
    int rc = check_dims(dims, rank, &nelem);
    if (rc) return rc;
    Identity id;
    id.blobs = {{values, (size_t)nelem * 8}};
    for (int d = 0; d < rank; ++d) {
        if (!grid[d]) return fail(CHEBPOL_CUDA_EINVAL, "grid[%d] is NULL", d);
        id.blobs.push_back({grid[d], (size_t)dims[d] * 8});
    }
    make_identity(id, PLAN_MLIP, dims, rank, 2);
    return run_sharded(rank, x, npts, ngpus, out, id, [&](int dev, chebpol_cuda_plan **p) {
        return chebpol_cuda_plan_mlip(grid, dims, rank, values, blend, dev, p);
    }, blend);
}

extern "C" int chebpol_cuda_fh(const double *vals, const double *const *grid,
                               const double *const *weights, const int *dims, int rank, const double *x,
                               int64_t npts, int ngpus, double *out)
{
    int64_t nelem = 0;
    if (!vals || !grid || !weights) return fail(CHEBPOL_CUDA_EINVAL, "vals, grid or weights is NULL");
    int rc = check_dims(dims, rank, &nelem);
    if (rc) return rc;
    Identity id;
    id.blobs = {{vals, (size_t)nelem * 8}};
    for (int d = 0; d < rank; ++d) {
        if (!grid[d] || !weights[d]) return fail(CHEBPOL_CUDA_EINVAL, "grid/weights[%d] is NULL", d);
        id.blobs.push_back({grid[d], (size_t)dims[d] * 8});
        id.blobs.push_back({weights[d], (size_t)dims[d] * 8});
    }
    make_identity(id, PLAN_FH, dims, rank, 3);
    return run_sharded(rank, x, npts, ngpus, out, id, [&](int dev, chebpol_cuda_plan **p) {
        return chebpol_cuda_plan_fh(vals, grid, weights, dims, rank, dev, p);
    });
}

extern "C" int chebpol_cuda_evalpolyh(const double *knots, int rank, int nknots, const double *weights,
                                      const double *lweights, double k, const double *x, int64_t npts,
                                      const double *norm_a, const double *norm_b, int ngpus, double *out)
{
    if (rank <= 0 || rank > CHEBPOL_CUDA_MAXRANK) return fail(CHEBPOL_CUDA_ERANK, "rank %d not in 1..%d", rank, CHEBPOL_CUDA_MAXRANK);
    if (nknots < 0) return fail(CHEBPOL_CUDA_ESIZE, "negative number of knots");
    if (!lweights || (nknots > 0 && (!knots || !weights))) return fail(CHEBPOL_CUDA_EINVAL, "knots, weights or lweights is NULL");
    int dims[CP_MAXR];
    for (int i = 0; i < rank; ++i) dims[i] = nknots;
    if ((norm_a == nullptr) != (norm_b == nullptr)) return fail(CHEBPOL_CUDA_EINVAL, "norm_a and norm_b go together");
    Identity id;
    const int32_t have[2] = {norm_a ? 1 : 0, nknots};
    id.blobs = {{have, sizeof have}, {&k, sizeof k}, {lweights, (size_t)(rank + 1) * 8}};
    if (nknots > 0) {
        id.blobs.push_back({knots, (size_t)nknots * rank * 8});
        id.blobs.push_back({weights, (size_t)nknots * 8});
    }
    if (norm_a) {
        id.blobs.push_back({norm_a, (size_t)rank * 8});
        id.blobs.push_back({norm_b, (size_t)rank * 8});
    }
    make_identity(id, PLAN_POLYH, dims, rank, 4);
    return run_sharded(rank, x, npts, ngpus, out, id, [&](int dev, chebpol_cuda_plan **p) {
        return chebpol_cuda_plan_polyh(knots, rank, nknots, weights, lweights, k, norm_a, norm_b, dev, p);
    });
}

static int stalk_stateless(int kind, const double *val, const double *const *grid, const int *dims, int rank,
                           const double *a, const double *b, const double *t1, const double *t2, int blend,
                           const double *x, int64_t npts, int ngpus, double *out)
{
    int64_t nelem = 0;
    if (!val || !grid || !b || !t1 || !t2 || (kind == 1 && !a)) return fail(CHEBPOL_CUDA_EINVAL, "a table pointer is NULL");
    int rc = check_dims(dims, rank, &nelem);
    if (rc) return rc;
    if (blend < 0 || blend > 5) return fail(CHEBPOL_CUDA_EBLEND, "blend %d not in 0..5", blend);
    Identity id;
    id.blobs = {{val, (size_t)nelem * 8}};
    if (kind == 1) id.blobs.push_back({a, (size_t)nelem * 8});
    id.blobs.push_back({b, (size_t)nelem * rank * 8});
    id.blobs.push_back({t1, (size_t)nelem * rank * 8});
    id.blobs.push_back({t2, (size_t)nelem * rank * 8});
    for (int d = 0; d < rank; ++d) {
        if (!grid[d]) return fail(CHEBPOL_CUDA_EINVAL, "grid[%d] is NULL", d);
        id.blobs.push_back({grid[d], (size_t)dims[d] * 8});
    }
    make_identity(id, PLAN_STALK, dims, rank, 5 + kind);
    return run_sharded(rank, x, npts, ngpus, out, id, [&](int dev, chebpol_cuda_plan **p) {
        return chebpol_cuda_plan_stalk(kind, val, grid, dims, rank, a, b, t1, t2, blend, dev, p);
    }, blend);
}

extern "C" int chebpol_cuda_evalstalk(const double *val, const double *const *grid, const int *dims, int rank,
                                      const double *b, const double *c, const double *r, int blend,
                                      const double *x, int64_t npts, int ngpus, double *out)
{
    return stalk_stateless(0, val, grid, dims, rank, nullptr, b, c, r, blend, x, npts, ngpus, out);
}

extern "C" int chebpol_cuda_evalhyp(const double *val, const double *const *grid, const int *dims, int rank,
                                    const double *a, const double *b, const double *c, const double *d, int blend,
                                    const double *x, int64_t npts, int ngpus, double *out)
{
    return stalk_stateless(1, val, grid, dims, rank, a, b, c, d, blend, x, npts, ngpus, out);
}

// ------------------------------------------------------------------ fit side
extern "C" int chebpol_cuda_evalsl(const double *x, int64_t npts, const double *knots, int dim, int nknots,
                                   const int *dtri, int numsimplex, const double *lumats, const int *pivots,
                                   const double *bbox, const double *val, int epol, int blend, int ngpus, double *out)
{
    int rc = check_sl(knots, dim, nknots, dtri, numsimplex, lumats, pivots, bbox, val);
    if (rc) return rc;
    if (blend < 0 || blend > 5) return fail(CHEBPOL_CUDA_EBLEND, "blend %d not in 0..5", blend);
    int dims[CP_MAXR];
    for (int i = 0; i < dim; ++i) dims[i] = nknots > 0 ? nknots : 1;
    const int N = dim + 1;
    Identity id;
    const int32_t counts[2] = {nknots, numsimplex};
    id.blobs = {{counts, sizeof counts},
                {knots, (size_t)nknots * dim * 8},
                {val, (size_t)nknots * 8},
                {dtri, (size_t)numsimplex * N * 4},
                {lumats, (size_t)numsimplex * N * N * 8},
                {pivots, (size_t)numsimplex * N * 4},
                {bbox, (size_t)numsimplex * 2 * dim * 8}};
    make_identity(id, PLAN_SL, dims, dim, 11);
    // blend and epol are per-call arguments of the closure (R/simplexlinear.R:34-38), not part of the key
    return run_sharded(dim, x, npts, ngpus, out, id, [&](int dev, chebpol_cuda_plan **p) {
        return chebpol_cuda_plan_sl(knots, dim, nknots, dtri, numsimplex, lumats, pivots, bbox, val, dev, p);
    }, blend, epol);
}

extern "C" int chebpol_cuda_evalstalker(const double *val, const double *const *grid, const int *dims, int rank,
                                        const double *degree, const double *const *det, const double *const *pmin,
                                        const double *const *pplus, const int *uniform, int blend, const double *x,
                                        int64_t npts, int ngpus, double *out)
{
    int64_t nelem = 0;
    if (!val || !grid || !degree || !det || !pmin || !pplus || !uniform) return fail(CHEBPOL_CUDA_EINVAL, "a table pointer is NULL");
    int rc = check_dims(dims, rank, &nelem);
    if (rc) return rc;
    if (blend < 0 || blend > 5) return fail(CHEBPOL_CUDA_EBLEND, "blend %d not in 0..5", blend);
    Identity id;
    id.blobs = {{val, (size_t)nelem * 8}, {degree, (size_t)rank * 8}, {uniform, (size_t)rank * 4}};
    for (int d = 0; d < rank; ++d) {
        if (!grid[d] || !det[d] || !pmin[d] || !pplus[d]) return fail(CHEBPOL_CUDA_EINVAL, "vector of dimension %d is NULL", d);
        id.blobs.push_back({grid[d], (size_t)dims[d] * 8});
        id.blobs.push_back({det[d], (size_t)dims[d] * 8});
        id.blobs.push_back({pmin[d], (size_t)dims[d] * 8});
        id.blobs.push_back({pplus[d], (size_t)dims[d] * 8});
    }
    make_identity(id, PLAN_OLDSTALK, dims, rank, 13);
    return run_sharded(rank, x, npts, ngpus, out, id, [&](int dev, chebpol_cuda_plan **p) {
        return chebpol_cuda_plan_oldstalk(val, grid, dims, rank, degree, det, pmin, pplus, uniform, blend, dev, p);
    }, blend);
}

extern "C" int chebpol_cuda_chebcoef(const double *val, const int *dims, int rank, int dct, int device, double *coef)
{
    if (!val || !coef) return fail(CHEBPOL_CUDA_EINVAL, "val or coef is NULL");
    int64_t nelem = 0;
    int rc = check_dims(dims, rank, &nelem);
    if (rc) return rc;
    int sms = 0;
    rc = select_device(device, &sms);
    if (rc) return rc;
    double *d_a = nullptr, *d_b = nullptr, *res = nullptr;
    int launches = 0;
    cudaError_t e = cudaMalloc(&d_a, sizeof(double) * nelem);
    if (e == cudaSuccess) e = cudaMalloc(&d_b, sizeof(double) * nelem);
    if (e == cudaSuccess) e = cudaMemcpy(d_a, val, sizeof(double) * nelem, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = fit_chebcoef_device(d_a, d_b, dims, rank, dct, nullptr, &res, &launches);
    if (e == cudaSuccess) e = cudaMemcpy(coef, res, sizeof(double) * nelem, cudaMemcpyDeviceToHost);
    cudaFree(d_a);
    cudaFree(d_b);
    g_launches.fetch_add(launches);
    if (e != cudaSuccess) return cuda_fail(e, "chebcoef");
    return CHEBPOL_CUDA_OK;
}

extern "C" int chebpol_cuda_fhweights(const double *grid, int nknots, int d, int device, double *w)
{
    if (!grid || !w) return fail(CHEBPOL_CUDA_EINVAL, "grid or w is NULL");
    if (nknots < 1 || d < 0 || d > nknots - 1) return fail(CHEBPOL_CUDA_ESIZE, "need 0 <= d <= nknots-1 (d=%d, nknots=%d)", d, nknots);
    int sms = 0;
    int rc = select_device(device, &sms);
    if (rc) return rc;
    double *d_g = nullptr, *d_w = nullptr;
    cudaError_t e = cudaMalloc(&d_g, sizeof(double) * nknots);
    if (e == cudaSuccess) e = cudaMalloc(&d_w, sizeof(double) * nknots);
    if (e == cudaSuccess) e = cudaMemcpy(d_g, grid, sizeof(double) * nknots, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = fit_fhweights_device(d_g, nknots, d, d_w, nullptr);
    if (e == cudaSuccess) e = cudaMemcpy(w, d_w, sizeof(double) * nknots, cudaMemcpyDeviceToHost);
    cudaFree(d_g);
    cudaFree(d_w);
    g_launches.fetch_add(1);
    if (e != cudaSuccess) return cuda_fail(e, "fhweights");
    return CHEBPOL_CUDA_OK;
}

// --------------------------------------------------- evaluation on a tensor grid
// kind 0 Chebyshev (pre-map optional), 1 Floater-Hormann, 2 multilinear
static int evalgrid_common(int kind, const double *tensor, const int *dims, int rank, const double *const *pts,
                           const int *gdims, const double *const *knots, const double *const *weights, int blend,
                           const PreMap *pm, int device, double *out)
{
    if (!tensor || !pts || !gdims || !out) return fail(CHEBPOL_CUDA_EINVAL, "NULL argument");
    int64_t nelem = 0;
    int rc = check_dims(dims, rank, &nelem);
    if (rc) return rc;
    int64_t nout = 1, maxbuf = 1;
    {
        // size of the intermediate after each mode product
        int64_t done = 1, rest = nelem;
        for (int d = 0; d < rank; ++d) {
            if (gdims[d] <= 0 || !pts[d]) return fail(CHEBPOL_CUDA_ESIZE, "grid %d is empty", d);
            if (kind == 2 && dims[d] < 2) return fail(CHEBPOL_CUDA_ESIZE, "multilinear grid %d needs at least 2 knots", d);
            rest /= dims[d];
            done *= gdims[d];
            if (done * rest > maxbuf) maxbuf = done * rest;
            if (done * rest >= (int64_t(1) << 34)) return fail(CHEBPOL_CUDA_ESIZE, "grid evaluation too large");
        }
        nout = done;
    }
    int sms = 0;
    rc = select_device(device, &sms);
    if (rc) return rc;
    std::vector<double *> dev;  // everything to free
    auto up = [&](const double *h, size_t n, double **d) -> cudaError_t {
        cudaError_t e = cudaMalloc(d, sizeof(double) * n);
        if (e != cudaSuccess) return e;
        dev.push_back(*d);
        return h ? cudaMemcpy(*d, h, sizeof(double) * n, cudaMemcpyHostToDevice) : cudaSuccess;
    };
    cudaError_t e = cudaSuccess;
    double *d_t = nullptr, *d_s0 = nullptr, *d_s1 = nullptr, *res = nullptr;
    std::vector<double *> d_pts(rank, nullptr), d_kn(rank, nullptr), d_w(rank, nullptr);
    e = up(tensor, (size_t)nelem, &d_t);
    if (e == cudaSuccess) e = up(nullptr, (size_t)maxbuf, &d_s0);
    if (e == cudaSuccess) e = up(nullptr, (size_t)maxbuf, &d_s1);
    for (int d = 0; d < rank && e == cudaSuccess; ++d) {
        e = up(pts[d], (size_t)gdims[d], &d_pts[d]);
        if (e == cudaSuccess && knots) e = up(knots[d], (size_t)dims[d], &d_kn[d]);
        if (e == cudaSuccess && weights) e = up(weights[d], (size_t)dims[d], &d_w[d]);
    }
    int launches = 0;
    if (e == cudaSuccess)
        e = fit_evalgrid_device(kind, d_t, dims, rank, d_pts.data(), gdims, knots ? d_kn.data() : nullptr,
                                weights ? d_w.data() : nullptr, blend, pm, d_s0, d_s1, nullptr, &res, &launches);
    if (e == cudaSuccess) e = cudaMemcpy(out, res, sizeof(double) * nout, cudaMemcpyDeviceToHost);
    for (double *d : dev) cudaFree(d);
    g_launches.fetch_add(launches);
    if (e != cudaSuccess) return cuda_fail(e, "evalgrid");
    return CHEBPOL_CUDA_OK;
}

extern "C" int chebpol_cuda_evalgrid_cheb(const double *coef, const int *dims, int rank, const double *const *pts,
                                          const int *gdims, const double *mid, const double *ispan, const int *ugm_n,
                                          int device, double *out)
{
    if ((mid == nullptr) != (ispan == nullptr)) return fail(CHEBPOL_CUDA_EINVAL, "mid and ispan must be given together");
    if (rank <= 0 || rank > CHEBPOL_CUDA_MAXRANK) return fail(CHEBPOL_CUDA_ERANK, "rank must be positive");
    PreMap pm;
    memset(&pm, 0, sizeof pm);
    for (int d = 0; d < rank; ++d) {
        pm.ispan[d] = 1.0;
        pm.nd[d] = 1.0;
        if (mid) { pm.mode |= 1; pm.mid[d] = mid[d]; pm.ispan[d] = ispan[d]; }
        if (ugm_n) { pm.mode |= 2; pm.nd[d] = (double)ugm_n[d]; }
    }
    return evalgrid_common(0, coef, dims, rank, pts, gdims, nullptr, nullptr, 0, &pm, device, out);
}

extern "C" int chebpol_cuda_evalgrid_mlip(const double *const *grid, const int *dims, int rank, const double *values,
                                          const double *const *pts, const int *gdims, int blend, int device, double *out)
{
    if (!grid) return fail(CHEBPOL_CUDA_EINVAL, "grid is NULL");
    if (blend < 0 || blend > 5) return fail(CHEBPOL_CUDA_EBLEND, "blend %d not in 0..5", blend);
    return evalgrid_common(2, values, dims, rank, pts, gdims, grid, nullptr, blend, nullptr, device, out);
}

extern "C" int chebpol_cuda_evalgrid_fh(const double *vals, const double *const *grid, const double *const *weights,
                                        const int *dims, int rank, const double *const *pts, const int *gdims, int device,
                                        double *out)
{
    if (!grid || !weights) return fail(CHEBPOL_CUDA_EINVAL, "grid or weights is NULL");
    return evalgrid_common(1, vals, dims, rank, pts, gdims, grid, weights, 0, nullptr, device, out);
}

extern "C" int chebpol_cuda_synth_points(uint64_t seed, int64_t first_point, int64_t npts, int rank, double lo, double hi,
                                         double *d_x, int device, void *stream)
{
    if (npts < 0 || rank <= 0) return fail(CHEBPOL_CUDA_EINVAL, "negative point count or non-positive rank");
    if (npts > 0 && !d_x) return fail(CHEBPOL_CUDA_EINVAL, "d_x is NULL");
    int cur = -1, sms = 0;
    CU(cudaGetDevice(&cur));
    if (cur != device) CU(cudaSetDevice(device));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    cudaError_t e = synth_fill_points(seed, first_point, npts, rank, lo, hi, d_x, sms, (cudaStream_t)stream);
    if (cur != device) cudaSetDevice(cur);
    if (e != cudaSuccess) return cuda_fail(e, "synthetic point generator");
    g_launches.fetch_add(1);
    return CHEBPOL_CUDA_OK;
}

extern "C" int chebpol_cuda_fp64_peak(int device, int repeats, double *tflops)
{
    if (!tflops) return fail(CHEBPOL_CUDA_EINVAL, "tflops is NULL");
    int sms = 0;
    int rc = select_device(device, &sms);
    if (rc) return rc;
    cudaError_t e = run_fp64_peak(sms, repeats, nullptr, tflops);
    if (e != cudaSuccess) return cuda_fail(e, "fp64 peak kernel");
    g_launches.fetch_add(1 + (repeats < 1 ? 1 : repeats));
    return CHEBPOL_CUDA_OK;
}
