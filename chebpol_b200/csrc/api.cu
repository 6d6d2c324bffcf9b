// api.cu -- the C ABI of libchebpol_cuda (include/chebpol_cuda.h): argument
// validation mirroring the reference's entry points, plan management, kernel
// selection, the chunked host<->device pipeline and the multi-GPU point
// partitioner.  No CPU evaluation path exists in this library.
#include "common.cuh"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

int cheb_slab_pick_n0p(int n0);

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
enum PlanKind { PLAN_CHEB = 1, PLAN_MLIP = 2, PLAN_FH = 3 };

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
    int blend = 0;
    // options
    int kernel_opt = CHEBPOL_CUDA_KERNEL_AUTO;
    int variant = 0;
    int64_t chunk_points = 0;
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
    int64_t ncells = 0;
    bool cells_failed = false;
    // statistics
    LaunchInfo last = {0, 0, 0, 0};
    int64_t launches = 0;
    int64_t tensor_bytes = 0;
    // host pipeline resources, created on first eval_host
    cudaStream_t st_in = nullptr, st_k = nullptr, st_out = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_k[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    double *d_x[2] = {nullptr, nullptr}, *d_o[2] = {nullptr, nullptr};
    int64_t buf_points = 0;
};

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
    cudaFree(p->d_tensor);
    cudaFree(p->d_padded);
    cudaFree(p->d_mma);
    cudaFree(p->d_ktab);
    cudaFree(p->d_cells);
    cudaFree(p->d_lut);
    cudaFree(p->d_grid);
    cudaFree(p->d_weights);
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
    const size_t rows = (size_t)m.nrest, rs = (size_t)m.rs, K = (size_t)m.K;
    std::vector<double> padded(rows * rs, 0.0);
    for (size_t r = 0; r < rows; ++r) {
        memcpy(&padded[r * rs], host_tensor + r * K, sizeof(double) * K);
        // basis-table rows (boff[d] + digit_d) of the rest-level digits of r, as int32
        int *info = reinterpret_cast<int *>(&padded[r * rs + m.kp]);
        size_t rem = r;
        for (int l = 0; l < m.nlev; ++l) {
            const int d = m.nfold + l;
            info[l] = m.boff[d] + (int)(rem % (size_t)m.dims[d]);
            rem /= (size_t)m.dims[d];
        }
    }
    std::vector<int> ktab((size_t)m.kp * 3, -1);
    for (int k = 0; k < m.K; ++k) {
        int rem = k;
        for (int l = 0; l < m.nfold; ++l) {
            ktab[(size_t)k * 3 + l] = m.boff[l] + rem % m.dims[l];
            rem /= m.dims[l];
        }
    }
    cudaError_t ek = cudaMalloc(&p->d_ktab, sizeof(int) * ktab.size());
    if (ek != cudaSuccess) return cuda_fail(ek, "cudaMalloc(ktab)");
    ek = cudaMemcpy(p->d_ktab, ktab.data(), sizeof(int) * ktab.size(), cudaMemcpyHostToDevice);
    if (ek != cudaSuccess) return cuda_fail(ek, "cudaMemcpy(ktab)");
    m.ktab = p->d_ktab;
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

extern "C" int chebpol_cuda_plan_cheb(const double *coef, const int *dims, int rank, const double *mid,
                                      const double *ispan, const int *ugm_n, int device,
                                      chebpol_cuda_plan **plan)
{
    if (!plan) return fail(CHEBPOL_CUDA_EINVAL, "plan is NULL");
    *plan = nullptr;
    if (!coef) return fail(CHEBPOL_CUDA_EINVAL, "coef is NULL");
    if ((mid == nullptr) != (ispan == nullptr)) return fail(CHEBPOL_CUDA_EINVAL, "mid and ispan must be given together");
    int64_t nelem = 0;
    int rc = check_dims(dims, rank, &nelem);
    if (rc) return rc;
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
    cudaError_t e = cudaMalloc(&p->d_tensor, sizeof(double) * nelem);
    if (e != cudaSuccess) return bail(cuda_fail(e, "cudaMalloc(tensor)"));
    e = cudaMemcpy(p->d_tensor, coef, sizeof(double) * nelem, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return bail(cuda_fail(e, "cudaMemcpy(tensor)"));
    p->tensor_bytes = sizeof(double) * nelem;
    configure_slab(p);
    if (p->slab_ok) {
        const int n0 = dims[0], n0p = p->slab.n0p;
        const int64_t rows = nelem / n0;
        const double *src = coef;
        std::vector<double> padded;
        if (n0p != n0) {
            padded.assign((size_t)rows * n0p, 0.0);
            for (int64_t r = 0; r < rows; ++r) memcpy(&padded[(size_t)r * n0p], coef + r * n0, sizeof(double) * n0);
            src = padded.data();
        }
        e = cudaMalloc(&p->d_padded, sizeof(double) * rows * n0p);
        if (e != cudaSuccess) return bail(cuda_fail(e, "cudaMalloc(padded tensor)"));
        e = cudaMemcpy(p->d_padded, src, sizeof(double) * rows * n0p, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return bail(cuda_fail(e, "cudaMemcpy(padded tensor)"));
        p->slab.cf = p->d_padded;
        p->tensor_bytes += sizeof(double) * rows * n0p;
    }
    rc = configure_mma(p, 0, coef);
    if (rc) return bail(rc);
    *plan = p;
    return CHEBPOL_CUDA_OK;
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
    // enough to amortise it; capped by CHEBPOL_CUDA_CELL_MAX_GB (default 16) and free memory
    p->cell_bytes = mlip_cell_bytes(rank, dims);
    p->ncells = 1;
    for (int d = 0; d < rank; ++d) p->ncells *= dims[d] - 1;
    {
        double cap_gb = 16.0;
        if (const char *env = getenv("CHEBPOL_CUDA_CELL_MAX_GB")) cap_gb = atof(env);
        if ((double)p->cell_bytes > cap_gb * 1e9) p->cell_bytes = 0;
    }
    *plan = p;
    return CHEBPOL_CUDA_OK;
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
        if (p->kind != PLAN_MLIP) return fail(CHEBPOL_CUDA_EINVAL, "blend applies to multilinear plans");
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
    default: return fail(CHEBPOL_CUDA_EINVAL, "unknown query %d", what);
    }
    return CHEBPOL_CUDA_OK;
}

// ------------------------------------------------------------------- launch
// The current device must be p->device.
static int launch_plan(chebpol_cuda_plan *p, const double *d_x, int64_t npts, double *d_out, cudaStream_t s)
{
    if (npts == 0) return CHEBPOL_CUDA_OK;
    cudaError_t e = cudaErrorNotSupported;
    LaunchInfo li = {0, 0, 0, 0};
    const bool want_fast = p->kernel_opt != CHEBPOL_CUDA_KERNEL_STRICT;
    const bool aligned16 = (((uintptr_t)d_x) & 15) == 0;
    // variant: 0 auto (DMMA engine, else DFMA slab kernel, else strict); 1..99999 tuning
    // code of the DFMA slab kernel; >= 100000 DMMA engine with tuning code variant-100000
    const bool force_slab = p->variant > 0 && p->variant < 100000;
    const int mma_variant = p->variant >= 100000 ? p->variant - 100000 : 0;
    if (p->kind == PLAN_CHEB) {
        if (want_fast && p->mma_ok && !force_slab) {
            MmaArgs a = p->mma;
            a.x = d_x;
            a.npts = npts;
            a.out = d_out;
            e = launch_contract_mma(a, mma_variant, p->num_sms, s, &li);
        }
        if (e == cudaErrorNotSupported && want_fast && p->slab_ok && p->variant < 100000) {
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
        // variant: 0 auto, 1 pair-gather kernel, 2 cell-major kernel (always expand),
        // >= 20 cell-major kernel with layout code warps*10 + stages
        const bool allow_cell = p->variant == 0 || p->variant >= 2;
        if (want_fast && allow_cell && p->cell_bytes > 0 && !p->d_cells && !p->cells_failed &&
            (p->variant >= 2 || npts >= p->ncells / 8)) {
            size_t free_b = 0, total_b = 0;
            cudaError_t ce = cudaMemGetInfo(&free_b, &total_b);
            if (ce == cudaSuccess && (size_t)p->cell_bytes < free_b / 2) ce = cudaMalloc(&p->d_cells, (size_t)p->cell_bytes);
            else if (ce == cudaSuccess) ce = cudaErrorMemoryAllocation;
            if (ce == cudaSuccess) ce = mlip_expand(a, p->d_tensor, p->d_cells, s);
            if (ce != cudaSuccess) {
                (void)cudaGetLastError();
                cudaFree(p->d_cells);
                p->d_cells = nullptr;
                p->cells_failed = true;  // fall back to the pair kernel for the plan's lifetime
            } else {
                p->tensor_bytes += p->cell_bytes;
                p->launches += 1;
                g_launches.fetch_add(1);
            }
        }
        if (want_fast && allow_cell && p->d_cells) e = launch_mlip_cell(a, p->d_cells, p->variant, p->num_sms, s, &li);
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
    int cur = -1;
    CU(cudaGetDevice(&cur));
    if (cur != p->device) CU(cudaSetDevice(p->device));

    int64_t chunk = p->chunk_points;
    if (chunk <= 0) {
        chunk = (int64_t(128) << 20) / (8 * p->rank);
        if (chunk < (1 << 20)) chunk = 1 << 20;
    }
    if (chunk > npts) chunk = npts;
    int rc = ensure_pipeline(p, chunk);
    if (rc) return rc;
    const int64_t nchunks = (npts + chunk - 1) / chunk;
    // chunk c uses buffer c&1:  H2D(c) -> kernel(c) -> D2H(c); D2H(c-1) is
    // enqueued after kernel(c) so that a blocking (pageable) copy-out overlaps
    // the next kernel.
    auto copy_out = [&](int64_t c) -> int {
        const int b = (int)(c & 1);
        const int64_t off = c * chunk, n = (off + chunk <= npts) ? chunk : npts - off;
        CU(cudaStreamWaitEvent(p->st_out, p->ev_k[b], 0));
        CU(cudaMemcpyAsync(out + off, p->d_o[b], sizeof(double) * n, cudaMemcpyDeviceToHost, p->st_out));
        CU(cudaEventRecord(p->ev_out[b], p->st_out));
        return CHEBPOL_CUDA_OK;
    };
    for (int64_t c = 0; c < nchunks; ++c) {
        const int b = (int)(c & 1);
        const int64_t off = c * chunk, n = (off + chunk <= npts) ? chunk : npts - off;
        if (c >= 2) CU(cudaStreamWaitEvent(p->st_in, p->ev_k[b], 0));  // x buffer free
        CU(cudaMemcpyAsync(p->d_x[b], x + off * p->rank, sizeof(double) * n * p->rank, cudaMemcpyHostToDevice, p->st_in));
        CU(cudaEventRecord(p->ev_in[b], p->st_in));
        CU(cudaStreamWaitEvent(p->st_k, p->ev_in[b], 0));
        if (c >= 2) CU(cudaStreamWaitEvent(p->st_k, p->ev_out[b], 0));  // out buffer free
        rc = launch_plan(p, p->d_x[b], n, p->d_o[b], p->st_k);
        if (rc) break;
        CU(cudaEventRecord(p->ev_k[b], p->st_k));
        if (c >= 1) {
            rc = copy_out(c - 1);
            if (rc) break;
        }
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

// -------------------------------------------------- stateless entry points
// Split [0, npts) into contiguous ranges over `ngpus` devices; one host thread
// per device builds a plan there (replicating the tensor) and runs the host
// pipeline on its slice.  No collective: every GPU writes its own slice.
template <class MakePlan>
static int run_sharded(int rank, const double *x, int64_t npts, int ngpus, double *out, MakePlan make_plan)
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

    if (ngpus == 1) {
        chebpol_cuda_plan *p = nullptr;
        rc = make_plan(0, &p);
        if (rc) return rc;
        rc = chebpol_cuda_plan_eval_host(p, x, npts, out);
        chebpol_cuda_plan_destroy(p);
        return rc;
    }
    std::vector<int> rcs(ngpus, 0);
    std::vector<std::string> msgs(ngpus);
    std::vector<std::thread> th;
    const int64_t per = (npts + ngpus - 1) / ngpus;
    for (int g = 0; g < ngpus; ++g) {
        th.emplace_back([&, g]() {
            const int64_t lo = (int64_t)g * per;
            const int64_t n = (lo >= npts) ? 0 : ((lo + per <= npts) ? per : npts - lo);
            chebpol_cuda_plan *p = nullptr;
            int r = make_plan(g, &p);
            if (!r && n > 0) r = chebpol_cuda_plan_eval_host(p, x + lo * rank, n, out + lo);
            chebpol_cuda_plan_destroy(p);
            rcs[g] = r;
            if (r) msgs[g] = g_last_error;
        });
    }
    for (auto &t : th) t.join();
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
    return run_sharded(rank, x, npts, ngpus, out, [&](int dev, chebpol_cuda_plan **p) {
        return chebpol_cuda_plan_cheb(coef, dims, rank, mid, ispan, ugm_n, dev, p);
    });
}

extern "C" int chebpol_cuda_evalmlip(const double *const *grid, const int *dims, int rank,
                                     const double *values, const double *x, int64_t npts, int blend,
                                     int ngpus, double *out)
{
    return run_sharded(rank, x, npts, ngpus, out, [&](int dev, chebpol_cuda_plan **p) {
        return chebpol_cuda_plan_mlip(grid, dims, rank, values, blend, dev, p);
    });
}

extern "C" int chebpol_cuda_fh(const double *vals, const double *const *grid,
                               const double *const *weights, const int *dims, int rank, const double *x,
                               int64_t npts, int ngpus, double *out)
{
    return run_sharded(rank, x, npts, ngpus, out, [&](int dev, chebpol_cuda_plan **p) {
        return chebpol_cuda_plan_fh(vals, grid, weights, dims, rank, dev, p);
    });
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
