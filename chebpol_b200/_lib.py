"""ctypes binding of libchebpol_cuda (include/chebpol_cuda.h).

This is the only way the package computes anything: there is no CPU
evaluation path.  If the shared library is missing, or no B200 is visible,
the calls raise -- they never fall back.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CHEBPOL_CUDA_LIB") or os.path.join(_HERE, "libchebpol_cuda.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_dpp = C.POINTER(_dp)
_vp = C.c_void_p

# option / query keys (include/chebpol_cuda.h)
OPT_KERNEL, OPT_BLEND, OPT_CHUNK_POINTS, OPT_VARIANT = 1, 2, 3, 4
OPT_EPOL, OPT_SL_CELLS = 5, 6
GET_KERNEL_USED, GET_LAUNCHES, GET_TENSOR_BYTES, GET_DEVICE, GET_SMEM_BYTES, GET_GRID = 101, 102, 103, 104, 105, 106
GET_RANK, GET_MMA_MACS = 107, 108
KERNEL_AUTO, KERNEL_STRICT, KERNEL_FAST = 0, 1, 2
VARIANT_AUTO, VARIANT_SLAB, VARIANT_MMA = 0, 1, 100000  # OPT_VARIANT families (csrc/api.cu: launch_plan)
VARIANT_MLIP_PAIR, VARIANT_MLIP_CELL, VARIANT_MLIP_QUAD, VARIANT_MLIP_BIN = 1, 2, 3, 4
KERNEL_NAMES = {
    0: "none", 10: "cheb_strict", 11: "cheb_slab", 12: "cheb_mma", 20: "mlip_strict", 21: "mlip_pair", 22: "mlip_cell", 23: "mlip_quad", 24: "mlip_bin",
    30: "fh_strict", 32: "fh_mma", 40: "polyh_strict", 41: "polyh_tile", 50: "stalk_strict", 51: "stalk_cell",
    52: "oldstalk", 60: "sl_strict", 61: "sl_cell",
}

EXPORTS = [
    "chebpol_cuda_strerror", "chebpol_cuda_last_error", "chebpol_cuda_device_count",
    "chebpol_cuda_version", "chebpol_cuda_evalcheb", "chebpol_cuda_evalmlip", "chebpol_cuda_fh",
    "chebpol_cuda_plan_cheb", "chebpol_cuda_plan_mlip", "chebpol_cuda_plan_fh",
    "chebpol_cuda_plan_destroy", "chebpol_cuda_plan_eval_device", "chebpol_cuda_plan_eval_host",
    "chebpol_cuda_plan_set", "chebpol_cuda_plan_get", "chebpol_cuda_fp64_peak",
    "chebpol_cuda_launch_count", "chebpol_cuda_cache_clear", "chebpol_cuda_debug_mma_geometry", "chebpol_cuda_debug_mma_row_pos", "chebpol_cuda_chebcoef", "chebpol_cuda_fhweights",
    "chebpol_cuda_evalgrid_cheb", "chebpol_cuda_evalgrid_mlip", "chebpol_cuda_evalgrid_fh",
    "chebpol_cuda_evalpolyh", "chebpol_cuda_plan_polyh",
    "chebpol_cuda_evalstalk", "chebpol_cuda_evalhyp", "chebpol_cuda_plan_stalk",
    "chebpol_cuda_evalchebg", "chebpol_cuda_plan_chebg",
    "chebpol_cuda_evalstalker", "chebpol_cuda_plan_oldstalk",
    "chebpol_cuda_evalsl", "chebpol_cuda_plan_sl", "chebpol_cuda_analyzesimplex", "chebpol_cuda_debug_sl_lists",
    "chebpol_cuda_synth_points", "chebpol_cuda_debug_cache_stats", "chebpol_cuda_debug_cache_weak_hash",
    "chebpol_cuda_debug_fingerprint", "chebpol_cuda_debug_cache_would_hit",
]


class ChebpolCudaError(RuntimeError):
    def __init__(self, code: int, what: str, detail: str):
        super().__init__(f"libchebpol_cuda: {what} (code {code}){': ' + detail if detail else ''}")
        self.code = code


_lib = None


def lib() -> C.CDLL:
    """Load libchebpol_cuda.so; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C chebpol_b200/csrc`). chebpol_b200 has no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    L.chebpol_cuda_strerror.restype = C.c_char_p
    L.chebpol_cuda_strerror.argtypes = [C.c_int]
    L.chebpol_cuda_last_error.restype = C.c_char_p
    L.chebpol_cuda_last_error.argtypes = []
    L.chebpol_cuda_device_count.argtypes = [_ip]
    L.chebpol_cuda_version.argtypes = []
    L.chebpol_cuda_evalcheb.argtypes = [_vp, _ip, C.c_int, _vp, C.c_int64, _vp, _vp, _vp, C.c_int, _vp]
    L.chebpol_cuda_evalmlip.argtypes = [_dpp, _ip, C.c_int, _vp, _vp, C.c_int64, C.c_int, C.c_int, _vp]
    L.chebpol_cuda_fh.argtypes = [_vp, _dpp, _dpp, _ip, C.c_int, _vp, C.c_int64, C.c_int, _vp]
    L.chebpol_cuda_plan_cheb.argtypes = [_vp, _ip, C.c_int, _vp, _vp, _vp, C.c_int, C.POINTER(_vp)]
    L.chebpol_cuda_plan_mlip.argtypes = [_dpp, _ip, C.c_int, _vp, C.c_int, C.c_int, C.POINTER(_vp)]
    L.chebpol_cuda_plan_fh.argtypes = [_vp, _dpp, _dpp, _ip, C.c_int, C.c_int, C.POINTER(_vp)]
    L.chebpol_cuda_plan_destroy.restype = None
    L.chebpol_cuda_plan_destroy.argtypes = [_vp]
    L.chebpol_cuda_plan_eval_device.argtypes = [_vp, _vp, C.c_int64, _vp, _vp]
    L.chebpol_cuda_plan_eval_host.argtypes = [_vp, _vp, C.c_int64, _vp]
    L.chebpol_cuda_plan_set.argtypes = [_vp, C.c_int, C.c_int64]
    L.chebpol_cuda_plan_get.argtypes = [_vp, C.c_int, C.POINTER(C.c_int64)]
    L.chebpol_cuda_fp64_peak.argtypes = [C.c_int, C.c_int, _dp]
    L.chebpol_cuda_chebcoef.argtypes = [_vp, _ip, C.c_int, C.c_int, C.c_int, _vp]
    L.chebpol_cuda_fhweights.argtypes = [_vp, C.c_int, C.c_int, C.c_int, _vp]
    L.chebpol_cuda_evalgrid_cheb.argtypes = [_vp, _ip, C.c_int, _dpp, _ip, _vp, _vp, _vp, C.c_int, _vp]
    L.chebpol_cuda_evalgrid_mlip.argtypes = [_dpp, _ip, C.c_int, _vp, _dpp, _ip, C.c_int, C.c_int, _vp]
    L.chebpol_cuda_evalgrid_fh.argtypes = [_vp, _dpp, _dpp, _ip, C.c_int, _dpp, _ip, C.c_int, _vp]
    L.chebpol_cuda_evalpolyh.argtypes = [_vp, C.c_int, C.c_int, _vp, _vp, C.c_double, _vp, C.c_int64, _vp, _vp,
                                         C.c_int, _vp]
    L.chebpol_cuda_plan_polyh.argtypes = [_vp, C.c_int, C.c_int, _vp, _vp, C.c_double, _vp, _vp, C.c_int,
                                          C.POINTER(_vp)]
    L.chebpol_cuda_evalstalk.argtypes = [_vp, _dpp, _ip, C.c_int, _vp, _vp, _vp, C.c_int, _vp, C.c_int64, C.c_int, _vp]
    L.chebpol_cuda_evalhyp.argtypes = [_vp, _dpp, _ip, C.c_int, _vp, _vp, _vp, _vp, C.c_int, _vp, C.c_int64, C.c_int,
                                       _vp]
    L.chebpol_cuda_plan_stalk.argtypes = [C.c_int, _vp, _dpp, _ip, C.c_int, _vp, _vp, _vp, _vp, C.c_int, C.c_int,
                                          C.POINTER(_vp)]
    L.chebpol_cuda_evalchebg.argtypes = [_vp, _ip, C.c_int, _vp, C.c_int64, _ip, _dpp, _dpp, _dpp, _dpp, _dpp, C.c_int, _vp]
    L.chebpol_cuda_plan_chebg.argtypes = [_vp, _ip, C.c_int, _ip, _dpp, _dpp, _dpp, _dpp, _dpp, C.c_int, C.POINTER(_vp)]
    L.chebpol_cuda_evalsl.argtypes = [_vp, C.c_int64, _vp, C.c_int, C.c_int, _vp, C.c_int, _vp, _vp, _vp, _vp, C.c_int,
                                      C.c_int, C.c_int, _vp]
    L.chebpol_cuda_plan_sl.argtypes = [_vp, C.c_int, C.c_int, _vp, C.c_int, _vp, _vp, _vp, _vp, C.c_int, C.POINTER(_vp)]
    L.chebpol_cuda_analyzesimplex.argtypes = [_vp, C.c_int, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]
    L.chebpol_cuda_debug_sl_lists.argtypes = [_vp, C.c_int, C.c_int, _vp, C.c_int, _vp, C.c_int, _ip, _vp, _vp, _vp,
                                              C.c_int64, _vp, C.c_int64, C.POINTER(C.c_int64)]
    L.chebpol_cuda_evalstalker.argtypes = [_vp, _dpp, _ip, C.c_int, _vp, _dpp, _dpp, _dpp, _ip, C.c_int, _vp, C.c_int64,
                                           C.c_int, _vp]
    L.chebpol_cuda_plan_oldstalk.argtypes = [_vp, _dpp, _ip, C.c_int, _vp, _dpp, _dpp, _dpp, _ip, C.c_int, C.c_int,
                                             C.POINTER(_vp)]
    L.chebpol_cuda_cache_clear.restype = None
    L.chebpol_cuda_cache_clear.argtypes = []
    L.chebpol_cuda_synth_points.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_int, C.c_double, C.c_double, _vp,
                                            C.c_int, _vp]
    _i64p = C.POINTER(C.c_int64)
    L.chebpol_cuda_debug_cache_stats.argtypes = [_i64p, _i64p, _i64p, _i64p, _i64p]
    L.chebpol_cuda_debug_cache_weak_hash.restype = None
    L.chebpol_cuda_debug_cache_weak_hash.argtypes = [C.c_int]
    L.chebpol_cuda_debug_fingerprint.argtypes = [_vp, C.c_int64, C.c_int, C.POINTER(C.c_uint64)]
    L.chebpol_cuda_debug_cache_would_hit.argtypes = [_vp, C.c_int64, _vp, C.c_int64, C.c_int, _ip]
    L.chebpol_cuda_launch_count.restype = C.c_int64
    L.chebpol_cuda_launch_count.argtypes = []
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        L = lib()
        raise ChebpolCudaError(rc, L.chebpol_cuda_strerror(rc).decode(), L.chebpol_cuda_last_error().decode())


def device_count() -> int:
    n = C.c_int(0)
    rc = lib().chebpol_cuda_device_count(C.byref(n))
    if rc != 0:
        return 0
    return n.value


def cache_clear() -> None:
    lib().chebpol_cuda_cache_clear()


def cache_stats() -> dict:
    """Facts of the stateless entry points' plan cache (csrc/api.cu)."""
    v = [C.c_int64(0) for _ in range(5)]
    check(lib().chebpol_cuda_debug_cache_stats(*[C.byref(t) for t in v]))
    return dict(zip(("entries", "hits", "misses", "rejected", "device_bytes"), (int(t.value) for t in v)))


def cache_weak_hash(on: bool) -> None:
    lib().chebpol_cuda_debug_cache_weak_hash(1 if on else 0)


def fingerprint(a: np.ndarray, threads: int = 1) -> tuple[int, int]:
    a = np.ascontiguousarray(a)
    out = (C.c_uint64 * 2)()
    check(lib().chebpol_cuda_debug_fingerprint(_addr(a), a.nbytes, threads, out))
    return int(out[0]), int(out[1])


def cache_would_hit(a: np.ndarray, b: np.ndarray, retain: bool = True) -> bool:
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    hit = C.c_int(0)
    check(lib().chebpol_cuda_debug_cache_would_hit(_addr(a), a.nbytes, _addr(b), b.nbytes, 1 if retain else 0,
                                                   C.byref(hit)))
    return bool(hit.value)


def synth_points_device(x_ptr: int, first_point: int, npts: int, rank: int, seed: int = 42, lo: float = -1.0,
                        hi: float = 1.0, device: int = 0, stream: int = 0) -> None:
    """Fill a device buffer with the counter-based synthetic batch (twin of chebpol_b200.synth.points)."""
    check(lib().chebpol_cuda_synth_points(seed, first_point, npts, rank, lo, hi, x_ptr, device, stream))


def launch_count() -> int:
    return int(lib().chebpol_cuda_launch_count())


def fp64_peak_tflops(device: int = 0, repeats: int = 5) -> float:
    v = C.c_double(0.0)
    check(lib().chebpol_cuda_fp64_peak(device, repeats, C.byref(v)))
    return v.value


# ---- helpers -----------------------------------------------------------------
def f64(a, order="C"):
    return np.require(a, dtype=np.float64, requirements=["C" if order == "C" else "F", "A"])


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _addr(a) -> int:
    return a.ctypes.data if a is not None else 0


def ptr_list(arrs):
    keep = [f64(np.asarray(a).ravel()) for a in arrs]
    arr = (_dp * len(keep))(*[k.ctypes.data_as(_dp) for k in keep])
    return arr, keep


def as_points(x, rank: int) -> np.ndarray:
    """Coerce to the (P, K) point-major layout of R's K x P column-major matrix.

    Accepts what the reference's closures accept (R/tools.R:9-20): a K x P
    matrix (numpy shape (K, P), any memory order -- a Fortran-ordered one is
    passed through without a copy), or a length-K vector.
    """
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        if rank == 1:
            return f64(x.reshape(-1, 1))
        if x.size == rank:
            return f64(x.reshape(1, rank))
        raise ValueError(f"Function should take {rank} arguments, you supplied a vector of length {x.size}")
    if x.ndim != 2 or x.shape[0] != rank:
        raise ValueError(f"argument must be a {rank} x P matrix, got shape {x.shape}")
    return f64(x.T)  # (P, K) C-contiguous == K x P column-major


class Plan:
    """A fitted interpolant resident on one GPU (chebpol_cuda_plan)."""

    def __init__(self, handle: int, rank: int, kind: str):
        self._h = _vp(handle)
        self.rank = rank
        self.kind = kind

    # -- constructors ----------------------------------------------------------
    @staticmethod
    def cheb(coef, mid=None, ispan=None, ugm_n=None, device: int = 0) -> "Plan":
        coef = np.asarray(coef, dtype=np.float64)
        dims = _i32(coef.shape)
        cf = f64(coef, order="F")
        mid_a = f64(mid) if mid is not None else None
        isp_a = f64(ispan) if ispan is not None else None
        ugm_a = _i32(ugm_n) if ugm_n is not None else None
        h = _vp()
        check(lib().chebpol_cuda_plan_cheb(_addr(cf), dims.ctypes.data_as(_ip), coef.ndim, _addr(mid_a),
                                           _addr(isp_a), _addr(ugm_a), device, C.byref(h)))
        return Plan(h.value, coef.ndim, "cheb")

    @staticmethod
    def chebg(coef, splines, device: int = 0) -> "Plan":
        """Chebyshev plan behind the 'general' grid maps; splines: per dimension (x, y, b, c, d)."""
        coef = np.asarray(coef, dtype=np.float64)
        dims = _i32(coef.shape)
        cf = f64(coef, order="F")
        if len(splines) != coef.ndim:
            raise ValueError(f"{len(splines)} grid maps for a rank-{coef.ndim} coefficient array")
        sn = _i32([len(t[0]) for t in splines])
        lists = [ptr_list([t[j] for t in splines]) for j in range(5)]
        h = _vp()
        check(lib().chebpol_cuda_plan_chebg(_addr(cf), dims.ctypes.data_as(_ip), coef.ndim, sn.ctypes.data_as(_ip),
                                            *[l[0] for l in lists], device, C.byref(h)))
        return Plan(h.value, coef.ndim, "cheb")

    @staticmethod
    def mlip(grid, values, blend: int = 0, device: int = 0) -> "Plan":
        dims = _i32([len(g) for g in grid])
        vals = f64(np.asarray(values, dtype=np.float64).ravel(order="F"))
        if vals.size != int(np.prod(dims.astype(np.int64))):
            raise ValueError(f"grid has size {int(np.prod(dims.astype(np.int64)))}, you supplied {vals.size} values")
        gl, keep = ptr_list(grid)
        h = _vp()
        check(lib().chebpol_cuda_plan_mlip(gl, dims.ctypes.data_as(_ip), len(grid), _addr(vals), blend, device,
                                           C.byref(h)))
        return Plan(h.value, len(grid), "mlip")

    @staticmethod
    def fh(vals, grid, weights, device: int = 0) -> "Plan":
        dims = _i32([len(g) for g in grid])
        v = f64(np.asarray(vals, dtype=np.float64).ravel(order="F"))
        if v.size != int(np.prod(dims.astype(np.int64))):
            raise ValueError(f"coefficient length({v.size}) does not match data length({int(np.prod(dims.astype(np.int64)))})")
        gl, k1 = ptr_list(grid)
        wl, k2 = ptr_list(weights)
        h = _vp()
        check(lib().chebpol_cuda_plan_fh(_addr(v), gl, wl, dims.ctypes.data_as(_ip), len(grid), device, C.byref(h)))
        return Plan(h.value, len(grid), "fh")

    @staticmethod
    def polyh(knots, weights, lweights, k: float, norm_a=None, norm_b=None, device: int = 0) -> "Plan":
        """knots: rank x nknots (column i is knot i, as the R matrix)."""
        kn = np.asarray(knots, dtype=np.float64)
        rank, nk = kn.shape
        knf = f64(kn, order="F")
        w, lw = f64(np.ravel(weights)), f64(np.ravel(lweights))
        if w.size != nk:
            raise ValueError(f"number of weights ({w.size}) should equal number of knots ({nk})")
        if lw.size != rank + 1:
            raise ValueError(f"linear weights ({lw.size}) should be one longer than rank({rank})")
        na = f64(np.ravel(norm_a)) if norm_a is not None else None
        nb = f64(np.ravel(norm_b)) if norm_b is not None else None
        h = _vp()
        check(lib().chebpol_cuda_plan_polyh(_addr(knf), rank, nk, _addr(w), _addr(lw), float(k), _addr(na), _addr(nb),
                                            device, C.byref(h)))
        return Plan(h.value, rank, "polyh")

    @staticmethod
    def stalk(kind: int, val, grid, tables, blend: int = 3, device: int = 0) -> "Plan":
        """kind 0: stalker, tables (b, c, r); kind 1: hstalker, tables (a, b, c, d).  b, c, r/d are
        rank x prod(dims) (column i belongs to grid point i in R order), a has prod(dims) entries."""
        dims = _i32([len(g) for g in grid])
        n = int(np.prod(dims.astype(np.int64)))
        v = f64(np.asarray(val, dtype=np.float64).ravel(order="F"))
        if v.size != n:
            raise ValueError(f"number of values ({v.size}) does not match grid size ({n})")
        a = f64(np.ravel(tables[0])) if kind == 1 else None
        b, t1, t2 = (f64(np.asarray(t, dtype=np.float64), order="F") for t in tables[-3:])
        for t in (b, t1, t2):
            if t.shape != (len(grid), n):
                raise ValueError(f"parameter table has shape {t.shape}, expected {(len(grid), n)}")
        gl, keep = ptr_list(grid)
        h = _vp()
        check(lib().chebpol_cuda_plan_stalk(kind, _addr(v), gl, dims.ctypes.data_as(_ip), len(grid), _addr(a), _addr(b),
                                            _addr(t1), _addr(t2), blend, device, C.byref(h)))
        return Plan(h.value, len(grid), "stalk")

    @staticmethod
    def oldstalk(val, grid, degree, ctx, uniform, blend: int = 0, device: int = 0) -> "Plan":
        """Deprecated recursive stalker evaluator; ctx = (det, pmin, pplus), lists of per-dimension vectors."""
        dims, v, deg, uni, lists = _oldstalk_args(val, grid, degree, ctx, uniform)
        gl, keep = ptr_list(grid)
        h = _vp()
        check(lib().chebpol_cuda_plan_oldstalk(_addr(v), gl, dims.ctypes.data_as(_ip), len(grid), _addr(deg),
                                               lists[0][0], lists[1][0], lists[2][0], uni.ctypes.data_as(_ip), blend,
                                               device, C.byref(h)))
        return Plan(h.value, len(grid), "oldstalk")

    @staticmethod
    def sl(knots, dtri, adata, val, device: int = 0) -> "Plan":
        """Simplex-linear plan.  knots: dim x nknots; dtri: (dim+1) x numsimplex, 1-based;
        adata = (lumats, pivots, bbox) as analyzesimplex returns them."""
        kn, tri, (lu, piv, bb), v = _sl_args(knots, dtri, adata, val)
        dim, nk = kn.shape
        h = _vp()
        check(lib().chebpol_cuda_plan_sl(_addr(kn), dim, nk, _addr(tri), tri.shape[1], _addr(lu), _addr(piv), _addr(bb),
                                         _addr(v), device, C.byref(h)))
        return Plan(h.value, dim, "sl")

    # -- evaluation --------------------------------------------------------------
    def eval_host(self, x_pk: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        """x_pk: (P, K) C-contiguous float64 host array (pinned or pageable)."""
        assert x_pk.dtype == np.float64 and x_pk.flags.c_contiguous and x_pk.shape[1] == self.rank
        n = x_pk.shape[0]
        if out is None:
            out = np.empty(n, dtype=np.float64)
        check(lib().chebpol_cuda_plan_eval_host(self._h, _addr(x_pk), n, _addr(out)))
        return out

    def eval_host_ptr(self, x_ptr: int, npts: int, out_ptr: int) -> None:
        check(lib().chebpol_cuda_plan_eval_host(self._h, x_ptr, npts, out_ptr))

    def eval_device_ptr(self, x_ptr: int, npts: int, out_ptr: int, stream: int = 0) -> None:
        check(lib().chebpol_cuda_plan_eval_device(self._h, x_ptr, npts, out_ptr, stream))

    def eval_torch(self, x, out=None):
        """x: CUDA float64 tensor of shape (P, K), contiguous, on the plan's device."""
        import torch

        assert x.is_cuda and x.dtype == torch.float64 and x.is_contiguous() and x.shape[1] == self.rank
        if out is None:
            out = torch.empty(x.shape[0], dtype=torch.float64, device=x.device)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        self.eval_device_ptr(x.data_ptr(), x.shape[0], out.data_ptr(), stream)
        return out

    # -- options -------------------------------------------------------------------
    def set(self, option: int, value: int) -> "Plan":
        check(lib().chebpol_cuda_plan_set(self._h, option, value))
        return self

    def get(self, what: int) -> int:
        v = C.c_int64(0)
        check(lib().chebpol_cuda_plan_get(self._h, what, C.byref(v)))
        return v.value

    @property
    def kernel_used(self) -> str:
        return KERNEL_NAMES.get(self.get(GET_KERNEL_USED), "?")

    def close(self) -> None:
        if self._h:
            lib().chebpol_cuda_plan_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- deprecated recursive stalker ---------------------------------------------
def _oldstalk_args(val, grid, degree, ctx, uniform):
    dims = _i32([len(g) for g in grid])
    n = int(np.prod(dims.astype(np.int64)))
    v = f64(np.asarray(val, dtype=np.float64).ravel(order="F"))
    if v.size != n:
        raise ValueError(f"number of values ({v.size}) does not match grid size ({n})")
    deg = f64(np.ravel(degree))
    if deg.size != len(grid):
        raise ValueError(f"degree length must match dimension ({len(grid)})")
    uni = _i32(np.ravel(uniform))
    lists = [ptr_list(c) for c in ctx]
    for c in ctx:
        if len(c) != len(grid) or any(len(a) != len(g) for a, g in zip(c, grid)):
            raise ValueError("stalker context does not match the grid")
    return dims, v, deg, uni, lists


def evalstalker(val, grid, degree, ctx, uniform, x_pk, blend: int = 0, ngpus: int = 0) -> np.ndarray:
    dims, v, deg, uni, lists = _oldstalk_args(val, grid, degree, ctx, uniform)
    gl, keep = ptr_list(grid)
    x = f64(x_pk)
    out = np.empty(x.shape[0], dtype=np.float64)
    check(lib().chebpol_cuda_evalstalker(_addr(v), gl, dims.ctypes.data_as(_ip), len(grid), _addr(deg), lists[0][0],
                                         lists[1][0], lists[2][0], uni.ctypes.data_as(_ip), blend, _addr(x), x.shape[0],
                                         ngpus, _addr(out)))
    return out


# ---- simplex-linear ------------------------------------------------------------
def _sl_args(knots, dtri, adata, val):
    kn = f64(np.asarray(knots, dtype=np.float64), order="F")
    tri = np.require(dtri, dtype=np.int32, requirements=["F", "A"])
    dim, nk = kn.shape
    if tri.ndim != 2 or tri.shape[0] != dim + 1:
        raise ValueError(f"dtri must be a {dim + 1} x numsimplex matrix, got shape {tri.shape}")
    ns = tri.shape[1]
    lu, piv, bb = adata
    lu, bb = f64(np.ravel(lu, order="F")), f64(np.ravel(bb, order="F"))
    piv = np.ascontiguousarray(np.ravel(piv, order="F"), dtype=np.int32)
    v = f64(np.ravel(val))
    if lu.size != ns * (dim + 1) ** 2 or piv.size != ns * (dim + 1) or bb.size != ns * 2 * dim:
        raise ValueError("adata does not match the triangulation")
    if v.size != nk:
        raise ValueError(f"number of values ({v.size}) should equal number of knots ({nk})")
    return kn, tri, (lu, piv, bb), v


def analyzesimplex(dtri, knots, device: int = 0):
    """R_analyzesimplex on the device: (lumats, pivots, bbox)."""
    kn = f64(np.asarray(knots, dtype=np.float64), order="F")
    tri = np.require(dtri, dtype=np.int32, requirements=["F", "A"])
    dim, nk = kn.shape
    N, ns = tri.shape
    if N != dim + 1:
        raise ValueError(f"dtri must be a {dim + 1} x numsimplex matrix, got shape {tri.shape}")
    lu = np.empty(ns * N * N)
    piv = np.empty(ns * N, dtype=np.int32)
    bb = np.empty(ns * 2 * dim)
    check(lib().chebpol_cuda_analyzesimplex(_addr(tri), ns, _addr(kn), dim, nk, device, _addr(lu), _addr(piv), _addr(bb)))
    return lu, piv, bb


def debug_sl_lists(knots, dtri, bbox, cells_per_dim: int):
    """Candidate lists of the sl_cell kernel, built on the host (no GPU): (G, lo, inv, start, items)."""
    kn = f64(np.asarray(knots, dtype=np.float64), order="F")
    tri = np.require(dtri, dtype=np.int32, requirements=["F", "A"])
    dim, nk = kn.shape
    ns = tri.shape[1]
    bb = f64(np.ravel(bbox, order="F"))
    g = max(1, int(cells_per_dim))
    start = np.zeros(min(g, 256) ** dim + 1 if g ** dim <= (1 << 24) else (1 << 24) + 1, dtype=np.int32)
    start = np.zeros(g ** dim + 1, dtype=np.int32)
    cap = 1 << 26
    items = np.zeros(cap, dtype=np.int32)
    lo, inv = np.zeros(dim), np.zeros(dim)
    gu, n = C.c_int(0), C.c_int64(0)
    check(lib().chebpol_cuda_debug_sl_lists(_addr(kn), dim, nk, _addr(tri), ns, _addr(bb), g, C.byref(gu), _addr(lo),
                                            _addr(inv), _addr(start), start.size, _addr(items), cap, C.byref(n)))
    G = gu.value
    return G, lo, inv, start[: G ** dim + 1], items[: n.value]


def evalsl(x_pk, knots, dtri, adata, val, epol: bool = False, blend: int = 3, ngpus: int = 0) -> np.ndarray:
    kn, tri, (lu, piv, bb), v = _sl_args(knots, dtri, adata, val)
    dim, nk = kn.shape
    x = f64(x_pk)
    out = np.empty(x.shape[0], dtype=np.float64)
    check(lib().chebpol_cuda_evalsl(_addr(x), x.shape[0], _addr(kn), dim, nk, _addr(tri), tri.shape[1], _addr(lu),
                                    _addr(piv), _addr(bb), _addr(v), int(bool(epol)), blend, ngpus, _addr(out)))
    return out


# ---- fit side ------------------------------------------------------------------
def chebcoef(val, dct: bool = False, device: int = 0) -> np.ndarray:
    """Chebyshev coefficients of values on a Chebyshev grid, on the device (chebpol_cuda_chebcoef)."""
    val = np.asarray(val, dtype=np.float64)
    if val.ndim == 0:
        val = val.reshape(1)
    v = f64(val, order="F")
    dims = _i32(v.shape)
    out = np.empty(v.shape, dtype=np.float64, order="F")
    check(lib().chebpol_cuda_chebcoef(_addr(v), dims.ctypes.data_as(_ip), v.ndim, int(dct), device, _addr(out)))
    return out


def fhweights(grid, d, device: int = 0):
    """Floater-Hormann weights per dimension, on the device (chebpol_cuda_fhweights)."""
    dd = np.resize(np.asarray(d, dtype=np.int64), len(grid))
    out = []
    for g, deg in zip(grid, dd):
        g = f64(np.asarray(g, dtype=np.float64).ravel())
        w = np.empty_like(g)
        check(lib().chebpol_cuda_fhweights(_addr(g), g.size, int(deg), device, _addr(w)))
        out.append(w)
    return out


def evalpolyh(knots, weights, lweights, k: float, x_pk, norm_a=None, norm_b=None, ngpus: int = 0) -> np.ndarray:
    """chebpol_cuda_evalpolyh: knots rank x nknots (column i is knot i), x_pk (P, rank)."""
    kn = np.asarray(knots, dtype=np.float64)
    rank, nk = kn.shape
    knf = f64(kn, order="F")
    w, lw = f64(np.ravel(weights)), f64(np.ravel(lweights))
    na = f64(np.ravel(norm_a)) if norm_a is not None else None
    nb = f64(np.ravel(norm_b)) if norm_b is not None else None
    x_pk = f64(x_pk)
    out = np.empty(x_pk.shape[0], dtype=np.float64)
    check(lib().chebpol_cuda_evalpolyh(_addr(knf), rank, nk, _addr(w), _addr(lw), float(k), _addr(x_pk),
                                       x_pk.shape[0], _addr(na), _addr(nb), ngpus, _addr(out)))
    return out


def evalstalk(val, grid, b, c, r, x_pk, blend: int = 3, ngpus: int = 0) -> np.ndarray:
    dims = _i32([len(g) for g in grid])
    v = f64(np.asarray(val, dtype=np.float64).ravel(order="F"))
    b, c, r = (f64(np.asarray(t, dtype=np.float64), order="F") for t in (b, c, r))
    x_pk = f64(x_pk)
    out = np.empty(x_pk.shape[0], dtype=np.float64)
    gl, keep = ptr_list(grid)
    check(lib().chebpol_cuda_evalstalk(_addr(v), gl, dims.ctypes.data_as(_ip), len(grid), _addr(b), _addr(c), _addr(r),
                                       blend, _addr(x_pk), x_pk.shape[0], ngpus, _addr(out)))
    return out


def evalhyp(val, grid, a, b, c, d, x_pk, blend: int = 3, ngpus: int = 0) -> np.ndarray:
    dims = _i32([len(g) for g in grid])
    v = f64(np.asarray(val, dtype=np.float64).ravel(order="F"))
    a = f64(np.ravel(a))
    b, c, d = (f64(np.asarray(t, dtype=np.float64), order="F") for t in (b, c, d))
    x_pk = f64(x_pk)
    out = np.empty(x_pk.shape[0], dtype=np.float64)
    gl, keep = ptr_list(grid)
    check(lib().chebpol_cuda_evalhyp(_addr(v), gl, dims.ctypes.data_as(_ip), len(grid), _addr(a), _addr(b), _addr(c),
                                     _addr(d), blend, _addr(x_pk), x_pk.shape[0], ngpus, _addr(out)))
    return out


# ---- evaluation on a tensor grid ---------------------------------------------------
def evalgrid_cheb(coef, pts, mid=None, ispan=None, ugm_n=None, device: int = 0) -> np.ndarray:
    coef = np.asarray(coef, dtype=np.float64)
    cf = f64(coef, order="F")
    dims = _i32(coef.shape)
    gdims = _i32([len(p) for p in pts])
    pl, keep = ptr_list(pts)
    out = np.empty(tuple(int(g) for g in gdims), dtype=np.float64, order="F")
    mid_a = f64(mid) if mid is not None else None
    isp_a = f64(ispan) if ispan is not None else None
    ugm_a = _i32(ugm_n) if ugm_n is not None else None
    check(lib().chebpol_cuda_evalgrid_cheb(_addr(cf), dims.ctypes.data_as(_ip), coef.ndim, pl, gdims.ctypes.data_as(_ip),
                                           _addr(mid_a), _addr(isp_a), _addr(ugm_a), device, _addr(out)))
    return out


def evalgrid_mlip(grid, values, pts, blend: int = 0, device: int = 0) -> np.ndarray:
    dims = _i32([len(g) for g in grid])
    vals = f64(np.asarray(values, dtype=np.float64).ravel(order="F"))
    gdims = _i32([len(p) for p in pts])
    gl, k1 = ptr_list(grid)
    pl, k2 = ptr_list(pts)
    out = np.empty(tuple(int(g) for g in gdims), dtype=np.float64, order="F")
    check(lib().chebpol_cuda_evalgrid_mlip(gl, dims.ctypes.data_as(_ip), len(grid), _addr(vals), pl,
                                           gdims.ctypes.data_as(_ip), blend, device, _addr(out)))
    return out


def evalgrid_fh(vals, grid, weights, pts, device: int = 0) -> np.ndarray:
    dims = _i32([len(g) for g in grid])
    v = f64(np.asarray(vals, dtype=np.float64).ravel(order="F"))
    gdims = _i32([len(p) for p in pts])
    gl, k1 = ptr_list(grid)
    wl, k2 = ptr_list(weights)
    pl, k3 = ptr_list(pts)
    out = np.empty(tuple(int(g) for g in gdims), dtype=np.float64, order="F")
    check(lib().chebpol_cuda_evalgrid_fh(_addr(v), gl, wl, dims.ctypes.data_as(_ip), len(grid), pl,
                                         gdims.ctypes.data_as(_ip), device, _addr(out)))
    return out


# ---- stateless entry points (what the R .Call glue binds) -------------------------
def evalcheb(coef, x_pk, mid=None, ispan=None, ugm_n=None, ngpus: int = 0) -> np.ndarray:
    coef = np.asarray(coef, dtype=np.float64)
    dims = _i32(coef.shape)
    cf = f64(coef, order="F")
    x_pk = f64(x_pk)
    out = np.empty(x_pk.shape[0], dtype=np.float64)
    mid_a = f64(mid) if mid is not None else None
    isp_a = f64(ispan) if ispan is not None else None
    ugm_a = _i32(ugm_n) if ugm_n is not None else None
    check(lib().chebpol_cuda_evalcheb(_addr(cf), dims.ctypes.data_as(_ip), coef.ndim, _addr(x_pk), x_pk.shape[0],
                                      _addr(mid_a), _addr(isp_a), _addr(ugm_a), ngpus, _addr(out)))
    return out


def evalchebg(coef, splines, x_pk, ngpus: int = 0) -> np.ndarray:
    """Stateless 'general' method evaluation (chebpol_cuda_evalchebg); splines: per dimension (x, y, b, c, d)."""
    coef = np.asarray(coef, dtype=np.float64)
    dims = _i32(coef.shape)
    cf = f64(coef, order="F")
    x = f64(x_pk)
    sn = _i32([len(t[0]) for t in splines])
    lists = [ptr_list([t[j] for t in splines]) for j in range(5)]
    out = np.empty(x.shape[0], dtype=np.float64)
    check(lib().chebpol_cuda_evalchebg(_addr(cf), dims.ctypes.data_as(_ip), coef.ndim, _addr(x), x.shape[0],
                                       sn.ctypes.data_as(_ip), *[l[0] for l in lists], ngpus, _addr(out)))
    return out


def evalmlip(grid, values, x_pk, blend: int = 0, ngpus: int = 0) -> np.ndarray:
    dims = _i32([len(g) for g in grid])
    vals = f64(np.asarray(values, dtype=np.float64).ravel(order="F"))
    if vals.size != int(np.prod(dims.astype(np.int64))):
        raise ValueError(f"grid has size {int(np.prod(dims.astype(np.int64)))}, you supplied {vals.size} values")
    x_pk = f64(x_pk)
    out = np.empty(x_pk.shape[0], dtype=np.float64)
    gl, keep = ptr_list(grid)
    check(lib().chebpol_cuda_evalmlip(gl, dims.ctypes.data_as(_ip), len(grid), _addr(vals), _addr(x_pk),
                                      x_pk.shape[0], blend, ngpus, _addr(out)))
    return out


def fh(vals, grid, weights, x_pk, ngpus: int = 0) -> np.ndarray:
    dims = _i32([len(g) for g in grid])
    v = f64(np.asarray(vals, dtype=np.float64).ravel(order="F"))
    x_pk = f64(x_pk)
    out = np.empty(x_pk.shape[0], dtype=np.float64)
    gl, k1 = ptr_list(grid)
    wl, k2 = ptr_list(weights)
    check(lib().chebpol_cuda_fh(_addr(v), gl, wl, dims.ctypes.data_as(_ip), len(grid), _addr(x_pk), x_pk.shape[0],
                                ngpus, _addr(out)))
    return out


class Stateless:
    """One stateless entry point with its interpolant arguments marshalled once: calling the object
    is the bare C call a ``.Call`` from the R closure makes (tensor re-passed every time, host
    pointers in and out), without per-call numpy conversions."""

    def __init__(self, fn, head, tail, rank, keep):
        self._fn, self._head, self._tail, self.rank, self._keep = fn, head, tail, rank, keep

    @staticmethod
    def cheb(coef, mid=None, ispan=None, ugm_n=None) -> "Stateless":
        coef = np.asarray(coef, dtype=np.float64)
        dims, cf = _i32(coef.shape), f64(coef, order="F")
        mid_a = f64(mid) if mid is not None else None
        isp_a = f64(ispan) if ispan is not None else None
        ugm_a = _i32(ugm_n) if ugm_n is not None else None
        return Stateless(lib().chebpol_cuda_evalcheb, (_addr(cf), dims.ctypes.data_as(_ip), coef.ndim),
                         (_addr(mid_a), _addr(isp_a), _addr(ugm_a)), coef.ndim, (cf, dims, mid_a, isp_a, ugm_a))

    @staticmethod
    def mlip(grid, values, blend: int = 0) -> "Stateless":
        dims = _i32([len(g) for g in grid])
        vals = f64(np.asarray(values, dtype=np.float64).ravel(order="F"))
        gl, keep = ptr_list(grid)
        return Stateless(lib().chebpol_cuda_evalmlip, (gl, dims.ctypes.data_as(_ip), len(grid), _addr(vals)), (blend,),
                         len(grid), (dims, vals, gl, keep))

    @staticmethod
    def fh(vals, grid, weights) -> "Stateless":
        dims = _i32([len(g) for g in grid])
        v = f64(np.asarray(vals, dtype=np.float64).ravel(order="F"))
        gl, k1 = ptr_list(grid)
        wl, k2 = ptr_list(weights)
        return Stateless(lib().chebpol_cuda_fh, (_addr(v), gl, wl, dims.ctypes.data_as(_ip), len(grid)), (), len(grid),
                         (dims, v, gl, k1, wl, k2))

    def __call__(self, x_pk: np.ndarray, out: np.ndarray | None = None, ngpus: int = 0) -> np.ndarray:
        assert x_pk.dtype == np.float64 and x_pk.flags.c_contiguous and x_pk.shape[1] == self.rank
        n = x_pk.shape[0]
        if out is None:
            out = np.empty(n, dtype=np.float64)
        check(self._fn(*self._head, _addr(x_pk), n, *self._tail, ngpus, _addr(out)))
        return out
