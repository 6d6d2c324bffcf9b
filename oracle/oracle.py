"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes front-end to the two CPU checkers:

* ``Oracle("port")``      -> oracle/liboracle.so, the C restatement
  (oracle/chebpol_oracle.c);
* ``Oracle("reference")`` -> oracle/_ref/libchebpol_ref.so, the reference's own
  src/chebpol.c compiled through the R-API shim (oracle/ref_wrap.c).

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  The product
(``chebpol_b200``) never does.

Points are passed the way R hands them to ``.Call``: a K x P column-major
matrix, i.e. a C-contiguous ``(P, K)`` numpy array (point-major).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_dpp = C.POINTER(_dp)


def build(verbose: bool = False) -> None:
    """Compile liboracle.so and (if /root/reference is present) _ref."""
    r = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout, r.stderr)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed")


def _as_pts(x, rank):
    x = np.ascontiguousarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x.reshape(1, -1) if x.size == rank else x.reshape(-1, rank)
    assert x.shape[1] == rank, (x.shape, rank)
    return x


def _ptr(a):
    return a.ctypes.data_as(_dp)


def _ptr_list(arrs):
    keep = [np.ascontiguousarray(a, dtype=np.float64) for a in arrs]
    arr = (_dp * len(keep))(*[_ptr(a) for a in keep])
    return arr, keep


class Oracle:
    def __init__(self, kind: str = "port"):
        self.kind = kind
        if kind == "port":
            path = os.path.join(_HERE, "liboracle.so")
            self.pfx = "oracle_"
        elif kind == "reference":
            path = os.path.join(_HERE, "_ref", "libchebpol_ref.so")
            self.pfx = "ref_"
        else:
            raise ValueError(kind)
        if not os.path.exists(path):
            build()
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.path = path

    @staticmethod
    def available(kind: str) -> bool:
        p = "liboracle.so" if kind == "port" else os.path.join("_ref", "libchebpol_ref.so")
        return os.path.exists(os.path.join(_HERE, p))

    # -- Chebyshev ---------------------------------------------------------
    def evalcheb(self, coef, x, threads: int = 1):
        """coef: numpy array in R layout (Fortran order, dim 0 fastest)."""
        coef = np.asfortranarray(coef, dtype=np.float64)
        dims = np.array(coef.shape, dtype=np.int32)
        rank = coef.ndim
        x = _as_pts(x, rank)
        out = np.empty(x.shape[0], dtype=np.float64)
        f = getattr(self.lib, self.pfx + "evalcheb")
        f.restype = None if self.kind == "reference" else C.c_int
        f.argtypes = [_dp, _ip, C.c_int, _dp, C.c_int64, C.c_int, _dp]
        f(coef.ctypes.data_as(_dp), dims.ctypes.data_as(_ip), rank, _ptr(x), x.shape[0], threads, _ptr(out))
        return out

    def imap(self, x, mid, ispan, threads: int = 1):
        rank = len(mid)
        x = _as_pts(x, rank)
        if self.kind == "reference":  # R code, not in the C reference: same restatement
            return (x - np.asarray(mid, dtype=np.float64)) * np.asarray(ispan, dtype=np.float64)
        out = np.empty_like(x)
        mid = np.ascontiguousarray(mid, dtype=np.float64)
        ispan = np.ascontiguousarray(ispan, dtype=np.float64)
        f = self.lib.oracle_imap
        f.restype = None
        f.argtypes = [_dp, C.c_int, C.c_int64, _dp, _dp, C.c_int, _dp]
        f(_ptr(x), rank, x.shape[0], _ptr(mid), _ptr(ispan), threads, _ptr(out))
        return out

    def ugm(self, x, n, mid=None, ispan=None, threads: int = 1):
        """R/uniform.R:6,63-72 (only in the port; the C reference has no such routine)."""
        lib = Oracle("port").lib if self.kind == "reference" else self.lib
        n = np.ascontiguousarray(n, dtype=np.int32)
        rank = n.size
        x = _as_pts(x, rank)
        out = np.empty_like(x)
        f = lib.oracle_ugm
        f.restype = None
        f.argtypes = [_dp, C.c_int, C.c_int64, _ip, _dp, _dp, C.c_int, _dp]
        if mid is None:
            f(_ptr(x), rank, x.shape[0], n.ctypes.data_as(_ip), None, None, threads, _ptr(out))
        else:
            mid = np.ascontiguousarray(mid, dtype=np.float64)
            ispan = np.ascontiguousarray(ispan, dtype=np.float64)
            f(_ptr(x), rank, x.shape[0], n.ctypes.data_as(_ip), _ptr(mid), _ptr(ispan), threads, _ptr(out))
        return out

    # -- 'general' grid map: R's splinefun(method='hyman'), port only (not in the C reference) ----
    def hyman_fit(self, x, y):
        """Tables (x, y, b, c, d) of stats::splinefun(x, y, method='hyman'); x strictly increasing."""
        lib = Oracle("port").lib if self.kind == "reference" else self.lib
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        b, c, d = np.empty_like(x), np.empty_like(x), np.empty_like(x)
        f = lib.oracle_hyman_fit
        f.restype = C.c_int
        f.argtypes = [C.c_int, _dp, _dp, _dp, _dp, _dp]
        if f(x.size, _ptr(x), _ptr(y), _ptr(b), _ptr(c), _ptr(d)):
            raise ValueError("hyman_fit needs at least 2 knots")
        return x, y, b, c, d

    def splinemap(self, x, tabs, threads: int = 1):
        """tabs: per dimension a tuple (x, y, b, c, d).  Returns the mapped P x K points."""
        lib = Oracle("port").lib if self.kind == "reference" else self.lib
        rank = len(tabs)
        x = _as_pts(x, rank)
        out = np.empty_like(x)
        sn = np.array([len(t[0]) for t in tabs], dtype=np.int32)
        lists = [_ptr_list([t[j] for t in tabs]) for j in range(5)]
        f = lib.oracle_splinemap
        f.restype = None
        f.argtypes = [_dp, C.c_int, C.c_int64, _ip, _dpp, _dpp, _dpp, _dpp, _dpp, C.c_int, _dp]
        f(_ptr(x), rank, x.shape[0], sn.ctypes.data_as(_ip), *[l[0] for l in lists], threads, _ptr(out))
        return out

    # -- multilinear -------------------------------------------------------
    def evalmlip(self, grid, values, x, blend: int = 0, threads: int = 1):
        rank = len(grid)
        dims = np.array([len(g) for g in grid], dtype=np.int32)
        values = np.ascontiguousarray(np.asarray(values, dtype=np.float64).ravel(order="F"))
        assert values.size == int(np.prod(dims.astype(np.int64)))
        x = _as_pts(x, rank)
        out = np.empty(x.shape[0], dtype=np.float64)
        gl, keep = _ptr_list(grid)
        f = getattr(self.lib, self.pfx + "evalmlip")
        f.restype = None if self.kind == "reference" else C.c_int
        f.argtypes = [C.c_int, _dpp, _ip, _dp, _dp, C.c_int64, C.c_int, C.c_int, _dp]
        f(rank, gl, dims.ctypes.data_as(_ip), _ptr(values), _ptr(x), x.shape[0], blend, threads, _ptr(out))
        return out

    def blendfun(self, w: float, blend: int) -> float:
        f = getattr(self.lib, self.pfx + "blendfun")
        f.restype = C.c_double
        f.argtypes = [C.c_double, C.c_int]
        return f(w, blend)

    def binsearch(self, arr, val: float) -> int:
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        f = getattr(self.lib, self.pfx + "binsearch")
        f.restype = C.c_int
        f.argtypes = [C.c_int, _dp, C.c_double]
        return f(arr.size, _ptr(arr), val)

    # -- Floater-Hormann -----------------------------------------------------
    def fh(self, vals, grid, weights, x, threads: int = 1):
        vals = np.asfortranarray(vals, dtype=np.float64)
        rank = len(grid)
        dims = np.array([len(g) for g in grid], dtype=np.int32)
        assert vals.size == int(np.prod(dims.astype(np.int64)))
        x = _as_pts(x, rank)
        out = np.empty(x.shape[0], dtype=np.float64)
        gl, k1 = _ptr_list(grid)
        wl, k2 = _ptr_list(weights)
        f = getattr(self.lib, self.pfx + "fh")
        f.restype = None if self.kind == "reference" else C.c_int
        f.argtypes = [_dp, _dp, _dpp, _ip, C.c_int, _dpp, C.c_int64, C.c_int, _dp]
        f(vals.ctypes.data_as(_dp), _ptr(x), gl, dims.ctypes.data_as(_ip), rank, wl, x.shape[0], threads, _ptr(out))
        return out

    def evalpolyh(self, knots, weights, lweights, k: float, x, threads: int = 1):
        """knots: rank x nknots (R matrix; any array whose [:, i] is knot i); x: P x rank."""
        knots = np.asfortranarray(knots, dtype=np.float64)
        rank, nk = knots.shape
        w = np.ascontiguousarray(weights, dtype=np.float64)
        lw = np.ascontiguousarray(lweights, dtype=np.float64)
        assert w.size == nk and lw.size == rank + 1
        x = _as_pts(x, rank)
        out = np.empty(x.shape[0], dtype=np.float64)
        f = getattr(self.lib, self.pfx + "evalpolyh")
        f.restype = None if self.kind == "reference" else C.c_int
        f.argtypes = [_dp, C.c_int64, _dp, _dp, _dp, C.c_int, C.c_int, C.c_double, C.c_int, _dp]
        f(_ptr(x), x.shape[0], knots.ctypes.data_as(_dp), _ptr(w), _ptr(lw), rank, nk, float(k), threads, _ptr(out))
        return out

    # -- stalker splines (fit + evaluation) ------------------------------------
    def _grid_args(self, grid):
        dims = np.array([len(g) for g in grid], dtype=np.int32)
        gl, keep = _ptr_list(grid)
        return dims, gl, keep

    def makestalk(self, val, grid):
        """makestalk (no GSL): tables b, c, r, each rank x prod(dims) (Fortran order)."""
        dims, gl, keep = self._grid_args(grid)
        rank, n = len(grid), int(np.prod(dims.astype(np.int64)))
        v = np.ascontiguousarray(np.asarray(val, dtype=np.float64).ravel(order="F"))
        assert v.size == n
        b, c, r = (np.zeros((rank, n), order="F") for _ in range(3))
        f = getattr(self.lib, self.pfx + "makestalk")
        f.restype = None if self.kind == "reference" else C.c_int
        f.argtypes = [C.c_int, _ip, _dpp, _dp, _dp, _dp, _dp]
        f(rank, dims.ctypes.data_as(_ip), gl, _ptr(v), b.ctypes.data_as(_dp), c.ctypes.data_as(_dp), r.ctypes.data_as(_dp))
        return b, c, r

    def makehyp(self, val, grid):
        dims, gl, keep = self._grid_args(grid)
        rank, n = len(grid), int(np.prod(dims.astype(np.int64)))
        v = np.ascontiguousarray(np.asarray(val, dtype=np.float64).ravel(order="F"))
        assert v.size == n
        a = np.zeros(n)
        b, c, d = (np.zeros((rank, n), order="F") for _ in range(3))
        f = getattr(self.lib, self.pfx + "makehyp")
        f.restype = None if self.kind == "reference" else C.c_int
        f.argtypes = [C.c_int, _ip, _dpp, _dp, _dp, _dp, _dp, _dp]
        f(rank, dims.ctypes.data_as(_ip), gl, _ptr(v), _ptr(a), b.ctypes.data_as(_dp), c.ctypes.data_as(_dp),
          d.ctypes.data_as(_dp))
        return a, b, c, d

    def evalstalk(self, val, grid, b, c, r, x, blend: int = 3, threads: int = 1):
        dims, gl, keep = self._grid_args(grid)
        rank = len(grid)
        v = np.ascontiguousarray(np.asarray(val, dtype=np.float64).ravel(order="F"))
        b, c, r = (np.asfortranarray(t, dtype=np.float64) for t in (b, c, r))
        x = _as_pts(x, rank)
        out = np.empty(x.shape[0], dtype=np.float64)
        f = getattr(self.lib, self.pfx + "evalstalk")
        f.restype = None if self.kind == "reference" else C.c_int
        f.argtypes = [_dp, C.c_int64, C.c_int, _ip, _dpp, _dp, C.c_int, _dp, _dp, _dp, C.c_int, _dp]
        f(_ptr(x), x.shape[0], rank, dims.ctypes.data_as(_ip), gl, _ptr(v), blend, b.ctypes.data_as(_dp),
          c.ctypes.data_as(_dp), r.ctypes.data_as(_dp), threads, _ptr(out))
        return out

    def evalhyp(self, val, grid, a, b, c, d, x, blend: int = 3, threads: int = 1):
        dims, gl, keep = self._grid_args(grid)
        rank = len(grid)
        v = np.ascontiguousarray(np.asarray(val, dtype=np.float64).ravel(order="F"))
        a = np.ascontiguousarray(a, dtype=np.float64)
        b, c, d = (np.asfortranarray(t, dtype=np.float64) for t in (b, c, d))
        x = _as_pts(x, rank)
        out = np.empty(x.shape[0], dtype=np.float64)
        f = getattr(self.lib, self.pfx + "evalhyp")
        f.restype = None if self.kind == "reference" else C.c_int
        f.argtypes = [_dp, C.c_int64, C.c_int, _ip, _dpp, _dp, C.c_int, _dp, _dp, _dp, _dp, C.c_int, _dp]
        f(_ptr(x), x.shape[0], rank, dims.ctypes.data_as(_ip), gl, _ptr(v), blend, _ptr(a), b.ctypes.data_as(_dp),
          c.ctypes.data_as(_dp), d.ctypes.data_as(_dp), threads, _ptr(out))
        return out

    # -- simplex-linear on a Delaunay triangulation -----------------------------
    def analyzesimplex(self, dtri, knots):
        """dtri: (dim+1) x numsimplex 1-based knot indices; knots: dim x nknots.
        Returns (lumats, pivots, bbox) as R_analyzesimplex lays them out."""
        knots = np.asfortranarray(knots, dtype=np.float64)
        dtri = np.asfortranarray(dtri, dtype=np.int32)
        dim, nk = knots.shape
        N, ns = dtri.shape
        assert N == dim + 1
        lu = np.empty(ns * N * N)
        piv = np.empty(ns * N, dtype=np.int32)
        bbox = np.empty(ns * 2 * dim)
        f = getattr(self.lib, self.pfx + "analyzesimplex")
        f.restype = None if self.kind == "reference" else C.c_int
        f.argtypes = [_ip, C.c_int, _dp, C.c_int, C.c_int, _dp, _ip, _dp]
        f(dtri.ctypes.data_as(_ip), ns, knots.ctypes.data_as(_dp), dim, nk, _ptr(lu), piv.ctypes.data_as(_ip), _ptr(bbox))
        return lu, piv, bbox

    def evalsl(self, x, knots, dtri, adata, val, epol: bool = False, blend: int = 3, threads: int = 1):
        knots = np.asfortranarray(knots, dtype=np.float64)
        dtri = np.asfortranarray(dtri, dtype=np.int32)
        dim, nk = knots.shape
        ns = dtri.shape[1]
        lu, piv, bbox = adata
        val = np.ascontiguousarray(val, dtype=np.float64)
        assert val.size == nk
        x = _as_pts(x, dim)
        out = np.empty(x.shape[0], dtype=np.float64)
        f = getattr(self.lib, self.pfx + "evalsl")
        f.restype = None if self.kind == "reference" else C.c_int
        f.argtypes = [_dp, C.c_int64, _dp, C.c_int, C.c_int, _ip, C.c_int, _dp, _ip, _dp, _dp, C.c_int, C.c_int, C.c_int, _dp]
        f(_ptr(x), x.shape[0], knots.ctypes.data_as(_dp), dim, nk, dtri.ctypes.data_as(_ip), ns, _ptr(lu),
          piv.ctypes.data_as(_ip), _ptr(bbox), _ptr(val), int(bool(epol)), threads, blend, _ptr(out))
        return out

    def sl_index(self, x, dim, adata, threads: int = 1):
        """0-based index of the simplex the reference accepts for each point (-1: outside the hull)."""
        lib = Oracle("port").lib if self.kind == "reference" else self.lib
        lu, piv, bbox = adata
        ns = bbox.size // (2 * dim)
        x = _as_pts(x, dim)
        out = np.empty(x.shape[0], dtype=np.int32)
        f = lib.oracle_sl_index
        f.restype = C.c_int
        f.argtypes = [_dp, C.c_int64, C.c_int, C.c_int, _dp, _ip, _dp, C.c_int, _ip]
        f(_ptr(x), x.shape[0], dim, ns, _ptr(lu), piv.ctypes.data_as(_ip), _ptr(bbox), threads, out.ctypes.data_as(_ip))
        return out

    def evalstalker_old(self, val, grid, degree, det, pmin, pplus, uniform, x, blend: int = 0, threads: int = 1):
        """The deprecated recursive stalker evaluator; reference library only (it IS the reference's code)."""
        assert self.kind == "reference"
        dims, gl, keep = self._grid_args(grid)
        rank = len(grid)
        v = np.ascontiguousarray(np.asarray(val, dtype=np.float64).ravel(order="F"))
        deg = np.ascontiguousarray(degree, dtype=np.float64)
        uni = np.ascontiguousarray(uniform, dtype=np.int32)
        dl, k1 = _ptr_list(det)
        pl, k2 = _ptr_list(pmin)
        ql, k3 = _ptr_list(pplus)
        x = _as_pts(x, rank)
        out = np.empty(x.shape[0], dtype=np.float64)
        f = self.lib.ref_evalstalker
        f.restype = None
        f.argtypes = [_dp, C.c_int64, C.c_int, _ip, _dpp, _dp, C.c_int, _dp, _dpp, _dpp, _dpp, _ip, C.c_int, _dp]
        f(_ptr(x), x.shape[0], rank, dims.ctypes.data_as(_ip), gl, _ptr(v), blend, _ptr(deg), dl, pl, ql,
          uni.ctypes.data_as(_ip), threads, _ptr(out))
        return out

    # -- fit side (fixtures) -------------------------------------------------
    def fhweights(self, grid, d):
        d = np.broadcast_to(np.asarray(d, dtype=np.int64), (len(grid),))
        f = getattr(self.lib, self.pfx + "fhweights")
        f.restype = None
        f.argtypes = [C.c_int, _dp, C.c_int, _dp]
        out = []
        for g, dd in zip(grid, d):
            g = np.ascontiguousarray(g, dtype=np.float64)
            w = np.empty_like(g)
            f(g.size, _ptr(g), int(dd), _ptr(w))
            out.append(w)
        return out

    def chebcoef(self, val, dct: bool = False):
        val = np.asfortranarray(val, dtype=np.float64)
        dims = np.array(val.shape, dtype=np.int32)
        F = np.empty(val.shape, dtype=np.float64, order="F")
        f = getattr(self.lib, self.pfx + "chebcoef")
        f.restype = None
        f.argtypes = [_dp, _ip, C.c_int, _dp, C.c_int]
        f(val.ctypes.data_as(_dp), dims.ctypes.data_as(_ip), val.ndim, F.ctypes.data_as(_dp), int(dct))
        return F
