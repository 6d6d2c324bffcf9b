"""Host-side mirror of chebpol's R interface for the batched-evaluation path.

R is not available in the build image, so the drop-in boundary that can be
executed is the C ABI (include/chebpol_cuda.h); this module reproduces, in
Python, the R layer that sits above it so that tests and users can drive the
GPU engine exactly the way the R closures would:

    ip = ipol(val, dims=..., intervals=..., grid=..., k=..., method=...)
    y  = ip(x, threads=...)            # x: K x P matrix or length-K vector

Reference (all under /root/reference/R): ipol.R:108-239 (method switch),
chebyshev.R:201-233 (chebappx.real, imap), multilinear.R:49-70 (mlappx.real,
blend names), floater-hormann.R:55-70 (fhappx.real), uniform.R:6,58-77
(ugm, ucappx.real), tools.R:2-25 (vectorfun argument coercion).

Evaluation always goes through libchebpol_cuda; the fit-side helpers
(chebcoef, fhweights, chebknots, evalongrid) run once per interpolant on the
host in numpy and are not part of the accelerated path (SURVEY.md section 2,
rows 8-10).
"""
from __future__ import annotations

import warnings
from typing import Callable, Sequence

import numpy as np

from . import _lib

__all__ = [
    "ipol", "chebappx", "chebappxf", "chebappxg", "chebappxgf", "hyman_spline", "slappx", "oldstalkerappx", "mlappx", "fhappx", "ucappx", "ucappxf",
    "chebeval",
    "chebknots", "chebcoef", "evalongrid", "evalongridV", "fhweights", "Interpolant",
]

BLENDS = {"linear": 0, "sigmoid": 1, "parodic": 2, "cubic": 3, "square": 4, "mean": 5}  # R/multilinear.R:64-65, R/simplexlinear.R:37


# ------------------------------------------------------------------ fit side
def _chebknots1(n: int, interval=None) -> np.ndarray:
    """R/chebyshev.R:4-9."""
    kn = np.cos(np.pi * (np.arange(1, n + 1) - 0.5) / n)
    if n % 2 == 1:
        kn[(n + 1) // 2 - 1] = 0.0
    if interval is None:
        return kn
    a, b = float(interval[0]), float(interval[1])
    return kn * (b - a) / 2 + (a + b) / 2


def chebknots(dims, intervals=None) -> list[np.ndarray]:
    """Chebyshev grid on a hypercube (R/chebyshev.R:41-51)."""
    dims = [int(d) for d in np.atleast_1d(dims)]
    if intervals is None:
        return [_chebknots1(d) for d in dims]
    if len(dims) == 1 and np.isscalar(intervals[0]):
        intervals = [intervals]
    return [_chebknots1(d, iv) for d, iv in zip(dims, intervals)]


def evalongrid(fun: Callable, dims=None, intervals=None, grid=None) -> np.ndarray:
    """Evaluate fun on a tensor grid (R/tools.R:122-128); array in R (Fortran) order."""
    if grid is None:
        grid = chebknots(dims, intervals)
    grid = [np.asarray(g, dtype=np.float64) for g in grid]
    shape = tuple(len(g) for g in grid)
    out = np.empty(shape, dtype=np.float64, order="F")
    flat = out.reshape(-1, order="F")
    idx = np.zeros(len(grid), dtype=np.int64)
    arg = np.array([g[0] for g in grid])
    for j in range(flat.size):
        flat[j] = fun(arg.copy())
        for d in range(len(grid)):  # odometer, dim 0 fastest
            idx[d] += 1
            if idx[d] < shape[d]:
                arg[d] = grid[d][idx[d]]
                break
            idx[d] = 0
            arg[d] = grid[d][0]
    return flat.reshape(shape, order="F")


def evalongridV(ip, dims=None, intervals=None, grid=None, blend=None) -> np.ndarray:
    """evalongridV for an interpolant (R/tools.R:132-137).  The reference evaluates
    fun(t(expand.grid(grid))) point by point; for the four evaluators of this package the
    tensor-grid structure is used instead (one mode product per dimension, on the device)."""
    if grid is None:
        grid = chebknots(dims, intervals)
    pts = [np.asarray(g, dtype=np.float64) for g in grid]
    t = ip.t
    if ip.kind in ("chebyshev", "uniform"):
        return _lib.evalgrid_cheb(t["coef"], pts, t.get("mid"), t.get("ispan"), t.get("ugm_n"))
    if ip.kind == "multilinear":
        b = BLENDS[blend or "linear"] if isinstance(blend or "linear", str) else int(blend)
        return _lib.evalgrid_mlip(t["grid"], t["values"], pts, b)
    if ip.kind == "fh":
        return _lib.evalgrid_fh(t["values"], t["grid"], t["weights"], pts)
    # any other interpolant: the reference's own route, fun(t(expand.grid(grid))), evaluated on the GPU
    mesh = np.meshgrid(*pts, indexing="ij")
    x = np.stack([m.ravel(order="F") for m in mesh])
    out = ip(x, blend=blend) if ip.kind in ("stalker", "hstalker") else ip(x)
    shape = tuple(len(p) for p in pts)
    return out.reshape(shape, order="F") if len(shape) > 1 else out


def chebcoef(val, dct: bool = False) -> np.ndarray:
    """Chebyshev coefficients of values on a Chebyshev grid (R/chebyshev.R:82-84,
    src/chebpol.c:8-141): DCT-II per dimension scaled by 1/prod(dims), index-0
    hyperplanes halved.  Fit side: host numpy/scipy."""
    from scipy.fft import dctn

    val = np.asarray(val, dtype=np.float64)
    if val.ndim == 0:
        val = val.reshape(1)
    F = dctn(val, type=2)
    if dct:
        return np.asfortranarray(F)
    F = F / val.size
    for d in range(val.ndim):
        sl = [slice(None)] * val.ndim
        sl[d] = 0
        F[tuple(sl)] *= 0.5
    return np.asfortranarray(F)


def fhweights(grid: Sequence[np.ndarray], d) -> list[np.ndarray]:
    """Floater-Hormann weights, eq. (18) of the FH paper (src/chebpol.c:267-282)."""
    dd = np.broadcast_to(np.asarray(d, dtype=np.int64), (len(grid),)) if np.ndim(d) == 0 else np.resize(
        np.asarray(d, dtype=np.int64), len(grid))
    out = []
    for g, deg in zip(grid, dd):
        g = np.asarray(g, dtype=np.float64)
        n = g.size - 1
        deg = int(deg)
        w = np.empty(n + 1)
        for k in range(n + 1):
            lo = 0 if k < deg else k - deg
            hi = k if k < n - deg else n - deg
            s = 0.0
            for i in range(lo, hi + 1):
                prod = 1.0
                for j in range(i, i + deg + 1):
                    if j != k:
                        prod *= g[k] - g[j]
                s += 1.0 / abs(prod)
            w[k] = s * (-1.0 if (k % 2 != deg % 2) else 1.0)
        out.append(w)
    return out


def _unsorted(g: np.ndarray) -> bool:
    d = np.diff(g)
    return not (np.all(d > 0) or np.all(d < 0))


# --------------------------------------------------------------- interpolant
class Interpolant:
    """The closure ipol() returns: call it on a K x P matrix or a length-K vector.

    ``threads`` is accepted for signature compatibility (R/interpolant.R:9-11)
    and ignored: points are spread over GPU threads, and over ``ngpus`` GPUs.
    """

    def __init__(self, kind: str, arity: int, domain, **tensors):
        self.kind = kind
        self.arity = arity
        self.domain = domain
        self.t = tensors
        self._plans: dict[int, _lib.Plan] = {}
        self.kernel = _lib.KERNEL_AUTO

    # save()/load() of the R closures (R/chebpol-package.R:18-22): an interpolant is its host tensors;
    # device plans are dropped on pickling and rebuilt on first use after loading
    def __getstate__(self):
        st = self.__dict__.copy()
        st["_plans"] = {}
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)
        self._plans = {}

    # device-resident copy, re-creatable from the host tensors at any time
    def plan(self, device: int = 0) -> _lib.Plan:
        p = self._plans.get(device)
        if p is None:
            t = self.t
            if self.kind in ("chebyshev", "uniform"):
                p = _lib.Plan.cheb(t["coef"], t.get("mid"), t.get("ispan"), t.get("ugm_n"), device)
            elif self.kind == "general":
                p = _lib.Plan.chebg(t["coef"], t["splines"], device)
            elif self.kind == "multilinear":
                p = _lib.Plan.mlip(t["grid"], t["values"], 0, device)
            elif self.kind == "fh":
                p = _lib.Plan.fh(t["values"], t["grid"], t["weights"], device)
            elif self.kind in ("stalker", "hstalker"):
                p = _lib.Plan.stalk(1 if self.kind == "hstalker" else 0, t["values"], t["grid"], t["tables"], 3, device)
            elif self.kind == "oldstalker":
                p = _lib.Plan.oldstalk(t["values"], t["grid"], t["degree"], t["ctx"], t["uniform"], 0, device)
            elif self.kind == "simplexlinear":
                p = _lib.Plan.sl(t["knots"], t["dtri"], t["adata"], t["values"], device)
            elif self.kind == "polyharmonic":
                p = _lib.Plan.polyh(t["knots"], t["weights"], t["lweights"], t["k"], t.get("norm_a"), t.get("norm_b"),
                                    device)
            else:
                raise ValueError(self.kind)
            self._plans[device] = p
        p.set(_lib.OPT_KERNEL, self.kernel)
        return p

    def _coerce(self, x) -> tuple[np.ndarray, bool]:
        """vectorfun, R/tools.R:9-20."""
        x = np.asarray(x, dtype=np.float64)
        K = self.arity
        if x.ndim == 2:
            if x.shape[0] != K:
                raise ValueError(f"coeffcient rank({K}) must match number of rows({x.shape[0]})")
            return _lib.f64(x.T), False
        x = x.ravel()
        if x.size == K:
            return _lib.f64(x.reshape(1, K)), True
        if K == 1:
            return _lib.f64(x.reshape(-1, 1)), False
        if x.size % K == 0:
            numvec = x.size // K
            if numvec > 1:
                warnings.warn(f"Coercing vector of length {x.size} into {K}x{numvec} matrix")
            return _lib.f64(x.reshape(numvec, K)), False
        raise ValueError(f"Function should take {K} arguments, you supplied a vector of length {x.size}")

    def __call__(self, x, threads=None, blend: str | int | None = None, ngpus: int = 1, device: int = 0,
                 epol: bool = False):
        xp, _single = self._coerce(x)
        blended = self.kind in ("multilinear", "stalker", "hstalker", "simplexlinear", "oldstalker")
        if blended:
            if blend is None:  # closure defaults: 'linear' (R/multilinear.R:63, R/stalker.R:36), 'cubic' (R/stalker.R:62,80; R/simplexlinear.R:36)
                blend = "linear" if self.kind in ("multilinear", "oldstalker") else "cubic"
            b = BLENDS[blend] if isinstance(blend, str) else int(blend)
        if self.kind == "simplexlinear":
            t = self.t
            if ngpus != 1:
                return _lib.evalsl(xp, t["knots"], t["dtri"], t["adata"], t["values"], epol, b, ngpus)
            p = self.plan(device)
            p.set(_lib.OPT_BLEND, b)
            p.set(_lib.OPT_EPOL, int(bool(epol)))
            return p.eval_host(xp)
        if ngpus != 1:
            t = self.t
            if self.kind in ("chebyshev", "uniform"):
                return _lib.evalcheb(t["coef"], xp, t.get("mid"), t.get("ispan"), t.get("ugm_n"), ngpus)
            if self.kind == "general":
                return _lib.evalchebg(t["coef"], t["splines"], xp, ngpus)
            if self.kind == "multilinear":
                return _lib.evalmlip(t["grid"], t["values"], xp, b, ngpus)
            if self.kind == "stalker":
                return _lib.evalstalk(t["values"], t["grid"], *t["tables"], xp, b, ngpus)
            if self.kind == "hstalker":
                return _lib.evalhyp(t["values"], t["grid"], *t["tables"], xp, b, ngpus)
            if self.kind == "oldstalker":
                return _lib.evalstalker(t["values"], t["grid"], t["degree"], t["ctx"], t["uniform"], xp, b, ngpus)
            if self.kind == "polyharmonic":
                return _lib.evalpolyh(t["knots"], t["weights"], t["lweights"], t["k"], xp, t.get("norm_a"),
                                      t.get("norm_b"), ngpus)
            return _lib.fh(t["values"], t["grid"], t["weights"], xp, ngpus)
        p = self.plan(device)
        if blended:
            p.set(_lib.OPT_BLEND, b)
        return p.eval_host(xp)

    def close(self):
        for p in self._plans.values():
            p.close()
        self._plans.clear()


# --------------------------------------------------------------- constructors
def chebappx(val, intervals=None, fit: str = "host") -> Interpolant:
    """chebappx.real, R/chebyshev.R:201-233.  fit="device" computes the coefficients with
    chebpol_cuda_chebcoef (the GPU counterpart of C_chebcoef) instead of host numpy."""
    val = np.asarray(val, dtype=np.float64)
    if val.ndim == 0:
        val = val.reshape(1)
    K = val.ndim
    if intervals is not None and np.ndim(intervals) == 1 and len(intervals) == 2 and np.isscalar(intervals[0]):
        intervals = [intervals]
    cf = _lib.chebcoef(val) if fit == "device" else chebcoef(val)
    if intervals is None:
        return Interpolant("chebyshev", K, [(-1.0, 1.0)] * K, coef=cf)
    if any(len(iv) != 2 for iv in intervals):
        raise ValueError("interval elements should have length 2")
    if len(intervals) != K:
        raise ValueError(f"values should have the same dimension as intervals {len(intervals)} {K}")
    ispan = np.array([2.0 / (float(iv[1]) - float(iv[0])) for iv in intervals])
    mid = np.array([(float(iv[0]) + float(iv[1])) / 2.0 for iv in intervals])
    return Interpolant("chebyshev", K, [tuple(iv) for iv in intervals], coef=cf, mid=mid, ispan=ispan)


def chebappxf(fun, dims, intervals=None) -> Interpolant:
    """chebappxf.real, R/chebyshev.R:243-245."""
    dims = [int(d) for d in np.atleast_1d(dims)]
    if intervals is not None and np.ndim(intervals) == 1 and len(intervals) == 2 and np.isscalar(intervals[0]):
        intervals = [intervals]
    return chebappx(evalongrid(fun, dims, intervals), intervals)


def hyman_spline(x, y):
    """Tables (x, y, b, c, d) of R's stats::splinefun(x, y, method='hyman') -- the fit behind the
    grid maps of chebappxg.real (R/chebyshev.R:349).  Fit side, host numpy: the "fmm" cubic spline
    (Forsythe, Malcolm & Moler; end conditions from the third divided differences of the first and
    last four points), Hyman's monotonicity filter on the slopes, and the c, d coefficients
    recomputed from the filtered slopes (R's spline_coef / hyman_filter / spl_coef_conv).  As
    splinefun does (regularize.values), the knots are sorted increasingly first."""
    x = np.asarray(x, dtype=np.float64).ravel()
    y = np.asarray(y, dtype=np.float64).ravel()
    if x.size != y.size or x.size < 2:
        raise ValueError("hyman_spline needs two vectors of the same length >= 2")
    o = np.argsort(x, kind="stable")
    x, y = x[o].copy(), y[o].copy()
    if np.any(np.diff(x) <= 0):
        raise ValueError("grid must be distinct ordered values")
    dy = np.diff(y)
    if not (np.all(dy >= 0) or np.all(dy <= 0)):
        raise ValueError("'y' must be increasing or decreasing")
    n = x.size
    b, c, d = np.zeros(n), np.zeros(n), np.zeros(n)
    if n < 3:
        b[:] = (y[1] - y[0]) / (x[1] - x[0])
    else:
        m = n - 1
        d[0] = x[1] - x[0]
        c[1] = (y[1] - y[0]) / d[0]
        for i in range(1, m):
            d[i] = x[i + 1] - x[i]
            b[i] = 2.0 * (d[i - 1] + d[i])
            c[i + 1] = (y[i + 1] - y[i]) / d[i]
            c[i] = c[i + 1] - c[i]
        b[0], b[m] = -d[0], -d[m - 1]
        c[0] = c[m] = 0.0
        if n > 3:
            c[0] = c[2] / (x[3] - x[1]) - c[1] / (x[2] - x[0])
            c[m] = c[m - 1] / (x[m] - x[m - 2]) - c[m - 2] / (x[m - 1] - x[m - 3])
            c[0] = c[0] * d[0] * d[0] / (x[3] - x[0])
            c[m] = -c[m] * d[m - 1] * d[m - 1] / (x[m] - x[m - 3])
        for i in range(1, m + 1):
            t = d[i - 1] / b[i - 1]
            b[i] = b[i] - t * d[i - 1]
            c[i] = c[i] - t * c[i - 1]
        c[m] = c[m] / b[m]
        for i in range(m - 1, -1, -1):
            c[i] = (c[i] - d[i] * c[i + 1]) / b[i]
        b[m] = (y[m] - y[m - 1]) / d[m - 1] + d[m - 1] * (c[m - 1] + 2.0 * c[m])
        for i in range(m):
            b[i] = (y[i + 1] - y[i]) / d[i] - d[i] * (c[i + 1] + 2.0 * c[i])
    # hyman_filter
    ss = np.diff(y) / np.diff(x)
    s0, s1 = np.concatenate([ss[:1], ss]), np.concatenate([ss, ss[-1:]])
    t1 = np.minimum(np.abs(s0), np.abs(s1))
    sig = np.where(s0 * s1 > 0, s1, b)
    pos = sig >= 0
    b = np.where(pos, np.minimum(np.maximum(0.0, b), 3.0 * t1), np.maximum(np.minimum(0.0, b), -3.0 * t1))
    # spl_coef_conv
    h, yy = np.diff(x), -np.diff(y)
    b0, b1 = b[:-1], b[1:]
    cc = -(3.0 * yy + (2.0 * b0 + b1) * h) / (h * h)
    c1 = (3.0 * yy[-1] + (b0[-1] + 2.0 * b1[-1]) * h[-1]) / (h[-1] * h[-1])
    dd = (2.0 * yy / h + b0 + b1) / (h * h)
    return x, y, b, np.concatenate([cc, [c1]]), np.concatenate([dd, dd[-1:]])


def chebappxg(val, grid=None) -> Interpolant:
    """chebappxg.real, R/chebyshev.R:333-359: Chebyshev interpolation of values given on an
    arbitrary tensor grid; every coordinate is first mapped to the Chebyshev grid by a monotone
    (Hyman) cubic spline.  The reference applies the maps in R on every call (mapply per point
    column); here they run in the kernel prologue."""
    if grid is None:
        return chebappx(val)
    val = np.asarray(val, dtype=np.float64)
    if np.isscalar(grid[0]):
        grid = [grid]
    grid = [np.asarray(g, dtype=np.float64) for g in grid]
    shape = tuple(len(g) for g in grid)
    if int(np.prod(shape)) != val.size:
        raise ValueError("grid size must match data length")
    if val.shape != shape:
        val = val.reshape(shape, order="F")
    splines = [hyman_spline(g, kn) for g, kn in zip(grid, chebknots(shape))]
    return Interpolant("general", len(grid), [(g.min(), g.max()) for g in grid], coef=chebcoef(val), splines=splines)


def chebappxgf(fun, grid) -> Interpolant:
    """chebappxgf.real, R/chebyshev.R:367-371."""
    if np.isscalar(grid[0]):
        grid = [grid]
    return chebappxg(evalongrid(fun, grid=grid), grid)


def chebeval(x, coef, intervals=None, threads=None):
    """chebeval, R/chebyshev.R:126-130."""
    coef = np.asarray(coef, dtype=np.float64)
    ip = Interpolant("chebyshev", coef.ndim, None, coef=np.asfortranarray(coef))
    if intervals is not None:
        ip.t["mid"] = np.array([np.mean(iv) for iv in intervals])
        ip.t["ispan"] = np.array([2.0 / (iv[1] - iv[0]) for iv in intervals])
    try:
        return ip(x)
    finally:
        ip.close()


def mlappx(val, grid) -> Interpolant:
    """mlappx.real, R/multilinear.R:49-70."""
    if np.isscalar(grid[0]):
        grid = [grid]
    grid = [np.asarray(g, dtype=np.float64) for g in grid]
    if any(np.any(np.diff(g) < 0) for g in grid):
        if not callable(val):
            raise ValueError("Grid points must be ordered in increasing order")
        grid = [np.sort(g) for g in grid]
    if callable(val):
        val = evalongrid(val, grid=grid)
    val = np.asarray(val, dtype=np.float64)
    gl = int(np.prod([len(g) for g in grid]))
    if val.size != gl:
        raise ValueError(f"length of values {val.size} do not match size of grid {gl}")
    values = val.ravel(order="F") if val.ndim > 1 else val
    return Interpolant("multilinear", len(grid), [(g.min(), g.max()) for g in grid], grid=grid,
                       values=np.ascontiguousarray(values))


def fhappx(val, grid=None, d=1, fit: str = "host") -> Interpolant:
    """fhappx.real, R/floater-hormann.R:55-70.  fit="device": weights from chebpol_cuda_fhweights."""
    if grid is None:
        raise ValueError("Must specify grid")
    if np.isscalar(grid[0]):
        grid = [grid]
    grid = [np.asarray(g, dtype=np.float64) for g in grid]
    if callable(val):
        val = evalongrid(val, grid=grid)
    val = np.asarray(val, dtype=np.float64)
    shape = tuple(len(g) for g in grid)
    if val.shape != shape:
        val = val.reshape(shape, order="F")
    dd = np.resize(np.asarray(d, dtype=np.int64), len(grid))
    weights = _lib.fhweights(grid, dd) if fit == "device" else fhweights(grid, dd)
    return Interpolant("fh", len(grid), [(g.min(), g.max()) for g in grid], grid=grid,
                       values=np.asfortranarray(val), weights=weights)


_MYEPS = 1e3 * np.sqrt(np.finfo(float).eps)  # MYEPS, src/stalker.c:8


def _stalk_neighbours(val, grid, i):
    """Interior slices of dimension i: v0, vplus, vmin, kplus, kmin broadcast over the array."""
    D = val.ndim
    sl = lambda a, b: tuple(slice(a, b) if d == i else slice(None) for d in range(D))
    v0 = val[sl(1, -1)]
    g = np.asarray(grid[i], dtype=np.float64)
    shp = [1] * D
    shp[i] = g.size - 2
    kplus = (g[2:] - g[1:-1]).reshape(shp)
    kmin = (g[:-2] - g[1:-1]).reshape(shp)
    return sl, v0, val[sl(2, None)] - v0, val[sl(0, -2)] - v0, kplus, kmin


def _stalk_borders(out_b, val, grid, i):
    """Linear one-sided slopes on the two border hyperplanes (src/stalker.c:305-313, 481-490)."""
    D = val.ndim
    at = lambda k: tuple(k if d == i else slice(None) for d in range(D))
    g = grid[i]
    out_b[at(0)] = (val[at(1)] - val[at(0)]) / (g[1] - g[0])
    out_b[at(-1)] = (val[at(-2)] - val[at(-1)]) / (g[-2] - g[-1])


def makehyp(val, grid):
    """makehyp, src/stalker.c:446-535: tables a (prod(dims)) and b, c, d (rank x prod(dims)) of the
    hyperbolic stalker; c = NaN marks a parabola.  Host fit (runs once per interpolant)."""
    val = np.asarray(val, dtype=np.float64)
    D = val.ndim
    b = np.zeros((D,) + val.shape)
    c = np.zeros_like(b)
    d = np.zeros_like(b)
    with np.errstate(all="ignore"):
        for i in range(D):
            _stalk_borders(b[i], val, grid, i)
            if val.shape[i] < 3:
                continue
            sl, v0, vp, vm, kp, km = _stalk_neighbours(val, grid, i)
            Dd = km * vp - kp * vm
            D2 = km * km * vp - kp * kp * vm
            mono = (np.sign(vp * vm) <= 0) | (np.abs(vp) <= _MYEPS) | (np.abs(vm) <= _MYEPS)
            const = mono & (np.abs(vp - vm) <= _MYEPS * kp)
            lin = mono & ~const & (np.abs(Dd) <= _MYEPS * kp)
            gen = mono & ~const & ~lin
            par = ~mono & (np.abs(D2) <= _MYEPS * kp)
            hyp = ~mono & ~par
            bi = np.zeros(v0.shape)
            ci = np.zeros(v0.shape)
            di = np.zeros(v0.shape)
            kpb, kmb = np.broadcast_to(kp, v0.shape), np.broadcast_to(km, v0.shape)
            bi[lin] = (vp / kpb)[lin]
            ci[gen] = (-Dd / (kmb * kpb * (vp - vm)))[gen]
            di[gen] = (-vp * vm * (kmb - kpb) / Dd)[gen]
            bi[par] = (vp / (kpb * kpb))[par]
            ci[par] = np.nan
            bi[hyp] = (vp * vm * (kmb - kpb) / D2)[hyp]
            ci[hyp] = (-D2 / (kmb * kpb * Dd))[hyp]
            di[hyp] = (-kmb * vp * kpb * vm * (kmb - kpb) * Dd / (D2 * D2))[hyp]
            b[i][sl(1, -1)], c[i][sl(1, -1)], d[i][sl(1, -1)] = bi, ci, di
    a = np.zeros(val.shape)
    for i in range(D):
        a = a - d[i]
    flat = lambda t: np.asfortranarray(np.stack([t[i].ravel(order="F") for i in range(D)]))
    return a.ravel(order="F").copy(), flat(b), flat(c), flat(d)


def makestalk(val, grid):
    """makestalk, src/stalker.c:271-364, as built WITHOUT GSL: tables b, c, r (rank x prod(dims)).
    Without GSL the reference uses degree 2 on non-uniform grids and warns (R_makestalk :711-718)."""
    val = np.asarray(val, dtype=np.float64)
    D = val.ndim
    for g in grid:
        if np.any(np.abs(np.diff(g)[1:] - (g[1] - g[0])) > _MYEPS):
            warnings.warn("Non-uniform grid requires linking chebpol with GSL")
            break
    b = np.zeros((D,) + val.shape)
    c = np.zeros_like(b)
    r = np.zeros_like(b)
    with np.errstate(all="ignore"):
        for i in range(D):
            _stalk_borders(b[i], val, grid, i)
            if val.shape[i] < 3:
                continue
            sl, v0, vp, vm, kp, km = _stalk_neighbours(val, grid, i)
            kpb, kmb = np.broadcast_to(kp, v0.shape), np.broadcast_to(km, v0.shape)
            uniform = np.abs(kpb + kmb) < _MYEPS * kpb
            mono = np.sign(vp * vm) <= 0
            const = mono & (np.abs(vp - vm) <= _MYEPS * kpb)
            lin = mono & ~const & (np.abs(kmb * vp - kpb * vm) <= _MYEPS * kpb)
            ru = np.abs((vp - vm) / (vp + vm))
            ru = np.where(ru < 1, 1.0, ru)
            ru = np.where(np.isnan(ru) | (ru > 2.0), 2.0, ru)
            ri = np.where(mono & ~const & ~lin & uniform, ru, 2.0)
            q = np.power(-kpb / kmb, ri)
            b[i][sl(1, -1)] = (vp - vm * q) / (kpb - kmb * q)
            c[i][sl(1, -1)] = (kpb * vm - kmb * vp) / (kpb * np.power(-kmb, ri) - kmb * np.power(kpb, ri))
            r[i][sl(1, -1)] = ri
    flat = lambda t: np.asfortranarray(np.stack([t[i].ravel(order="F") for i in range(D)]))
    return flat(b), flat(c), flat(r)


def _stalk_front(val, grid):
    if np.isscalar(grid[0]):
        grid = [grid]
    grid = [np.asarray(g, dtype=np.float64) for g in grid]
    if any(np.any(np.diff(g) <= 0) for g in grid):
        if not callable(val):
            raise ValueError("Grid points must be ordered in increasing order")
        grid = [np.sort(g) for g in grid]
    if callable(val):
        val = evalongrid(val, grid=grid)
    val = np.asarray(val, dtype=np.float64)
    shape = tuple(len(g) for g in grid)
    if val.size != int(np.prod(shape)):
        raise ValueError(f"number of values ({val.size}) does not match grid size ({int(np.prod(shape))})")
    if val.shape != shape:
        val = val.reshape(shape, order="F")
    return val, grid


def hstalkerappx(val, grid) -> Interpolant:
    """hstalkerappx, R/stalker.R:53-69: hyperbolic stalker spline; evaluator chebpol_cuda_evalhyp."""
    val, grid = _stalk_front(val, grid)
    return Interpolant("hstalker", len(grid), [(g.min(), g.max()) for g in grid], grid=grid,
                       values=np.asfortranarray(val), tables=makehyp(val, grid))


def stalkerappx(val, grid) -> Interpolant:
    """stalkerappx, R/stalker.R:71-87: stalker spline; evaluator chebpol_cuda_evalstalk."""
    val, grid = _stalk_front(val, grid)
    return Interpolant("stalker", len(grid), [(g.min(), g.max()) for g in grid], grid=grid,
                       values=np.asfortranarray(val), tables=makestalk(val, grid))


def stalkercontext(grid, r):
    """stalkercontext, R/stalker.R:2-10, with the dmin / dplus vectors of oldstalkerappx (:26-27):
    (det, pmin, pplus), one vector per dimension.  R's `^` gives 1 for x^0 even when x is NaN."""
    det, pmin, pplus = [], [], []
    with np.errstate(invalid="ignore", divide="ignore"):
        for g, ri in zip(grid, r):
            dmin = np.concatenate([[np.nan], np.diff(g)])
            dplus = np.concatenate([dmin[1:], [np.nan]])
            pm = np.ones_like(dmin) if ri == 0 else dmin ** ri
            pp = np.ones_like(dplus) if ri == 0 else dplus ** ri
            det.append(1.0 / (dmin * pp + dplus * pm))
            pmin.append(pm)
            pplus.append(pp)
    return det, pmin, pplus


def oldstalkerappx(val, grid, r=2) -> Interpolant:
    """oldstalkerappx, R/stalker.R:11-49 (deprecated in the reference; ipol(method='oldstalker'))."""
    if np.isscalar(grid[0]):
        grid = [grid]
    grid = [np.asarray(g, dtype=np.float64) for g in grid]
    if any(np.any(np.diff(g) < 0) for g in grid):
        if not callable(val):
            raise ValueError("Grid points must be ordered in increasing order")
        grid = [np.sort(g) for g in grid]
    if callable(val):
        val = evalongrid(val, grid=grid)
    val = np.asarray(val, dtype=np.float64)
    gl = int(np.prod([len(g) for g in grid]))
    if val.size != gl:
        raise ValueError(f"length of values {val.size} do not match size of grid {gl}")
    deg = np.resize(np.asarray(r, dtype=np.float64), len(grid))
    uniform = np.array([bool(np.all(np.abs(np.diff(g, n=2)) < 1e-8)) for g in grid])
    if np.any(np.isnan(deg) & ~uniform):
        raise ValueError("Non-uniform grid and varying degree needs GSL. Recompile with GSL.")
    values = val.ravel(order="F") if val.ndim > 1 else val
    return Interpolant("oldstalker", len(grid), [(g.min(), g.max()) for g in grid], grid=grid,
                       values=np.ascontiguousarray(values), degree=deg, ctx=stalkercontext(grid, deg),
                       uniform=uniform.astype(np.int32))


def analyzesimplex(dtri, knots):
    """Host (numpy) counterpart of C_analyzesimplex, src/chebpol.c:830-875, vectorised over the
    simplices: the matrix [vertices; 1 ... 1] of every simplex LU-factorised with partial pivoting
    (first largest pivot, multipliers scaled by the pivot's reciprocal, as LAPACK dgetrf does) and
    the bounding boxes.  Returns (lumats, pivots, bbox) in the reference's layout."""
    knots = np.asarray(knots, dtype=np.float64)
    dtri = np.asarray(dtri)
    dim, N, ns = knots.shape[0], knots.shape[0] + 1, dtri.shape[1]
    V = knots[:, dtri - 1]  # dim x N x ns
    M = np.ones((ns, N, N))
    M[:, :dim, :] = V.transpose(2, 0, 1)  # row j, column v: coordinate j of vertex v; last row ones
    bbox = np.empty((ns, 2 * dim))
    bbox[:, 0::2] = V.min(axis=1).T
    bbox[:, 1::2] = V.max(axis=1).T
    piv = np.empty((ns, N), dtype=np.int32)
    idx = np.arange(ns)
    for j in range(N):
        jp = np.argmax(np.abs(M[:, j:, j]), axis=1) + j
        piv[:, j] = jp + 1
        rj, rp = M[idx, j, :].copy(), M[idx, jp, :].copy()
        M[idx, j, :], M[idx, jp, :] = rp, rj
        if j < N - 1:
            with np.errstate(divide="ignore", invalid="ignore"):
                r = 1.0 / M[:, j, j]
                M[:, j + 1:, j] *= r[:, None]
                M[:, j + 1:, j + 1:] += M[:, j + 1:, j][:, :, None] * (-M[:, j, j + 1:])[:, None, :]
    return np.ascontiguousarray(M.transpose(0, 2, 1)).reshape(-1), piv.reshape(-1), bbox.reshape(-1)


def slappx(val, knots, fit: str = "device") -> Interpolant:
    """slappx.real, R/simplexlinear.R:27-40: piecewise linear interpolation on the Delaunay
    triangulation of scattered knots (dim x nknots matrix, one knot per column).  The
    triangulation is qhull's, as in the reference (geometry::delaunayn(..., "Qt Pp") there,
    scipy.spatial.Delaunay here); the per-simplex LU factors and bounding boxes
    (C_analyzesimplex) are computed on the device."""
    from scipy.spatial import Delaunay

    knots = np.asarray(knots, dtype=np.float64)
    if knots.ndim == 1:
        knots = knots.reshape(1, -1)
    dim, nk = knots.shape
    if callable(val):
        val = np.array([val(knots[:, i].copy()) for i in range(nk)], dtype=np.float64)
    val = np.asarray(val, dtype=np.float64).ravel()
    if val.size != nk:
        raise ValueError(f"number of values ({val.size}) should equal number of knots ({nk})")
    if dim == 1:  # qhull needs >= 2 dimensions: consecutive sorted knots are the 1-simplices
        o = np.argsort(knots[0], kind="stable")
        simplices = np.stack([o[:-1], o[1:]], axis=1)
    else:
        simplices = Delaunay(knots.T, qhull_options="Qt Pp").simplices
    dtri = np.asfortranarray((simplices.T + 1).astype(np.int32))
    adata = _lib.analyzesimplex(dtri, knots) if fit == "device" else analyzesimplex(dtri, knots)
    return Interpolant("simplexlinear", dim, [(knots[d].min(), knots[d].max()) for d in range(dim)],
                       knots=np.asfortranarray(knots), dtri=dtri, adata=adata, values=val)


def _phifunc(d2: np.ndarray, k) -> np.ndarray:
    """R_phifunc, src/chebpol.c:886-921, on squared distances."""
    d2 = np.asarray(d2, dtype=np.float64)
    if k < 0:
        return np.exp(k * d2)
    ki = int(k)
    out = np.zeros_like(d2)
    pos = d2 > 0.0
    r = np.sqrt(d2[pos])
    out[pos] = r ** ki if ki % 2 == 1 else 0.5 * np.log(d2[pos]) * r ** ki
    return out


def polyh(val, knots, k=2, normalize=None, nowarn: bool = False) -> Interpolant:
    """polyh.real, R/polyharmonic.R:48-131: polyharmonic spline on scattered knots (columns of
    ``knots``).  The fit (a dense (N+M+1)-square solve, R's solve()/lm.fit) runs on the host with
    LAPACK as in R; the evaluator is chebpol_cuda_evalpolyh with the closure's normalisation
    x -> a*(x-b) applied inside the kernel."""
    knots = np.asarray(knots, dtype=np.float64)
    if knots.ndim == 1:
        knots = knots.reshape(1, -1)
    if callable(val):
        val = np.array([float(val(knots[:, i])) for i in range(knots.shape[1])])
    val = np.asarray(val, dtype=np.float64).ravel()
    M, N = knots.shape
    if k > 0:
        if abs(round(k) - k) > np.sqrt(np.finfo(float).eps):
            raise ValueError(f"k is positive, but not an integer: {k:.17f}")
        k = int(round(k))
    norm_a = norm_b = None
    if normalize is None:
        normalize = bool(knots.min() < 0 or knots.max() > 1)
    if not isinstance(normalize, (bool, np.bool_)):
        nm = np.asarray(normalize, dtype=np.float64)
        if nm.size == 2 and nm.ndim == 1:
            nm = nm.reshape(1, 2)
        if nm.ndim != 2 or nm.shape[1] != 2 or M % nm.shape[0] != 0:
            raise ValueError(f"normalize must but be a n x 2 matrix with {M} divisible by n({nm.shape[0]})")
        norm_a = np.resize(nm[:, 0], M)
        norm_b = np.resize(nm[:, 1], M)
        kn = knots  # as in R: with an explicit matrix the knots are NOT transformed (R/polyharmonic.R:72-78)
        normalize = True
    elif normalize:
        norm_a = 1.0 / (knots.max(axis=1) - knots.min(axis=1))
        norm_b = knots.min(axis=1)
        kn = norm_a[:, None] * (knots - norm_b[:, None])
    else:
        kn = knots
    diff = kn[:, :, None] - kn[:, None, :]
    A = _phifunc(np.einsum("kij,kij->ij", diff, diff), k)
    B = np.vstack([np.ones((1, N)), kn])
    mat = np.zeros((N + M + 1, N + M + 1))
    mat[:N, :N] = A
    mat[N:, :N] = B
    mat[:N, N:] = B.T
    rhs = np.concatenate([val, np.zeros(M + 1)])
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("error")
            wv = np.linalg.solve(mat, rhs)
        if not np.all(np.isfinite(wv)) or 1.0 / np.linalg.cond(mat, 1) < np.finfo(float).eps:
            raise np.linalg.LinAlgError("computationally singular")
    except (np.linalg.LinAlgError, Warning):
        if not nowarn:
            warnings.warn("Failed to fit exactly, fallback to least squares fit."
                          + ("" if normalize else " You could try normalize=TRUE."))
        wv = np.linalg.lstsq(mat, rhs, rcond=None)[0]
        wv[~np.isfinite(wv)] = 0.0
    return Interpolant("polyharmonic", M, [(r.min(), r.max()) for r in kn], knots=np.asfortranarray(kn),
                       weights=wv[:N].copy(), lweights=wv[N:].copy(), k=float(k), norm_a=norm_a, norm_b=norm_b)


def ucappx(val, intervals=None) -> Interpolant:
    """ucappx.real, R/uniform.R:58-77: Chebyshev series of the values with the
    per-coordinate map x -> sin(0.5*pi*x*(1-n)/n) (ugm, R/uniform.R:6) applied,
    after the optional interval map, on every call -- here inside the kernel."""
    val = np.asarray(val, dtype=np.float64)
    if val.ndim == 0:
        val = val.reshape(1)
    dims = val.shape
    cf = chebcoef(val)
    K = val.ndim
    if intervals is None:
        return Interpolant("uniform", K, [(-1.0, 1.0)] * K, coef=cf, ugm_n=np.array(dims, dtype=np.int32))
    if np.ndim(intervals) == 1 and len(intervals) == 2 and np.isscalar(intervals[0]):
        intervals = [intervals]
    mid = np.array([(float(iv[0]) + float(iv[1])) / 2.0 for iv in intervals])
    ispan = np.array([2.0 / (float(iv[1]) - float(iv[0])) for iv in intervals])
    return Interpolant("uniform", K, [tuple(iv) for iv in intervals], coef=cf, mid=mid, ispan=ispan,
                       ugm_n=np.array(dims, dtype=np.int32))


def ucappxf(fun, dims, intervals=None) -> Interpolant:
    """ucappxf.real, R/uniform.R:86-96."""
    dims = [int(d) for d in np.atleast_1d(dims)]
    if intervals is None:
        return ucappx(evalongrid(fun, grid=[np.linspace(-1, 1, d) for d in dims]))
    if np.ndim(intervals) == 1 and len(intervals) == 2 and np.isscalar(intervals[0]):
        intervals = [intervals]
    grid = [np.linspace(min(iv), max(iv), d) for d, iv in zip(dims, intervals)]
    return ucappx(evalongrid(fun, grid=grid), intervals)


_METHODS = ["chebyshev", "multilinear", "fh", "uniform", "general", "polyharmonic", "simplexlinear",
            "hstalker", "stalker", "crbf", "oldstalker"]
_ON_PATH = {"chebyshev", "multilinear", "fh", "uniform", "general", "polyharmonic", "simplexlinear", "stalker", "hstalker",
            "oldstalker"}


def ipol(val, dims=None, intervals=None, grid=None, knots=None, k=None, method="chebyshev", **kw) -> Interpolant:
    """ipol, R/ipol.R:108-239, for the four evaluators on the accelerated path."""
    cand = [m for m in _METHODS if m.startswith(method)]  # match.arg partial matching
    if method in _METHODS:
        cand = [method]
    if len(cand) != 1:
        raise ValueError(f"Unknown interpolation method: {method}")
    method = cand[0]
    if method not in _ON_PATH:
        raise NotImplementedError(
            f"method '{method}' is outside the GPU evaluation path (SURVEY.md section 2); use the R package")
    if method in ("stalker", "hstalker"):  # R/ipol.R:198-218
        if grid is None:
            raise ValueError("grid must be specified for stalker interpolation")
        if np.isscalar(grid[0]):
            grid = [grid]
        if any(_unsorted(np.asarray(g, dtype=np.float64)) for g in grid):
            raise ValueError("grid must be distinct ordered values")
        return (hstalkerappx if method == "hstalker" else stalkerappx)(val, grid)
    if method == "oldstalker":  # R/ipol.R:187-197
        if grid is None:
            raise ValueError("grid must be specified for stalker interpolation")
        if np.isscalar(grid[0]):
            grid = [grid]
        if any(_unsorted(np.asarray(g, dtype=np.float64)) for g in grid):
            raise ValueError("grid must be distinct ordered values")
        return oldstalkerappx(val, grid, r=2 if k is None else k)
    if method == "simplexlinear":  # R/ipol.R:131-142
        if knots is None:
            if grid is None:
                raise ValueError("knots must be specified for simplex linear interpolation")
            gl = [np.asarray(g, dtype=np.float64) for g in ([grid] if np.isscalar(grid[0]) else grid)]
            mesh = np.meshgrid(*gl, indexing="ij")
            knots = np.stack([m.ravel(order="F") for m in mesh])  # t(expand.grid(grid))
        return slappx(val, knots, fit=kw.get("fit", "device"))
    if method == "general":  # R/ipol.R:161-168
        if grid is None:
            raise ValueError("grid must be specified for general interpolation")
        if np.isscalar(grid[0]):
            grid = [grid]
        grid = [np.asarray(g, dtype=np.float64) for g in grid]
        if any(_unsorted(g) for g in grid):
            raise ValueError("grid must be distinct ordered values")
        return chebappxgf(val, grid) if callable(val) else chebappxg(val, grid)
    if method == "polyharmonic":  # R/ipol.R:169-186
        if knots is None:
            if grid is None:
                raise ValueError("Must specify knots for polyharmonic splines.")
            gl = [np.asarray(g, dtype=np.float64) for g in ([grid] if np.isscalar(grid[0]) else grid)]
            mesh = np.meshgrid(*gl, indexing="ij")
            knots = np.stack([m.ravel(order="F") for m in mesh])  # t(expand.grid(grid))
        return polyh(val, knots, 3 if k is None else k, kw.get("normalize"), kw.get("nowarn", False))
    if method == "chebyshev":
        if callable(val):
            if dims is None:
                raise ValueError("dims must be specified")
            return chebappxf(val, dims, intervals)
        val = np.asarray(val, dtype=np.float64)
        if val.ndim <= 1 and dims is not None:
            val = val.reshape(tuple(int(d) for d in np.atleast_1d(dims)), order="F")
        return chebappx(val, intervals)
    if method == "multilinear":
        if grid is None:
            raise ValueError("grid must be specified for multi linear interpolation")
        if np.isscalar(grid[0]):
            grid = [grid]
        grid = [np.asarray(g, dtype=np.float64) for g in grid]
        if any(_unsorted(g) for g in grid):
            raise ValueError("grid must be distinct ordered values")
        return mlappx(val, grid)
    if method == "fh":
        if grid is None:
            raise ValueError("grid must be specified for Floater-Hormann interpolation")
        if np.isscalar(grid[0]):
            grid = [grid]
        grid = [np.asarray(g, dtype=np.float64) for g in grid]
        if any(_unsorted(g) for g in grid):
            raise ValueError("grid must be distinct ordered values")
        if k is None:
            k = [min(4, len(g) - 1) for g in grid]
        return fhappx(val, grid, d=k)
    # uniform
    if callable(val) and dims is None:
        raise ValueError("Must specify dims for uniform intervals")
    if intervals is None and grid is not None:
        intervals = [(np.min(g), np.max(g)) for g in grid]
        warnings.warn("intervals constructed from ranges of grid-argument")
    if callable(val):
        return ucappxf(val, dims, intervals)
    val = np.asarray(val, dtype=np.float64)
    if val.ndim <= 1 and dims is not None:
        val = val.reshape(tuple(int(d) for d in np.atleast_1d(dims)), order="F")
    return ucappx(val, intervals)
