"""Seeded synthetic interpolants and point sets.  Test infrastructure."""
import numpy as np

from chebpol_b200 import interp


def gfun(x):
    """g(x) = exp(-sum x_d^2) * prod cos(d*x_d): smooth, decaying spectrum, max|g| ~ 1."""
    x = np.asarray(x, dtype=np.float64)
    d = np.arange(1, x.shape[-1] + 1)
    return np.exp(-np.sum(x * x, axis=-1)) * np.prod(np.cos(d * x), axis=-1)


def grid_values(grids):
    """g on the tensor grid, as an R (Fortran-ordered) array -- vectorised evalongrid."""
    mesh = np.meshgrid(*grids, indexing="ij")
    pts = np.stack(mesh, axis=-1)
    return np.asfortranarray(gfun(pts))


def points(npts, rank, seed=42, lo=-1.0, hi=1.0):
    rng = np.random.Generator(np.random.Philox(key=seed))
    return np.ascontiguousarray(lo + (hi - lo) * rng.random((npts, rank)))


def cheb_fitted(dims):
    """Chebyshev coefficients of g on the Chebyshev grid (host fit, chebpol_b200.chebcoef)."""
    grids = interp.chebknots(dims)
    return interp.chebcoef(grid_values(grids))


def cheb_stress(dims, seed=43):
    """c ~ U(-1,1)/(1+sum i_d): slowly decaying adversarial tensor."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    c = rng.uniform(-1, 1, size=tuple(dims))
    idx = np.indices(tuple(dims)).sum(axis=0)
    return np.asfortranarray(c / (1.0 + idx))


def mlip_case(dims, seed=44, uniform=True):
    rng = np.random.Generator(np.random.Philox(key=seed))
    if uniform:
        grid = [np.linspace(-1, 1, n) for n in dims]
    else:
        grid = [np.sort(np.concatenate([[-1.0, 1.0], rng.uniform(-1, 1, n - 2)])) for n in dims]
    values = np.asfortranarray(rng.random(tuple(dims)))
    return grid, values


def fh_case(dims, d=4, kind="uniform"):
    if kind == "uniform":
        grid = [np.linspace(-1, 1, n) for n in dims]
    elif kind == "mixed":  # as tests/all.R:100-101: uniform, cubed, decreasing Chebyshev
        grid = []
        for i, n in enumerate(dims):
            if i % 3 == 0:
                grid.append(np.linspace(-1, 1, n))
            elif i % 3 == 1:
                grid.append(np.linspace(-1, 1, n) ** 3)
            else:
                grid.append(interp.chebknots([n])[0])
    vals = grid_values(grid)
    dd = [min(d, n - 1) for n in dims]
    weights = interp.fhweights(grid, dd)
    return vals, grid, weights


def polyh_case(rank, nknots, k=3, seed=45):
    """Polyharmonic spline of gfun on `nknots` seeded random knots of [-1,1]^rank, fitted as
    polyh.real does (R/polyharmonic.R:48-131) with normalize=FALSE: (knots rank x nknots,
    weights, lweights)."""
    from chebpol_b200 import interp

    rng = np.random.default_rng(seed)
    knots = rng.uniform(-1.0, 1.0, (rank, nknots))
    ip = interp.polyh(lambda x: float(gfun(np.asarray(x).reshape(1, -1))[0]), knots, k=k, normalize=False)
    return ip.t["knots"], ip.t["weights"], ip.t["lweights"]


def sl_case(dim, nknots, seed=46):
    """Scattered knots in [-1,1]^dim, their Delaunay triangulation (qhull "Qt Pp", as
    R/simplexlinear.R:30) and the values of gfun on them: (knots dim x nknots, dtri (dim+1) x ns
    1-based int32, values)."""
    from scipy.spatial import Delaunay

    rng = np.random.default_rng(seed)
    knots = np.asfortranarray(rng.uniform(-1.0, 1.0, (dim, nknots)))
    if dim == 1:
        o = np.argsort(knots[0], kind="stable")
        simplices = np.stack([o[:-1], o[1:]], axis=1)
    else:
        simplices = Delaunay(knots.T, qhull_options="Qt Pp").simplices
    dtri = np.asfortranarray((simplices.T + 1).astype(np.int32))
    return knots, dtri, gfun(knots.T)
