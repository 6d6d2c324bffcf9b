"""GPU parity tests (run with -m gpu on a B200): libchebpol_cuda, called through its
C ABI, against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star: "at most 1e-12 relative to max|f|, or a few ULPs
accounting for summation order"):
  * STRICT kernels (reference operation order, no FMA): BIT-IDENTICAL to the oracle,
    which is itself bit-identical to the reference C (tests/test_oracle.py);
    where libm is involved (exp in blends 1/2, sin in the uniform map) 1e-14 relative.
  * fast kernels (basis contraction + FMA, pair-wise gathers): |diff| <= TOL_REL * scale,
    TOL_REL = 1e-12, scale = max|f| over the evaluated points for fitted interpolants;
    for the adversarial "stress" tensor (slowly decaying random coefficients, points
    outside the domain) scale = sum|c| * max_k|T_k(x)|, the natural conditioning bound.
"""
import os

import numpy as np
import pytest

import cases
import chebpol_b200 as cb
from chebpol_b200 import _lib

pytestmark = pytest.mark.gpu
TOL_REL = 1e-12


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def assert_close(got, want, tol_rel):
    """Finite results within tol_rel*max|finite want|; non-finite results (blends 1/2 overflow
    when extrapolating: exp(2 - 1/w) with w < 0) must be non-finite on both sides."""
    fin = np.isfinite(want)
    assert np.array_equal(fin, np.isfinite(got))
    assert np.array_equal(np.isnan(want), np.isnan(got))
    if fin.any():
        scale = np.abs(want[fin]).max()
        assert np.abs(got[fin] - want[fin]).max() <= tol_rel * scale, (np.abs(got[fin] - want[fin]).max(), scale)


def eval_plan(plan, x, kernel, variant=0):
    plan.set(_lib.OPT_KERNEL, kernel)
    plan.set(_lib.OPT_VARIANT, variant)
    out = plan.eval_host(np.ascontiguousarray(x))
    return out, plan.kernel_used


# ----------------------------------------------------------------- Chebyshev
CHEB_SHAPES = [(7,), (20, 20), (8, 7, 6), (10, 10, 10), (5, 4, 3, 2), (10,) * 5, (3, 3, 3, 3, 3, 3), (1, 4), (4, 1, 3),
               (2,) * 7, (33, 5), (12, 12, 12, 3), (17, 9), (31, 2, 3)]


@pytest.mark.parametrize("dims", CHEB_SHAPES)
def test_cheb_strict_bit_exact(gpu_ok, port, dims):
    cf = cases.cheb_stress(dims)
    x = cases.points(3001, len(dims), lo=-1.3, hi=1.3)
    plan = _lib.Plan.cheb(cf)
    got, k = eval_plan(plan, x, _lib.KERNEL_STRICT)
    assert k == "cheb_strict"
    want = port.evalcheb(cf, x, threads=4)
    assert np.array_equal(bits(got), bits(want))
    plan.close()


FAST = [("mma", _lib.VARIANT_MMA), ("slab", _lib.VARIANT_SLAB)]


def expected_kernel(family, dims):
    """DMMA engine when dims[0] <= 192 (the folded K), DFMA slab kernel when dims[0] <= 32,
    otherwise the strict kernel serves the shape (csrc/api.cu: launch_plan)."""
    if family == "mma":
        return "cheb_mma" if dims[0] <= 192 else "cheb_strict"
    return "cheb_slab" if dims[0] <= 32 else "cheb_strict"


@pytest.mark.parametrize("family,variant", FAST)
@pytest.mark.parametrize("dims", CHEB_SHAPES)
def test_cheb_fast_fitted(gpu_ok, port, dims, family, variant):
    cf = cases.cheb_fitted(dims)
    x = cases.points(20011, len(dims))
    plan = _lib.Plan.cheb(cf)
    got, k = eval_plan(plan, x, _lib.KERNEL_AUTO, variant)
    assert k == expected_kernel(family, dims)
    want = port.evalcheb(cf, x, threads=4)
    scale = np.abs(want).max()
    assert np.abs(got - want).max() <= TOL_REL * scale, (np.abs(got - want).max(), scale)
    plan.close()


@pytest.mark.parametrize("family,variant", FAST)
@pytest.mark.parametrize("dims", [(20, 20), (10,) * 5, (12, 12, 12, 3), (6, 5, 4, 3, 2, 2), (40, 7), (150,), (8, 7, 6)])
def test_cheb_fast_stress_outside_domain(gpu_ok, port, dims, family, variant):
    """The reference evaluates outside [-1,1] too (R/chebyshev.R:152-153; C_evalcheb's recurrence is
    valid for any real x).  Adversarial case: slowly decaying random coefficients, points up to
    |x| = 1.2.  The bar is north_star's: |diff| <= 1e-12 * max|f| with max|f| the largest |reference
    value| over the evaluated points -- NOT the conditioning product sum|c| * prod max|T_k|, which is
    10^3-10^4 times larger for (8,7,6); that product is printed beside the measured error."""
    cf = cases.cheb_stress(dims)
    x = cases.points(5003, len(dims), lo=-1.2, hi=1.2)
    plan = _lib.Plan.cheb(cf)
    got, k = eval_plan(plan, x, _lib.KERNEL_AUTO, variant)
    assert k == expected_kernel(family, dims)
    want = port.evalcheb(cf, x, threads=4)
    err, maxf = np.abs(got - want).max(), np.abs(want).max()
    cond = np.abs(cf).sum() * np.prod([np.cosh((n - 1) * np.arccosh(1.2)) for n in dims])
    print(f"stress {dims} {k}: |diff| = {err:.3e}, max|f| = {maxf:.3e}, |diff|/max|f| = {err / maxf:.2e}, "
          f"conditioning product = {cond:.3e}")
    assert err <= TOL_REL * maxf, (err, maxf, err / maxf)
    # and inside the domain alone, against max|f| there
    xi = cases.points(5003, len(dims), seed=7)
    gi, _ = eval_plan(plan, xi, _lib.KERNEL_AUTO, variant)
    wi = port.evalcheb(cf, xi, threads=4)
    assert np.abs(gi - wi).max() <= TOL_REL * np.abs(wi).max()
    plan.close()


@pytest.mark.parametrize("family,variant", FAST)
def test_cheb_12pow6_strided_sample(gpu_ok, port, family, variant):
    """BASELINE config 5's tensor (6-D, 12^6 = 2 985 984 coefficients): >= 1e5 strided points of a 1e7-point
    counter-based batch through each fast kernel.  The CPU oracle manages ~1e3 evaluations/s on this
    shape (3.3e6 recurrence steps each), so it checks the first 4000 sample points; the strict kernel is
    required to be BIT-IDENTICAL to the oracle on those and then stands in for it on all 1e5."""
    import torch

    from chebpol_b200 import synth

    dims = (12,) * 6
    cf = cases.cheb_fitted(dims)
    n, stride = 100_000, 100
    idx = np.arange(n, dtype=np.uint64) * np.uint64(stride)
    x = synth.points_at(idx, 6)
    plan = _lib.Plan.cheb(cf)
    got, k = eval_plan(plan, x, _lib.KERNEL_AUTO, variant)
    assert k == ("cheb_mma" if family == "mma" else "cheb_slab")
    strict, ks = eval_plan(plan, x, _lib.KERNEL_STRICT)
    assert ks == "cheb_strict"
    n_orc = 4000
    want = port.evalcheb(cf, x[:n_orc], threads=os.cpu_count() or 4)
    assert np.array_equal(bits(strict[:n_orc]), bits(want))
    maxf = np.abs(strict).max()
    err = np.abs(got - strict).max()
    print(f"12^6 {k}: |diff| = {err:.3e} on {n} points, max|f| = {maxf:.3e}, ratio {err / maxf:.2e}; Kahan-free checksum "
          f"{int(bits(got).astype(np.uint64).sum(dtype=np.uint64)):016x}")
    assert err <= TOL_REL * maxf, (err, maxf)
    assert np.abs(got[:n_orc] - want).max() <= TOL_REL * np.abs(want).max()
    plan.close()


def test_cheb_auto_kernel_choice(gpu_ok, port):
    """AUTO picks the DFMA slab kernel for small tensors, the DMMA engine for large ones, and the
    strict kernel for small tensors the slab kernel cannot take (csrc/api.cu launch_plan;
    crossover measured by tools/crossover.py)."""
    for dims, want in [((20, 20), "cheb_slab"), ((8, 7, 6), "cheb_slab"), ((12,) * 4, "cheb_slab"),
                       ((16,) * 4, "cheb_mma"), ((10,) * 5, "cheb_mma"), ((8,) * 5, "cheb_mma"), ((6,) * 6, "cheb_mma"),
                       ((40, 7), "cheb_strict"), ((40, 40, 9), "cheb_mma")]:
        cf = cases.cheb_fitted(dims)
        x = cases.points(3000, len(dims))
        plan = _lib.Plan.cheb(cf)
        got, k = eval_plan(plan, x, _lib.KERNEL_AUTO)
        assert k == want, (dims, k)
        ref = port.evalcheb(cf, x, threads=4)
        assert np.abs(got - ref).max() <= TOL_REL * np.abs(ref).max()
        plan.close()


def test_cheb_intervals_and_uniform_premaps(gpu_ok, port):
    dims = (9, 8, 7)
    iv = [(1.0, 2.0), (1.0, 4.0), (-3.0, 3.0)]
    grids = cb.chebknots(dims, iv)
    cf = cb.chebcoef(cases.grid_values(grids))
    mid = np.array([np.mean(i) for i in iv])
    ispan = np.array([2.0 / (i[1] - i[0]) for i in iv])
    rng = np.random.default_rng(3)
    x = np.ascontiguousarray(np.stack([rng.uniform(a, b, 4000) for a, b in iv], axis=1))
    for kernel, exact in [(_lib.KERNEL_STRICT, True), (_lib.KERNEL_AUTO, False)]:
        plan = _lib.Plan.cheb(cf, mid=mid, ispan=ispan)
        got, _ = eval_plan(plan, x, kernel)
        want = port.evalcheb(cf, port.imap(x, mid, ispan))
        if exact:
            assert np.array_equal(bits(got), bits(want))
        else:
            assert np.abs(got - want).max() <= TOL_REL * np.abs(want).max()
        plan.close()
        # uniform evaluator: ugm after the interval map (R/uniform.R:68)
        plan = _lib.Plan.cheb(cf, mid=mid, ispan=ispan, ugm_n=dims)
        got, _ = eval_plan(plan, x, kernel)
        want = port.evalcheb(cf, port.ugm(x, dims, mid=mid, ispan=ispan))
        # device sin vs glibc sin: <= 1 ulp in the mapped coordinate
        assert np.abs(got - want).max() <= 1e-13 * np.abs(want).max()
        plan.close()
        plan = _lib.Plan.cheb(cf, ugm_n=dims)
        xu = cases.points(4000, 3)
        got, _ = eval_plan(plan, xu, kernel)
        want = port.evalcheb(cf, port.ugm(xu, dims))
        assert np.abs(got - want).max() <= 1e-13 * np.abs(want).max()
        plan.close()


def test_cheb_long_1d_golden(gpu_ok):
    """tests/all.R:10-13 -> all.Rout.save:57 through the GPU path (n = 50000 falls to the strict kernel)."""
    n = 50000
    kn = cb.chebknots([n])[0]
    with np.errstate(divide="ignore", invalid="ignore"):
        val = np.where(kn == 0, 0.0, np.sin(1.0 / kn))
    ch = cb.ipol(val, dims=n, method="chebyshev")
    got = ch(np.array([0.03, 0.031]))
    assert abs(got[0] - 0.9400964) < 5e-8 and abs(got[1] - 0.7465345) < 5e-8


def test_cheb_goldens_through_ipol(gpu_ok):
    """all.Rout.save:69,71 and :74 via the host mirror of the R interface."""
    f = lambda x: np.exp(-np.sum(x ** 2))
    ch = cb.ipol(f, dims=[8, 7, 6], intervals=[(1, 2), (1, 4), (1, 3)], method="chebyshev")
    for kern in (_lib.KERNEL_AUTO, _lib.KERNEL_STRICT):
        ch.kernel = kern
        s = np.array([1.4, 2.3, 1.9])
        assert abs((ch(s)[0] - f(s)) - (-3.115592e-06)) < 5e-13
        a = np.array([1.01, 1.01, 1.01])
        assert abs((ch(a)[0] - f(a)) - 0.0001730159) < 5e-11
    ch.close()


# --------------------------------------------------------------- multilinear
@pytest.mark.parametrize("blend", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("dims", [(10,), (7, 9), (11, 11, 11), (16, 16, 16, 16), (5, 4, 3, 6, 4), (3, 4, 3, 2, 5, 3), (3,) * 8])
def test_mlip_strict(gpu_ok, port, dims, blend):
    grid, values = cases.mlip_case(dims, uniform=False)
    # blends 1/2 are exp(2 - 1/w): outside the grid (w < 0) they overflow and the result is
    # numerically meaningless in the reference too, so those two are compared inside the domain
    ext = 1.0 if blend in (1, 2) else 1.2
    x = cases.points(4001, len(dims), lo=-ext, hi=ext)
    plan = _lib.Plan.mlip(grid, values, blend)
    got, k = eval_plan(plan, x, _lib.KERNEL_STRICT)
    assert k == "mlip_strict"
    want = port.evalmlip(grid, values, x, blend=blend, threads=4)
    if blend in (1, 2):  # device exp vs glibc exp
        assert_close(got, want, 1e-14)
    else:
        assert np.array_equal(bits(got), bits(want))
    plan.close()


@pytest.mark.parametrize("family,variant", [("mlip_pair", _lib.VARIANT_MLIP_PAIR), ("mlip_cell", _lib.VARIANT_MLIP_CELL),
                                            ("mlip_quad", _lib.VARIANT_MLIP_QUAD), ("mlip_bin", _lib.VARIANT_MLIP_BIN)])
@pytest.mark.parametrize("blend", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("dims", [(10,), (7, 9), (11, 11, 11), (16, 16, 16, 16), (5, 4, 3, 6, 4), (3, 4, 3, 2, 5, 3)])
def test_mlip_fast(gpu_ok, port, dims, blend, family, variant):
    grid, values = cases.mlip_case(dims, uniform=(len(dims) % 2 == 0))
    # linear blending extrapolates linearly (src/chebpol.c:628-629) and is tested outside the
    # grid; the other blends are only well conditioned for w in [0,1] (outside, the corner
    # weights change sign and sum(cw) can cancel to ~0 before the division)
    ext = 1.2 if blend == 0 else 1.0
    x = cases.points(20001, len(dims), lo=-ext, hi=ext)
    # points exactly on knots and on the domain edge
    x[:50, 0] = np.resize(grid[0], 50)
    plan = _lib.Plan.mlip(grid, values, blend)
    got, k = eval_plan(plan, x, _lib.KERNEL_AUTO, variant)
    if family in ("mlip_quad", "mlip_bin") and len(dims) == 1:
        assert k == "mlip_strict"  # both need a second dimension (rows to interleave / upper cells to bin by)
    else:
        assert k == family
    want = port.evalmlip(grid, values, x, blend=blend, threads=4)
    assert_close(got, want, TOL_REL)
    plan.close()


def test_mlip_auto_picks_cell_layout_when_amortised(gpu_ok, port):
    """AUTO expands the table to the cell-major layout once a batch is large enough (>= cells/8)
    and keeps using it; small batches before that use the pair-gather kernel."""
    grid, values = cases.mlip_case((16, 16, 16, 16), uniform=False)
    plan = _lib.Plan.mlip(grid, values, 0)
    small = cases.points(1000, 4)
    _, k = eval_plan(plan, small, _lib.KERNEL_AUTO)
    assert k == "mlip_pair"
    big = cases.points(30000, 4, lo=-1.2, hi=1.2)
    got, k = eval_plan(plan, big, _lib.KERNEL_AUTO)
    assert k == "mlip_cell"
    assert_close(got, port.evalmlip(grid, values, big, threads=4), TOL_REL)
    _, k = eval_plan(plan, small, _lib.KERNEL_AUTO)
    assert k == "mlip_cell"
    plan.close()


def test_mlip_auto_uses_quads_when_cells_do_not_fit(gpu_ok, port, monkeypatch):
    """AUTO without the cell-major table (here: capped away, as for a 64^5-class table whose 2^D-fold
    expansion does not fit): small batches gather pairs from the values as given, a batch of at least
    half as many points as the table has values builds the row-pair-interleaved table (2x) and stays on it."""
    monkeypatch.setenv("CHEBPOL_CUDA_CELL_MAX_GB", "0")
    grid, values = cases.mlip_case((12, 11, 10, 9), uniform=False)
    plan = _lib.Plan.mlip(grid, values, 0)
    small = cases.points(1000, 4)
    _, k = eval_plan(plan, small, _lib.KERNEL_AUTO)
    assert k == "mlip_pair"
    b0 = plan.get(_lib.GET_TENSOR_BYTES)
    big = cases.points(8000, 4, lo=-1.2, hi=1.2)
    big[:12, 0] = grid[0]
    got, k = eval_plan(plan, big, _lib.KERNEL_AUTO)
    assert k == "mlip_quad"
    assert plan.get(_lib.GET_TENSOR_BYTES) - b0 == 2 * 12 * 10 * 10 * 9 * 8  # 2 * n0 * (n1 - 1) * rest doubles
    assert_close(got, port.evalmlip(grid, values, big, threads=4), TOL_REL)
    _, k = eval_plan(plan, small, _lib.KERNEL_AUTO)
    assert k == "mlip_quad"
    nan = small.copy()
    nan[3, 2] = np.nan
    out, _ = eval_plan(plan, nan, _lib.KERNEL_AUTO)
    assert np.isnan(out[3]) and np.isfinite(np.delete(out, 3)).all()
    plan.close()


def test_mlip_binned_path_handles_clustered_and_ragged_batches(gpu_ok, port):
    """The binned path (points reordered by upper-dimension cell, sub-tables staged in shared memory):
    uniform batches, a batch clustered in one cell (every bin but one empty, most points overflow the
    bin capacity and take the gather kernel), points on knots / outside the grid / NaN, batch sizes
    around the bin count, and results restored to the caller's order."""
    grid, values = cases.mlip_case((12, 10, 9, 8), uniform=False)
    plan = _lib.Plan.mlip(grid, values, 0)
    plan.set(_lib.OPT_VARIANT, _lib.VARIANT_MLIP_BIN)
    rng = np.random.default_rng(11)
    for n in (1, 7, 500, 4001, 60_000):
        x = cases.points(n, 4, seed=n, lo=-1.2, hi=1.2)
        x[: min(n, 12), 0] = grid[0][: min(n, 12)]
        got = plan.eval_host(x)
        assert plan.kernel_used == "mlip_bin"
        assert_close(got, port.evalmlip(grid, values, x, threads=4), TOL_REL)
    cl = np.ascontiguousarray(np.array([0.31, -0.42, 0.11, 0.73]) + 1e-3 * rng.standard_normal((50_000, 4)))
    cl[17, 1] = np.nan
    got = plan.eval_host(cl)
    want = port.evalmlip(grid, values, np.where(np.isnan(cl), 0.0, cl), threads=4)
    assert np.isnan(got[17]) and np.abs(np.delete(got - want, 17)).max() <= TOL_REL
    for blend in (3, 5):
        plan.set(_lib.OPT_BLEND, blend)
        x = cases.points(30_000, 4, seed=blend)
        assert_close(plan.eval_host(x), port.evalmlip(grid, values, x, blend=blend, threads=4), TOL_REL)
    plan.close()


@pytest.mark.parametrize("dims,m", [((9, 8, 7, 6), 3), ((9, 8, 7, 6), 2), ((6, 5, 4, 5, 4), 3), ((6, 5, 4, 5, 4), 4), ((4, 3, 4, 3, 3, 3), 3),
                                    ((7, 6, 5), 1), ((7, 6, 5), 2)])
def test_mlip_bricks_partial_cell_major(gpu_ok, port, monkeypatch, dims, m):
    """Bricks: only the first m dimensions expanded (2^m x the table instead of 2^D x), a point reads
    2^(D-m) bricks.  Same kernel, same results as the full cell-major table."""
    monkeypatch.setenv("CHEBPOL_CUDA_CELL_M", str(m))
    grid, values = cases.mlip_case(dims, uniform=False)
    x = cases.points(20_011, len(dims), lo=-1.2, hi=1.2)
    x[:40, 0] = np.resize(grid[0], 40)
    x[40:80, -1] = np.resize(grid[-1], 40)
    plan = _lib.Plan.mlip(grid, values, 0)
    b0 = plan.get(_lib.GET_TENSOR_BYTES)
    got, k = eval_plan(plan, x, _lib.KERNEL_AUTO, _lib.VARIANT_MLIP_CELL)
    assert k == "mlip_cell"
    bricks = int(np.prod([n - 1 for n in dims[:m]]) * np.prod(dims[m:], dtype=np.int64))
    assert plan.get(_lib.GET_TENSOR_BYTES) - b0 == bricks * 2 ** m * 8
    assert_close(got, port.evalmlip(grid, values, x, threads=4), TOL_REL)
    for blend in (1, 3, 4):
        xi = cases.points(5000, len(dims), seed=blend)
        plan.set(_lib.OPT_BLEND, blend)
        gi, _ = eval_plan(plan, xi, _lib.KERNEL_AUTO, _lib.VARIANT_MLIP_CELL)
        assert_close(gi, port.evalmlip(grid, values, xi, blend=blend, threads=4), TOL_REL)
    plan.close()


def test_mlip_golden_and_blend_names(gpu_ok):
    """chebpol-Ex.Rout.save:451-452: ml1(su) - exp(su)."""
    import math

    su = np.array([i * (1.0 / 9.0) for i in range(10)])
    su[-1] = 1.0
    s = np.array([math.pow(v, 3.0) for v in su])
    ml1 = cb.ipol(np.array([math.exp(v) for v in s]), grid=[s], method="multilin")
    want = [0.0, 0.0007963441, 0.0023666771, 0.0036681447, 0.0028930495, 0.0111173403, 0.0064666457, 0.0191999204,
            0.0246303860, 0.0]
    assert np.allclose(ml1(su) - np.exp(su), want, rtol=0, atol=6e-11)
    for name in cb.interp.BLENDS:
        assert np.all(np.isfinite(ml1(su, blend=name)))
    ml1.close()


def test_mlip_nan_terminates(gpu_ok):
    grid, values = cases.mlip_case((6, 5))
    for kern in (_lib.KERNEL_AUTO, _lib.KERNEL_STRICT):
        plan = _lib.Plan.mlip(grid, values, 0)
        x = np.array([[np.nan, 0.1], [0.2, np.nan], [0.3, 0.4]])
        got, _ = eval_plan(plan, x, kern)
        assert np.isnan(got[0]) and np.isnan(got[1]) and np.isfinite(got[2])
        plan.close()


# ------------------------------------------------------------ Floater-Hormann
@pytest.mark.parametrize("dims,kind,d", [((10,), "uniform", 4), ((12, 9), "mixed", 3), ((10, 8, 7), "mixed", 3),
                                         ((50, 30, 20), "mixed", 4), ((5, 4, 3, 4), "uniform", 2), ((40, 40, 40), "uniform", 4)])
def test_fh_strict_bit_exact(gpu_ok, port, dims, kind, d):
    vals, grid, weights = cases.fh_case(dims, d=d, kind=kind)
    x = cases.points(1501, len(dims), lo=-1.05, hi=1.05)
    for j in range(60):  # exact and near knot hits (absolute 10*eps test, src/chebpol.c:195-197)
        dd = j % len(dims)
        x[j, dd] = grid[dd][j % len(grid[dd])] + (0.0, 1e-16, -1e-15, 3e-15)[j % 4]
    plan = _lib.Plan.fh(vals, grid, weights)
    got, k = eval_plan(plan, x, _lib.KERNEL_STRICT)
    assert k == "fh_strict"
    want = port.fh(vals, grid, weights, x, threads=4)
    assert np.array_equal(bits(got), bits(want))
    plan.close()


@pytest.mark.parametrize("dims,kind,d", [((10,), "uniform", 4), ((12, 9), "mixed", 3), ((10, 8, 7), "mixed", 3),
                                         ((50, 30, 20), "mixed", 4), ((5, 4, 3, 4), "uniform", 2), ((40, 40, 40), "uniform", 4),
                                         ((150, 3), "uniform", 4), ((3, 3, 3, 3, 3), "uniform", 2)])
def test_fh_fast(gpu_ok, port, dims, kind, d):
    """DMMA engine with the barycentric basis against the reference's nested num/den recursion.

    Bar: |diff| <= 1e-12*max|f| + 16*eps*Lambda(x)*max|v| pointwise, where
    Lambda(x) = prod_d sum_i |b_{d,i}(x_d)| is the Lebesgue function of the interpolant.
    The second term is the "few ULPs accounting for summation order" of the north_star scaled
    by the conditioning: FH weights alternate in sign, and on irregular grids (cubed, or 5%
    outside the domain) Lambda reaches 1e4, where the REFERENCE ITSELF is only accurate to
    ~2e-12 against exact arithmetic (measured with mpmath; DESIGN.md, parity section)."""
    vals, grid, weights = cases.fh_case(dims, d=d, kind=kind)
    x = cases.points(6007, len(dims), lo=-1.05, hi=1.05)
    for j in range(120):  # exact hits, sub-threshold and just-above-threshold distances to knots
        dd = j % len(dims)
        x[j, dd] = grid[dd][j % len(grid[dd])] + (0.0, 1e-16, -1e-15, 3e-15, 1e-13, -1e-11)[j % 6]
    plan = _lib.Plan.fh(vals, grid, weights)
    got, k = eval_plan(plan, x, _lib.KERNEL_AUTO)
    assert k == "fh_mma"
    want = port.fh(vals, grid, weights, x, threads=4)
    lam = np.ones(x.shape[0])
    for dd in range(len(dims)):
        diff = x[:, dd:dd + 1] - grid[dd][None, :]
        hit = np.abs(diff) < 10 * np.finfo(float).eps
        with np.errstate(divide="ignore", invalid="ignore"):
            p = weights[dd][None, :] / diff
            b = np.abs(p) / np.abs(p.sum(axis=1, keepdims=True))
        lam *= np.where(hit.any(axis=1), 1.0, b.sum(axis=1))
    tol = TOL_REL * np.abs(want).max() + 16 * np.finfo(float).eps * lam * np.abs(vals).max()
    assert np.all(np.abs(got - want) <= tol), (np.abs(got - want).max(), lam.max())
    # and inside the domain, away from the ill-conditioned fringe, the plain bar holds
    inside = np.all(np.abs(x) <= 1.0, axis=1) & (lam < 50)
    assert np.abs(got - want)[inside].max() <= TOL_REL * np.abs(want).max()
    plan.close()


def test_fh_long_grid_falls_back_to_strict(gpu_ok, port):
    """dims[0] > 192 exceeds the folded K of the DMMA engine: AUTO must serve it with the strict kernel."""
    vals, grid, weights = cases.fh_case((200, 3), d=4)
    x = cases.points(3000, 2)
    plan = _lib.Plan.fh(vals, grid, weights)
    got, k = eval_plan(plan, x, _lib.KERNEL_AUTO)
    assert k == "fh_strict"
    assert np.array_equal(bits(got), bits(port.fh(vals, grid, weights, x, threads=4)))
    plan.close()


def test_fh_fast_far_extrapolation_and_nan(gpu_ok, port):
    """Far outside the grid sum_i w_i/(x-k_i) cancels catastrophically (the w_i alternate and
    nearly sum to zero), so neither the reference nor any reordering of it has correct digits
    there; the requirement is only: finite where the reference is finite, NaN in -> NaN out,
    and -- mildly outside (Lambda still moderate) -- agreement within the conditioning bound."""
    vals, grid, weights = cases.fh_case((30, 20), d=4, kind="uniform")
    x = np.array([[50.0, 0.3], [-1e6, 2.0], [1e200, 0.0], [0.2, -3e3], [np.nan, 0.1], [0.1, np.nan], [1.02, -1.01]])
    plan = _lib.Plan.fh(vals, grid, weights)
    got, k = eval_plan(plan, x, _lib.KERNEL_AUTO)
    assert k == "fh_mma"
    want = port.fh(vals, grid, weights, x)
    assert np.isnan(got[4]) and np.isnan(got[5])
    fin = np.isfinite(want[:4])
    assert np.all(np.isfinite(got[:4][fin]))
    assert abs(got[6] - want[6]) <= 1e-9 * max(1.0, abs(want[6]))
    plan.close()


def test_fh_goldens_through_ipol(gpu_ok):
    """chebpol-Ex.Rout.save:416-424 (set.seed(1)): ip(c(0.3,0.8)) and six more values."""
    from r_rng import RRng

    r = RRng(1)
    g = np.array([i * (1.0 / 9.0) for i in range(10)])
    g[-1] = 1.0
    val = r.runif(100).reshape((10, 10), order="F")
    ip = cb.ipol(val, grid=[g, g], method="fh")
    assert abs(ip(np.array([0.3, 0.8]))[0] - 0.2284498) < 5e-8
    x = r.runif(12).reshape((2, 6), order="F")
    got = ip(x, threads=2)
    assert np.allclose(got, [0.74713282, 0.83340490, 0.00568654, 0.70100518, 0.39261432, 0.71810846], rtol=0, atol=5e-9)
    ip.close()


# ----------------------------------------------------- boundary / host pipeline
def test_stateless_entry_points_and_chunking(gpu_ok, port):
    dims = (10, 10, 10)
    cf = cases.cheb_fitted(dims)
    x = cases.points(300007, 3)
    want = port.evalcheb(cf, x, threads=8)
    got = _lib.evalcheb(cf, x, ngpus=1)
    assert np.abs(got - want).max() <= TOL_REL * np.abs(want).max()
    # many small chunks through the double-buffered pipeline, odd tail
    plan = _lib.Plan.cheb(cf)
    plan.set(_lib.OPT_CHUNK_POINTS, 4099)
    got2 = plan.eval_host(x)
    assert np.array_equal(bits(got), bits(got2))
    assert plan.get(_lib.GET_LAUNCHES) == (300007 + 4098) // 4099
    plan.close()
    grid, values = cases.mlip_case((9, 8, 7))
    xm = cases.points(100003, 3)
    assert np.abs(_lib.evalmlip(grid, values, xm, 0, 1) - port.evalmlip(grid, values, xm, threads=8)).max() <= 1e-14
    vals, g, w = cases.fh_case((9, 8, 7))
    assert np.array_equal(bits(_lib.fh(vals, g, w, xm[:5000], 1)), bits(port.fh(vals, g, w, xm[:5000], threads=8))) or \
        np.abs(_lib.fh(vals, g, w, xm[:5000], 1) - port.fh(vals, g, w, xm[:5000], threads=8)).max() <= 1e-12


def test_empty_and_single_point(gpu_ok, port):
    cf = cases.cheb_fitted((6, 5))
    plan = _lib.Plan.cheb(cf)
    assert plan.eval_host(np.empty((0, 2))).shape == (0,)
    one = plan.eval_host(np.array([[0.25, -0.5]]))
    assert abs(one[0] - port.evalcheb(cf, np.array([[0.25, -0.5]]))[0]) <= 1e-14
    plan.close()
    ip = cb.chebappx(cases.grid_values(cb.chebknots([6, 5])))
    assert ip(np.array([0.25, -0.5])).shape == (1,)
    with pytest.raises(ValueError):
        ip(np.array([0.1, 0.2, 0.3]))
    with pytest.warns(UserWarning):
        assert ip(np.array([0.1, 0.2, 0.3, 0.4])).shape == (2,)
    ip.close()


def test_device_resident_torch(gpu_ok, port):
    import torch

    dims = (10,) * 5
    cf = cases.cheb_fitted(dims)
    x = cases.points(50000, 5)
    plan = _lib.Plan.cheb(cf)
    xd = torch.from_numpy(x).cuda()
    out = plan.eval_torch(xd)
    torch.cuda.synchronize()
    want = port.evalcheb(cf, x, threads=8)
    assert np.abs(out.cpu().numpy() - want).max() <= TOL_REL * np.abs(want).max()
    assert plan.kernel_used == "cheb_mma"
    plan.close()


def test_full_size_properties(gpu_ok, port):
    """BASELINE-size shapes: properties that need no full CPU pass.
    cfg2 tensor (10^5), 2e6 points: (a) partition independence -- evaluating two halves
    equals evaluating the whole, bitwise; (b) linearity in the coefficients;
    (c) a strided sample of 2000 points against the oracle."""
    import torch

    dims = (10,) * 5
    cf = cases.cheb_fitted(dims)
    g = torch.Generator(device="cuda").manual_seed(42)
    xd = torch.rand((2_000_000, 5), generator=g, device="cuda", dtype=torch.float64) * 2 - 1
    plan = _lib.Plan.cheb(cf)
    whole = plan.eval_torch(xd)
    a = plan.eval_torch(xd[:777_777].contiguous())
    b = plan.eval_torch(xd[777_777:].contiguous())
    torch.cuda.synchronize()
    assert torch.equal(whole, torch.cat([a, b]))
    plan2 = _lib.Plan.cheb(2.0 * cf)
    twice = plan2.eval_torch(xd)
    torch.cuda.synchronize()
    assert torch.equal(twice, 2.0 * whole)  # scaling by 2 is exact in binary FP
    idx = torch.arange(0, 2_000_000, 1000, device="cuda")
    want = port.evalcheb(cf, xd[idx].cpu().numpy(), threads=8)
    assert np.abs(whole[idx].cpu().numpy() - want).max() <= TOL_REL * np.abs(want).max()
    plan.close()
    plan2.close()


# ------------------------------------------------- full-size shapes, size-independent properties
def test_full_size_multilinear_properties(gpu_ok, port):
    """BASELINE config 3 table (64^4, cell-major 2 GB), 4e6 points: (a) partition independence,
    bitwise; (b) exact reproduction at the knots (interpolation property); (c) affine
    invariance: scaling the table by 2 and adding nothing doubles the result exactly;
    (d) a strided sample against the oracle."""
    import torch

    dims = (64,) * 4
    grid, values = cases.mlip_case(dims)
    plan = _lib.Plan.mlip(grid, values, 0)
    g = torch.Generator(device="cuda").manual_seed(7)
    xd = torch.rand((4_000_000, 4), generator=g, device="cuda", dtype=torch.float64) * 2 - 1
    whole = plan.eval_torch(xd)
    assert plan.kernel_used == "mlip_cell"
    a = plan.eval_torch(xd[:1_234_567].contiguous())
    b = plan.eval_torch(xd[1_234_567:].contiguous())
    torch.cuda.synchronize()
    assert torch.equal(whole, torch.cat([a, b]))
    # knots: random grid nodes must return the table value exactly (weights are exactly 0/1)
    rng = np.random.default_rng(1)
    idx = rng.integers(0, 64, size=(5000, 4))
    xk = np.stack([grid[d][idx[:, d]] for d in range(4)], axis=1)
    got = plan.eval_host(np.ascontiguousarray(xk))
    assert np.array_equal(got, values[idx[:, 0], idx[:, 1], idx[:, 2], idx[:, 3]])
    plan2 = _lib.Plan.mlip(grid, 2.0 * values, 0)
    twice = plan2.eval_torch(xd)
    torch.cuda.synchronize()
    assert torch.equal(twice, 2.0 * whole)
    sel = torch.arange(0, 4_000_000, 2000, device="cuda")
    want = port.evalmlip(grid, values, xd[sel].cpu().numpy(), threads=8)
    assert np.abs(whole[sel].cpu().numpy() - want).max() <= TOL_REL * np.abs(want).max()
    plan.close()
    plan2.close()


def test_full_size_fh_properties(gpu_ok, port):
    """BASELINE config 4 (40^3, d=4), 1e6 points: partition independence, reproduction of the
    values at the knots (one-hot basis on a knot hit, src/chebpol.c:195-197), constants are
    reproduced (barycentric weights sum to one), strided sample against the oracle."""
    import torch

    vals, grid, weights = cases.fh_case((40, 40, 40), d=4)
    plan = _lib.Plan.fh(vals, grid, weights)
    g = torch.Generator(device="cuda").manual_seed(8)
    xd = torch.rand((1_000_000, 3), generator=g, device="cuda", dtype=torch.float64) * 2 - 1
    whole = plan.eval_torch(xd)
    assert plan.kernel_used == "fh_mma"
    a = plan.eval_torch(xd[:333_333].contiguous())
    b = plan.eval_torch(xd[333_333:].contiguous())
    torch.cuda.synchronize()
    assert torch.equal(whole, torch.cat([a, b]))
    rng = np.random.default_rng(2)
    idx = rng.integers(0, 40, size=(3000, 3))
    xk = np.stack([grid[d][idx[:, d]] for d in range(3)], axis=1)
    got = plan.eval_host(np.ascontiguousarray(xk))
    assert np.array_equal(got, vals[idx[:, 0], idx[:, 1], idx[:, 2]])
    pc = _lib.Plan.fh(np.full((40, 40, 40), 3.25), grid, weights)
    const = pc.eval_torch(xd[:100_000].contiguous()).cpu().numpy()
    assert np.abs(const - 3.25).max() <= 1e-13
    sel = torch.arange(0, 1_000_000, 500, device="cuda")
    want = port.fh(vals, grid, weights, xd[sel].cpu().numpy(), threads=8)
    assert np.abs(whole[sel].cpu().numpy() - want).max() <= TOL_REL * np.abs(want).max()
    plan.close()
    pc.close()


def test_multi_gpu_sharding_matches_single(gpu_ok, port):
    """The in-process partitioner (csrc/api.cu run_sharded): contiguous ranges over all visible
    GPUs, replicated tensor; with one GPU visible this degenerates to the single-device path."""
    ngpu = cb.device_count()
    cf = cases.cheb_fitted((10, 10, 10))
    x = cases.points(600_011, 3)
    one = _lib.evalcheb(cf, x, ngpus=1)
    allg = _lib.evalcheb(cf, x, ngpus=0)
    assert np.array_equal(bits(one), bits(allg))
    if ngpu >= 2:
        two = _lib.evalcheb(cf, x, ngpus=2)
        assert np.array_equal(bits(one), bits(two))
        # the plan of EVERY device is found again on the next call (entries are keyed per device, the
        # identity of the interpolant is not): no rebuild, no rejected candidate
        _lib.cache_clear()
        _lib.evalcheb(cf, x, ngpus=ngpu)
        before = _lib.cache_stats()
        _lib.evalcheb(cf, x, ngpus=ngpu)
        after = _lib.cache_stats()
        assert after["hits"] - before["hits"] == ngpu and after["misses"] == before["misses"], (before, after)
        assert after["rejected"] == before["rejected"] and after["entries"] == ngpu, (before, after)
    grid, values = cases.mlip_case((12, 11, 10))
    # (the kernel chosen per device can differ with the shard size, so this one is tolerance-based)
    assert np.abs(_lib.evalmlip(grid, values, x, 0, 1) - _lib.evalmlip(grid, values, x, 0, 0)).max() <= 1e-14


def test_stateless_plan_cache_and_pageable_staging(gpu_ok, port):
    """(a) The stateless entry points cache small interpolants by CONTENT: a second call with the
    same tensor (even at another address) must not rebuild; a modified tensor must miss.
    (b) Large pageable batches take the threaded pinned-staging path and give the same bits as
    the pinned path."""
    import torch

    cf = cases.cheb_fitted((10, 10, 10))
    x = cases.points(5000, 3)
    _lib.cache_clear()
    a = _lib.evalcheb(cf, x, ngpus=1)
    b = _lib.evalcheb(cf.copy(order="F"), x, ngpus=1)  # same content, different address: cache hit
    assert np.array_equal(bits(a), bits(b))
    cf2 = cf.copy(order="F")
    cf2[1, 2, 3] += 0.5
    c = _lib.evalcheb(cf2, x, ngpus=1)
    assert np.abs(c - port.evalcheb(cf2, x)).max() <= TOL_REL * np.abs(c).max()
    assert not np.array_equal(bits(a), bits(c))
    # blends are a per-call argument on a cached multilinear plan
    grid, values = cases.mlip_case((9, 8, 7))
    for blend in (0, 3, 0, 4):
        got = _lib.evalmlip(grid, values, x, blend, 1)
        assert np.abs(got - port.evalmlip(grid, values, x, blend=blend)).max() <= 1e-13
    _lib.cache_clear()
    # pageable vs pinned, > 16 MB so that staging engages; odd size for a ragged last chunk
    n = 3_000_017
    xp = torch.empty((n, 3), dtype=torch.float64, pin_memory=True)
    xp.uniform_(-1, 1, generator=torch.Generator().manual_seed(5))
    op = torch.empty(n, dtype=torch.float64, pin_memory=True)
    plan = _lib.Plan.cheb(cf)
    plan.eval_host_ptr(xp.data_ptr(), n, op.data_ptr())
    xg = xp.numpy().copy()
    og = plan.eval_host(xg)
    assert np.array_equal(bits(og), bits(op.numpy()))
    sel = slice(0, n, 3001)
    assert np.abs(og[sel] - port.evalcheb(cf, xg[sel])).max() <= TOL_REL * np.abs(og).max()
    plan.close()


# ----------------------------------------------------------------------------- fit side
@pytest.mark.parametrize("dims", [(7,), (2, 3, 4), (8, 7, 6), (10,) * 5, (33, 5), (1, 4), (1500,)])
def test_chebcoef_on_device(gpu_ok, port, dims):
    """chebpol_cuda_chebcoef against the reference's matrix path (src/chebpol.c:65-140), with and
    without dct, and the golden table of tests/all.Rout.save:28-48 for the 2x3x4 case."""
    rng = np.random.default_rng(11)
    a = np.asfortranarray(rng.standard_normal(dims))
    for dct in (False, True):
        got = _lib.chebcoef(a, dct)
        want = port.chebcoef(a, dct)
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= 1e-13 * max(1.0, np.abs(want).max())
    if dims == (2, 3, 4):
        from r_rng import RRng

        a = RRng(42).rnorm(24).reshape((2, 3, 4), order="F")
        cf = _lib.chebcoef(a)
        assert np.allclose(cf[:, :, 0], [[0.1163838, -0.1845058, -0.05441405], [0.1003040, -0.1881047, -0.34460793]], rtol=0, atol=6e-8)


def test_fhweights_on_device_and_fit_on_device_ipol(gpu_ok, port):
    grid = [np.linspace(-1, 1, 50), np.linspace(-1, 1, 30) ** 3, cb.chebknots([20])[0]]
    got = _lib.fhweights(grid, [3, 4, 4])
    want = port.fhweights(grid, [3, 4, 4])
    for g, w in zip(got, want):
        assert np.array_equal(bits(g), bits(w))  # same operation order: bit-identical
    # ipol end to end with the fit on the device: golden values of tests/all.Rout.save:69,71
    f = lambda x: np.exp(-np.sum(x ** 2))
    iv = [(1, 2), (1, 4), (1, 3)]
    val = cb.evalongrid(f, [8, 7, 6], iv)
    ch = cb.chebappx(val, iv, fit="device")
    s = np.array([1.4, 2.3, 1.9])
    assert abs((ch(s)[0] - f(s)) - (-3.115592e-06)) < 5e-13
    a = np.array([1.01, 1.01, 1.01])
    assert abs((ch(a)[0] - f(a)) - 0.0001730159) < 5e-11
    ch.close()
    fh = cb.fhappx(lambda x: np.sum(np.log(1 + x ** 2)), grid=[g[:12] for g in grid], d=[3, 4, 4], fit="device")
    x = cases.points(500, 3, lo=0.0, hi=0.9)
    host = cb.fhappx(lambda x: np.sum(np.log(1 + x ** 2)), grid=[g[:12] for g in grid], d=[3, 4, 4])
    assert np.array_equal(bits(fh(x.T)), bits(host(x.T)))
    fh.close()
    host.close()


# --------------------------------------------------------------- evaluation on a tensor grid
def _expand_grid(pts):
    mesh = np.meshgrid(*pts, indexing="ij")
    return np.stack([m.ravel(order="F") for m in mesh], axis=1)  # first factor fastest, as expand.grid


def test_evalongridV_matches_pointwise(gpu_ok, port):
    """evalongridV(interpolant, grid) by mode products against the reference evaluated point by
    point on expand.grid(grid) (R/tools.R:132-137), for all four evaluators."""
    rng = np.random.default_rng(4)
    # Chebyshev with intervals, uniform
    dims = (7, 6, 5)
    iv = [(0.0, 1.0), (-2.0, 2.0), (1.0, 3.0)]
    val = cases.grid_values(cb.chebknots(dims, iv))
    pts = [np.sort(rng.uniform(a, b, m)) for (a, b), m in zip(iv, (9, 4, 11))]
    ch = cb.chebappx(val, iv)
    got = cb.evalongridV(ch, grid=pts)
    X = _expand_grid(pts)
    want = port.evalcheb(ch.t["coef"], port.imap(X, ch.t["mid"], ch.t["ispan"])).reshape(got.shape, order="F")
    assert got.shape == (9, 4, 11)
    assert np.abs(got - want).max() <= TOL_REL * np.abs(want).max()
    uc = cb.ucappx(val)
    ptsu = [np.linspace(-1, 1, m) for m in (5, 6, 7)]
    gotu = cb.evalongridV(uc, grid=ptsu)
    wantu = port.evalcheb(uc.t["coef"], port.ugm(_expand_grid(ptsu), dims)).reshape(gotu.shape, order="F")
    assert np.abs(gotu - wantu).max() <= 1e-13 * np.abs(wantu).max()
    # the grid of the fit reproduces the values (interpolation property)
    back = cb.evalongridV(ch, dims=dims, intervals=iv)
    assert np.abs(back - val).max() <= 1e-13
    # multilinear, two blends
    grid, values = cases.mlip_case((8, 7, 6), uniform=False)
    ml = cb.mlappx(values, grid)
    ptsm = [np.sort(rng.uniform(-1, 1, m)) for m in (10, 3, 9)]
    for name, b in (("linear", 0), ("cubic", 3)):
        gotm = cb.evalongridV(ml, grid=ptsm, blend=name)
        wantm = port.evalmlip(grid, values, _expand_grid(ptsm), blend=b).reshape(gotm.shape, order="F")
        assert np.abs(gotm - wantm).max() <= TOL_REL * np.abs(wantm).max()
    # Floater-Hormann, including grid nodes (knot hits); the example of chebpol-Ex.Rout.save:319-328
    # evaluates g on its own grid and gets the values back
    vals, g, w = cases.fh_case((9, 8, 7), d=3, kind="mixed")
    fh = cb.fhappx(vals, grid=g, d=3)
    gotf = cb.evalongridV(fh, grid=g)
    assert np.array_equal(gotf, vals)
    ptsf = [np.sort(rng.uniform(-0.95, 0.95, m)) for m in (6, 7, 8)]
    gotf = cb.evalongridV(fh, grid=ptsf)
    wantf = port.fh(vals, g, fh.t["weights"], _expand_grid(ptsf)).reshape(gotf.shape, order="F")
    assert np.abs(gotf - wantf).max() <= 1e-11 * np.abs(wantf).max()  # cubed grid: Lebesgue constant ~1e2
    # interpolants without a tensor-grid kernel take the reference's route (expand.grid, pointwise)
    hs = cb.ipol(cases.grid_values([np.linspace(-1, 1, 7)] * 2), grid=[np.linspace(-1, 1, 7)] * 2, method="hstalker")
    ptsh = [np.linspace(-0.9, 0.9, 5), np.linspace(-1, 1, 4)]
    goth = cb.evalongridV(hs, grid=ptsh)
    wanth = port.evalhyp(hs.t["values"], hs.t["grid"], *hs.t["tables"], _expand_grid(ptsh), 3).reshape((5, 4), order="F")
    assert goth.shape == (5, 4) and np.abs(goth - wanth).max() <= TOL_REL * np.abs(wanth).max()
    for ip in (ch, uc, ml, fh, hs):
        ip.close()


# --------------------------------------------------------------- polyharmonic splines
def _polyh_case(rank, nk, seed=0):
    rng = np.random.default_rng(seed)
    return rng.uniform(0, 1, (rank, nk)), rng.standard_normal(nk), rng.standard_normal(rank + 1)


def _polyh_scale(kn, w, lw, k, x):
    """sum_i |w_i phi_i(x)| + |affine terms|: the conditioning bound of the radial sum."""
    d2 = np.maximum((x ** 2).sum(axis=1)[:, None] + (kn ** 2).sum(axis=0)[None, :] - 2.0 * x @ kn, 0.0)
    from chebpol_b200.interp import _phifunc
    return (np.abs(w)[None, :] * np.abs(_phifunc(d2, k))).sum(axis=1) + np.abs(lw[0]) + np.abs(x * lw[1:]).sum(axis=1)


POLYH_SHAPES = [(1, 17), (2, 40), (3, 1000), (4, 33), (5, 64), (6, 1000), (7, 9), (8, 100), (11, 50), (16, 300), (20, 1000),
                (31, 20), (3, 1), (24, 3000)]


@pytest.mark.parametrize("rank,nk", POLYH_SHAPES)
@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 0, -1.0, -0.37])
def test_polyh_strict_and_tile(gpu_ok, port, rank, nk, k):
    """Both kernels against the oracle; points inside and outside the knots' hull plus exact knot
    hits (distance zero is skipped for k >= 0, src/chebpol.c:400,408).
    strict, odd k (sqrt and multiplies only): BIT-IDENTICAL.  strict, even k / k < 0 (device
    log/exp within 1 ulp of libm) and the tile kernel (FMA, phi reformulated):
    |diff| <= 1e-12*max|f| + 8*eps*sum_i|w_i phi_i| pointwise."""
    kn, w, lw = _polyh_case(rank, nk, seed=rank * 100 + nk)
    rng = np.random.default_rng(9)
    x = np.ascontiguousarray(np.vstack([rng.uniform(-0.3, 1.3, (2500, rank)), kn.T[: min(nk, 11)]]))
    want = port.evalpolyh(kn, w, lw, k, x, threads=4)
    scale = _polyh_scale(kn, w, lw, k, x)
    tol = TOL_REL * np.abs(want).max() + 8 * np.finfo(float).eps * scale
    plan = _lib.Plan.polyh(kn, w, lw, k)
    got, kern = eval_plan(plan, x, _lib.KERNEL_STRICT)
    assert kern == "polyh_strict"
    if k > 0 and int(k) % 2 == 1:
        assert np.array_equal(bits(got), bits(want))
    else:
        assert np.all(np.abs(got - want) <= tol)
    got, kern = eval_plan(plan, x, _lib.KERNEL_AUTO)
    assert kern == "polyh_tile"
    assert np.all(np.abs(got - want) <= tol), (np.abs(got - want).max(), tol.min())
    plan.close()


def test_polyh_normalisation_nan_and_empty(gpu_ok, port):
    kn, w, lw = _polyh_case(4, 50, seed=3)
    a = np.array([0.05, 2.0, 1.0, 0.5])
    b = np.array([-3.0, 0.25, 0.0, 1.0])
    rng = np.random.default_rng(2)
    x = np.ascontiguousarray(rng.uniform(-3, 17, (3000, 4)))
    xn = a * (x - b)  # normfun, R/polyharmonic.R:77
    for kern, exact in ((_lib.KERNEL_STRICT, True), (_lib.KERNEL_AUTO, False)):
        plan = _lib.Plan.polyh(kn, w, lw, 3, norm_a=a, norm_b=b)
        got, _ = eval_plan(plan, x, kern)
        want = port.evalpolyh(kn, w, lw, 3, xn, threads=2)
        if exact:
            assert np.array_equal(bits(got), bits(want))
        else:
            assert np.all(np.abs(got - want) <= TOL_REL * np.abs(want).max() + 8e-16 * _polyh_scale(kn, w, lw, 3, xn))
        xx = x[:5].copy()
        xx[1, 2] = np.nan
        got, _ = eval_plan(plan, xx, kern)
        assert np.isnan(got[1]) and np.all(np.isfinite(got[[0, 2, 3, 4]]))
        plan.close()
    # no knots: the affine part alone
    plan = _lib.Plan.polyh(np.zeros((2, 0)), np.zeros(0), np.array([1.0, 2.0, -3.0]), 3)
    got, kern = eval_plan(plan, np.array([[0.5, 0.25], [1.0, 1.0]]), _lib.KERNEL_AUTO)
    assert kern == "polyh_strict" and np.array_equal(got, [1.25, 0.0])
    plan.close()


def test_polyh_goldens_through_ipol(gpu_ok):
    """all.R:82-95 -> all.Rout.save:195,203,207,208 through fit + GPU evaluation (ipol front end)."""
    from r_rng import RRng

    r = RRng(42)
    r.rnorm(24); r.runif(3, -1, 1); r.runif(1); r.runif(3); r.runif(3); r.rnorm(24); r.runif(10); r.runif(1, -1, 1)
    f = lambda x: 10.0 / (10.0 + np.sum(np.sqrt(x)))
    knots = r.runif(6000, 0, 20).reshape((6, 1000), order="F")
    a = r.runif(6, 0, 20)
    assert abs(f(a) - 0.3536114) < 5e-8
    phs = cb.ipol(f, knots=knots, k=3, method="poly")
    assert abs(phs(a)[0] - 0.3539679) < 5e-7
    assert np.abs(phs(knots) - np.array([f(knots[:, i]) for i in range(1000)])).max() < 5e-7
    assert abs(cb.ipol(f, knots=knots, k=5, method="poly", normalize=True)(a)[0] - 0.3539086) < 5e-7
    assert abs(cb.ipol(f, knots=knots, k=-1, method="poly", normalize=False)(a)[0] - 0.3561987) < 5e-7
    nphs = cb.ipol(f, knots=knots, k=-1, method="poly", normalize=True)
    assert abs(nphs(a)[0] - 0.3523617) < 5e-7
    for kern in (_lib.KERNEL_STRICT, _lib.KERNEL_AUTO):
        nphs.kernel = kern
        assert abs(nphs(a, threads=2)[0] - 0.3523617) < 5e-7


def test_polyh_stateless_and_sharded(gpu_ok, port):
    kn, w, lw = _polyh_case(6, 1000, seed=8)
    rng = np.random.default_rng(1)
    x = np.ascontiguousarray(rng.uniform(0, 1, (300_000, 6)))
    want = port.evalpolyh(kn, w, lw, 3, x[:4000], threads=8)
    tol = TOL_REL * np.abs(want).max() + 8e-16 * _polyh_scale(kn, w, lw, 3, x[:4000])
    got1 = _lib.evalpolyh(kn, w, lw, 3, x, ngpus=1)
    assert np.all(np.abs(got1[:4000] - want) <= tol)
    got2 = _lib.evalpolyh(kn, w, lw, 3, x, ngpus=0)  # all visible GPUs; second call hits the plan cache
    assert np.array_equal(got1, got2)
    a = np.full(6, 0.5)
    b = np.full(6, -0.25)
    gotn = _lib.evalpolyh(kn, w, lw, 2, x[:5000], norm_a=a, norm_b=b)
    wantn = port.evalpolyh(kn, w, lw, 2, a * (x[:5000] - b), threads=8)
    assert np.all(np.abs(gotn - wantn) <= TOL_REL * np.abs(wantn).max() + 8e-16 * _polyh_scale(kn, w, lw, 2, a * (x[:5000] - b)))


def test_polyh_fitted_spline_error_against_extended_precision(gpu_ok, port):
    """A fitted spline sums 1000 terms w_i phi_i of mixed sign that are ~1e3 times larger than f,
    so |GPU - reference| (3e-13 on max|f| = 0.15) is above 1e-12*max|f| -- but so is the
    reference's own distance from the exactly rounded sum.  Measured against an 80-bit evaluation:
    the tile kernel is as close to the exact value as the reference C is."""
    kn, w, lw = cases.polyh_case(6, 1000, 3)
    x = cases.points(400, 6)
    ld = np.longdouble
    d2 = ((x[:, :, None].astype(ld) - kn[None, :, :].astype(ld)) ** 2).sum(axis=1)
    exact = (w.astype(ld)[None, :] * d2 * np.sqrt(d2)).sum(axis=1) + ld(lw[0]) + (x.astype(ld) * lw[1:].astype(ld)).sum(axis=1)
    ref = port.evalpolyh(kn, w, lw, 3, x, threads=4)
    plan = _lib.Plan.polyh(kn, w, lw, 3)
    got, kern = eval_plan(plan, x, _lib.KERNEL_AUTO)
    assert kern == "polyh_tile"
    err_ref = np.abs(ref.astype(ld) - exact).max()
    err_gpu = np.abs(got.astype(ld) - exact).max()
    scale = _polyh_scale(kn, w, lw, 3, x).max()
    assert err_ref > 1e-13 * np.abs(ref).max()          # the reference is not within 1e-13*max|f| of exact either
    assert err_gpu <= 2.0 * err_ref + 2e-16 * scale, (float(err_gpu), float(err_ref), scale)
    plan.close()


# --------------------------------------------------------------- stalker splines
def _stalk_case(dims, uniform, smooth, seed=0):
    rng = np.random.default_rng(seed)
    grid = [np.linspace(-1, 1, n) if uniform else np.sort(rng.uniform(-1, 1, n)) for n in dims]
    val = cases.grid_values(grid) if smooth else np.asfortranarray(rng.standard_normal(dims))
    x = np.ascontiguousarray(rng.uniform(-1.15, 1.15, (4000, len(dims))))
    x[:3] = np.array([g[min(1, len(g) - 1)] for g in grid])  # knot hits
    x[3] = [g[-1] for g in grid]
    x[4] = [g[0] for g in grid]
    return grid, val, x


def _close(got, want, tol):
    """Same non-finite pattern (the hyperbola's pole 1 + c*z == 0 is reached by exact knot hits under
    the blends that keep every corner: 0/0 in the reference too), finite values within tol."""
    fin = np.isfinite(want)
    return (np.array_equal(fin, np.isfinite(got)) and np.array_equal(np.isnan(want), np.isnan(got))
            and (not fin.any() or np.abs(got[fin] - want[fin]).max() <= tol))


def _stalk_scale(kind, val, grid, tables, x):
    """Magnitude of the terms a corner sums (the cancellation the evaluator is exposed to)."""
    rank = len(grid)
    H = max(float(np.max(np.diff(g))) for g in grid) + 0.2
    if kind == 1:
        a, b, c, d = tables
        return np.abs(val).max() + np.abs(a).max() + rank * (np.abs(b).max() * H * max(H, 1.0) + np.abs(d).max())
    b, c, r = tables
    return np.abs(val).max() + rank * (np.abs(b).max() * H + np.abs(c).max() * max(H, H ** 2))


STALK_SHAPES = [((9,), True), ((12, 7), False), ((8, 9, 7), True), ((6, 5, 7), False), ((5, 6, 4, 5), True),
                ((4, 3, 5, 3, 4), False), ((3, 4, 3, 3, 2, 3), True), ((2, 2), True)]


@pytest.mark.parametrize("dims,uniform", STALK_SHAPES)
@pytest.mark.parametrize("smooth", [True, False])
def test_hstalker_strict_bit_exact_and_cell(gpu_ok, port, dims, uniform, smooth):
    """evalhyp uses only + - * /: the strict kernel is BIT-IDENTICAL to the reference for the
    blends without exp (0, 3, 4, 5) and within 1e-14 for 1, 2; the packed-record kernel within
    1e-12*max|f| plus 32*eps times the size of the terms that cancel in a corner."""
    grid, val, x = _stalk_case(dims, uniform, smooth, seed=len(dims))
    th = port.makehyp(val, grid)
    plan = _lib.Plan.stalk(1, val, grid, th)
    for blend in range(6):
        want = port.evalhyp(val, grid, *th, x, blend, threads=4)
        plan.set(_lib.OPT_BLEND, blend)
        got, k = eval_plan(plan, x, _lib.KERNEL_STRICT)
        assert k == "stalk_strict"
        if blend in (0, 3, 4, 5):
            assert np.array_equal(bits(got), bits(want)), blend
        else:
            assert _close(got, want, 1e-13 * _stalk_scale(1, val, grid, th, x))
        got, k = eval_plan(plan, x, _lib.KERNEL_AUTO)
        assert k == "stalk_cell"
        tol = TOL_REL * np.nanmax(np.abs(want)) + 32 * np.finfo(float).eps * _stalk_scale(1, val, grid, th, x)
        assert _close(got, want, tol), (blend, np.nanmax(np.abs(got - want)), tol)
    plan.close()


@pytest.mark.parametrize("dims,uniform", STALK_SHAPES)
@pytest.mark.parametrize("degrees", ["fit", "random"])
def test_stalker_strict_and_cell(gpu_ok, port, dims, uniform, degrees):
    """evalstalk: fitted tables (degrees 2, or |dv/sv| in [1,2] on uniform grids) and random
    degrees in (1,2) on every grid point (the pow branch).  Device pow is within 2 ulp of libm."""
    import warnings

    grid, val, x = _stalk_case(dims, uniform, True, seed=10 + len(dims))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        b, c, r = port.makestalk(val, grid)
    if degrees == "random":
        rng = np.random.default_rng(4)
        r = np.asfortranarray(np.where(c != 0, rng.uniform(1.0, 2.0, r.shape), r))
    ts = (b, c, r)
    plan = _lib.Plan.stalk(0, val, grid, ts)
    scale = _stalk_scale(0, val, grid, ts, x)
    for blend in (3, 0, 1, 4):
        want = port.evalstalk(val, grid, *ts, x, blend, threads=4)
        plan.set(_lib.OPT_BLEND, blend)
        for kern, name in ((_lib.KERNEL_STRICT, "stalk_strict"), (_lib.KERNEL_AUTO, "stalk_cell")):
            got, k = eval_plan(plan, x, kern)
            assert k == name
            tol = TOL_REL * np.nanmax(np.abs(want)) + 32 * np.finfo(float).eps * scale
            assert _close(got, want, tol), (blend, name, np.nanmax(np.abs(got - want)), tol)
    plan.close()


def test_stalker_high_rank_nan_and_golden(gpu_ok, port):
    # rank 7: outside the packed kernel, served by the strict one
    dims = (3, 2, 3, 2, 3, 2, 3)
    grid, val, x = _stalk_case(dims, True, True, seed=2)
    th = port.makehyp(val, grid)
    plan = _lib.Plan.stalk(1, val, grid, th)
    got, k = eval_plan(plan, x[:500], _lib.KERNEL_AUTO)
    assert k == "stalk_strict"
    assert np.array_equal(bits(got), bits(port.evalhyp(val, grid, *th, x[:500], 3)))
    plan.close()
    # NaN coordinates give NaN, the other points are unaffected
    grid, val, x = _stalk_case((8, 9, 7), False, True, seed=5)
    th = port.makehyp(val, grid)
    plan = _lib.Plan.stalk(1, val, grid, th)
    xx = x[:6].copy()
    xx[2, 1] = np.nan
    for kern in (_lib.KERNEL_STRICT, _lib.KERNEL_AUTO):
        got, _ = eval_plan(plan, xx, kern)
        assert np.isnan(got[2]) and np.all(np.isfinite(np.delete(got, 2)))
    plan.close()
    # chebpol-Ex.Rout.save:465,486: hst <- ipol(f, grid=grid, method='hstalker'); hst(m)
    import math
    from r_rng import RRng

    su = np.array([i * (1.0 / 9.0) for i in range(10)])
    su[-1] = 1.0
    s = np.array([math.pow(v, 3.0) for v in su])
    f = lambda x: 10.0 / (1.0 + 25.0 * np.mean(np.asarray(x) ** 2))
    r = RRng(1)
    r.runif(3000)
    m = r.runif(21).reshape((3, 7), order="F")
    hst = cb.ipol(f, grid=[s, su, np.sort(1.0 - s)], method="hstalker")
    assert np.allclose(hst(m), [1.166729, 0.6569404, 1.552856, 0.9696696, 0.5326136, 1.416647, 0.7832519], rtol=6e-7)
    assert not np.allclose(hst(m, blend="linear"), hst(m), rtol=1e-6)  # the blend argument reaches the kernel
    st = cb.ipol(f, grid=[su, su, su], method="stalker")
    want = port.evalstalk(st.t["values"], st.t["grid"], *st.t["tables"], m.T, 3)
    assert np.abs(st(m) - want).max() <= 1e-12 * np.abs(want).max()


def test_stalker_stateless_and_sharded(gpu_ok, port):
    grid, val, _ = _stalk_case((20, 18, 16), False, True, seed=6)
    x = cases.points(300_000, 3)
    th = port.makehyp(val, grid)
    want = port.evalhyp(val, grid, *th, x[:5000], 3, threads=8)
    tol = TOL_REL * np.abs(want).max() + 32 * np.finfo(float).eps * _stalk_scale(1, val, grid, th, x)
    g1 = _lib.evalhyp(val, grid, *th, x, blend=3, ngpus=1)
    assert np.abs(g1[:5000] - want).max() <= tol
    g2 = _lib.evalhyp(val, grid, *th, x, blend=3, ngpus=0)  # all GPUs; plan cache hit on device 0
    assert np.array_equal(g1, g2)
    g0 = _lib.evalhyp(val, grid, *th, x[:5000], blend=0)  # blend is per call, not part of the cached plan
    assert np.abs(g0 - port.evalhyp(val, grid, *th, x[:5000], 0, threads=8)).max() <= tol
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ts = port.makestalk(val, grid)
    gs = _lib.evalstalk(val, grid, *ts, x[:5000], blend=3)
    ws = port.evalstalk(val, grid, *ts, x[:5000], 3, threads=8)
    assert np.abs(gs - ws).max() <= TOL_REL * np.abs(ws).max() + 32 * np.finfo(float).eps * _stalk_scale(0, val, grid, ts, x)


def test_empty_and_single_point_batches_every_evaluator(gpu_ok, port):
    """P = 0 returns an empty vector and P = 1 (a length-K vector in R, src/chebpol.c:368) one value,
    through plans and through the stateless entry points, for every evaluator."""
    import warnings

    cf = cases.cheb_fitted((6, 5, 4))
    grid, values = cases.mlip_case((6, 5, 4))
    vals, g, w = cases.fh_case((6, 5, 4))
    kn, wk, lw = _polyh_case(3, 30)
    sgrid, sval, _ = _stalk_case((6, 5, 4), True, True)
    th = port.makehyp(sval, sgrid)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ts = port.makestalk(sval, sgrid)
    x1 = np.array([[0.3, -0.2, 0.5]])
    x0 = np.empty((0, 3))
    todo = [
        (_lib.Plan.cheb(cf), lambda x: _lib.evalcheb(cf, x), port.evalcheb(cf, x1)),
        (_lib.Plan.mlip(grid, values, 0), lambda x: _lib.evalmlip(grid, values, x, 0), port.evalmlip(grid, values, x1)),
        (_lib.Plan.fh(vals, g, w), lambda x: _lib.fh(vals, g, w, x), port.fh(vals, g, w, x1)),
        (_lib.Plan.polyh(kn, wk, lw, 3), lambda x: _lib.evalpolyh(kn, wk, lw, 3, x), port.evalpolyh(kn, wk, lw, 3, x1)),
        (_lib.Plan.stalk(1, sval, sgrid, th), lambda x: _lib.evalhyp(sval, sgrid, *th, x), port.evalhyp(sval, sgrid, *th, x1, 3)),
        (_lib.Plan.stalk(0, sval, sgrid, ts), lambda x: _lib.evalstalk(sval, sgrid, *ts, x), port.evalstalk(sval, sgrid, *ts, x1, 3)),
    ]
    for plan, stateless, want in todo:
        assert plan.eval_host(x0).shape == (0,) and stateless(x0).shape == (0,)
        for got in (plan.eval_host(x1), stateless(x1)):
            assert got.shape == (1,) and abs(got[0] - want[0]) <= 1e-12 * max(1.0, abs(want[0]))
        plan.close()


def test_cheb_and_fh_random_shapes_sweep(gpu_ok, port):
    """Sixty seeded random shapes (rank 1-7, dims 1-14) through both fast Chebyshev kernels and
    the DMMA Floater-Hormann path: odd step counts, partial slabs, padded level-A / column tiles,
    rank-deficient foldings -- whatever the layout logic does with them must agree with the oracle."""
    rng = np.random.default_rng(2024)
    n_mma = n_slab = n_fh = 0
    for _ in range(60):
        rank = int(rng.integers(1, 8))
        dims = tuple(int(v) for v in rng.integers(1, 15, rank))
        while np.prod(dims) > 300_000:
            dims = tuple(max(1, d - 2) for d in dims)
        cf = cases.cheb_stress(dims, seed=int(rng.integers(1 << 30)))
        x = cases.points(int(rng.integers(1, 3000)), rank, seed=int(rng.integers(1 << 30)))
        want = port.evalcheb(cf, x, threads=4)
        scale = np.abs(cf).sum()
        plan = _lib.Plan.cheb(cf)
        for variant in (_lib.VARIANT_MMA, _lib.VARIANT_SLAB, 0):
            got, k = eval_plan(plan, x, _lib.KERNEL_AUTO, variant)
            assert np.abs(got - want).max() <= TOL_REL * scale, (dims, k, np.abs(got - want).max(), scale)
            n_mma += k == "cheb_mma"
            n_slab += k == "cheb_slab"
        plan.close()
        if rank <= 4 and min(dims) >= 2:
            g = [np.sort(rng.uniform(-1, 1, n)) for n in dims]
            w = port.fhweights(g, [min(3, n - 1) for n in dims])
            vals = np.asfortranarray(rng.standard_normal(dims))
            xin = np.ascontiguousarray(np.stack([rng.uniform(gg[0], gg[-1], x.shape[0]) for gg in g], axis=1))
            wantf = port.fh(vals, g, w, xin, threads=4)
            # Lebesgue-function bound of the product basis (see test_fh_fast_irregular_grids)
            lam = np.ones(xin.shape[0])
            for d in range(rank):
                p = w[d][None, :] / (xin[:, d, None] - g[d][None, :])
                lam *= np.abs(p).sum(axis=1) / np.abs(p.sum(axis=1))
            planf = _lib.Plan.fh(vals, g, w)
            gotf, k = eval_plan(planf, xin, _lib.KERNEL_AUTO)
            n_fh += k == "fh_mma"
            # random knots come arbitrarily close, Lambda reaches 1e4 and |f| ~ Lambda*max|v|: a relative
            # perturbation of the poles moves the numerator by Lambda*max|v| and, through the cancelling
            # denominator, f by Lambda*|f| -- for the reference's roundings as much as for ours
            tol = TOL_REL * np.abs(wantf).max() + 16 * np.finfo(float).eps * lam * (np.abs(vals).max() + np.abs(wantf))
            assert np.all(np.abs(gotf - wantf) <= tol), (dims, k, np.abs(gotf - wantf).max())
            planf.close()
    assert n_mma >= 40 and n_slab >= 40 and n_fh >= 15, (n_mma, n_slab, n_fh)


# ------------------------------------------------- 'general' grid method (chebappxg)
def general_case(dims, seed=5):
    """Values of g on an irregular tensor grid (mixed increasing / decreasing, clustered knots)."""
    rng = np.random.default_rng(seed)
    grid = []
    for i, n in enumerate(dims):
        g = np.sort(np.concatenate([[-1.0, 1.0], rng.uniform(-1, 1, n - 2)])) if i % 2 == 0 else np.linspace(-1, 1, n) ** 3
        grid.append(g[::-1].copy() if i % 3 == 2 else g)
    return grid, cases.grid_values(grid)


@pytest.mark.parametrize("dims", [(10,), (12, 9), (8, 7, 6), (10,) * 5, (40, 7)])
def test_general_method_all_kernels(gpu_ok, port, dims):
    """chebappxg.real (R/chebyshev.R:333-359): hyman-spline grid map fused into the Chebyshev kernels.
    Strict kernel bit-identical to oracle map + oracle Clenshaw; fast kernels within 1e-12 max|f|.
    Points reach 5 % outside the grid (the first / last cubic piece extrapolates) and hit knots."""
    grid, val = general_case(dims)
    ip = cb.ipol(val, grid=grid, method="general")
    K = len(dims)
    x = cases.points(20011, K, lo=-1.05, hi=1.05)
    for k in range(K):
        x[: dims[k], k] = grid[k]
    want = port.evalcheb(ip.t["coef"], port.splinemap(x, ip.t["splines"], threads=4), threads=4)
    ip.kernel = _lib.KERNEL_STRICT
    got = ip(x.T)
    assert ip.plan().kernel_used == "cheb_strict"
    assert np.array_equal(bits(got), bits(want))
    plan = ip.plan()
    for variant in (_lib.VARIANT_MMA, _lib.VARIANT_SLAB, 0):
        got, k = eval_plan(plan, x, _lib.KERNEL_AUTO, variant)
        assert np.abs(got - want).max() <= TOL_REL * np.abs(want).max(), (k, np.abs(got - want).max())
    # the interpolant reproduces the data on its grid
    mesh = np.meshgrid(*grid, indexing="ij")
    gx = np.stack([m.ravel(order="F") for m in mesh])
    ip.kernel = _lib.KERNEL_AUTO
    assert np.abs(ip(gx) - val.ravel(order="F")).max() <= 1e-12
    # stateless entry point, spread over all GPUs, equals the plan path; NaN in, NaN out
    xs = cases.points(200000, K, seed=3)
    assert np.array_equal(bits(_lib.evalchebg(ip.t["coef"], ip.t["splines"], xs, 0)), bits(ip(xs.T)))
    xn = x[:4].copy()
    xn[1, 0] = np.nan
    out = ip(xn.T)
    assert np.isnan(out[1]) and np.isfinite(out[[0, 2, 3]]).all()
    ip.close()


def test_general_goldens_through_ipol(gpu_ok):
    """tests/all.R:36-59 -> all.Rout.save:109,121,126 through ipol(method='general') on the GPU."""
    from r_rng import RRng

    f = lambda x: np.exp(-np.sum(np.asarray(x) ** 2))
    r = RRng(42)
    r.rnorm(24)
    r.runif(3, -1, 1)
    s = (np.arange(10) * (1.0 / 9)) ** 3
    ch = cb.ipol(np.exp(s), grid=[s], method="general")
    rr = r.runif(1)
    assert abs(ch(rr)[0] - 1.418184) < 5e-7
    ch.close()
    su = np.linspace(0, 1, 11)
    grid = [su, su ** 2, su ** 3]
    ch = cb.ipol(f, grid=grid, method="general")
    for kern in (_lib.KERNEL_AUTO, _lib.KERNEL_STRICT):
        ch.kernel = kern
        rs = RRng(42)
        rs.rnorm(24); rs.runif(3, -1, 1); rs.runif(1)
        assert abs(ch(rs.runif(3))[0] - 0.460965) < 5e-7
        assert abs(ch(rs.runif(3))[0] - 0.3501301) < 5e-7
    # evalongridV falls to the expand-grid route for this method
    pts = [np.linspace(0.1, 0.9, 4), np.linspace(0.2, 0.8, 3), np.linspace(0.0, 1.0, 5)]
    G = cb.evalongridV(ch, grid=pts)
    assert G.shape == (4, 3, 5) and abs(G[1, 2, 3] - ch(np.array([pts[0][1], pts[1][2], pts[2][3]]))[0]) == 0.0
    ch.close()
    with pytest.raises(cb.ChebpolCudaError):
        t = ch.t["splines"]
        bad = [(t[0][0][::-1].copy(),) + tuple(t[0][1:])] + list(t[1:])
        _lib.Plan.chebg(ch.t["coef"], bad)


# ------------------------------------------------- simplex-linear (slappx)
def nan_aware_bits_equal(got, want):
    """Bit-identical, NaN payloads included (NA_real_ outside the hull)."""
    return np.array_equal(bits(got), bits(want))


@pytest.mark.parametrize("dim,nk", [(1, 50), (2, 400), (3, 1000), (4, 200), (5, 60), (6, 30), (7, 20)])
def test_simplexlinear_strict_and_cell(gpu_ok, port, dim, nk):
    """R_evalsl (src/chebpol.c:738-828): both kernels issue the reference's operations in its order, so
    both must be BIT-IDENTICAL to the oracle: every blend, with / without extrapolation, points
    outside the hull (NA_real_ payload included), exact knot hits, NaN coordinates."""
    knots, dtri, val = cases.sl_case(dim, nk)
    ad = port.analyzesimplex(dtri, knots)
    got_ad = _lib.analyzesimplex(dtri, knots)  # the fit step on the device
    assert all(np.array_equal(u, v) for u, v in zip(ad, got_ad))
    x = cases.points(30011, dim, lo=-1.1, hi=1.1)
    x[:nk] = knots.T
    x[nk, 0] = np.nan
    x[nk + 1, dim - 1] = np.nan
    plan = _lib.Plan.sl(knots, dtri, ad, val)
    for epol in (0, 1):
        plan.set(_lib.OPT_EPOL, epol)
        for blend in range(6):
            plan.set(_lib.OPT_BLEND, blend)
            want = port.evalsl(x, knots, dtri, ad, val, epol, blend, threads=4)
            got, k = eval_plan(plan, x, _lib.KERNEL_STRICT)
            assert k == "sl_strict"
            libm = blend in (1, 2)  # device exp vs libm exp
            if libm:
                assert np.array_equal(np.isnan(got), np.isnan(want))
                f = np.isfinite(want)
                assert np.abs(got[f] - want[f]).max() <= 1e-14 * np.abs(want[f]).max()
            else:
                assert nan_aware_bits_equal(got, want), (epol, blend)
            if dim <= 6:
                for variant in (0, 804, 900, 2):  # default, batched, per-lane state machine, 100-register build
                    got2, k = eval_plan(plan, x, _lib.KERNEL_FAST, variant)
                    assert k == "sl_cell"
                    assert nan_aware_bits_equal(got2, got), (epol, blend, variant)  # the kernels agree bit for bit
    if dim > 6:
        got, k = eval_plan(plan, x, _lib.KERNEL_AUTO)
        assert k == "sl_strict"
        with pytest.raises(cb.ChebpolCudaError):
            eval_plan(plan, x, _lib.KERNEL_FAST)
    plan.close()


def test_simplexlinear_cell_resolutions_and_stateless(gpu_ok, port, monkeypatch):
    """The candidate lists give the reference's answer at any grid resolution (1 cell = the plain
    scan ... much finer than the simplices), and the stateless entry point, sharded over all GPUs,
    equals the plan path."""
    knots, dtri, val = cases.sl_case(3, 800, seed=9)
    ad = port.analyzesimplex(dtri, knots)
    x = cases.points(200000, 3, lo=-1.05, hi=1.05, seed=2)
    want = port.evalsl(x, knots, dtri, ad, val, True, 3, threads=8)
    for cells in ("1", "2", "7", "64", "0"):
        monkeypatch.setenv("CHEBPOL_CUDA_SL_CELLS", cells)
        plan = _lib.Plan.sl(knots, dtri, ad, val)
        plan.set(_lib.OPT_EPOL, 1)
        got, k = eval_plan(plan, x, _lib.KERNEL_FAST)
        assert k == "sl_cell" and nan_aware_bits_equal(got, want), cells
        plan.close()
    monkeypatch.delenv("CHEBPOL_CUDA_SL_CELLS")
    # AUTO: a small first batch goes through the plain scan, a large one builds the lists, and a much
    # larger one rebuilds them finer; every route gives the same bits
    plan = _lib.Plan.sl(knots, dtri, ad, val)
    plan.set(_lib.OPT_EPOL, 1)
    got, k = eval_plan(plan, x[:500], _lib.KERNEL_AUTO)
    assert k == "sl_strict" and nan_aware_bits_equal(got, want[:500])
    got, k = eval_plan(plan, x[:20000], _lib.KERNEL_FAST)
    assert k == "sl_cell" and nan_aware_bits_equal(got, want[:20000])
    got, k = eval_plan(plan, x[:500], _lib.KERNEL_AUTO)  # the lists exist now
    assert k == "sl_cell" and nan_aware_bits_equal(got, want[:500])
    xx = np.tile(x, (20, 1))
    got, k = eval_plan(plan, xx, _lib.KERNEL_AUTO)
    assert k == "sl_cell" and nan_aware_bits_equal(got, np.tile(want, 20))
    for cells in (33, 0, 5):  # pinned grid, back to automatic, pinned again
        plan.set(_lib.OPT_SL_CELLS, cells)
        got, k = eval_plan(plan, x, _lib.KERNEL_AUTO)
        assert k == "sl_cell" and nan_aware_bits_equal(got, want), cells
    plan.close()
    got = _lib.evalsl(x, knots, dtri, ad, val, True, 3, 0)
    assert nan_aware_bits_equal(got, want)
    got = _lib.evalsl(x, knots, dtri, ad, val, False, 0, 1)  # per-call blend / epol on a cached plan
    assert nan_aware_bits_equal(got, port.evalsl(x, knots, dtri, ad, val, False, 0, threads=8))
    # empty batch, single point, empty triangulation
    assert _lib.evalsl(np.empty((0, 3)), knots, dtri, ad, val).shape == (0,)
    one = _lib.evalsl(x[:1], knots, dtri, ad, val, True, 3)
    assert nan_aware_bits_equal(one, want[:1])
    with pytest.raises(cb.ChebpolCudaError):
        bad = dtri.copy()
        bad[0, 0] = 0
        _lib.Plan.sl(knots, bad, ad, val)


def test_simplexlinear_goldens_through_ipol(gpu_ok):
    """chebpol-Ex.Rout.save:469,482-486, row 'sl': ipol(f, knots=knots, method='simplexlinear') at the
    seven points of the example (set.seed(1) stream), default blend 'cubic'."""
    from r_rng import RRng

    f = lambda x: 10.0 / (1.0 + 25.0 * np.mean(np.asarray(x) ** 2))
    r = RRng(1)
    knots = r.runif(3000).reshape((3, 1000), order="F")
    m = r.runif(21).reshape((3, 7), order="F")
    sl = cb.ipol(f, knots=knots, method="simplexlinear")
    want = [1.186031, 0.6786952, 1.550582, 0.9776420, 0.5247542, 1.478330, 0.7961787]
    for kern in (_lib.KERNEL_AUTO, _lib.KERNEL_STRICT):
        sl.kernel = kern
        assert np.allclose(sl(m), want, rtol=6e-7)
    assert np.isnan(sl(np.array([2.0, 2.0, 2.0]))[0]) and np.isfinite(sl(np.array([2.0, 2.0, 2.0]), epol=True)[0])
    lin = sl(m, blend="linear")
    assert np.all(np.abs(lin - want) < 0.05) and not np.allclose(lin, want, rtol=1e-9)
    G = cb.evalongridV(sl, grid=[np.linspace(0.2, 0.8, 4)] * 3)
    assert G.shape == (4, 4, 4) and np.isfinite(G).all()
    sl.close()


def test_full_size_simplexlinear_properties(gpu_ok, port):
    """Size-independent properties at 2e7 points (the oracle's linear scan would take minutes):
    with the linear blend a simplex-linear interpolant reproduces every affine function inside the
    hull; the set of points outside the hull (NA) and every value agree with the plain-scan kernel on
    a strided sample; a contiguous slice evaluates to the same bits as the same points inside the
    big batch (chunking / sharding independence)."""
    import torch

    knots, dtri, _ = cases.sl_case(3, 5000, seed=4)
    ad = _lib.analyzesimplex(dtri, knots)
    coef = np.array([0.7, -1.3, 0.4])
    val = knots.T @ coef + 0.25  # affine in the knots
    n = 20_000_000
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.rand((n, 3), generator=g, device="cuda", dtype=torch.float64) * 2.1 - 1.05
    plan = _lib.Plan.sl(knots, dtri, ad, val)
    plan.set(_lib.OPT_BLEND, 0)
    out = plan.eval_torch(x)
    assert plan.kernel_used == "sl_cell"
    want = x @ torch.tensor(coef, device="cuda") + 0.25
    inside = ~torch.isnan(out)
    frac = float(inside.double().mean())
    assert 0.5 < frac < 0.95  # the hull of 5000 knots in [-1,1]^3 against points of [-1.05,1.05]^3
    assert float((out[inside] - want[inside]).abs().max()) <= 1e-12
    # strided sample against the plain scan (bit for bit, NA included)
    idx = torch.arange(0, n, 997, device="cuda")
    xs = x[idx].contiguous()
    plan.set(_lib.OPT_KERNEL, _lib.KERNEL_STRICT)
    ref = plan.eval_torch(xs)
    assert plan.kernel_used == "sl_strict"
    assert torch.equal(ref.view(torch.int64), out[idx].view(torch.int64))
    # a slice evaluated on its own (host path, chunked) gives the same bits
    plan.set(_lib.OPT_KERNEL, _lib.KERNEL_AUTO)
    lo, hi = 7_000_001, 7_300_003
    got = plan.eval_host(np.ascontiguousarray(x[lo:hi].cpu().numpy()))
    assert np.array_equal(bits(got), bits(out[lo:hi].cpu().numpy()))
    # extrapolation: affine functions are reproduced outside the hull too (barycentric combination)
    plan.set(_lib.OPT_EPOL, 1)
    oute = plan.eval_torch(x[:2_000_000].contiguous())
    assert not bool(torch.isnan(oute).any())
    assert float((oute - want[:2_000_000]).abs().max()) <= 1e-7  # slivers on the hull boundary amplify rounding
    plan.close()


# ------------------------------------------------- the deprecated recursive stalker ('oldstalker')
@pytest.mark.parametrize("dims,uniform", [((9,), True), ((7, 6), False), ((6, 5, 7), True), ((4, 5, 4, 3), False),
                                          ((2, 2), True)])
def test_oldstalker_matches_reference(gpu_ok, ref, dims, uniform):
    """R_evalstalker (src/stalker.c:617-678, 834-891) against the compiled reference: every degree the
    no-GSL build accepts (2, 1, fractional, 0 = hyperbolic, NaN = per-basis on uniform grids), blends
    0-4, points outside the grid, exact knot hits.  Same operation order, so only pow / exp differ."""
    rng = np.random.default_rng(12)
    grid = [np.linspace(-1, 1, n) if uniform else np.sort(rng.uniform(-1, 1, n)) for n in dims]
    val = np.asfortranarray(rng.standard_normal(dims))
    K = len(dims)
    x = cases.points(4001, K, lo=-1.15, hi=1.15)
    x[:4] = np.array([g[min(1, len(g) - 1)] for g in grid])
    x[4] = [g[-1] for g in grid]
    x[5] = [g[0] for g in grid]
    degs = [[2.0], [1.0], [1.5], [0.0], list(np.resize([2.0, 1.3, 0.0], K))] + ([[np.nan]] if uniform else [])
    for r in degs:
        ip = cb.ipol(val, grid=grid, k=r, method="oldstalker")
        t = ip.t
        lo_in = np.array([g[0] + 1e-3 * (g[-1] - g[0]) for g in grid])
        hi_in = np.array([g[-1] - 1e-3 * (g[-1] - g[0]) for g in grid])
        for blend, name in enumerate(["linear", "sigmoid", "parodic", "cubic", "square"]):
            # sigmoid / parodic: exp(2 - 1/w) with w < 0 outside the grid overflows (and cancels) in the
            # reference too, so those two blends are compared inside the grid
            xb = np.clip(x, lo_in, hi_in) if blend in (1, 2) else x
            want = ref.evalstalker_old(t["values"], t["grid"], t["degree"], *t["ctx"], t["uniform"], xb, blend=blend, threads=4)
            got = ip(xb.T, blend=name)
            assert ip.plan().kernel_used == "oldstalk"
            fin = np.isfinite(want)
            assert np.array_equal(fin, np.isfinite(got)), (r, blend)
            exact = (r == [2.0] or r == [1.0]) and blend in (0, 3, 4)
            if exact and dims != (2, 2):
                pass  # pow(|x|, 2) and pow(|x|, 1) are exact in both libraries, but keep one tolerance for all
            # Outside the grid the blend weight of a narrow boundary cell is far from [0,1] (cubic: ~1e4), and
            # the reference itself moves by that factor times an ulp when x moves by an ulp.  The bar is
            # 1e-12 of the scale plus the reference's own sensitivity to a 1-ulp change of the point.
            up = ref.evalstalker_old(t["values"], t["grid"], t["degree"], *t["ctx"], t["uniform"], np.nextafter(xb, np.inf), blend=blend, threads=4)
            dn = ref.evalstalker_old(t["values"], t["grid"], t["degree"], *t["ctx"], t["uniform"], np.nextafter(xb, -np.inf), blend=blend, threads=4)
            with np.errstate(invalid="ignore"):
                sens = np.nan_to_num(np.maximum(np.abs(up - want), np.abs(dn - want)), nan=0.0, posinf=0.0)
            if blend == 4:
                sens[:] = 0.0  # the step blend is discontinuous: a neighbouring point may sit on the other side
            scale = max(1.0, np.abs(want[fin]).max())
            err = np.abs(got[fin] - want[fin])
            ok = err <= TOL_REL * scale + 16.0 * sens[fin]
            if blend == 4:  # ... so allow the (measure-zero) points that sit exactly on a switch
                assert ok.mean() > 0.999, (r, blend)
            else:
                assert ok.all(), (r, blend, err.max(), (err - 16.0 * sens[fin]).max())
        # stateless entry point (all GPUs) equals the plan path
        st = _lib.evalstalker(t["values"], t["grid"], t["degree"], t["ctx"], t["uniform"], x, 3, 0)
        assert np.array_equal(bits(st), bits(ip(x.T, blend="cubic")))
        ip.close()
    with pytest.raises(cb.ChebpolCudaError):  # NaN degree on a non-uniform grid: refused like the no-GSL reference
        g2 = [np.array([0.0, 0.1, 0.4, 1.0])]
        _lib.Plan.oldstalk(np.zeros(4), g2, [np.nan], cb.interp.stalkercontext(g2, [2.0]), [0])


# ------------------------------------------------- round 2: cache, generator, stream ordering
def test_synth_points_device_equals_numpy_twin(gpu_ok):
    import torch

    from chebpol_b200 import synth

    for first, n, k, lo, hi in [(0, 100_003, 5, -1.0, 1.0), (10**9 - 7, 7, 6, -1.0, 1.0), (123_456_789, 4099, 3, -1.2, 1.2)]:
        d = torch.empty((n, k), dtype=torch.float64, device="cuda")
        _lib.synth_points_device(d.data_ptr(), first, n, k, seed=42, lo=lo, hi=hi)
        torch.cuda.synchronize()
        assert np.array_equal(bits(d.cpu().numpy()), bits(synth.points(first, n, k, seed=42, lo=lo, hi=hi)))


def test_checksum_is_independent_of_sharding(gpu_ok):
    """What bench.py's cross-rank checksum relies on: the wrapping 64-bit sum of the result bit patterns of a
    counter-based batch does not depend on how the batch is cut into shards (here 1, 2 and 3 pieces)."""
    import torch

    for plan, k in [(_lib.Plan.cheb(cases.cheb_fitted((10,) * 5)), 5), (_lib.Plan.mlip(*cases.mlip_case((16,) * 4), 0), 4),
                    (_lib.Plan.fh(*cases.fh_case((12, 11, 10))), 3)]:
        n = 400_003
        sums = []
        for pieces in (1, 2, 3):
            tot = 0
            per = (n + pieces - 1) // pieces
            for r in range(pieces):
                lo, hi = min(n, r * per), min(n, (r + 1) * per)
                d = torch.empty((hi - lo, k), dtype=torch.float64, device="cuda")
                _lib.synth_points_device(d.data_ptr(), lo, hi - lo, k)
                o = plan.eval_torch(d)
                tot = (tot + int(o.view(torch.int64).sum().item())) & 0xFFFFFFFFFFFFFFFF
            sums.append(tot)
        assert sums[0] == sums[1] == sums[2], sums
        plan.close()


def test_cache_tells_apart_tensors_with_equal_fingerprints(gpu_ok, port):
    """VERDICT r1 item 7 on the device path: every fingerprint forced equal, so two coefficient tensors of
    the same shape collide in the key; the byte comparison must keep them apart."""
    cf1 = cases.cheb_fitted((8, 7, 6))
    cf2 = cf1.copy(order="F")
    cf2[3, 2, 1] += 0.25
    x = cases.points(4000, 3)
    _lib.cache_clear()
    _lib.cache_weak_hash(True)
    try:
        s0 = _lib.cache_stats()
        a1 = _lib.evalcheb(cf1, x, ngpus=1)
        a2 = _lib.evalcheb(cf2, x, ngpus=1)   # same key as cf1: must be rejected by the byte comparison
        b1 = _lib.evalcheb(cf1.copy(order="F"), x, ngpus=1)
        b2 = _lib.evalcheb(cf2, x, ngpus=1)
        s1 = _lib.cache_stats()
    finally:
        _lib.cache_weak_hash(False)
        _lib.cache_clear()
    assert np.array_equal(bits(a1), bits(b1)) and np.array_equal(bits(a2), bits(b2))
    assert not np.array_equal(bits(a1), bits(a2))
    for got, cf in ((a1, cf1), (a2, cf2)):
        want = port.evalcheb(cf, x)
        assert np.abs(got - want).max() <= TOL_REL * np.abs(want).max()
    assert s1["rejected"] - s0["rejected"] >= 1 and s1["hits"] - s0["hits"] >= 2


def test_cache_keeps_large_tables_and_their_cell_tables(gpu_ok, port):
    """an interpolant above the 32 MB retention limit (48^4 doubles = 42 MB) is cached by
    fingerprint; the second call neither re-uploads it nor rebuilds the cell-major table."""
    grid, values = cases.mlip_case((48,) * 4)
    call = _lib.Stateless.mlip(grid, values, 0)
    x = cases.points(700_000, 4)  # >= cells/8: the first call expands the cell-major table
    _lib.cache_clear()
    s0, l0 = _lib.cache_stats(), cb.launch_count()
    a = call(x, ngpus=1)
    s1, l1 = _lib.cache_stats(), cb.launch_count()
    b = call(x[:1000], ngpus=1)
    s2, l2 = _lib.cache_stats(), cb.launch_count()
    assert s1["misses"] - s0["misses"] == 1 and s1["entries"] == 1
    assert s1["device_bytes"] >= values.nbytes * 17  # table + 16x cell-major expansion held
    assert s2["hits"] - s1["hits"] == 1 and s2["entries"] == 1
    assert l1 - l0 >= 2 and l2 - l1 == 1  # expansion + evaluation, then one evaluation only
    assert np.array_equal(bits(a[:1000]), bits(b)) or np.abs(a[:1000] - b).max() <= 1e-14  # (kernel may differ with batch size)
    want = port.evalmlip(grid, values, x[:20000])
    assert np.abs(a[:20000] - want).max() <= TOL_REL
    values2 = values.copy(order="F")
    values2[5, 6, 7, 8] += 1.0
    c = _lib.evalmlip(grid, values2, x[:20000], 0, 1)
    assert np.abs(c - port.evalmlip(grid, values2, x[:20000])).max() <= TOL_REL
    assert _lib.cache_stats()["misses"] - s2["misses"] == 1
    _lib.cache_clear()
    assert _lib.cache_stats()["entries"] == 0


def test_lazy_tables_are_ordered_across_streams(gpu_ok, port):
    """a derived table built lazily on the first evaluation's stream must be
    complete before an evaluation on ANOTHER stream reads it (chebpol_cuda_plan_eval_device takes
    the caller's stream)."""
    import torch

    grid, values = cases.mlip_case((40,) * 4)
    x = cases.points(400_000, 4)
    want = port.evalmlip(grid, values, x[:20000])
    xd = torch.from_numpy(x).cuda()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(3):
        plan = _lib.Plan.mlip(grid, values, 0)
        plan.set(_lib.OPT_VARIANT, _lib.VARIANT_MLIP_CELL)  # expand on first use
        o1 = torch.empty(x.shape[0], dtype=torch.float64, device="cuda")
        o2 = torch.empty(x.shape[0], dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        plan.eval_device_ptr(xd.data_ptr(), x.shape[0], o1.data_ptr(), s1.cuda_stream)  # builds the table on s1
        plan.eval_device_ptr(xd.data_ptr(), x.shape[0], o2.data_ptr(), s2.cuda_stream)  # must wait for it
        torch.cuda.synchronize()
        assert plan.kernel_used == "mlip_cell"
        assert torch.equal(o1, o2)
        assert np.abs(o2[:20000].cpu().numpy() - want).max() <= TOL_REL
        plan.close()
    # the padded slab tensor of the DFMA Chebyshev kernel is built the same way
    cf = cases.cheb_fitted((10, 10, 10))
    x3 = cases.points(300_000, 3)
    x3d = torch.from_numpy(x3).cuda()
    plan = _lib.Plan.cheb(cf)
    plan.set(_lib.OPT_VARIANT, _lib.VARIANT_SLAB)
    o1 = torch.empty(x3.shape[0], dtype=torch.float64, device="cuda")
    o2 = torch.empty(x3.shape[0], dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    plan.eval_device_ptr(x3d.data_ptr(), x3.shape[0], o1.data_ptr(), s1.cuda_stream)
    plan.eval_device_ptr(x3d.data_ptr(), x3.shape[0], o2.data_ptr(), s2.cuda_stream)
    torch.cuda.synchronize()
    assert torch.equal(o1, o2)
    w3 = port.evalcheb(cf, x3[:20000])
    assert np.abs(o2[:20000].cpu().numpy() - w3).max() <= TOL_REL * np.abs(w3).max()
    plan.close()
