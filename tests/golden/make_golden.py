#!/usr/bin/env python
"""tests/golden/make_golden.py -- regenerates tests/golden/ref_vectors.npz.

TEST INFRASTRUCTURE.  Run in the build container, where /root/reference is mounted:

    make -C oracle            # builds oracle/_ref/libchebpol_ref.so from the reference's own C
    python tests/golden/make_golden.py

Every output in the fixture comes from the REFERENCE's code (src/chebpol.c, src/stalker.c,
compiled unmodified through oracle/shim), never from this repository's kernels or from the
restatement.  The fixture is what travels: the GPU box has no /root/reference, so
tests/test_golden.py checks the restatement (CPU) and the CUDA kernels (GPU) against these
stored vectors.  Inputs are stored next to the outputs, so nothing depends on a random
generator's stream staying stable across numpy versions.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.oracle import Oracle  # noqa: E402


def main():
    if not os.path.exists("/root/reference/src/chebpol.c"):
        raise SystemExit("needs /root/reference (the fixture is made from the reference's own code)")
    ref = Oracle("reference")
    rng = np.random.default_rng(20261020)
    out = {}

    def pts(n, k, lo=-1.0, hi=1.0):
        return np.ascontiguousarray(rng.uniform(lo, hi, (n, k)))

    # ---- Chebyshev: random decaying tensors, points inside and outside [-1,1]
    for name, dims, lo, hi in [("cheb_a", (20, 20), -1.0, 1.0), ("cheb_b", (8, 7, 6), -1.2, 1.2),
                               ("cheb_c", (6, 5, 4, 3, 2), -1.0, 1.0), ("cheb_d", (37,), -1.1, 1.1)]:
        c = rng.uniform(-1, 1, dims) / (1.0 + np.indices(dims).sum(axis=0)) ** 2
        c = np.asfortranarray(c)
        x = pts(257, len(dims), lo, hi)
        out[name + "_coef"] = c
        out[name + "_x"] = x
        out[name + "_y"] = ref.evalcheb(c, x)

    # ---- chebcoef (matrix path of the reference, src/chebpol.c:65-122)
    v = np.asfortranarray(rng.standard_normal((5, 4, 3)))
    out["coef_val"] = v
    out["coef_out"] = ref.chebcoef(v)
    out["coef_dct"] = ref.chebcoef(v, dct=True)

    # ---- multilinear: non-uniform grid, every blend, points reaching outside the grid
    grid = [np.sort(np.concatenate([[-1.0, 1.0], rng.uniform(-1, 1, n - 2)])) for n in (9, 6, 5)]
    vals = np.asfortranarray(rng.random((9, 6, 5)))
    x = pts(257, 3, -1.05, 1.05)
    x[:9, 0] = grid[0]  # exact knot hits
    for d, g in enumerate(grid):
        out[f"mlip_grid{d}"] = g
    out["mlip_values"] = vals
    out["mlip_x"] = x
    for blend in range(6):
        xi = x if blend not in (1, 2) else np.clip(x, -0.999, 0.999)  # exp(2-1/w) overflows outside
        out[f"mlip_y_blend{blend}"] = ref.evalmlip(grid, vals, xi, blend=blend)
    out["mlip_x_inside"] = np.clip(x, -0.999, 0.999)

    # ---- Floater-Hormann: uniform, cubed and decreasing Chebyshev grids (tests/all.R:100-101)
    n = (11, 9, 8)
    fg = [np.linspace(-1, 1, n[0]), np.linspace(-1, 1, n[1]) ** 3,
          np.cos(np.pi * (2 * np.arange(1, n[2] + 1) - 1) / (2 * n[2]))]
    fw = ref.fhweights(fg, [3, 4, 4])
    fv = np.asfortranarray(rng.standard_normal(n))
    x = pts(257, 3)
    x[:8, 2] = fg[2]  # knot hits on the decreasing grid
    x[8:19, 0] = fg[0]
    for d in range(3):
        out[f"fh_grid{d}"] = fg[d]
        out[f"fh_w{d}"] = fw[d]
    out["fh_vals"] = fv
    out["fh_x"] = x
    out["fh_y"] = ref.fh(fv, fg, fw, x)

    # ---- polyharmonic: odd, even and negative k
    kn = np.asfortranarray(rng.uniform(-1, 1, (3, 60)))
    w = rng.standard_normal(60)
    lw = rng.standard_normal(4)
    x = pts(129, 3)
    out["polyh_knots"], out["polyh_w"], out["polyh_lw"], out["polyh_x"] = kn, w, lw, x
    for k in (3, 2, 1, -1.5):
        out[f"polyh_y_k{k}"] = ref.evalpolyh(kn, w, lw, k, x)

    # ---- stalker / hstalker: the reference's own fit and evaluation
    sg = [np.sort(np.concatenate([[-1.0, 1.0], rng.uniform(-1, 1, m - 2)])) for m in (7, 6)]
    mesh = np.meshgrid(*sg, indexing="ij")
    sv = np.asfortranarray(np.exp(-mesh[0] ** 2) * np.cos(2 * mesh[1]) + 0.05 * rng.standard_normal((7, 6)))
    x = pts(129, 2, -1.02, 1.02)
    for d in range(2):
        out[f"stalk_grid{d}"] = sg[d]
    out["stalk_val"], out["stalk_x"] = sv, x
    b, c, r = ref.makestalk(sv, sg)
    out["stalk_b"], out["stalk_c"], out["stalk_r"] = b, c, r
    a, hb, hc, hd = ref.makehyp(sv, sg)
    out["hyp_a"], out["hyp_b"], out["hyp_c"], out["hyp_d"] = a, hb, hc, hd
    for blend in (0, 3, 4, 5):
        out[f"stalk_y_blend{blend}"] = ref.evalstalk(sv, sg, b, c, r, x, blend=blend)
        out[f"hyp_y_blend{blend}"] = ref.evalhyp(sv, sg, a, hb, hc, hd, x, blend=blend)

    # ---- simplex-linear: the reference's R_analyzesimplex / R_evalsl (LAPACK dgetrf/dgetrs restated
    # in oracle/ref_wrap.c); triangulation by qhull through scipy, stored with the vectors
    from scipy.spatial import Delaunay

    kn = np.asfortranarray(rng.uniform(-1, 1, (3, 80)))
    dtri = np.asfortranarray((Delaunay(kn.T, qhull_options="Qt Pp").simplices.T + 1).astype(np.float64))
    kv = rng.standard_normal(80)
    x = pts(257, 3, -1.1, 1.1)
    x[:80] = kn.T
    lu, piv, bb = ref.analyzesimplex(dtri.astype(np.int32), kn)
    out["sl_knots"], out["sl_dtri"], out["sl_val"], out["sl_x"] = kn, dtri, kv, x
    out["sl_lumats"], out["sl_pivots"], out["sl_bbox"] = lu, piv.astype(np.float64), bb
    for epol in (0, 1):
        for blend in (0, 3, 4, 5):
            out[f"sl_y_epol{epol}_blend{blend}"] = ref.evalsl(x, kn, dtri.astype(np.int32), (lu, piv, bb), kv, epol, blend)

    # ---- the deprecated recursive stalker evaluator (reference's evalstalker, no restatement exists)
    from chebpol_b200 import interp

    og = [np.linspace(-1.0, 1.0, 6), np.sort(np.concatenate([[-1.0, 1.0], rng.uniform(-1, 1, 3)]))]
    ov = np.asfortranarray(rng.standard_normal((6, 5)))
    ox = pts(129, 2, -1.02, 1.02)
    out["oldst_grid0"], out["oldst_grid1"], out["oldst_val"], out["oldst_x"] = og[0], og[1], ov, ox
    for tag, r in (("r2", [2.0, 2.0]), ("r15", [1.5, 1.0]), ("r0", [0.0, 2.0])):
        t = interp.oldstalkerappx(ov, og, r=r).t
        for blend in (0, 3):
            out[f"oldst_y_{tag}_blend{blend}"] = ref.evalstalker_old(t["values"], t["grid"], t["degree"], *t["ctx"],
                                                                     t["uniform"], ox, blend=blend)

    path = os.path.join(HERE, "ref_vectors.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
