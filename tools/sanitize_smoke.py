"""Small run of every kernel for compute-sanitizer (memcheck): odd sizes, ragged tails."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import cases
from chebpol_b200 import _lib

def run(plan, x, variants):
    for kern, var in variants:
        plan.set(_lib.OPT_KERNEL, kern); plan.set(_lib.OPT_VARIANT, var)
        out = plan.eval_host(x)
        assert np.all(np.isfinite(out)), (kern, var)
        print("  ", plan.kernel_used, out[:2])

for dims in [(10,) * 5, (7, 5, 3), (20, 20), (13,), (6,) * 6]:
    cf = cases.cheb_fitted(dims)
    x = cases.points(1237, len(dims))
    p = _lib.Plan.cheb(cf)
    print("cheb", dims)
    run(p, x, [(0, 0), (0, _lib.VARIANT_MMA), (0, _lib.VARIANT_SLAB), (1, 0)])
    p.close()
for dims in [(9,), (7, 6), (6, 5, 4), (5, 5, 5, 5), (4, 3, 4, 3, 4), (3, 3, 3, 3, 3, 3)]:
    grid, values = cases.mlip_case(dims, uniform=False)
    x = cases.points(1237, len(dims), lo=-1.1, hi=1.1)
    for blend in (0, 3):
        p = _lib.Plan.mlip(grid, values, blend)
        print("mlip", dims, blend)
        run(p, x, [(0, _lib.VARIANT_MLIP_PAIR), (0, _lib.VARIANT_MLIP_CELL), (1, 0)])
        p.close()
for dims in [(11,), (9, 8), (12, 11, 10), (5, 4, 3, 4)]:
    vals, grid, w = cases.fh_case(dims, d=3)
    x = cases.points(1237, len(dims))
    x[:5, 0] = grid[0][:5]
    p = _lib.Plan.fh(vals, grid, w)
    print("fh", dims)
    run(p, x, [(0, 0), (1, 0)])
    p.close()
a = np.asfortranarray(np.random.default_rng(0).standard_normal((5, 4, 3)))
print("chebcoef", _lib.chebcoef(a)[0, 0, 0], _lib.fhweights([np.linspace(0, 1, 9)], [3])[0][:2])
print("stateless", _lib.evalcheb(cases.cheb_fitted((6, 5)), cases.points(300, 2), ngpus=1)[:2])
print("sanitize smoke done")
