import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, cases, chebpol_b200 as cb
from chebpol_b200 import _lib
dims = (10,) * 5
ch = cb.chebappx(cases.grid_values(cb.chebknots(dims)))
m = 24
pts = [np.linspace(-1, 1, m)] * 5
t0 = time.perf_counter(); a = cb.evalongridV(ch, grid=pts); t1 = time.perf_counter()
a = cb.evalongridV(ch, grid=pts); t2 = time.perf_counter()
mesh = np.meshgrid(*pts, indexing="ij"); X = np.stack([g.ravel(order="F") for g in mesh], axis=0)
b = ch(X); t3 = time.perf_counter(); b = ch(X); t4 = time.perf_counter()
print(f"5-D Chebyshev 10^5 on a {m}^5 = {m**5} point grid: mode products {1e3*(t2-t1):.1f} ms, pointwise (GPU) {1e3*(t4-t3):.1f} ms, max diff {np.abs(a.ravel(order='F')-b).max():.2e}")
