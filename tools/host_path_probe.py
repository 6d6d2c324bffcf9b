"""Probe of the host-pointer path: pageable vs pinned throughput, small-call latency."""
import sys, time
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import cases, chebpol_b200 as cb
from chebpol_b200 import _lib

def t(f, n=3):
    f(); best = 1e9
    for _ in range(n):
        t0 = time.perf_counter(); f(); best = min(best, time.perf_counter() - t0)
    return best

cf = cases.cheb_fitted((10,) * 5)
for npts in (1, 1000, 100000, 10_000_000):
    x = cases.points(npts, 5)
    dt = t(lambda: _lib.evalcheb(cf, x, ngpus=1))
    print(f"stateless evalcheb 10^5 tensor, {npts:>9} pts: {dt*1e3:8.3f} ms  ({npts/dt:.3e} evals/s)")
plan = _lib.Plan.cheb(cf)
for npts in (1, 1000, 100000):
    x = cases.points(npts, 5)
    out = np.empty(npts)
    dt = t(lambda: plan.eval_host(x, out))
    print(f"plan.eval_host, {npts:>9} pts: {dt*1e3:8.3f} ms")
grid, values = cases.mlip_case((64,) * 4)
pm = _lib.Plan.mlip(grid, values, 0)
n = 50_000_000
xp = torch.empty((n, 4), dtype=torch.float64, pin_memory=True); xp.uniform_(-1, 1)
op = torch.empty(n, dtype=torch.float64, pin_memory=True)
xg = xp.numpy().copy(); og = np.empty(n)
pm.eval_host_ptr(xp.data_ptr(), n, op.data_ptr())
dt = t(lambda: pm.eval_host_ptr(xp.data_ptr(), n, op.data_ptr()))
print(f"mlip 64^4, {n} pts, pinned  : {dt*1e3:8.1f} ms  {n/dt:.3e} evals/s  {(n*40)/dt/1e9:.1f} GB/s over PCIe")
dt = t(lambda: pm.eval_host(xg, og))
print(f"mlip 64^4, {n} pts, pageable: {dt*1e3:8.1f} ms  {n/dt:.3e} evals/s  {(n*40)/dt/1e9:.1f} GB/s")
assert np.array_equal(og, op.numpy())
x1 = cases.points(1000, 4)
dt = t(lambda: _lib.evalmlip(grid, values, x1, 0, 1))
print(f"stateless evalmlip 64^4 (134 MB upload), 1000 pts: {dt*1e3:8.2f} ms")
