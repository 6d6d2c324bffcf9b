"""One-off probe on the GPU box: Chebyshev and Floater-Hormann on the same 40^3 shape, device-resident,
to separate what the shape costs the DMMA engine from what the FH basis prologue costs."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import cases, chebpol_b200 as cb
from chebpol_b200 import _lib

dims = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "40,40,40").split(","))
n = 10_000_000
x = torch.rand((n, len(dims)), device="cuda", dtype=torch.float64) * 2 - 1
out = torch.empty(n, device="cuda", dtype=torch.float64)
peak = cb.fp64_peak_tflops(0, 5)
S = 0; p = 1
for m in reversed(dims):
    p *= m; S += p
def run(plan, name):
    for _ in range(3): plan.eval_torch(x, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): plan.eval_torch(x, out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(name, dims, plan.kernel_used, "%.4g evals/s" % (n / ms * 1e3), "frac %.3f" % (2 * S * n / (ms * 1e-3) / 1e12 / peak))
run(_lib.Plan.cheb(cases.cheb_fitted(dims)), "cheb")
vals, grid, w = cases.fh_case(dims, d=4)
run(_lib.Plan.fh(vals, grid, w), "fh  ")
