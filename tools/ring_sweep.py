#!/usr/bin/env python
"""Default layout of the DMMA engine over a set of shapes, for one ring geometry (CHEBPOL_CUDA_MMA_RINGDIV,
CHEBPOL_CUDA_MMA_SLAB_KB are read once per process: run it once per setting).  Needs a B200."""
import json
import os
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import torch  # noqa: E402

import cases  # noqa: E402
import chebpol_b200 as cb  # noqa: E402
from chebpol_b200 import _lib  # noqa: E402

SHAPES = [("cheb", (10,) * 5, 20_000_000), ("fh", (40, 40, 40), 10_000_000), ("cheb", (12,) * 6, 2_000_000),
          ("cheb", (40, 40, 40), 10_000_000), ("fh", (20,) * 4, 4_000_000), ("cheb", (16,) * 4, 10_000_000),
          ("cheb", (8,) * 6, 4_000_000), ("cheb", (24,) * 3, 20_000_000), ("cheb", (6,) * 7, 4_000_000),
          ("cheb", (32,) * 3, 10_000_000), ("cheb", (20,) * 4, 4_000_000), ("fh", (100, 100), 20_000_000)]


def main():
    peak = cb.fp64_peak_tflops(0, 5)
    for kind, dims, n in SHAPES:
        S, p = 0, 1
        for m in reversed(dims):
            p *= m
            S += p
        if kind == "cheb":
            plan = _lib.Plan.cheb(cases.cheb_fitted(dims))
        else:
            vals, grid, w = cases.fh_case(dims, d=4)
            plan = _lib.Plan.fh(vals, grid, w)
        plan.set(_lib.OPT_VARIANT, _lib.VARIANT_MMA)
        x = torch.rand((n, len(dims)), device="cuda", dtype=torch.float64) * 2 - 1
        out = torch.empty(n, device="cuda", dtype=torch.float64)
        for _ in range(2):
            plan.eval_torch(x, out)
        torch.cuda.synchronize()
        reps = 5 if S * n < 4e12 else 2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            plan.eval_torch(x, out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(json.dumps({"ringdiv": os.environ.get("CHEBPOL_CUDA_MMA_RINGDIV", "3"), "slab_kb": os.environ.get("CHEBPOL_CUDA_MMA_SLAB_KB", "64"),
                          "kind": kind, "dims": list(dims), "points": n, "kernel": plan.kernel_used, "ms": ms, "evals_per_s": n / ms * 1e3,
                          "frac_fp64_peak": 2 * S * n / (ms * 1e-3) / 1e12 / peak, "checksum": float(out.sum())}), flush=True)
        plan.close()


if __name__ == "__main__":
    main()
