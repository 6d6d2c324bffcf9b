#!/usr/bin/env python
"""A/B of the DMMA engine's staggered warp groups (csrc/contract_mma.cuh) on a B200: for each shape,
lock-step passes (groups = 1) against 2 and 4 staggered groups, device-resident points, CUDA events.
Also checks that every variant gives the same values as lock-step (the sum over tensor rows is only
re-ordered at r'-group granularity) and the lock-step engine against the strict kernel on a sample.

    python tools/group_sweep.py [--quick] > gpurun_out/group_sweep.jsonl
"""
import json
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np  # noqa: E402
import torch  # noqa: E402

import cases  # noqa: E402
import chebpol_b200 as cb  # noqa: E402
from chebpol_b200 import _lib  # noqa: E402

SHAPES = [("fh", (40, 40, 40), 10_000_000), ("cheb", (10,) * 5, 20_000_000), ("cheb", (12,) * 6, 2_000_000),
          ("cheb", (40, 40, 40), 10_000_000), ("fh", (20,) * 4, 4_000_000), ("fh", (100, 100), 20_000_000),
          ("cheb", (16,) * 4, 10_000_000), ("cheb", (8,) * 6, 4_000_000), ("cheb", (24,) * 3, 20_000_000)]
GROUPS = [1, 2, 4]


def main():
    quick = "--quick" in sys.argv
    peak = cb.fp64_peak_tflops(0, 5)
    for kind, dims, n in SHAPES[:3] if quick else SHAPES:
        S, p = 0, 1
        for m in reversed(dims):
            p *= m
            S += p
        if kind == "cheb":
            plan = _lib.Plan.cheb(cases.cheb_fitted(dims))
            x = torch.rand((n, len(dims)), device="cuda", dtype=torch.float64) * 2 - 1
        else:
            vals, grid, w = cases.fh_case(dims, d=4)
            plan = _lib.Plan.fh(vals, grid, w)
            x = torch.rand((n, len(dims)), device="cuda", dtype=torch.float64) * 2 - 1
        out = torch.empty(n, device="cuda", dtype=torch.float64)
        base = None
        for grp in GROUPS:
            plan.set(_lib.OPT_VARIANT, 100000 + 1000000 * grp)
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
            rec = {"kind": kind, "dims": list(dims), "points": n, "groups": grp, "kernel": plan.kernel_used, "ms": ms,
                   "evals_per_s": n / ms * 1e3, "frac_fp64_peak": 2 * S * n / (ms * 1e-3) / 1e12 / peak}
            if base is None:
                base = out.clone()
                # lock-step engine against the strict kernel on a strided sample
                m = min(n, 20000)
                xs = x[:: n // m][:m].contiguous()
                plan.set(_lib.OPT_VARIANT, 0)
                plan.set(_lib.OPT_KERNEL, _lib.KERNEL_STRICT)
                ref = plan.eval_torch(xs)
                plan.set(_lib.OPT_KERNEL, _lib.KERNEL_AUTO)
                got = base[:: n // m][:m]
                rec["max_rel_err_vs_strict"] = float((got - ref).abs().max() / ref.abs().max())
            else:
                rec["max_abs_diff_vs_lockstep"] = float((out - base).abs().max())
                rec["max_abs_f"] = float(base.abs().max())
            print(json.dumps(rec), flush=True)
        plan.close()


if __name__ == "__main__":
    main()
