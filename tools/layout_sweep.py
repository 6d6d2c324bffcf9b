#!/usr/bin/env python
"""Time the DMMA contraction engine's CTA layouts on a set of tensor shapes (needs a B200).

  python tools/layout_sweep.py [npts]

Prints evaluations/s per (shape, variant); variant = 100000 + OCC*10000 + MT*100 + NW
(include/chebpol_cuda.h, CHEBPOL_CUDA_OPT_VARIANT), 0 = the library's default choice.
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from chebpol_b200 import _lib  # noqa: E402

SHAPES = [("cheb", (10,) * 5), ("cheb", (12,) * 6), ("cheb", (16,) * 4), ("cheb", (24,) * 3), ("cheb", (32,) * 3),
          ("cheb", (8,) * 6), ("cheb", (20,) * 4), ("cheb", (6,) * 7), ("cheb", (64, 64)), ("cheb", (40, 40, 40)),
          ("fh", (40, 40, 40)), ("fh", (20, 20, 20, 20)), ("fh", (100, 100))]
VARIANTS = [0, 100208, 100116, 100108, 100212, 100408]


def main():
    npts = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    rng = np.random.default_rng(0)
    for kind, dims in SHAPES:
        t = rng.standard_normal(dims)
        n = npts if np.prod(dims) < 5e5 else npts // 8
        if kind == "cheb":
            plan = _lib.Plan.cheb(t)
            x = rng.uniform(-1, 1, (n, len(dims)))
        else:
            grid = [np.linspace(0, 1, d) for d in dims]
            w = _lib.fhweights(grid, 3)
            plan = _lib.Plan.fh(t, grid, w)
            x = rng.uniform(0, 1, (n, len(dims)))
        xd = torch.from_numpy(x).cuda()
        od = torch.empty(n, dtype=torch.float64, device="cuda")
        row = []
        for v in VARIANTS:
            try:
                plan.set(_lib.OPT_VARIANT, v)
                for _ in range(2):
                    plan.eval_device_ptr(xd.data_ptr(), n, od.data_ptr())
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3):
                    plan.eval_device_ptr(xd.data_ptr(), n, od.data_ptr())
                e1.record()
                torch.cuda.synchronize()
                row.append("%d:%.3e(%s)" % (v, n / (e0.elapsed_time(e1) / 3) * 1e3, plan.kernel_used[:8]))
            except Exception as ex:  # layout not instantiated / does not fit
                row.append("%d:- [%s]" % (v, str(ex)[:60]))
        print(kind, dims, "  ".join(row), flush=True)
        plan.close()


if __name__ == "__main__":
    main()
