#!/usr/bin/env python
"""Multilinear kernel variants on tables larger than L2: 1 pair gather on the values as
given, 3 row-pair-interleaved quads (2x), 4 binned / cache-blocked path (1x + workspace), 2 bricks of 2^m corners
(2^m x; CHEBPOL_CUDA_CELL_M pins m, m = rank is the cell-major table).  One process per setting:

    CHEBPOL_CUDA_CELL_M=3 python tools/mlip_variants.py [--shape 64,64,64,64] [--points 5e7] [--variants 1,3,4,2]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch

    from chebpol_b200 import _lib, synth

    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="64,64,64,64")
    ap.add_argument("--points", type=float, default=5e7)
    ap.add_argument("--variants", default="1,3,2")
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    dims = tuple(int(t) for t in args.shape.split(","))
    k, npts = len(dims), int(args.points)
    rng = np.random.default_rng(44)
    grid = [np.linspace(-1.0, 1.0, n) for n in dims]
    values = np.asfortranarray(rng.random(dims))
    plan = _lib.Plan.mlip(grid, values, 0)
    x = torch.empty((npts, k), dtype=torch.float64, device="cuda")
    _lib.synth_points_device(x.data_ptr(), 0, npts, k)
    out = torch.empty(npts, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream()
    from oracle.oracle import Oracle

    orc = Oracle("reference") if Oracle.available("reference") else Oracle("port")
    idx = np.arange(0, npts, max(1, npts // 20000))[:20000]
    want = orc.evalmlip(grid, values, synth.points_at(idx, k), threads=os.cpu_count() or 1)
    alg = 8.0 * (k + 1 + 2 ** k)
    for v in [int(t) for t in args.variants.split(",")]:
        plan.set(_lib.OPT_VARIANT, v)
        try:
            for _ in range(3):
                plan.eval_device_ptr(x.data_ptr(), npts, out.data_ptr(), st.cuda_stream)
        except _lib.ChebpolCudaError as e:
            print(json.dumps({"shape": dims, "variant": v, "error": str(e)}), flush=True)
            continue
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(args.steps):
            plan.eval_device_ptr(x.data_ptr(), npts, out.data_ptr(), st.cuda_stream)
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        err = float(np.abs(out[torch.from_numpy(idx).cuda()].cpu().numpy() - want).max())
        rate = npts / (ms * 1e-3)
        print(json.dumps({"shape": dims, "points": npts, "variant": v, "kernel": plan.kernel_used, "ms": ms, "evals_per_s": rate,
                          "alg_GBs": rate * alg / 1e9, "frac_of_6450.9": rate * alg / 1e9 / 6450.9, "max_abs_err": err,
                          "table_MB": values.nbytes / 1e6, "plan_MB": plan.get(_lib.GET_TENSOR_BYTES) / 1e6,
                          "cell_m": os.environ.get("CHEBPOL_CUDA_CELL_M", "auto"),
                          "cell_max_gb": os.environ.get("CHEBPOL_CUDA_CELL_MAX_GB", "16 (default)")}), flush=True)
    plan.close()


if __name__ == "__main__":
    main()
