#!/usr/bin/env python
"""What bounds the path R takes: one process, PAGEABLE numpy arrays, the stateless C ABI.
Sweeps the host threads that feed the pinned staging buffers and the non-temporal copy switch; every
setting runs in its own process (the library reads both when it is loaded).

    python tools/dropin_probe.py [--config 3] [--points 5e7] [--ngpus 1] [--threads 1,2,4,8,12,16]
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def child(cfg_id, npts, ngpus):
    import numpy as np

    import bench
    from chebpol_b200 import _lib

    cfg = bench.CONFIGS[cfg_id]
    k = bench.cfg_rank(cfg)
    case = bench.build_case(cfg)
    call = bench.make_stateless(cfg, case)
    x = bench.fill_pageable(npts, k)
    out = np.zeros(npts)
    call(x, out, ngpus)
    t0 = time.perf_counter()
    steps = 5
    for _ in range(steps):
        call(x, out, ngpus)
    dt = (time.perf_counter() - t0) / steps
    # the same batch from PINNED memory through the same stateless call (no staging copies)
    import torch

    xp = torch.empty((npts, k), dtype=torch.float64, pin_memory=True)
    op = torch.empty(npts, dtype=torch.float64, pin_memory=True)
    xp.numpy()[:] = x
    call(xp.numpy(), op.numpy(), ngpus)
    t0 = time.perf_counter()
    for _ in range(steps):
        call(xp.numpy(), op.numpy(), ngpus)
    dtp = (time.perf_counter() - t0) / steps
    print(json.dumps({"cfg": cfg_id, "ngpus": ngpus, "points": npts, "host_threads": os.environ.get("CHEBPOL_CUDA_HOST_THREADS"),
                      "nt_stores": os.environ.get("CHEBPOL_CUDA_STAGING_NT", "1"), "pageable_evals_per_s": npts / dt,
                      "pageable_host_GBs": npts * (k + 1) * 8 / dt / 1e9, "pinned_evals_per_s": npts / dtp,
                      "pinned_host_GBs": npts * (k + 1) * 8 / dtp / 1e9, "cores": os.cpu_count()}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=3)
    ap.add_argument("--points", type=float, default=5e7)
    ap.add_argument("--ngpus", type=int, default=1)
    ap.add_argument("--threads", default="1,2,4,8,12,16")
    ap.add_argument("--child", action="store_true")
    args = ap.parse_args()
    if args.child:
        return child(args.config, int(args.points), args.ngpus)
    for nt in ("1", "0"):
        for t in args.threads.split(","):
            env = dict(os.environ, CHEBPOL_CUDA_HOST_THREADS=t, CHEBPOL_CUDA_STAGING_NT=nt)
            subprocess.run([sys.executable, __file__, "--child", "--config", str(args.config), "--points", str(args.points),
                            "--ngpus", str(args.ngpus)], env=env, check=False)


if __name__ == "__main__":
    main()
