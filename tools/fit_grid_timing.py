#!/usr/bin/env python
"""Timing lines for SURVEY 8(f) rows 1 and 2: the fit step on the device and
evalongridV of an interpolant, each beside what it replaces.  One JSON line per measurement:

    python tools/fit_grid_timing.py >> profiles/bench_all_r02.jsonl

  chebcoef      chebpol_cuda_chebcoef (host array in, host array out: upload + separable DCT-II on the
                device + download) against the reference's chebcoef (src/chebpol.c:8-141, matrix path,
                through oracle/_ref) on the host
  evalongridV   chebpol_cuda_evalgrid_* (one mode product per dimension) against the same interpolant
                evaluated point by point on the expanded grid -- on the GPU (the stateless path) and by
                the reference C on all host cores (what R/tools.R:132-137 does)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def best(fn, reps=5):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        ts.append(time.perf_counter() - t0)
    return min(ts), r


def main():
    import cases
    import chebpol_b200 as cb
    from chebpol_b200 import _lib
    from oracle.oracle import Oracle

    orc = Oracle("reference") if Oracle.available("reference") else Oracle("port")
    cores = os.cpu_count() or 1
    for dims in [(20, 20), (10,) * 5, (12,) * 6, (40, 40, 40)]:
        val = cases.grid_values(cb.chebknots(dims))
        tg, cg = best(lambda: _lib.chebcoef(val))
        tc, cc = best(lambda: orc.chebcoef(val), reps=2 if np.prod(dims) > 1e6 else 5)
        print(json.dumps({"what": "chebcoef (SURVEY 8f row 1)", "dims": dims, "values": int(np.prod(dims)),
                          "gpu_ms": tg * 1e3, "gpu_path": "chebpol_cuda_chebcoef: upload + separable DCT-II on the device + download",
                          "cpu_ms": tc * 1e3, "cpu_path": f"reference chebcoef (matrix path) via oracle/{orc.kind}, 1 thread",
                          "speedup": tc / tg, "max_abs_diff": float(np.abs(cg - cc).max())}), flush=True)
    # evalongridV: 5-D Chebyshev 10^5 on a 24^5 grid, multilinear 64^4 on a 48^4 grid, FH 40^3 on a 100^3 grid
    jobs = []
    dims = (10,) * 5
    ch = cb.chebappx(cases.grid_values(cb.chebknots(dims)))
    jobs.append(("Chebyshev 10^5", ch, [np.linspace(-1, 1, 24)] * 5, lambda X: orc.evalcheb(ch.t["coef"], X, threads=cores)))
    g4, v4 = cases.mlip_case((64,) * 4)
    ml = cb.mlappx(v4, g4)
    jobs.append(("multilinear 64^4", ml, [np.linspace(-1, 1, 48)] * 4, lambda X: orc.evalmlip(g4, v4, X, threads=cores)))
    vals, gf, wf = cases.fh_case((40, 40, 40), d=4)
    fh = cb.fhappx(vals, gf, 4)
    jobs.append(("Floater-Hormann 40^3", fh, [np.linspace(-1, 1, 100)] * 3, lambda X: orc.fh(vals, gf, wf, X, threads=cores)))
    for name, ip, pts, cpu in jobs:
        npt = int(np.prod([len(p) for p in pts]))
        tg, a = best(lambda: cb.evalongridV(ip, grid=pts), reps=3)
        mesh = np.meshgrid(*pts, indexing="ij")
        X = np.stack([g.ravel(order="F") for g in mesh], axis=0)  # K x P, expand.grid order
        tp, b = best(lambda: ip(X), reps=3)
        Xc = np.ascontiguousarray(X.T)
        nc = min(npt, 200_000)
        t0 = time.perf_counter()
        c = cpu(Xc[:nc])
        tcpu = (time.perf_counter() - t0) * npt / nc
        print(json.dumps({"what": "evalongridV of an interpolant (SURVEY 8f row 2)", "interpolant": name, "grid_points": npt,
                          "gpu_mode_products_ms": tg * 1e3, "gpu_pointwise_ms": tp * 1e3,
                          "cpu_pointwise_ms_all_cores": tcpu * 1e3, "cpu_sample": nc, "cores": cores,
                          "speedup_vs_gpu_pointwise": tp / tg, "speedup_vs_cpu_pointwise": tcpu / tg,
                          "max_abs_diff_vs_pointwise": float(np.abs(np.asarray(a).ravel(order="F") - b).max()),
                          "max_abs_diff_vs_reference": float(np.abs(np.asarray(a).ravel(order="F")[:nc] - c).max())}), flush=True)
        ip.close()


if __name__ == "__main__":
    main()
