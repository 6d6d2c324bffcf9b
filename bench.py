#!/usr/bin/env python
"""bench.py -- interpolant evaluations/sec (FP64) on B200, BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C] [--impl ours|reference]

A "step" is one evaluation pass of the fitted interpolant over one batch of synthetic
points.  Default workload = BASELINE.json configs[1]: 5-D Chebyshev, dims = rep(10,5),
1e8 uniform random points per GPU (weak scaling: every rank evaluates its own 1e8 points
against its replica of the coefficient tensor; no data-path collective).

One JSON line is printed by rank 0:
  value      evaluations/s over all GPUs, inputs resident in HBM, CUDA-event timed
  e2e        the same metric through the host-pointer C ABI (chebpol_cuda_plan_eval_host)
             with pinned HOST buffers: H2D of the points and D2H of the results are inside
             the timed region
  roofline   the dominant kernel against the FP64-FMA roofline (Chebyshev, FH) or the HBM
             roofline (multilinear); the FP64 peak is measured in-run with a DFMA-only
             kernel (MEASURED_PEAKS.json has no FP64 figure), HBM peak from that file
  cpu_baseline  the reference's own C evaluator (oracle/_ref, compiled from
             /root/reference/src/chebpol.c) or its C restatement, all host cores, on a
             bounded prefix of the same points

--impl reference times only that CPU evaluator (rank 0), same metric / config / unit.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "interpolant evaluations/sec (FP64) at 1/2/4/8 B200 vs FP64/HBM roofline"  # BASELINE.json "metric", verbatim
UNIT = "evals/s"

CONFIGS = {
    1: dict(kind="cheb", dims=(20, 20), points=1_000_000, name="cfg1: 2-D Chebyshev dims=c(20,20), 1e6 points"),
    2: dict(kind="cheb", dims=(10,) * 5, points=100_000_000, name="cfg2: 5-D Chebyshev dims=rep(10,5), 1e8 points"),
    3: dict(kind="mlip", dims=(64,) * 4, points=100_000_000, name="cfg3: 4-D multilinear 64^4 grid, 1e8 points"),
    4: dict(kind="fh", dims=(40, 40, 40), points=10_000_000, name="cfg4: 3-D Floater-Hormann dims=c(40,40,40) d=4, 1e7 points"),
    6: dict(kind="polyh", dims=(6,), rank=6, nknots=1000, k=3, points=20_000_000,
            name="cfg6 (SURVEY 8f row 4): 6-D polyharmonic spline, 1000 knots, k=3, 2e7 points"),
    7: dict(kind="hstalk", dims=(64, 64, 64), points=50_000_000,
            name="cfg7 (SURVEY 8f row 3): 3-D hyperbolic stalker spline on a 64^3 grid, cubic blend, 5e7 points"),
    8: dict(kind="sl", dims=(3,), rank=3, nknots=20000, points=50_000_000,
            name="cfg8 (beyond SURVEY 8f): 3-D simplex-linear on the Delaunay triangulation of 20000 scattered knots, cubic blend, 5e7 points"),
    10: dict(kind="sl", dims=(2,), rank=2, nknots=100000, points=100_000_000,
             name="cfg10 (beyond SURVEY 8f): 2-D simplex-linear on the Delaunay triangulation of 100000 scattered knots, cubic blend, 1e8 points"),
    9: dict(kind="chebg", dims=(10,) * 5, points=100_000_000,
            name="cfg9 (beyond SURVEY 8f): 5-D 'general' method (hyman-spline grid maps + Chebyshev) on an irregular 10^5 grid, 1e8 points"),
    5: dict(kind="cheb", dims=(12,) * 6, points=20_000_000, name="cfg5: 6-D Chebyshev dims=rep(12,6); default 2e7 points per GPU, --total-points 1000000000 for the full sharded batch"),
}


def steps_per_eval(dims):
    """S = sum_k prod_{j>=k} n_j: coefficient-touching steps of the nested recursion."""
    s, p = 0, 1
    for n in reversed(dims):
        p *= n
        s += p
    return s


def load_traffic(kernel, cfg_id, npts):
    """DRAM bytes per launch from the committed ncu capture of this kernel (None if not captured for this config)."""
    p = os.path.join(ROOT, "profiles", "traffic_r01.json")
    try:
        with open(p) as f:
            t = json.load(f).get(kernel)
        if t and t.get("cfg") == cfg_id:
            return t["bytes_per_point"] * npts
    except (OSError, ValueError):
        pass
    return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


# ------------------------------------------------------------------ workloads
def build_case(cfg):
    """Host tensors of the named interpolant (fit side: host numpy, untimed)."""
    import cases

    kind, dims = cfg["kind"], cfg["dims"]
    if kind == "cheb":
        return dict(coef=cases.cheb_fitted(dims))
    if kind == "chebg":
        from chebpol_b200 import interp

        grid = [np.linspace(-1.0, 1.0, n) ** 3 if d % 2 == 0 else np.linspace(-1.0, 1.0, n) for d, n in enumerate(dims)]
        ip = interp.chebappxg(cases.grid_values(grid), grid)
        return dict(coef=ip.t["coef"], splines=ip.t["splines"])
    if kind == "mlip":
        grid, values = cases.mlip_case(dims)
        return dict(grid=grid, values=values)
    if kind == "hstalk":
        from chebpol_b200 import interp

        grid = [np.linspace(-1.0, 1.0, n) for n in dims]
        val = cases.grid_values(grid)
        return dict(grid=grid, val=val, tables=interp.makehyp(val, grid))
    if kind == "polyh":
        knots, w, lw = cases.polyh_case(cfg["rank"], cfg["nknots"], cfg["k"])
        return dict(knots=knots, w=w, lw=lw, k=cfg["k"])
    if kind == "sl":
        from chebpol_b200 import interp

        knots, dtri, val = cases.sl_case(cfg["rank"], cfg["nknots"])
        return dict(knots=knots, dtri=dtri, val=val, adata=interp.analyzesimplex(dtri, knots))
    vals, grid, weights = cases.fh_case(dims, d=4)
    return dict(vals=vals, grid=grid, weights=weights)


def make_plan(cfg, case, device):
    from chebpol_b200 import _lib

    if cfg["kind"] == "cheb":
        return _lib.Plan.cheb(case["coef"], device=device)
    if cfg["kind"] == "chebg":
        return _lib.Plan.chebg(case["coef"], case["splines"], device=device)
    if cfg["kind"] == "mlip":
        return _lib.Plan.mlip(case["grid"], case["values"], 0, device=device)
    if cfg["kind"] == "hstalk":
        return _lib.Plan.stalk(1, case["val"], case["grid"], case["tables"], 3, device=device)
    if cfg["kind"] == "polyh":
        return _lib.Plan.polyh(case["knots"], case["w"], case["lw"], case["k"], device=device)
    if cfg["kind"] == "sl":
        # the candidate grid is part of the plan (untimed, like the fit): sized as for a long-lived
        # interpolant instead of letting it refine itself in the middle of the timed steps
        p = _lib.Plan.sl(case["knots"], case["dtri"], case["adata"], case["val"], device=device)
        ns, dim = case["dtri"].shape[1], cfg["rank"]
        cells = int(np.ceil(min((8.0 if dim <= 2 else 128.0) * ns, float(1 << 24)) ** (1.0 / dim)))  # the library's own upper size (sl_wanted_G)
        return p.set(_lib.OPT_SL_CELLS, int(os.environ.get("CHEBPOL_CUDA_SL_CELLS", cells)))
    return _lib.Plan.fh(case["vals"], case["grid"], case["weights"], device=device)


def oracle_eval(orc, cfg, case, x, threads):
    if cfg["kind"] == "cheb":
        return orc.evalcheb(case["coef"], x, threads=threads)
    if cfg["kind"] == "chebg":
        return orc.evalcheb(case["coef"], orc.splinemap(x, case["splines"], threads=threads), threads=threads)
    if cfg["kind"] == "mlip":
        return orc.evalmlip(case["grid"], case["values"], x, threads=threads)
    if cfg["kind"] == "hstalk":
        return orc.evalhyp(case["val"], case["grid"], *case["tables"], x, blend=3, threads=threads)
    if cfg["kind"] == "polyh":
        return orc.evalpolyh(case["knots"], case["w"], case["lw"], case["k"], x, threads=threads)
    if cfg["kind"] == "sl":
        return orc.evalsl(x, case["knots"], case["dtri"], case["adata"], case["val"], False, 3, threads=threads)
    return orc.fh(case["vals"], case["grid"], case["weights"], x, threads=threads)


def get_oracle():
    from oracle.oracle import Oracle

    if Oracle.available("reference"):
        return Oracle("reference"), "reference"
    return Oracle("port"), "port"


def cpu_points(npts, rank_dims, seed=42):
    rng = np.random.Generator(np.random.Philox(key=seed))
    return np.ascontiguousarray(rng.random((npts, rank_dims)) * 2.0 - 1.0)


def time_cpu(orc, cfg, case, x, threads):
    t0 = time.perf_counter()
    oracle_eval(orc, cfg, case, x, threads)
    return time.perf_counter() - t0


def cpu_sample_size(orc, cfg, case, threads, target_s):
    """Calibrate a prefix length that takes about target_s seconds of CPU work."""
    k = cfg.get("rank", len(cfg["dims"]))
    n = 2000
    x = cpu_points(n, k)
    dt = max(time_cpu(orc, cfg, case, x, threads), 1e-4)
    n2 = int(min(max(n * 0.5 / dt, 2000), 4e6))  # ~0.5 s probe (first-touch and thread start-up amortised)
    x = cpu_points(n2, k)
    dt = max(time_cpu(orc, cfg, case, x, threads), 1e-4)
    return int(min(max(n2 * target_s / dt, 2000), 4e8))


# --------------------------------------------------------------- clock sampler
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [s.strip() for s in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                if int(parts[0]) != self.gpu_index:
                    continue
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[4:8]):
                if v == "Active":
                    reasons.add(nm)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except (ValueError, IndexError):
            pass
    return local_rank


# ------------------------------------------------------------------ reference arm
def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    orc, kind = get_oracle()
    case = build_case(cfg)
    threads = os.cpu_count() or 1
    k = cfg.get("rank", len(cfg["dims"]))
    n = min(cpu_sample_size(orc, cfg, case, threads, target_s=args.ref_seconds), 50_000_000)
    x = cpu_points(n, k)
    for _ in range(args.warmup):
        time_cpu(orc, cfg, case, x[: max(1000, n // 8)], threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_eval(orc, cfg, case, x, threads)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["name"], "points_per_step": n,
                   "note": "reference C evaluator (OpenMP over points) on the host cores; bounded sample of the workload"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{n} points/step x {args.steps} steps, first points of the seed-42 stream"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------ ours
def run_ours(args, cfg):
    import torch

    import chebpol_b200 as cb
    from chebpol_b200 import _lib, shard

    rank, local_rank, world = shard.dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; chebpol_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        shard.init_process_group("nccl")
    npts = args.points if args.points else cfg["points"]
    if args.total_points:  # strong scaling: a fixed batch split into contiguous shards, one per GPU
        npts = (args.total_points + world - 1) // world
    k = cfg.get("rank", len(cfg["dims"]))
    case = build_case(cfg)
    plan = make_plan(cfg, case, local_rank)
    if args.variant:
        plan.set(_lib.OPT_VARIANT, args.variant)
    if args.strict:
        plan.set(_lib.OPT_KERNEL, _lib.KERNEL_STRICT)

    # synthetic points generated on the device, per-rank seed (weak scaling: npts per GPU)
    g = torch.Generator(device=dev).manual_seed(42 + rank)
    x = torch.empty((npts, k), dtype=torch.float64, device=dev)
    blk = 1 << 24
    for lo in range(0, npts, blk):
        hi = min(npts, lo + blk)
        x[lo:hi] = torch.rand((hi - lo, k), generator=g, device=dev, dtype=torch.float64) * 2.0 - 1.0
    out = torch.empty(npts, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev)

    def step():
        plan.eval_device_ptr(x.data_ptr(), npts, out.data_ptr(), stream.cuda_stream)

    wpts = min(npts, args.warmup_points) if args.warmup_points else npts
    for _ in range(args.warmup):
        plan.eval_device_ptr(x.data_ptr(), wpts, out.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize(dev)

    # FP64 peak (roofline denominator), measured now so clocks match the run
    fp64_peak = cb.fp64_peak_tflops(local_rank, 5) if rank == 0 else 0.0
    torch.cuda.synchronize(dev)

    sampler = ClockSampler(physical_gpu_index(local_rank))
    shard.barrier()
    torch.cuda.synchronize(dev)
    sampler.start()
    l0 = cb.launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record(stream)
    for i in range(args.steps):
        step()
        evs[i + 1].record(stream)
    torch.cuda.synchronize(dev)
    shard.barrier()
    launches = int(shard.sum_over_ranks(cb.launch_count() - l0, dev))
    clocks = sampler.stop()
    total_ms = evs[0].elapsed_time(evs[-1])
    step_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    total_ms = shard.max_over_ranks(total_ms, dev)
    value = npts * world * args.steps / (total_ms * 1e-3)
    kernel_ms = float(np.mean(step_ms))  # one kernel launch per step
    kernel_used = plan.kernel_used

    # ---- parity spot check (untimed): strided sample against the oracle
    parity = None
    if rank == 0:
        orc, okind = get_oracle()
        idx = torch.arange(0, npts, max(1, npts // 2000), device=dev)[:2000]
        xs = x[idx].cpu().numpy()
        want = oracle_eval(orc, cfg, case, xs, os.cpu_count() or 1)
        got = out[idx].cpu().numpy()
        fin = np.isfinite(want)  # simplex-linear: NA outside the convex hull, on both sides
        parity = {"n": int(len(want)), "max_abs_err": float(np.abs(got[fin] - want[fin]).max()) if fin.any() else 0.0,
                  "max_abs_f": float(np.abs(want[fin]).max()) if fin.any() else 0.0, "oracle": okind,
                  "non_finite": int((~fin).sum()), "non_finite_agree": bool(np.array_equal(fin, np.isfinite(got)))}

    # ---- end to end through the host-pointer C ABI, pinned host buffers
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    epts = min(npts, args.e2e_points) if args.e2e_points else npts  # bounds the pinned host memory per rank
    hx = torch.empty((epts, k), dtype=torch.float64, pin_memory=True)
    ho = torch.empty(epts, dtype=torch.float64, pin_memory=True)
    hx.copy_(x[:epts])
    torch.cuda.synchronize(dev)
    del x, out
    torch.cuda.empty_cache()
    plan.eval_host_ptr(hx.data_ptr(), min(epts, 1 << 22), ho.data_ptr())  # warm the pipeline buffers
    shard.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        plan.eval_host_ptr(hx.data_ptr(), epts, ho.data_ptr())
    e2e_s = time.perf_counter() - t0
    shard.barrier()
    e2e_s = shard.max_over_ranks(e2e_s, dev)
    e2e_value = epts * world * e2e_steps / e2e_s

    # ---- CPU baseline: rank 0, N = 1 only
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        orc, okind = get_oracle()
        threads = os.cpu_count() or 1
        n = min(cpu_sample_size(orc, cfg, case, threads, target_s=10.0), epts)
        xs = hx[:n].numpy()
        dt = time_cpu(orc, cfg, case, np.ascontiguousarray(xs), threads)
        cpu_baseline = {"value": n / dt, "unit": UNIT, "cores": threads, "kind": okind,
                        "sample": f"first {n} points of the rank-0 batch, {dt:.1f} s wall"}

    if rank == 0:
        peaks, peak_src = load_peaks()
        if cfg["kind"] == "hstalk":
            # 2^D corners, each one packed record of 3D+1 doubles, plus the point and the result.  The
            # 21 MB record table is L2-resident, so this is the L2-side byte count set against the HBM
            # peak: a reference line for the gather rate, not a bound.
            byts = 8.0 * (k + 1 + 2 ** k * (3 * k + 1))
            achieved = byts * npts / (kernel_ms * 1e-3) / 1e9
            roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": achieved / peaks["hbm_gbs"], "traffic": load_traffic(kernel_used, args.config, npts),
                        "peak_source": f"MEASURED_PEAKS.json ({peak_src}); bytes are gathered from L2, not HBM",
                        "bytes_per_eval": byts, "kernel": kernel_used, "kernel_ms": kernel_ms}
        elif cfg["kind"] == "sl":
            # streamed point + result, plus the one packed simplex record that is accepted; the candidate
            # records rejected on the way and the cell lists are L2-side traffic.  A reference line for
            # a gather-latency kernel, not a bound.
            n1 = k + 1
            byts = 8.0 * (k + 1) + 8.0 * (((2 * k + n1 * n1 + n1 + 1) + 3) // 4 * 4)
            achieved = byts * npts / (kernel_ms * 1e-3) / 1e9
            roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": achieved / peaks["hbm_gbs"], "traffic": load_traffic(kernel_used, args.config, npts),
                        "peak_source": f"MEASURED_PEAKS.json ({peak_src}); the simplex records are gathered from L2",
                        "bytes_per_eval": byts, "kernel": kernel_used, "kernel_ms": kernel_ms}
        elif cfg["kind"] in ("cheb", "chebg", "fh", "polyh"):
            # polyharmonic: per knot a subtract and a multiply-add per coordinate plus the weight
            # multiply-add, 3*rank + 2 flop as the reference executes them; the basis function
            # (sqrt / log / exp) is not counted.  Half of those instruction slots are not FMAs, so
            # this fraction tops out near 0.45; ncu's FP64-pipe-active figure is in profiles/.
            flop = (2.0 * steps_per_eval(cfg["dims"]) if cfg["kind"] != "polyh"
                    else float(cfg["nknots"]) * (3.0 * cfg["rank"] + 2.0))
            achieved = flop * npts / (kernel_ms * 1e-3) / 1e12
            # "tensor": the kernel's multiply-adds run as DMMA (FP64 tensor-core instructions); the DFMA
            # kernels are bounded by the same FP64 peak (B200's DMMA and DFMA rates are equal)
            bound = "tensor" if kernel_used in ("cheb_mma", "fh_mma") else "fp64"
            roofline = {"bound": bound, "precision": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                        "frac": achieved / fp64_peak if fp64_peak else None,
                        "traffic": load_traffic(kernel_used, args.config, npts),
                        "peak_source": "best of a DFMA-only and a DMMA-only kernel measured in this run (MEASURED_PEAKS.json has no FP64 peak); "
                                       "nominal 148 SM x 64 FMA/clk x 2 x 1.965 GHz = 37.2",
                        "flop_per_eval": flop, "kernel": kernel_used, "kernel_ms": kernel_ms}
        else:
            byts = 8.0 * (k + 1 + 2 ** k)
            achieved = byts * npts / (kernel_ms * 1e-3) / 1e9
            roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": achieved / peaks["hbm_gbs"], "traffic": load_traffic(kernel_used, args.config, npts),
                        "peak_source": f"MEASURED_PEAKS.json ({peak_src})",
                        "bytes_per_eval": byts, "kernel": kernel_used, "kernel_ms": kernel_ms}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.total_points else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["name"], "points_per_gpu": npts, "parallelism": f"points sharded over {world} GPU(s), tensor replicated",
                       **({"total_points": args.total_points} if args.total_points else {}),
                       **({"warmup_points": wpts} if wpts != npts else {}),
                       "l2": "inputs larger than L2 (point batch %.1f GB)" % (npts * k * 8 / 1e9) if npts * k * 8 > 126e6
                       else "inputs smaller than L2 (latency regime, reported not targeted)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": epts * k * 8, "d2h_bytes_per_step": epts * 8,
                    **({"points_per_gpu": epts} if epts != npts else {}),
                    "steps": e2e_steps, "api": "chebpol_cuda_plan_eval_host (pinned host buffers)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "parity": parity,
        }
        print(json.dumps(line), flush=True)
    plan.close()
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--points", type=int, default=0, help="override points per GPU")
    ap.add_argument("--total-points", type=int, default=0, help="strong scaling: this many points split over the GPUs")
    ap.add_argument("--warmup-points", type=int, default=0, help="points per warm-up step (default: a full step)")
    ap.add_argument("--e2e-points", type=int, default=0, help="cap the points per GPU of the end-to-end leg (pinned host memory)")
    ap.add_argument("--variant", type=int, default=0, help="kernel tuning variant (0 = default)")
    ap.add_argument("--strict", action="store_true", help="use the reference-order kernels")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--ref-seconds", type=float, default=8.0, help="--impl reference: CPU seconds per step the sample is sized for")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        return run_reference(args, cfg)
    return run_ours(args, cfg)


if __name__ == "__main__":
    sys.exit(main())
