#!/usr/bin/env python
"""Small-call latency of the drop-in path and the batch size P* at which it overtakes the reference C
(VERDICT r1 item 3).  The reference's closures are routinely called on one or a few points
(tests/all.R:21-25; numvec = 1 at src/chebpol.c:368), where a GPU call is pure fixed cost.

For each BASELINE config 1-4: wall time of ONE stateless C-ABI call (pageable numpy in/out, tensor re-passed,
plan cache warm = the second and later calls of a closure) at P = 1, 1e3, 1e5, 1e6 beside the reference C
(oracle/_ref) on the same points with one thread and with all host cores; the first (cold) call; and P*.
For the 64^4 table additionally: the plan-handle API and the opt-in address-trust cache mode.

    python tools/crossover_table.py > profiles/crossover_r02.jsonl
"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def med_time(fn, reps):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts))


def run_config(cfg_id):
    import bench
    from chebpol_b200 import _lib, synth

    cfg = bench.CONFIGS[cfg_id]
    k = bench.cfg_rank(cfg)
    case = bench.build_case(cfg)
    orc, okind = bench.get_oracle()
    cores = os.cpu_count() or 1
    call = bench.make_stateless(cfg, case)
    sizes = [1, 10, 100, 1000, 10_000, 100_000, 1_000_000]
    _lib.cache_clear()
    x1 = synth.points(0, 1000, k)
    o1 = np.empty(1000)
    t0 = time.perf_counter()
    call(x1, o1, 1)
    cold_ms = (time.perf_counter() - t0) * 1e3
    rows = []
    for P in sizes:
        x = synth.points(0, P, k)
        out = np.empty(P)
        call(x, out, 1)
        reps = 200 if P <= 1000 else (30 if P <= 100_000 else 7)
        gpu = med_time(lambda: call(x, out, 1), reps)
        creps = max(3, min(reps, int(0.5 / max(1e-6, med_time(lambda: bench.oracle_eval(orc, cfg, case, x, 1), 1))) + 1))
        cpu1 = med_time(lambda: bench.oracle_eval(orc, cfg, case, x, 1), creps)
        cpuN = med_time(lambda: bench.oracle_eval(orc, cfg, case, x, cores), creps)
        want = bench.oracle_eval(orc, cfg, case, x, cores)
        rows.append({"P": P, "gpu_stateless_ms": gpu * 1e3, "cpu_1thread_ms": cpu1 * 1e3, f"cpu_{cores}threads_ms": cpuN * 1e3,
                     "max_abs_err": float(np.abs(out - want).max())})

    def crossover(key):
        prev = None
        for r in rows:
            if r["gpu_stateless_ms"] < r[key]:
                if prev is None:
                    return r["P"]
                # log-linear interpolation of the two time curves between the bracketing sizes
                a0, b0 = np.log(prev["gpu_stateless_ms"]), np.log(prev[key])
                a1, b1 = np.log(r["gpu_stateless_ms"]), np.log(r[key])
                t = (a0 - b0) / ((a0 - b0) - (a1 - b1))
                return float(np.exp(np.log(prev["P"]) + t * (np.log(r["P"]) - np.log(prev["P"]))))
            prev = r
        return None

    rec = {"cfg": cfg_id, "workload": cfg["name"], "oracle": okind, "cores": cores, "cold_first_call_P1000_ms": cold_ms,
           "rows": rows, "crossover_P_vs_1thread": crossover("cpu_1thread_ms"),
           "crossover_P_vs_all_cores": crossover(f"cpu_{cores}threads_ms"), "cache": _lib.cache_stats(),
           "cache_mode": "address+sample (opt-in)" if os.environ.get("CHEBPOL_CUDA_CACHE_TRUST_ADDRESS") else "exact content (default)"}
    if cfg["kind"] == "mlip":
        plan = bench.make_plan(cfg, case, 0)
        plan.eval_host(x1, o1)
        rec["plan_handle_P1000_ms"] = med_time(lambda: plan.eval_host(x1, o1), 200) * 1e3
        plan.close()
        vals = case["values"]
        rec["fingerprint_ms_of_the_table"] = med_time(lambda: _lib.fingerprint(vals.ravel(order="F"), min(16, max(1, cores // 2))), 5) * 1e3
        rec["table_MB"] = vals.nbytes / 1e6
    print(json.dumps(rec), flush=True)
    _lib.cache_clear()


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        return run_config(int(sys.argv[2]))
    for c in (1, 2, 3, 4):
        subprocess.run([sys.executable, __file__, "--child", str(c)], check=False)
    # the 64^4 table again with the address-trust cache key
    subprocess.run([sys.executable, __file__, "--child", "3"], env=dict(os.environ, CHEBPOL_CUDA_CACHE_TRUST_ADDRESS="1"), check=False)


if __name__ == "__main__":
    main()
