#!/usr/bin/env python
"""Condense the ncu launch-list CSV (tools/profile.sh launches ...) into profiles/launches_*.csv:
one row per kernel launch of the process, in order, with its duration and whether it is this
library's kernel or torch's (synthetic-input generation).

    python tools/launch_list.py gpurun_out/launches_cfg2_r01d.csv profiles/launches_cfg2_r01.csv "<command>"
"""
import csv
import sys

OURS = ("contract_mma", "cheb_", "mlip_", "fh_", "polyh", "stalk", "sl_", "oldstalk", "dfma_peak", "dmma_peak", "fit_", "chebcoef")


def main():
    src, dst = sys.argv[1], sys.argv[2]
    cmd = sys.argv[3] if len(sys.argv) > 3 else ""
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    out, tot, mine = [], 0.0, 0.0
    for r in rows[1:]:
        if r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        name = r[ix["Kernel Name"]]
        owner = "ours" if any(k in name for k in OURS) else "torch"
        tot += ms
        mine += ms if owner == "ours" else 0.0
        out.append((r[ix["ID"]], name[:90], r[ix["Grid Size"]], r[ix["Block Size"]], ms, owner))
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv {cmd}\n")
        f.write(f"# every launch of the process in order; ms per launch (cold-cache, serialised). total {tot:.2f} ms, "
                f"library kernels {mine:.2f} ms\n")
        f.write("id,kernel,grid,block,ms,owner\n")
        w = csv.writer(f)
        for o in out:
            w.writerow([o[0], o[1], o[2], o[3], f"{o[4]:.4f}", o[5]])
    print(f"{len(out)} launches, total {tot:.2f} ms, ours {mine:.2f} ms -> {dst}")


if __name__ == "__main__":
    main()
