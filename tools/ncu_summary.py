#!/usr/bin/env python
"""Distil an .ncu-rep (captured on the GPU box, read here with `ncu -i`) into a small text
summary for profiles/: key raw metrics, stall mix and the hottest SASS lines.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/prof_r01.txt [units_per_launch [launch_index]]

launch_index picks one launch of a report that holds several (0 = the first); the output file is appended to
when launch_index > 0, so a report of a multi-kernel path becomes one summary file.
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sass__inst_executed_shared_loads", "sass__inst_executed_global_loads",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    units = float(sys.argv[3]) if len(sys.argv) > 3 else None
    which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    raw = ncu_csv(rep, "raw")
    hdr, unit, val = raw[0], raw[1], raw[2 + which]
    m = {h: (v, u) for h, u, v in zip(hdr, unit, val)}
    lines = [f"# ncu summary of {rep.split('/')[-1]} (ncu --set full --clock-control none, one launch)",
             f"kernel: {m.get('Kernel Name', ('?',))[0]}", ""]
    for k in KEYS:
        if k in m:
            lines.append(f"{k:86s} {m[k][0]:>18s} {m[k][1]}")
    lines.append("")
    lines.append("stall reasons (warp-cycles per issued instruction):")
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            v = float(m[h][0])
            if v >= 0.02:
                lines.append(f"  {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:24s} {v:8.3f}")
    if units:
        def num(k):
            return float(m[k][0]) if k in m else 0.0

        def to_bytes(k):
            v, u = m.get(k, ("0", "byte"))
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)
            return float(v) * scale

        tr = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
        lines += ["", f"units (points) in this launch: {units:.0f}",
                  f"DRAM traffic per launch: {tr:.4g} B = {tr / units:.1f} B/point"]
    src = ncu_csv(rep, "source")
    starts = [i for i, r in enumerate(src) if r and r[0] == "Kernel Name"]
    lo = starts[which] if which < len(starts) else 0
    hi = starts[which + 1] if which + 1 < len(starts) else len(src)
    src = src[lo:hi]
    shdr, data = src[1], [r for r in src[2:] if len(r) > 10]
    ix = {h: i for i, h in enumerate(shdr)}
    tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
    lines += ["", f"hottest SASS lines (of {tot} warp samples):"]
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:25]:
        s = int(r[ix["# Samples"]] or 0)
        lines.append(f"  {100.0 * s / max(tot, 1):5.2f}%  {r[ix['Source']][:90]}")
    byop = {}
    for r in data:
        toks = r[ix["Source"]].split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        op = op.split(".")[0]
        byop[op] = byop.get(op, 0) + int(r[ix["# Samples"]] or 0)
    lines += ["", "samples by opcode:"]
    for k, v in sorted(byop.items(), key=lambda kv: -kv[1])[:14]:
        lines.append(f"  {100.0 * v / max(tot, 1):5.2f}%  {k}")
    open(dst, "a" if which > 0 else "w").write("\n".join(lines) + "\n\n")
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
