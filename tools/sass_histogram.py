#!/usr/bin/env python
"""Opcode histogram per kernel of the shipped library (evidence for profiles/sass_rNN.txt):
    python tools/sass_histogram.py chebpol_b200/libchebpol_cuda.so > profiles/sass_r02.txt
Shows, per cubin function, the counts of the instructions that carry the design: DMMA.8x8x4 (FP64 tensor
core), UBLKCP (TMA bulk copy, cp.async.bulk), SYNCS (mbarrier), LDGSTS (cp.async), DFMA/DMUL/DADD, MUFU,
ATOMG/RED, LDG/STG/LDS/STS, SHFL."""
import collections
import re
import subprocess
import sys

KEYS = ["DMMA", "UBLKCP", "UTMALDG", "SYNCS", "LDGSTS", "DFMA", "DMUL", "DADD", "MUFU", "ATOMG", "RED", "ATOMS", "LDG", "STG",
        "LDS", "STS", "SHFL", "BAR", "UTCHMMA", "LDTM"]


def main():
    so = sys.argv[1] if len(sys.argv) > 1 else "chebpol_b200/libchebpol_cuda.so"
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    arch = collections.Counter(re.findall(r"arch = (sm_\w+)", txt))
    per, cur = {}, None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            per[cur][m.group(1)] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
    print(f"# {so}: cubin architectures {dict(arch)}; {len(per)} functions; cuobjdump -sass opcode counts")
    tot = collections.Counter()
    rows = []
    for name, nice in zip(per, demangled):
        c = per[name]
        agg = collections.Counter()
        for op, n in c.items():
            for k in KEYS:
                if op == k or op.startswith(k + "."):
                    agg[k] += n
        tot.update(agg)
        nice = re.sub(r"\(anonymous namespace\)::", "", nice)
        nice = re.sub(r"\((int|bool)\)", "", nice)
        nice = re.sub(r"\([^()]*\)$", "", nice).replace("void ", "")[:100]
        rows.append((nice, sum(c.values()), agg))
    print("total: " + ", ".join(f"{k} {tot[k]}" for k in KEYS if tot[k]))
    print()
    for nice, n, agg in sorted(rows):
        print(f"{nice}: {n} instr; " + ", ".join(f"{k} {agg[k]}" for k in KEYS if agg[k]))
    dm = collections.Counter()
    for c in per.values():
        for op, n in c.items():
            if op.startswith("DMMA") or op.startswith("UBLKCP") or op.startswith("LDGSTS") or op.startswith("SYNCS") or op.startswith("ATOMG"):
                dm[op] += n
    print("\nexact mnemonics: " + ", ".join(f"{k} {v}" for k, v in sorted(dm.items())))


if __name__ == "__main__":
    main()
