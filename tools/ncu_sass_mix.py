#!/usr/bin/env python
"""Instruction mix / stall-sample share per SASS opcode from `ncu -i X.ncu-rep --page source --csv --kernel-id :::K`."""
import csv, collections, re, sys

def main(path, top=40):
    rows = list(csv.reader(open(path)))
    hdr = next(r for r in rows if r and r[0] == "Address")
    iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
    ops, smp = collections.Counter(), collections.Counter()
    tot_i = tot_s = 0
    for r in rows:
        if len(r) < 10 or not r[0].startswith("0x"):
            continue
        s = r[1].strip()
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", s)
        op = m.group(2) if m else s[:10]
        op = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "STS", "LDG", "UTC", "LDT", "MUFU", "BAR", "SYNCS")) else op.split(".")[0]
        n, k = int(r[iI] or 0), int(r[iS] or 0)
        ops[op] += n; smp[op] += k; tot_i += n; tot_s += k
    print(f"warp instructions {tot_i}, stall samples {tot_s}")
    for op, n in ops.most_common(top):
        print(op.ljust(14), f"{100*n/tot_i:5.1f}% inst  {100*smp[op]/max(tot_s,1):5.1f}% samples")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
