#!/usr/bin/env python
"""Hot source lines from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-id :::K`."""
import csv, sys

def main(path, top=40, key="inst"):
    rows = list(csv.reader(open(path)))
    cur, hdr, out = None, None, []
    for r in rows:
        if not r: continue
        if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
        if r[0] == "Line No": hdr = r; continue
        if hdr and r[0].isdigit():
            num = lambda x: int(x) if x.strip().isdigit() else 0
            out.append((cur, int(r[0]), r[1].strip(), num(r[hdr.index("# Samples")]), num(r[hdr.index("Instructions Executed")])))
    ts, ti = sum(o[3] for o in out), sum(o[4] for o in out)
    print(f"lines {len(out)}  samples {ts}  warp-inst {ti}")
    k = 4 if key == "inst" else 3
    for f, ln, src, s, i in sorted(out, key=lambda o: -o[k])[:top]:
        print(f"{f[:16]:16s} {ln:5d} {100*i/ti:5.1f}%i {100*s/ts:5.1f}%s  {src[:120]}")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40, sys.argv[3] if len(sys.argv) > 3 else "inst")
