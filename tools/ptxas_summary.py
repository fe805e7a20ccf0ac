#!/usr/bin/env python
"""Summarise `nvcc -Xptxas -v` output: one line per kernel (regs / spills / stack / smem)."""
import re, subprocess, sys
txt = sys.stdin.read()
names = re.findall(r"Compiling entry function '(\S+)'", txt)
blocks = re.split(r"ptxas info\s+: Compiling entry function ", txt)[1:]
dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines() if names else []
for name, blk in zip(dem, blocks):
    short = re.sub(r"pn::\(anonymous namespace\)::", "", name)
    short = re.sub(r"\(.*", "", short).replace("void ", "")
    regs = re.search(r"Used (\d+) registers", blk)
    spill = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", blk)
    smem = re.search(r"(\d+) bytes smem", blk)
    print(f"{short:70s} regs={regs.group(1) if regs else '?':>4} stack={spill.group(1) if spill else '?':>5} "
          f"spill={spill.group(2) if spill else '?'}/{spill.group(3) if spill else '?'} smem={smem.group(1) if smem else 0}")
for l in txt.splitlines():
    if "error" in l or "warning" in l:
        print(l[:240])
