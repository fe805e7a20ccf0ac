#!/usr/bin/env python
"""Per-kernel SASS opcode summary of libpnslam.so (cuobjdump -sass): instruction count and the mnemonics that prove
which hardware paths a kernel uses (UTCHMMA/UTCQMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit,
UBLKCP = cp.async.bulk (TMA engine, 1-D), UTMALDG/UTMASTG = cp.async.bulk.tensor, REDG = red.global, SYNCS = mbarrier).
usage: sass_summary.py LIB.so OUT.txt [--dump REGEX OUTFILE]   (the optional pair writes the full listing of the kernels
whose demangled name matches REGEX)"""
import collections, re, subprocess, sys

KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "REDG", "RED", "ATOMG", "SYNCS", "MUFU", "HMMA", "FFMA",
        "LDG", "STG", "LDS", "STS", "SHFL", "BAR", "DADD", "DMUL", "DFMA"]


def main():
    lib, out = sys.argv[1], sys.argv[2]
    dump_re = dump_out = None
    if "--dump" in sys.argv:
        i = sys.argv.index("--dump")
        dump_re, dump_out = re.compile(sys.argv[i + 1]), open(sys.argv[i + 2], "w")
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    names = {}
    try:
        import shutil
        if shutil.which("cu++filt"):
            mangled = sorted(set(re.findall(r"Function : (\S+)", txt)))
            dem = subprocess.run(["cu++filt"] + mangled, capture_output=True, text=True).stdout.splitlines()
            names = dict(zip(mangled, dem))
    except Exception:
        pass
    funcs, cur, dumping = collections.OrderedDict(), None, False
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = names.get(m.group(1), m.group(1))
            funcs[cur] = collections.Counter()
            dumping = bool(dump_re and dump_re.search(cur))
            if dumping:
                dump_out.write(f"\n// ======== {cur}\n")
            continue
        if cur is None:
            continue
        if dumping:
            dump_out.write(line + "\n")
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            funcs[cur][m.group(2)] += 1
    with open(out, "w") as f:
        f.write(f"# cuobjdump -sass {lib.split('/')[-1]} (sm_100a): static instruction counts per kernel; tools/sass_summary.py\n")
        tot = collections.Counter()
        for name, c in funcs.items():
            short = re.sub(r"\([^<>]*\)$", "", name).replace("pn::<unnamed>::", "").replace("void ", "")
            short = short.replace("(int)", "").replace("(bool)", "")
            n = sum(c.values())
            keys = "  ".join(f"{k}={c[k]}" for k in KEYS if c.get(k))
            f.write(f"{short[:70]:70s} {n:7d} instr  {keys}\n")
            tot.update(c)
        f.write("\n# whole library: " + "  ".join(f"{k}={tot[k]}" for k in KEYS if tot.get(k)) + "\n")
    print(open(out).read()[-600:])


if __name__ == "__main__":
    main()
