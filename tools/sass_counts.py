#!/usr/bin/env python
"""Per-kernel SASS opcode counts of the shipped libscrappie_b200.so (cuobjdump, no GPU needed): the mnemonics that show
the Blackwell-native paths -- UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit), UBLKCP
(cp.async.bulk, TMA), SYNCS (mbarrier), FFMA2 / FADD2 / FMUL2 (packed fp32), MUFU -- next to registers and spills from
the ptxas logs.  Usage: python tools/sass_counts.py [out.md]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "scrappie_b200", "libscrappie_b200.so")
OPS = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "SYNCS", "LDGSTS", "FFMA2", "FADD2", "FMUL2", "FFMA", "MUFU", "REDUX"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            counts[cur][m.group(1)] += 1
            counts[cur]["_total"] += 1
    regs = {}
    for log in os.listdir(os.path.join(ROOT, "scrappie_b200", "csrc", "build")):
        if not log.endswith(".ptxas.log"):
            continue
        txt = open(os.path.join(ROOT, "scrappie_b200", "csrc", "build", log)).read()
        for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?(\d+) bytes spill stores.*?\n.*?Used (\d+) registers", txt, re.S):
            regs[m.group(1)] = (int(m.group(3)), int(m.group(2)))
    names = demangle(list(counts))
    rows = []
    for fn, c in counts.items():
        short = re.sub(r"\(.*", "", names.get(fn, fn).replace("(anonymous namespace)::", "")).replace("void ", "").replace("sb2::", "")
        r = regs.get(fn, ("?", "?"))
        rows.append((short, c["_total"], r[0], r[1]) + tuple(c[o] for o in OPS))
    rows.sort(key=lambda r: (-r[4], r[0]))
    lines = ["# SASS opcode counts per kernel (`cuobjdump -sass scrappie_b200/libscrappie_b200.so`)", "",
             "Static instruction counts of the whole kernel (all warp roles), registers / spill bytes from `ptxas -v`.", "",
             "| kernel | instr | regs | spill B | " + " | ".join(OPS) + " |", "|---|---|---|---|" + "---|" * len(OPS)]
    for r in rows:
        lines.append("| `%s` | %s |" % (r[0], " | ".join(str(x) for x in r[1:])))
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    lines += ["", "Totals over the library: " + ", ".join("%s %d" % (o, tot[o]) for o in OPS) + "; %d kernels." % len(counts)]
    text = "\n".join(lines) + "\n"
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text)
    else:
        sys.stdout.write(text)


if __name__ == "__main__":
    main()
