#!/usr/bin/env python
"""Trimmed SASS listing of the core library: one row per kernel with the
counts of the mnemonics that show what the kernel is made of (TMA, mbarrier,
two-wide float32 arithmetic, shared / global memory operations, fences).

    python tools/sass_listing.py [library] > profiles/rNN_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["UTMALDG", "UTMAPF", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FFMA", "DFMA", "LDS", "STS",
        "LDG", "STG", "MEMBAR", "BAR", "MUFU", "ATOMG", "REDG"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(
        REPO, "simwave_b200", "lib", "libsimwave_b200.so")
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    names = {}
    rows = []
    for block in out.split("Function : ")[1:]:
        mangled = block.split("\n", 1)[0].strip()
        ops = collections.Counter()
        n = 0
        for line in block.split("\n"):
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
            if m:
                ops[m.group(1)] += 1
                n += 1
        rows.append((mangled, n, ops))
    dem = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True,
                         text=True).stdout.split("\n")
    print("# %s: %d kernels, architectures %s" % (os.path.basename(lib), len(rows), arch))
    tot = collections.Counter()
    for r in rows:
        tot.update(r[2])
    print("# library totals: " + ", ".join("%s %d" % (k, tot[k]) for k in WANT if tot[k]))
    print("%-110s %7s " % ("kernel", "instr") + " ".join("%7s" % w for w in WANT))
    for (mangled, n, ops), name in sorted(zip(rows, dem), key=lambda x: x[1]):
        name = re.sub(r"\(.*", "", name.replace("void ", "").replace("sw::", ""))
        print("%-110s %7d " % (name[:110], n) + " ".join("%7d" % ops[w] for w in WANT))


if __name__ == "__main__":
    main()
