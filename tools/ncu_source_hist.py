"""Opcode histogram and the most-sampled instructions of an `ncu --page source --csv` export.

usage: ncu -i x.ncu-rep --page source --csv > x.csv ; python tools/ncu_source_hist.py x.csv [kernel-index]
"""
import collections
import csv
import re
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    num = lambda s: int(s) if s.isdigit() else 0
    tot = sum(num(r[isamp]) for r in data)
    ops, samp = collections.Counter(), collections.Counter()
    for r in data:
        m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ia].strip())
        op = m.group(2).split('.')[0] if m else r[ia].strip()
        ops[op] += num(r[iex])
        samp[op] += num(r[isamp])
    te = sum(ops.values())
    print("instructions %d, executed %d, samples %d" % (len(data), te, tot))
    for op, c in ops.most_common(30):
        print("%-12s exec %5.1f%%  samples %5.1f%%" % (op, c / te * 100, samp[op] / tot * 100))
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    for r in sorted(data, key=lambda r: -num(r[isamp]))[:top]:
        why = sorted(((num(r[i]), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
        print(r[isamp], r[iex], r[ia].strip()[:80], why)


if __name__ == "__main__":
    main(sys.argv[1])
