"""Static instruction mix of the hottest loop of a kernel, from `cuobjdump -sass`.

usage: python tools/sass_loop_mix.py <object-or-library> <substring of the mangled kernel name> [planes per trip]

Finds the longest backward branch of the kernel (the unrolled plane loop of the
tiled kernels) and prints the opcode counts of its body, divided by the number
of planes per trip when given.  A way to compare formulations on the CPU box
before spending GPU time.
"""
import collections
import re
import subprocess
import sys


def kernel_sass(path, needle):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    blocks = out.split("Function : ")
    hits = [b for b in blocks[1:] if needle in b.split("\n", 1)[0]]
    if not hits:
        raise SystemExit("no kernel matching %r" % needle)
    return hits[0]


def main():
    path, needle = sys.argv[1], sys.argv[2]
    per = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    text = kernel_sass(path, needle)
    ins = []
    for line in text.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    best = None
    for addr, s in ins:
        m = re.search(r"\bBRA(?:\.U)?(?:\.\w+)*\s+(?:!?U?P\d+,\s*)?`?\(?(0x[0-9a-f]+)", s)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < addr and (best is None or addr - tgt > best[1] - best[0]):
                best = (tgt, addr)
    print(text.split("\n", 1)[0].strip()[:150])
    print("instructions in kernel: %d; loop %#x..%#x" % (len(ins), best[0], best[1]))
    body = [s for a, s in ins if best[0] <= a <= best[1]]
    mix = collections.Counter()
    for s in body:
        s = re.sub(r"^@!?U?P\d+\s+", "", s)
        op = s.split()[0]
        base = op.split(".")[0]
        if base == "IMAD" and ".MOV" in op:
            base = "IMAD.MOV"
        if base in ("LDS", "LDG", "STG", "STS", "LD", "ST"):
            base = ".".join(op.split(".")[:1]) + ("." + op.split(".")[-1] if op.split(".")[-1].isdigit() else "")
        mix[base] += 1
    print("loop body: %d instructions (%.1f per plane)" % (len(body), len(body) / per))
    for op, c in mix.most_common(28):
        print("  %-10s %5d  %7.1f" % (op, c, c / per))


if __name__ == "__main__":
    main()
