"""Registers / spills of every tiled-kernel instantiation from an `nvcc -Xptxas -v` log."""
import re
import sys

t = open(sys.argv[1]).read()
pat = (r"Compiling entry function '(\S+)'.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores"
       r".*?\n.*?Used (\d+) registers")
for m in re.finditer(pat, t):
    k = re.search(r"kernelILi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)ELb(\d)ELi(\d+)",
                  m.group(1))
    if k:
        print("R%s PM%s TX%s TY%s PF%s PS%s MATH%s MINB%s VD%s UNR%s" % k.groups(),
              "regs", m.group(4), "spill", m.group(3))
