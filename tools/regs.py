"""Registers / spills per kernel instantiation from a ptxas -v log.   python tools/regs.py <log> [filter]"""
import re
import subprocess
import sys
log = open(sys.argv[1]).read()
flt = sys.argv[2] if len(sys.argv) > 2 else ""
blocks = re.findall(r"Function properties for (\S+)\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers", log)
names = subprocess.run(["c++filt"] + [b[0] for b in blocks], capture_output=True, text=True).stdout.strip().split("\n")
for b, n in zip(blocks, names):
    n = n.split("(")[0].replace("void sesrq::", "")
    if flt in n:
        print("%4s regs  stack %4s  spill st/ld %4s/%4s  %s" % (b[4], b[1], b[2], b[3], n))
