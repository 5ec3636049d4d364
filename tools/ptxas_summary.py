#!/usr/bin/env python
"""Registers / stack / spills per kernel from the `-Xptxas -v` logs the Makefile keeps (csrc/*.ptxas.log)."""
import glob, os, re, subprocess, sys
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mcmurchie-davidson_b200", "csrc")
rows = []
for f in sorted(glob.glob(os.path.join(root, "*.ptxas.log"))):
    txt = open(f).read().split("Compiling entry function '")
    for blk in txt[1:]:
        name = blk.split("'")[0]
        m1 = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", blk)
        m2 = re.search(r"Used (\d+) registers", blk)
        m3 = re.search(r"(\d+) bytes smem", blk)
        rows.append((name, int(m2.group(1)) if m2 else -1, int(m1.group(1)), int(m1.group(2)), int(m1.group(3)), int(m3.group(1)) if m3 else 0, os.path.basename(f)))
names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
pat = sys.argv[1] if len(sys.argv) > 1 else ""
for n, r in zip(names, rows):
    n = n.replace("mmdb::", "").replace("(mmdb::EriArgs)", "").replace("void ", "")
    if pat in n:
        print("%-60s regs=%3d stack=%5d spill_st=%5d spill_ld=%5d smem=%d  [%s]" % (n[:60], r[1], r[2], r[3], r[4], r[5], r[6]))
