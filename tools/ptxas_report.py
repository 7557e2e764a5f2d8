#!/usr/bin/env python3
"""Registers / spills / static smem per kernel from the ptxas logs of the last build (csrc/build/*.ptxas.log)."""
import re
import subprocess
import sys
import os

root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "corticall_b200", "csrc", "build")
pat = sys.argv[1] if len(sys.argv) > 1 else ""
for f in sorted(os.listdir(root)):
    if not f.endswith(".ptxas.log"):
        continue
    log = open(os.path.join(root, f)).read()
    ents = re.findall(r"Compiling entry function '([^']+)' for 'sm_100a'\n[^\n]*\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                      r"ptxas info\s*: Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes smem)?", log)
    if not ents:
        continue
    names = subprocess.run(["c++filt"] + [e[0] for e in ents], capture_output=True, text=True).stdout.split("\n")
    for e, n in zip(ents, names):
        n = re.sub(r"\(.*", "", n.replace("(anonymous namespace)::", "")).replace("void cc::", "")
        if pat in n:
            print("%-70s regs %3s stack %4s spill %s/%s smem %s" % (n[:70], e[4], e[1], e[2], e[3], e[5] or 0))
