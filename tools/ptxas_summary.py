"""Summarise `-Xptxas -v` logs written by mage_b200/csrc/Makefile: registers, smem, spills per kernel."""
import glob
import os
import re
import sys

d = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "..", "mage_b200", "csrc")
pat = re.compile(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores, "
                 r"(\d+) bytes spill loads\n.*?Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes smem)?")
for f in sorted(glob.glob(os.path.join(d, "*.ptxas.log"))):
    for m in pat.finditer(open(f).read()):
        name = re.sub(r"^_ZN\d+_GLOBAL__N__\w+?_cu_\w{8}\d*", "", m.group(1))[:80]
        print(f"{os.path.basename(f)[:-10]:14s} {name:82s} regs={m.group(5):>3s} smem={m.group(6)} spill={m.group(3)}/{m.group(4)}")
