#!/usr/bin/env python
"""Summarise nvcc -Xptxas -v logs: kernel, registers, spills, smem.  Usage: ptxas_summary.py build/*.ptxas.log"""
import re, subprocess, sys
rows = []
for path in sys.argv[1:]:
    txt = open(path).read()
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n(?:.*\n)*?ptxas info\s+: Used (\d+) registers(.*)", txt):
        name, regs, rest = m.group(1), int(m.group(2)), m.group(3)
        blk = txt[m.start():m.end()]
        sp = re.search(r"(\d+) bytes spill stores", blk)
        sm = re.search(r"(\d+) bytes smem", rest)
        try:
            dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        except Exception:
            dem = name
        dem = re.sub(r"\(anonymous namespace\)::", "", dem)
        dem = dem.split("(")[0][-70:]
        rows.append((dem, regs, int(sp.group(1)) if sp else 0, int(sm.group(1)) if sm else 0))
for r in rows:
    print(f"{r[0]:72s} regs={r[1]:4d} spill={r[2]:5d} smem={r[3]}")
