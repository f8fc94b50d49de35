#!/usr/bin/env python
"""Attribute ncu stall samples to CUDA source lines: joins `ncu --page source --csv` (SASS rows, in address order) with
`nvdisasm -g -c` of the cubin (same instructions in the same order, annotated with //## File ..., line N inlining info).
Usage: ncu_line_summary.py src.csv file.cubin <mangled-substring> [topN]"""
import csv, re, subprocess, sys, collections
src_csv, cubin, key = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and key in l)
lines = []      # (line chain string) per instruction
cur = "?"
for l in dis[start + 1:]:
    if l.startswith("//---") or l.startswith(".text."): break
    m = re.search(r"//## File \"([^\"]+)\", line (\d+)(.*)", l)
    if m:
        inl = re.findall(r"inlined at \"[^\"]+\", line (\d+)", l)
        cur = (m.group(1).split("/")[-1], int(m.group(2)), tuple(int(x) for x in inl))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
if len(lines) != len(data):
    print("warning: instruction count mismatch", len(lines), len(data))
tot = sum(float(r[idx["# Samples"]] or 0) for r in data)
by = collections.Counter(); ex = collections.Counter()
for r, ln in zip(data, lines):
    # attribute to the outermost location inside the kernel file (last inlined-at) and the innermost
    outer = ln[2][-1] if isinstance(ln, tuple) and ln[2] else (ln[1] if isinstance(ln, tuple) else -1)
    inner = f"{ln[0]}:{ln[1]}" if isinstance(ln, tuple) else "?"
    by[(outer, inner)] += float(r[idx["# Samples"]] or 0)
    ex[(outer, inner)] += float(r[idx["Instructions Executed"]] or 0)
print(f"total samples {tot:.0f}")
print("by (outermost kernel line <- innermost location):")
for (o, i), c in by.most_common(topn):
    print(f"  {c/tot*100:5.2f}%  exec={ex[(o,i)]:.3g}  kernel line {o} <- {i}")
outer = collections.Counter()
for (o, i), c in by.items(): outer[o] += c
print("by outermost kernel line:")
for o, c in outer.most_common(25): print(f"  {c/tot*100:5.2f}%  line {o}")
