#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` dump: top SASS instructions by stall samples, stall-reason totals, opcode mix.
Usage: ncu -i rep --page source --csv --kernel-name regex:X > src.csv ; ncu_source_summary.py src.csv [topN]"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
def f(r, k):
    try: return float(r[idx[k]])
    except Exception: return 0.0
tot = sum(f(r, "# Samples") for r in data)
print("kernel:", rows[0][1][:100]); print("SASS instructions:", len(data), " samples:", tot)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
st = {s: sum(f(r, s) for r in data) for s in stalls}
print("stall totals:", ", ".join(f"{k[6:]}={v/tot*100:.1f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1]) if v > 0.005 * tot))
ops = collections.Counter(); opx = collections.Counter()
for r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[idx["Source"]])
    op = m.group(2).split(".")[0] if m else "?"
    ops[op] += f(r, "Instructions Executed"); opx[op] += f(r, "# Samples")
ti = sum(ops.values())
print("opcode mix (executed warp-instr %, samples %):")
for op, c in ops.most_common(18): print(f"   {op:10s} {c/ti*100:5.1f}%  {opx[op]/tot*100:5.1f}%")
print(f"shared wavefronts: {sum(f(r,'L1 Wavefronts Shared') for r in data):.3g} ideal {sum(f(r,'L1 Wavefronts Shared Ideal') for r in data):.3g}")
print("top instructions by samples:")
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:topn]:
    main = max(stalls, key=lambda s: f(r, s))
    print(f"  {f(r,'# Samples')/tot*100:5.2f}%  exec={f(r,'Instructions Executed'):.3g} {main[6:]:12s} wf={f(r,'L1 Wavefronts Shared'):.3g}/{f(r,'L1 Wavefronts Shared Ideal'):.3g}  {r[idx['Source']][:90]}")
