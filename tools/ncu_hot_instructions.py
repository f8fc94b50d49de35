"""The hottest SASS instructions of a kernel from an .ncu-rep (--page source --csv): instruction mix of the instructions executed at least
`frac` x the maximum count (the hot loop), and the top stall samplers."""
import csv, subprocess, sys, collections
rep = sys.argv[1]; frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
ins = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0] == 'Address' or not r[0].startswith('0x'): continue
    try: ins.append((r[idx['Source']].strip(), int(r[idx['Instructions Executed']]), int(r[idx['# Samples']] or 0)))
    except ValueError: pass
mx = max(c for _, c, _ in ins)
hot = [(s, c, n) for s, c, n in ins if c >= frac * mx]
print('# %s: %d instructions, max executed %d; hot set (>= %.2f x max): %d instructions, %d stall samples of %d' % (rep, len(ins), mx, frac, len(hot), sum(n for _, _, n in hot), sum(n for _, _, n in ins)))
mix = collections.Counter()
for s, c, n in hot:
    t = s.split()
    op = t[1] if t[0].startswith('@') else t[0]
    mix[op.split('.')[0]] += 1
print('# mix of the hot set:', ', '.join('%s %d' % kv for kv in mix.most_common(25)))
print('# top stall samples:')
for s, c, n in sorted(ins, key=lambda x: -x[2])[:25]: print('%8d samples  %12d exec  %s' % (n, c, s))
