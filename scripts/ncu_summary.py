"""Summarise an `ncu --page source --csv` export: stall reasons, opcode mix, hottest source lines."""
import collections
import csv
import re
import sys

path = sys.argv[1]
rows = list(csv.reader(open(path)))
hdr = rows[1]; data = rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def f(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


tot_samples = sum(f(r[col['# Samples']]) for r in data)
tot_inst = sum(f(r[col['Instructions Executed']]) for r in data)
print(f"sass lines {len(data)}  samples {tot_samples:.0f}  warp instructions executed {tot_inst:.4g}")
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {s: sum(f(r[col[s]]) for r in data) for s in stalls}
for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {s:26s} {100 * v / tot_samples:5.1f}%")
ops = collections.Counter()
for r in data:
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[col['Source']])
    if m:
        ops[m.group(2).split('.')[0]] += f(r[col['Instructions Executed']])
print("opcode mix (executed warp instructions):")
for k, v in ops.most_common(18):
    print(f"  {k:10s} {100 * v / tot_inst:5.1f}%")
if 'L1 Wavefronts Shared' in col:
    w = sum(f(r[col['L1 Wavefronts Shared']]) for r in data); wi = sum(f(r[col['L1 Wavefronts Shared Ideal']]) for r in data)
    print(f"shared wavefronts {w:.4g} (ideal {wi:.4g})")
