"""Per-source-line instruction / sample totals from `ncu --page source --print-source cuda,sass --csv`."""
import csv, sys, collections, subprocess
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 45
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None; hdr = None; lines = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; ncol = len(r); continue
    if hdr and r[0] not in ("", "-"):
        try:
            ln = int(r[0])
        except ValueError:
            continue
        def fl(x):
            try:
                return float(x)
            except ValueError:
                return 0.0
        if len(r) != ncol:
            continue
        ie = fl(r[hdr["Instructions Executed"]]); sm = fl(r[hdr["# Samples"]])
        lines.append((cur_file, ln, r[1].strip()[:90], ie, sm))
ti = sum(l[3] for l in lines); ts = sum(l[4] for l in lines)
print(f"total inst {ti:.4g} samples {ts:.0f}")
byfile = collections.defaultdict(lambda: [0, 0])
for f, ln, src, ie, sm in lines:
    byfile[f][0] += ie; byfile[f][1] += sm
for f, (ie, sm) in sorted(byfile.items(), key=lambda kv: -kv[1][1]):
    print(f"  {f:40s} inst {100*ie/ti:5.1f}%  samples {100*sm/ts:5.1f}%")
print("hottest lines by samples:")
for f, ln, src, ie, sm in sorted(lines, key=lambda l: -l[4])[:topn]:
    print(f"  {100*sm/ts:5.2f}% smp {100*ie/ti:5.2f}% inst  {f}:{ln}  {src}")
# phase breakdown for nmpc_ipm.cuh by line ranges given as extra args: name:lo-hi
if len(sys.argv) > 3:
    print("phase breakdown (samples / instructions):")
    for spec in sys.argv[3:]:
        name, rng = spec.split(":"); fn, rr = rng.split("@"); lo, hi = map(int, rr.split("-"))
        sm = sum(l[4] for l in lines if l[0] == fn and lo <= l[1] <= hi); ie = sum(l[3] for l in lines if l[0] == fn and lo <= l[1] <= hi)
        print(f"  {name:28s} {100*sm/ts:5.1f}% smp {100*ie/ti:5.1f}% inst")
