"""Static SASS size + executed instructions + samples per source region (cuda,sass view)."""
import csv, sys, subprocess, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
regions = []
for spec in sys.argv[2:]:
    name, rng = spec.split(":"); fn, rr = rng.split("@"); lo, hi = map(int, rr.split("-")); regions.append((name, fn, lo, hi))
cur_file = None; hdr = None; cur_line = None
stat = collections.defaultdict(lambda: [0, 0.0, 0.0])
fl = lambda x: float(x) if x not in ("", "-") else 0.0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; ia = r.index("Address"); continue
    if hdr is None: continue
    if r[0] not in ("", "-"):
        try: cur_line = int(r[0])
        except ValueError: cur_line = None
        continue
    if cur_line is None: continue
    # a SASS row belonging to (cur_file, cur_line)
    key = "other:" + cur_file
    for name, fn, lo, hi in regions:
        if fn == cur_file and lo <= cur_line <= hi: key = name; break
    st = stat[key]; st[0] += 1
    # columns after the second "Source" are shifted by 2 relative to hdr of line rows
    st[1] += fl(r[hdr["Instructions Executed"]]); st[2] += fl(r[hdr["# Samples"]])
ts = sum(v[0] for v in stat.values()); ti = sum(v[1] for v in stat.values()); tsm = sum(v[2] for v in stat.values())
print(f"static SASS {ts}  executed {ti:.4g}  samples {tsm:.0f}")
for k, v in sorted(stat.items(), key=lambda kv: -kv[1][2]):
    print(f"  {k:28s} static {v[0]:6d} ({100*v[0]/ts:4.1f}%)  exec {100*v[1]/ti:5.1f}%  samples {100*v[2]/tsm:5.1f}%")
