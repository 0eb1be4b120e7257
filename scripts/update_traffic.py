"""Rewrites profiles/{fused,backsolve}_kernel_traffic.json (read by bench.py for `roofline*.traffic`) from the ncu digests of
the current round: python scripts/update_traffic.py profiles/r02_fused_fp64_summary.txt profiles/r02_backsolve_summary.txt"""
import json, os, re, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def dram(path):
    out = {}
    for ln in open(path):
        m = re.match(r"\s*dram__bytes_(read|write)\.sum \[(\w+)\] = ([0-9.]+)", ln)
        if m:
            out[m.group(1)] = int(round(float(m.group(3)) * UNIT[m.group(2)]))
    return out


for path, name, kernel, batch, note in (
        (sys.argv[1], "fused_kernel_traffic.json", "nmpc_ipm_kernel<double,20>", 4096,
         "writes (11.5 MB of results) were still resident in L2 when the kernel ended"),
        (sys.argv[2], "backsolve_kernel_traffic.json", "kkt_backsolve_kernel<double,20>", 16384,
         "the P slices are fetched twice by TMA (second time an L2 hit): DRAM still sees every factor byte once")):
    d = dram(path)
    rec = {"kernel": kernel, "batch": batch, "dram_bytes_read": d["read"], "dram_bytes_write": d["write"],
           "dram_bytes_per_launch": d["read"] + d["write"],
           "source": f"ncu --set full --clock-control none, digest {os.path.relpath(path, ROOT)} (round 2, current kernel)", "note": note}
    json.dump(rec, open(os.path.join(ROOT, "profiles", name), "w"), indent=1)
    print(name, rec["dram_bytes_per_launch"])
