"""SASS digest of libnmpc_b200.so (cuobjdump -sass): per kernel the instruction count and the mnemonics that prove what the
code uses -- TMA bulk copies (UBLKCP), mbarrier traffic (SYNCS), fp64 / fp32 FMA, shared-memory and global accesses,
CTA barriers, warp shuffles.  Run in the build container (no GPU needed):  python scripts/sass_digest.py > profiles/r02_sass_digest.md"""
import collections, os, re, subprocess, sys
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "forces_resilient_planner_b200", "libnmpc_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
cur = None
for ln in res.splitlines():
    m = re.match(r"\s*Function (\S+):", ln)
    if m:
        cur = m.group(1); continue
    if cur and "REG:" in ln:
        usage[cur] = ln.strip(); cur = None
kern = None
counts = collections.OrderedDict()
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        kern = m.group(1); counts[kern] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if kern and m:
        op = m.group(1)
        counts[kern]["_total"] += 1
        counts[kern][op.split(".")[0]] += 1
        if op.startswith("SYNCS") or op.startswith("UBLKCP") or op.startswith("BAR"):
            counts[kern][op] += 1
def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    except Exception:
        return n
print("# SASS digest of libnmpc_b200.so (sm_100a; `cuobjdump -sass`, `cuobjdump -res-usage`)\n")
print("| kernel | SASS instr | UBLKCP (TMA bulk) | SYNCS (mbarrier) | DFMA | FFMA | LDS | STS | LDG | BAR.SYNC | SHFL | MUFU | resources |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for k, c in counts.items():
    if c["_total"] < 50:
        continue
    syncs = sum(v for o, v in c.items() if o.startswith("SYNCS."))
    print(f"| `{demangle(k)[:110]}` | {c['_total']} | {c['UBLKCP']} | {syncs} | {c['DFMA']} | {c['FFMA']} | {c['LDS']} | {c['STS']} | {c['LDG']} | {c['BAR']} | {c['SHFL']} | {c['MUFU']} | {usage.get(k, '')} |")
print("\nmbarrier / TMA mnemonics seen: " + ", ".join(sorted({o for c in counts.values() for o in c if o.startswith(('SYNCS.', 'UBLKCP'))})))
