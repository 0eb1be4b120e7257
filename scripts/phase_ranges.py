"""Prints `name:file@lo-hi` phase specs for scripts/ncu_lines.py, derived from the current sources (function start lines),
so that the per-phase breakdown of an ncu report follows the code as it is edited.
  python scripts/phase_ranges.py nmpc_ipm.cuh nmpc_ipm_mixed.cuh nmpc_model.cuh"""
import os, re, sys
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "forces_resilient_planner_b200", "csrc")
pat = re.compile(r"^\s*(?:template <[^>]*>\s*)?(?:__device__|__global__)[^;]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(")
for fn in sys.argv[1:]:
    lines = open(os.path.join(root, fn)).read().splitlines()
    starts = []
    for i, ln in enumerate(lines, 1):
        m = pat.match(ln)
        if m and not ln.strip().endswith(";"):
            starts.append((i, m.group(1)))
    for (lo, name), nxt in zip(starts, starts[1:] + [(len(lines) + 1, None)]):
        if nxt[0] - lo >= 6:
            print(f"{name}:{fn}@{lo}-{nxt[0] - 1}")
