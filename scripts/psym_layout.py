"""Search for a bank-conflict-free packed layout of a symmetric 13x13 matrix (91 words).

Row-per-lane products y = P x read, for each column j, the 13 words {addr(i, j), i = 0..12} with one
lane per row.  The words of such a "star" must fall into 13 different 8-byte banks (16 per 128-byte
wavefront), for every j.  That is a proper edge colouring of K13 with loops by the 16 residues
mod 16, with residue r used exactly as often as there are addresses == r (mod 16) in 0..90
(6 times for r < 11, 5 times otherwise).  Randomised greedy + repair; prints the table
PSYM[i][j] used by csrc/nmpc_model.cuh.
"""
import random
import sys

n, nb = 13, 16
pairs = [(i, j) for i in range(n) for j in range(i + 1)]
cap = [len(range(r, 91, nb)) for r in range(nb)]


def attempt(rng):
    colour = {}
    used = [set() for _ in range(n)]          # colours present at each vertex
    left = cap[:]
    order = pairs[:]
    rng.shuffle(order)
    for (i, j) in order:
        cands = [c for c in range(nb) if left[c] > 0 and c not in used[i] and c not in used[j]]
        if not cands:
            return None
        # prefer the colour with most capacity left (keeps the tight counts feasible)
        m = max(left[c] for c in cands)
        c = rng.choice([c for c in cands if left[c] == m])
        colour[(i, j)] = c
        used[i].add(c); used[j].add(c)
        left[c] -= 1
    return colour


rng = random.Random(13)
for it in range(200000):
    col = attempt(rng)
    if col:
        break
else:
    sys.exit("no layout found")
slots = {r: list(range(r, 91, nb)) for r in range(nb)}
addr = {}
for p in pairs:
    addr[p] = slots[col[p]].pop(0)
assert sorted(addr.values()) == list(range(91))
tab = [[addr[(max(i, j), min(i, j))] for j in range(n)] for i in range(n)]
for j in range(n):
    assert len({tab[i][j] % nb for i in range(n)}) == n
print(f"// found after {it + 1} attempts")
for i in range(n):
    print("    {" + ", ".join(f"{v:2d}" for v in tab[i]) + ", 0, 0, 0},")
