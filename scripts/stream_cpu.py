import sys
sys.path.insert(0, ".")
import numpy as np
from oracle import oracle as O
from forces_resilient_planner_b200 import stream as ST, prep, workloads as W
B = 256; b = W.config2(B)
rng = np.random.Generator(np.random.PCG64(W.SEED + 5))
ext = b.hdr[:, 0, 3:6].copy()
A = b.rows[:, 1, :, 0:3][:, None]; braw = (b.rows[:, 1, :, 3] + np.linalg.norm(b.rows[:, 1, :, 0:3] * W.EGO_E, axis=-1))[:, None]
pm = b.nrows[:, 1:2].astype(np.int32); pidx = np.zeros((B, b.N), np.int32)
E = np.tile(np.diag(W.EGO_E).reshape(1, 1, 9), (B, b.N, 1))
xinit, z0 = b.xinit.copy(), b.z0.copy()
mu0w = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
hist = []
for step in range(int(sys.argv[2]) if len(sys.argv) > 2 else 120):
    ref, yaw, ext = ST.synthetic_refs(b, step, rng, ext)
    hdr, rows, nrows = prep.pack_params_reference(ref, yaw, ext, E, A, braw, pm, pidx, (7.0, 1.0, 80.0, 12.0, 0.5), b.mcap)
    cb = W.Batch(xinit, z0, hdr, rows, nrows, 0)
    c = O.solve_batch(cb, opts=O.default_opts(mu0=1.0 if step == 0 else mu0w))
    bad = np.nonzero(c["flag"] != 1)[0]
    hist.append((step, c["it"].mean(), c["it"].max(), len(bad)))
    if len(bad) and step < 60:
        i = bad[0]
        print("step", step, "bad", bad[:5], "flags", c["flag"][bad[:5]], "it", c["it"][bad[:5]], "res", c["info_real"][i][:6], "xinit vel", xinit[i, 3:6], "pos-ref0", xinit[i, :3] - ref[i, 0])
    z = c["z"].copy()
    # failure policy: cold restart at predicted state
    for i in bad:
        z[i] = 0; z[i, :, 3] = 7.3; z[i, :, 7] = 7.3; z[i, :, 8:17] = c["z"][i, 1, 8:17]
    xinit, z0 = W.shift_warm_start(z)
h = np.array(hist)
print("mean it %.2f  max it %d  failed %d  steps with it>40: %d" % (h[1:, 1].mean(), h[:, 2].max(), h[:, 3].sum(), (h[:, 2] > 40).sum()))
print(h[:45].astype(int).tolist())
