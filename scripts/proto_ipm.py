"""Algorithm prototype (dense KKT, numpy) used to freeze the IPM spec before writing C / CUDA.
Scratch tool, not shipped on any path.  Usage: python scripts/proto_ipm.py [nprob]
"""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import model_np as M
from forces_resilient_planner_b200 import workloads as W

TOL = 1e-4
MU_FLOOR = 1e-5


def solve(xinit, z0, hdr, rows, nrows, variant="normal", mu0=0.1, kappa=1e-2, maxit=200, verbose=False,
          mehrotra=True, ls=True, sigma_fixed=0.2):
    N = z0.shape[0]
    z = z0.copy()
    z[0, 8:] = xinit
    free = np.ones((N, 17), bool)
    free[0, 8:] = False
    lb, ub = M.LB, M.UB
    # push into interior
    for k in range(N):
        for i in range(17):
            if free[k, i]:
                pl = min(kappa * max(1, abs(lb[i])), kappa * (ub[i] - lb[i]))
                pu = min(kappa * max(1, abs(ub[i])), kappa * (ub[i] - lb[i]))
                z[k, i] = min(max(z[k, i], lb[i] + pl), ub[i] - pu)
    # corridor rows (stage>=1)
    A = [rows[k, :nrows[k], 0:3] if k > 0 else np.zeros((0, 3)) for k in range(N)]
    bb = [rows[k, :nrows[k], 3] + M.HU if k > 0 else np.zeros(0) for k in range(N)]
    s = [np.maximum(bb[k] - A[k] @ z[k, 8:11], 1e-2) for k in range(N)]   # slack floor
    zl = np.where(free, mu0 / np.maximum(z - lb, 1e-300), 0.0)
    zu = np.where(free, mu0 / np.maximum(ub - z, 1e-300), 0.0)
    lc = [mu0 / s[k] for k in range(N)]
    y = np.zeros((N, 13))     # y[k] multiplies (c(z_{k-1}) - E z_k), k>=1 ; c-ordering [x(9);u(4)]
    nineq = 2 * free.sum() + sum(len(x) for x in s)
    nz = N * 17
    iters = 0
    hist = []
    nbt_total = 0
    for it in range(maxit + 1):
        # ---- model
        f = 0.0
        g = np.zeros((N, 17)); H = np.zeros((N, 17, 17)); c = np.zeros((N, 13)); J = np.zeros((N, 13, 17))
        for k in range(N):
            fk, g[k], H[k] = M.objective(z[k], np.concatenate([hdr[k], np.zeros(120)]), k, variant, N)
            f += fk
            if k < N - 1:
                c[k], J[k] = M.dynamics(z[k], hdr[k, 3:6])
        # defects d_k = c(z_k) - E z_{k+1}
        d = np.zeros((N, 13))
        for k in range(N - 1):
            d[k] = c[k] - np.concatenate([z[k + 1, 8:17], z[k + 1, 4:8]])
        rc = [A[k] @ z[k, 8:11] - bb[k] + s[k] for k in range(N)]
        # stationarity
        rs = g - zl + zu
        for k in range(N):
            if k < N - 1:
                rs[k] += J[k].T @ y[k + 1]
            if k > 0:
                rs[k, 8:17] -= y[k, 0:9]
                rs[k, 4:8] -= y[k, 9:13]
            rs[k, 8:11] += A[k].T @ lc[k]
        rs_n = np.max(np.abs(rs[free]))
        req_n = np.max(np.abs(d[:N - 1]))
        rin_n = max([0.0] + [np.max(np.maximum(A[k] @ z[k, 8:11] - bb[k], 0), initial=0) for k in range(N)])
        sl, su = z - lb, ub - z
        comp = np.concatenate([(sl * zl)[free], (su * zu)[free]] + [s[k] * lc[k] for k in range(N)])
        mu = comp.sum() / nineq
        rcomp = comp.max()
        rcres = max([0.0] + [np.max(np.abs(r), initial=0) for r in rc])
        hist.append((rs_n, req_n, rin_n, rcomp))
        if verbose:
            print(f"it {it:3d} f={f:.6f} rs={rs_n:.2e} req={req_n:.2e} rin={rin_n:.2e} rcres={rcres:.2e} comp={rcomp:.2e} mu={mu:.2e}")
        if rs_n <= TOL and req_n <= TOL and rin_n <= TOL and rcomp <= TOL:
            return dict(z=z, it=it, flag=1, f=f, nbt=nbt_total, hist=hist)
        if it == maxit:
            break
        # ---- KKT matrix
        Phi = np.zeros((nz, nz))
        for k in range(N):
            Pk = H[k].copy()
            Sig = np.where(free[k], zl[k] / np.where(free[k], sl[k], 1) + zu[k] / np.where(free[k], su[k], 1), 0)
            Pk += np.diag(Sig)
            Pk[8:11, 8:11] += A[k].T @ np.diag(lc[k] / s[k]) @ A[k]
            Phi[k * 17:(k + 1) * 17, k * 17:(k + 1) * 17] = Pk
        neq = 9 + 13 * (N - 1)
        Je = np.zeros((neq, nz)); re = np.zeros(neq)
        Je[0:9, 8:17] = np.eye(9); re[0:9] = 0.0
        for k in range(N - 1):
            r0 = 9 + 13 * k
            Je[r0:r0 + 13, k * 17:(k + 1) * 17] = J[k]
            Je[r0:r0 + 9, (k + 1) * 17 + 8:(k + 1) * 17 + 17] -= np.eye(9)
            Je[r0 + 9:r0 + 13, (k + 1) * 17 + 4:(k + 1) * 17 + 8] -= np.eye(4)
            re[r0:r0 + 13] = d[k]
        K = np.block([[Phi, Je.T], [Je, np.zeros((neq, neq))]])

        def kkt_solve(tl, tu, tc):
            """targets t = sigma*mu - corr for each complementarity pair."""
            gt = g.copy()
            gt -= np.where(free, tl / np.where(free, sl, 1), 0)
            gt += np.where(free, tu / np.where(free, su, 1), 0)
            for k in range(N):
                gt[k, 8:11] += A[k].T @ ((tc[k] + lc[k] * rc[k]) / s[k])
            rhs = np.concatenate([-gt.reshape(-1), -re])
            sol = np.linalg.solve(K, rhs)
            dz = sol[:nz].reshape(N, 17)
            ynew = sol[nz:]
            dzl = np.where(free, (tl - zl * dz) / np.where(free, sl, 1) - zl, 0)
            dzu = np.where(free, (tu + zu * dz) / np.where(free, su, 1) - zu, 0)
            ds = [-rc[k] - A[k] @ dz[k, 8:11] for k in range(N)]
            dlc = [(tc[k] - lc[k] * ds[k]) / s[k] - lc[k] for k in range(N)]
            return dz, ynew, dzl, dzu, ds, dlc

        def max_step(dz, dzl, dzu, ds, dlc, tau):
            ap = 1.0; ad = 1.0
            m = free & (dz < 0);
            if m.any(): ap = min(ap, np.min(-tau * sl[m] / dz[m]))
            m = free & (dz > 0)
            if m.any(): ap = min(ap, np.min(tau * su[m] / dz[m]))
            for k in range(N):
                m = ds[k] < 0
                if m.any(): ap = min(ap, np.min(-tau * s[k][m] / ds[k][m]))
                m = dlc[k] < 0
                if m.any(): ad = min(ad, np.min(-tau * lc[k][m] / dlc[k][m]))
            m = free & (dzl < 0)
            if m.any(): ad = min(ad, np.min(-tau * zl[m] / dzl[m]))
            m = free & (dzu < 0)
            if m.any(): ad = min(ad, np.min(-tau * zu[m] / dzu[m]))
            return ap, ad

        zero_c = [np.zeros_like(x) for x in s]
        if mehrotra:
            dz, yn, dzl, dzu, ds, dlc = kkt_solve(np.zeros((N, 17)), np.zeros((N, 17)), zero_c)
            ap, ad = max_step(dz, dzl, dzu, ds, dlc, 1.0)
            comp_aff = np.concatenate([((sl + ap * dz) * (zl + ad * dzl))[free], ((su - ap * dz) * (zu + ad * dzu))[free]]
                                      + [(s[k] + ap * ds[k]) * (lc[k] + ad * dlc[k]) for k in range(N)])
            mu_aff = comp_aff.sum() / nineq
            sigma = min(1.0, max((mu_aff / mu) ** 3, 1e-6)) if mu > 0 else 0.0
            mu_t = max(sigma * mu, MU_FLOOR)
            tl = mu_t - dz * dzl
            tu = mu_t + dz * dzu
            tc = [mu_t - ds[k] * dlc[k] for k in range(N)]
        else:
            sigma = sigma_fixed
            mu_t = max(sigma * mu, MU_FLOOR)
            tl = np.full((N, 17), mu_t); tu = tl.copy(); tc = [np.full_like(x, mu_t) for x in s]
        dz, yn, dzl, dzu, ds, dlc = kkt_solve(tl, tu, tc)
        tau = min(max(0.995, 1 - mu), 0.99999)
        ap, ad = max_step(dz, dzl, dzu, ds, dlc, tau)
        # ---- line search (filter-lite on theta / barrier objective)
        def theta_phi(zt, st):
            ft = 0.0; th = 0.0
            for k in range(N):
                fk = M.objective(zt[k], np.concatenate([hdr[k], np.zeros(120)]), k, variant, N)[0]
                ft += fk
                if k < N - 1:
                    ck = M.dynamics(zt[k], hdr[k, 3:6], jac=False)
                    th += np.sum(np.abs(ck - np.concatenate([zt[k + 1, 8:17], zt[k + 1, 4:8]])))
                th += np.sum(np.abs(A[k] @ zt[k, 8:11] - bb[k] + st[k]))
            bar = -mu_t * (np.sum(np.log((zt - lb)[free])) + np.sum(np.log((ub - zt)[free])) + sum(np.sum(np.log(x)) for x in st))
            return th, ft + bar
        nbt = 0
        if ls:
            th0, ph0 = theta_phi(z, s)
            a = ap
            while True:
                zt = z + a * dz
                st = [s[k] + a * ds[k] for k in range(N)]
                th, ph = theta_phi(zt, st)
                if th <= (1 - 1e-5) * th0 or ph <= ph0 - 1e-5 * th0 + 1e-12 * abs(ph0):
                    break
                nbt += 1
                if nbt >= 10:
                    break
                a *= 0.5
            ap_used = a
        else:
            ap_used = ap
        nbt_total += nbt
        z = z + ap_used * dz
        s = [s[k] + ap_used * ds[k] for k in range(N)]
        zl = zl + ad * dzl; zu = zu + ad * dzu
        lc = [lc[k] + ad * dlc[k] for k in range(N)]
        ynew = np.zeros((N, 13))
        for k in range(N - 1):
            ynew[k + 1] = yn[9 + 13 * k: 9 + 13 * k + 13]
        y = y + ap_used * (ynew - y)
        if verbose:
            print(f"      sigma={sigma:.3e} ap={ap:.3f} used={ap_used:.3f} ad={ad:.3f} nbt={nbt}")
    return dict(z=z, it=maxit, flag=0, f=f, nbt=nbt_total, hist=hist)


if __name__ == "__main__":
    nprob = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    b = W.config1()
    r = solve(b.xinit[0], b.z0[0], b.hdr[0], b.rows[0], b.nrows[0], verbose=True)
    print("config1:", r["it"], r["flag"], r["f"])
    b = W.config2(nprob)
    for kw in (dict(mehrotra=True), dict(mehrotra=False)):
        its = []; flags = []; nbts = []
        for i in range(nprob):
            r = solve(b.xinit[i], b.z0[i], b.hdr[i], b.rows[i], b.nrows[i], **kw)
            its.append(r["it"]); flags.append(r["flag"]); nbts.append(r["nbt"])
        print(kw, "iters mean/max", np.mean(its), np.max(its), "ok", np.mean(flags), "backtracks", np.sum(nbts))
