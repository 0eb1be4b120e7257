"""Stand-alone Riccati factor + KKT backsolve kernels and the pack / shift kernels, against the
oracle's Schur-complement solve and numpy restatements.  All need a GPU."""
import numpy as np
import pytest

from forces_resilient_planner_b200 import kkt, prep, workloads as W
from oracle import oracle as O
from oracle import prep_np as PN

pytestmark = pytest.mark.gpu


def _t(a, dtype=None):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


@pytest.mark.parametrize("N", [20, 40])
def test_factor_and_backsolve_match_schur_oracle(N):
    import torch
    B = 37
    phi, jc, g, d = kkt.random_kkt_problems(B, N, seed=N)
    fac, status = kkt.riccati_factor(_t(phi), _t(jc))
    assert torch.all(status == 0)
    dz, y = kkt.kkt_backsolve(fac, _t(g), _t(d))
    dz, y = dz.cpu().numpy(), y.cpu().numpy()
    for b in range(0, B, 6):
        Phi, C = kkt.dense_from_compact(phi[b], jc[b])
        rc, dz_ref, y_ref = O.kkt_solve(Phi, g[b], C, d[b, :N - 1])
        assert rc == 0
        assert np.max(np.abs(dz[b] - dz_ref)) < 1e-9 * max(1.0, np.max(np.abs(dz_ref)))
        assert np.max(np.abs(y[b, 1:] - y_ref[1:])) < 1e-8 * max(1.0, np.max(np.abs(y_ref)))
        assert np.all(y[b, 0] == 0) and np.all(dz[b, 0, 8:17] == 0)


def test_backsolve_is_linear_in_the_rhs_and_reuses_the_factor():
    """Size-independent property at the bench size: one factor, many right-hand sides."""
    import torch
    B, N = 4096, 20
    phi, jc, g, d = kkt.random_kkt_problems(B, N, seed=7)
    fac, status = kkt.riccati_factor(_t(phi), _t(jc))
    assert torch.all(status == 0)
    g1, d1 = _t(g), _t(d)
    g2, d2 = _t(np.roll(g, 1, axis=0)), _t(np.roll(d, 1, axis=0))
    z1, y1 = kkt.kkt_backsolve(fac, g1, d1)
    z2, y2 = kkt.kkt_backsolve(fac, g2, d2)
    z3, y3 = kkt.kkt_backsolve(fac, 2.0 * g1 - 0.5 * g2, 2.0 * d1 - 0.5 * d2)
    scale = float(z1.abs().max())
    assert float((z3 - (2.0 * z1 - 0.5 * z2)).abs().max()) < 1e-10 * scale
    assert float((y3 - (2.0 * y1 - 0.5 * y2)).abs().max()) < 1e-9 * float(y1.abs().max())
    # residual of the KKT system itself for a few problems: Phi dz + g + J'y+ - E'y = 0, E dz+ = J dz + d
    z1n, y1n = z1.cpu().numpy(), y1.cpu().numpy()
    for b in (0, 1234, 4095):
        Phi, C = kkt.dense_from_compact(phi[b], jc[b])
        for k in range(N):
            r = Phi[k] @ z1n[b, k] + g[b, k]
            if k < N - 1:
                r += C[k].T @ y1n[b, k + 1]
                e = np.concatenate([z1n[b, k + 1, 8:17], z1n[b, k + 1, 4:8]]) - C[k] @ z1n[b, k] - d[b, k]
                assert np.max(np.abs(e)) < 1e-10
            if k > 0:
                r[8:17] -= y1n[b, k, 0:9]; r[4:8] -= y1n[b, k, 9:13]
                assert np.max(np.abs(r)) < 1e-8
            else:
                assert np.max(np.abs(r[:8])) < 1e-8


def test_backsolve_fp32_within_tolerance():
    import torch
    B, N = 64, 20
    phi, jc, g, d = kkt.random_kkt_problems(B, N, seed=3)
    f64, _ = kkt.riccati_factor(_t(phi), _t(jc))
    z64, _ = kkt.kkt_backsolve(f64, _t(g), _t(d))
    f32, st = kkt.riccati_factor(_t(phi, torch.float32), _t(jc, torch.float32))
    assert torch.all(st == 0)
    z32, _ = kkt.kkt_backsolve(f32, _t(g, torch.float32), _t(d, torch.float32))
    # stated fp32 tolerance: 1e-4 relative to the largest step component
    assert float((z32.double() - z64).abs().max()) < 1e-4 * float(z64.abs().max())


def test_factor_flags_indefinite_blocks():
    import torch
    phi, jc, g, d = kkt.random_kkt_problems(3, 20, seed=1)
    phi[1, 5, 0:4] = -1e3          # negative curvature in the input block of one stage
    _, status = kkt.riccati_factor(_t(phi), _t(jc))
    assert status.cpu().tolist() == [0, -5, 0]


def test_pack_params_matches_reference_loop():
    rng = np.random.default_rng(0)
    B, N, P, M, mcap = 33, 20, 3, 40, 30          # polytopes with more rows than the capacity: truncated at 30
    ref_pos = rng.normal(size=(B, N, 3)); ref_yaw = rng.normal(size=(B, N)); ext = rng.normal(size=(B, 3))
    Emat = rng.normal(size=(B, N, 9)) * 0.3
    A = rng.normal(size=(B, P, M, 3)); A /= np.linalg.norm(A, axis=-1, keepdims=True)
    bb = rng.uniform(0.5, 3, (B, P, M)); pm = rng.integers(0, M + 1, (B, P)).astype(np.int32)
    pidx = rng.integers(0, P, (B, N)).astype(np.int32)
    w5 = (7.0, 1.0, 80.0, 12.0, 0.5)
    hdr, rows, nrows = prep.pack_params(_t(ref_pos), _t(ref_yaw), _t(ext), _t(Emat), _t(A), _t(bb), _t(pm), _t(pidx), w5, mcap)
    h0, r0, n0 = PN.pack_params_reference(ref_pos, ref_yaw, ext, Emat, A, bb, pm, pidx, w5, mcap)
    assert np.array_equal(nrows.cpu().numpy(), n0) and n0.max() == 30
    assert np.array_equal(hdr.cpu().numpy(), h0)
    assert np.max(np.abs(rows.cpu().numpy() - r0)) < 1e-14


def test_shift_warm_start_matches_reference_shift():
    rng = np.random.default_rng(1)
    z = rng.normal(size=(50, 20, 17)); z[:, :, 16] = rng.uniform(-2 * np.pi, 2 * np.pi, (50, 20))
    xinit, z0 = prep.shift_warm_start(_t(z), wrap_yaw=False)
    x_ref, z_ref = W.shift_warm_start(z)
    assert np.array_equal(z0.cpu().numpy(), z_ref) and np.array_equal(xinit.cpu().numpy(), x_ref)
    xinit, z0 = prep.shift_warm_start(_t(z), wrap_yaw=True)
    x_ref, z_ref = W.shift_warm_start(PN.wrap_yaw(z))          # the reference's wrap, with its own PI = 3.1415926
    assert np.allclose(z0.cpu().numpy(), z_ref, atol=0, rtol=0) and np.array_equal(xinit.cpu().numpy(), x_ref)
    zd = _t(z)
    assert prep.wrap_yaw(zd) is zd and np.array_equal(zd.cpu().numpy(), PN.wrap_yaw(z))       # updateFORCESResults, in place


def test_sample_reference_matches_getCurTraj_and_calculate_yaw():
    """Rank 3 of SURVEY §8f: reference sampling + yaw reference, against the loop restatement."""
    import torch
    rng = np.random.default_rng(5)
    B, P, N, Ts = 257, 64, 20, 0.05
    size = rng.integers(1, P + 1, B).astype(np.int32); size[:4] = (1, 2, 6, P)
    step = rng.normal(scale=0.06, size=(B, P, 3)); step[:, :, 2] *= 0.2
    step[::7] *= 0.01                                      # nearly stationary paths: the "dir too short" branch
    path = np.cumsum(step, axis=1) + rng.uniform(-3, 3, (B, 1, 3))
    path[1::5, :, 0] = -np.abs(path[1::5, :, 0]) - 1.0     # headings near +-pi: the unwrap branch
    path[1::5, :, 1] = 0.02 * np.sin(np.arange(P))[None, :] * rng.choice([-1, 1], (len(path[1::5]), 1))
    path[1::5, :, 0] -= 0.1 * np.arange(P)[None, :]
    t_off = rng.uniform(0, 2.5, B); t_off[:8] = 0.0
    last = rng.uniform(-3.1, 3.1, B)
    pos1 = path[np.arange(B), np.minimum((t_off / Ts).astype(int), size - 1)] + rng.normal(scale=0.7, size=(B, 3))
    rp, ry, far = prep.sample_reference(_t(path), torch.from_numpy(size).cuda(), _t(t_off), _t(last), N, Ts, pos1=_t(pos1))
    rp0, ry0, far0 = PN.sample_reference(path, size, t_off, last, N, Ts, pos1=pos1)
    assert np.max(np.abs(rp.cpu().numpy() - rp0)) < 1e-13
    assert np.max(np.abs(ry.cpu().numpy() - ry0)) < 1e-12
    assert np.array_equal(far.cpu().numpy(), far0) and 0 < far0.sum() < B
    assert np.any(np.abs(np.diff(ry0, axis=1)) > 1.0) or np.any(np.abs(ry0) > PN.REF_PI)   # an unwrap happened


def test_ellipsoid_propagation_matches_the_literal_restatement():
    """Rank 2 of SURVEY §8f: E_i along a plan, series formulation on the device vs the Schur/Sylvester +
    Pade restatement of setFORCESParams / updateMatrix / getDistrEllipsoid (oracle/ellipsoid_np.py)."""
    from oracle import ellipsoid_np as EN
    b = W.config2(48)
    z = O.solve_batch(b)["z"]                      # realistic plans: tilted attitudes, non-zero velocities
    rng = np.random.default_rng(2)
    z[40:] += rng.normal(scale=0.05, size=z[40:].shape)     # and some off-trajectory states
    E = prep.propagate_ellipsoids(_t(z)).cpu().numpy().reshape(48, 20, 3, 3)
    E0 = EN.propagate_batch(z)
    assert np.max(np.abs(E - E0)) < 1e-11 * np.max(np.abs(E0))
    assert np.allclose(E, np.swapaxes(E, -1, -2), atol=1e-14)               # symmetric square roots
    assert np.all(np.linalg.eigvalsh(E) > 0)
    # stage 0 is the ego ellipsoid rotated into the world frame: E_0 E_0 = R diag(r^2, r^2, h^2) R'
    R = EN.euler_to_rot(z[0, 0, 14:17])
    assert np.allclose(E[0, 0] @ E[0, 0], R @ np.diag([0.27 ** 2, 0.27 ** 2, 0.0425 ** 2]) @ R.T, atol=1e-14)
    # other constants, and the result feeds pack_params unchanged
    c = dict(mass=0.74, ext_noise_bound=0.3, Ts=0.05)
    E2 = prep.propagate_ellipsoids(_t(z[:5]), consts=c).cpu().numpy().reshape(5, 20, 3, 3)
    E20 = EN.propagate_batch(z[:5], EN.EllipsoidConsts(mass=0.74, ext_noise_bound=0.3))
    assert np.max(np.abs(E2 - E20)) < 1e-11 * np.max(np.abs(E20))
    assert np.all(np.trace(E2, axis1=-2, axis2=-1)[:, 1:] < np.trace(E[:5], axis1=-2, axis2=-1)[:, 1:])   # smaller noise bound


def test_corridor_selection_matches_the_decomp_restatement():
    """Rank 4 of SURVEY §8f: getSikangConst + EllipsoidDecomp::dilate per agent on the device vs the loop
    restatement (oracle/corridor_np.py): same polytopes (rows in the same order), same stage -> polytope map."""
    import torch
    from oracle import corridor_np as CN
    from test_prep_host import _scene
    rng = np.random.default_rng(7)
    B, N, M, P, R = 12, 20, 512, 20, 40
    refs, yaws, clouds, cn = [], [], np.zeros((B, M, 3)), np.zeros(B, np.int32)
    for a in range(B):
        ref, yaw, cloud = _scene(rng, n_pts=rng.integers(0, 450) if a else 0, clear=rng.uniform(0.02, 0.5))
        ref = ref + rng.normal(scale=0.3, size=3)          # off the centre of the clear tube
        if a == 3:                                   # obstacles inside the 5 cm seed sphere of stage 0
            cloud = np.concatenate([cloud, ref[0] + np.array([[0.05, 0.03, 0.0], [0.06, -0.02, 0.02]])])
        refs.append(ref); yaws.append(yaw); cn[a] = len(cloud); clouds[a, :len(cloud)] = cloud
    refs, yaws = np.stack(refs), np.stack(yaws)
    E = np.tile(np.diag([0.27, 0.27, 0.0425]).reshape(1, 1, 3, 3), (B, N, 1, 1)) * (1 + 0.04 * np.arange(N))[None, :, None, None]
    pa, pb, pm, pidx, npoly, ovf = prep.select_corridors(_t(clouds), torch.from_numpy(cn).cuda(), _t(refs), _t(yaws),
                                                         _t(E.reshape(B, N, 9)), max_polys=P, max_rows=R)
    pa, pb, pm, pidx, npoly, ovf = (x.cpu().numpy() for x in (pa, pb, pm, pidx, npoly, ovf))
    assert np.all(ovf == 0)
    total = 0
    for a in range(B):
        polys, idx = CN.select_corridors(refs[a], yaws[a], E[a], clouds[a, :cn[a]])
        assert npoly[a] == len(polys) and np.array_equal(pidx[a], idx), a
        for k, (A, b) in enumerate(polys):
            assert pm[a, k] == len(b), (a, k)
            assert np.max(np.abs(pa[a, k, :len(b)] - A)) < 1e-9 and np.max(np.abs(pb[a, k, :len(b)] - b)) < 1e-9, (a, k)
            assert np.all(pa[a, k, len(b):] == 0)
            total += 1
    assert total > 2 * B
    # one cloud shared by all agents gives the same answer as B copies of it
    shared = clouds[5, :cn[5]].copy()
    out_s = prep.select_corridors(_t(shared), torch.tensor([len(shared)], dtype=torch.int32).cuda(), _t(refs), _t(yaws),
                                  _t(E.reshape(B, N, 9)), max_polys=P, max_rows=R)
    rep = np.zeros((B, M, 3)); rep[:, :len(shared)] = shared
    out_r = prep.select_corridors(_t(rep), torch.full((B,), len(shared), dtype=torch.int32).cuda(), _t(refs), _t(yaws),
                                  _t(E.reshape(B, N, 9)), max_polys=P, max_rows=R)
    for x, y in zip(out_s, out_r):
        assert torch.equal(x, y)
    # the polytopes feed pack_params unchanged
    hdr, rows, nrows = prep.pack_params(_t(refs), _t(yaws), _t(np.zeros((B, 3))), _t(E.reshape(B, N, 9)), out_r[0], out_r[1],
                                        out_r[2], out_r[3], (7.0, 1.0, 80.0, 12.0, 0.5), 30)
    assert int(nrows.max()) <= 30 and int(nrows.min()) >= 6
