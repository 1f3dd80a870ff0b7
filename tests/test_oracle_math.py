"""Cross-checks of the oracle's linear algebra and closed-form models against NumPy
("cross-check the linear algebra against NumPy/SciPy")."""
import numpy as np

from scenelib2_b200 import synth


def _spd(rng, n, scale=1e-4):
    a = rng.standard_normal((n, n))
    return a @ a.T * scale + 1e-6 * np.eye(n)


def test_kalman_update_dense_matches_numpy(oracle):
    rng = np.random.default_rng(1)
    for n, m in ((25, 8), (73, 20), (163, 100)):
        P = _spd(rng, n)
        H = np.zeros((m, n))
        for k in range(m // 2):
            H[2 * k:2 * k + 2, :7] = rng.standard_normal((2, 7)) * 50
            pos = 13 + 3 * int(rng.integers(0, (n - 13) // 3))
            H[2 * k:2 * k + 2, pos:pos + 3] = rng.standard_normal((2, 3)) * 300
        R = np.diag(rng.uniform(1, 4, m))
        nu = rng.standard_normal(m)
        x = rng.standard_normal(n)
        x1, P1 = oracle.kalman_update_dense(x, P, H, R, nu)
        S = H @ P @ H.T + R
        W = P @ H.T @ np.linalg.inv(S)
        assert np.allclose(x1, x + W @ nu, rtol=1e-9, atol=1e-12)
        Pn = P - W @ S @ W.T
        d = np.sqrt(np.diag(Pn))
        assert np.abs(P1 - Pn).max() <= 1e-9 * np.abs(P).max()
        assert (np.abs(P1 - Pn) <= 1e-7 * d[:, None] * d[None, :] + 1e-18).all()


def test_motion_model_jacobian_and_Q(oracle):
    rng = np.random.default_rng(2)
    dt = 0.033333333
    xv = np.array([0.1, -0.2, -0.6, 0.9, 0.1, -0.3, 0.2, 0.05, -0.1, 0.2, 0.3, -0.2, 0.4])
    xv[3:7] /= np.linalg.norm(xv[3:7])
    fv, F, Q = oracle.motion(xv, dt)
    assert np.allclose(fv[:3], xv[:3] + xv[7:10] * dt) and np.allclose(fv[7:], xv[7:])
    assert abs(np.linalg.norm(fv[3:7]) - 1) < 1e-12
    num = np.zeros((13, 13))
    for j in range(13):
        e = np.zeros(13)
        e[j] = 1e-6
        num[:, j] = (oracle.motion(xv + e, dt)[0] - oracle.motion(xv - e, dt)[0]) / 2e-6
    assert np.abs(F - num).max() < 1e-8
    assert np.allclose(Q, Q.T, atol=1e-18) and np.linalg.eigvalsh(Q).min() > -1e-15
    # Q = G Pnn G^T with Pnn = diag(sdA^2 dt^2 x3, sdAlpha^2 dt^2 x3) (motion_model.cpp:157-160)
    assert np.isclose(Q[7, 7], 16 * dt * dt) and np.isclose(Q[10, 10], 36 * dt * dt)
    assert np.isclose(Q[0, 0], 16 * dt ** 4) and np.isclose(Q[0, 7], 16 * dt ** 3)
    # quirk Q5: omega = 0 gives NaN in F (division by |omega|)
    xv0 = xv.copy()
    xv0[10:] = 0
    assert np.isnan(oracle.motion(xv0, dt)[1]).any()


def test_dqnorm_quirk(oracle):
    xv = np.zeros(13)
    xv[3:7] = [0.8, 0.1, -0.2, 0.3]
    J = oracle.dxvnorm_by_dxv(xv)
    q = xv[3:7]
    qq = q @ q
    exp = np.eye(13)
    for i in range(4):
        for j in range(4):
            exp[3 + i, 3 + j] = (1 - q[i] ** 2 / qq ** 2) / qq if i == j else -q[i] * q[j] / qq ** 3
    assert np.allclose(J, exp, rtol=0, atol=1e-15)


def test_feature_prediction(oracle):
    rng = np.random.default_rng(3)
    cam8 = synth.camera_params(320, 240)
    xv = np.array([0.02, -0.01, -0.6, 0.995, 0.03, -0.05, 0.02, 0, 0, 0, 0, 0, 0.01])
    y = np.array([0.08, -0.05, 0.1])
    Pxx = _spd(rng, 13)
    Pyy = _spd(rng, 3)
    Pxy = rng.standard_normal((13, 3)) * 1e-5
    h, dxv, dy, R, S = oracle.predict_feature(cam8, xv, y, Pxx, Pxy, Pyy)

    def hfun(xv_, y_):
        return oracle.predict_feature(cam8, xv_, y_, Pxx, Pxy, Pyy)[0]
    xv[3:7] /= np.linalg.norm(xv[3:7])
    h, dxv, dy, R, S = oracle.predict_feature(cam8, xv, y, Pxx, Pxy, Pyy)
    for j in (0, 1, 2):
        e = np.zeros(13)
        e[j] = 1e-6
        num = (hfun(xv + e, y) - hfun(xv - e, y)) / 2e-6
        assert np.allclose(dxv[:, j], num, atol=2e-5), j
    # quaternion columns: the reference's analytic Jacobian (feature_model.cpp:187-238) is that of
    # the homogeneous rotation form, so it agrees with finite differences only along directions
    # tangent to the unit sphere (d . q = 0), which is where a normalised filter moves.
    for _ in range(4):
        d = rng.standard_normal(4)
        d -= (d @ xv[3:7]) * xv[3:7]
        d /= np.linalg.norm(d)
        e = np.zeros(13)
        e[3:7] = 1e-6 * d
        num = (hfun(xv + e, y) - hfun(xv - e, y)) / 2e-6
        assert np.allclose(dxv[:, 3:7] @ d, num, atol=2e-4)
    assert (dxv[:, 7:] == 0).all()
    for j in range(3):
        e = np.zeros(3)
        e[j] = 1e-6
        assert np.allclose(dy[:, j], (hfun(xv, y + e) - hfun(xv, y - e)) / 2e-6, atol=2e-5)
    S_np = dxv @ Pxx @ dxv.T + dxv @ Pxy @ dy.T + (dxv @ Pxy @ dy.T).T + dy @ Pyy @ dy.T + R
    assert np.allclose(S, S_np, rtol=1e-12)
    # numpy projection used by the synthetic generator agrees with the oracle's camera
    q = xv[3:7]
    assert np.allclose(h, _project_np(cam8, xv, y), atol=1e-9)
    dist = np.hypot(h[0] - cam8[4], h[1] - cam8[5]) / np.hypot(cam8[4], cam8[5])
    assert np.isclose(R[0, 0], (1 + dist) ** 2) and R[0, 1] == 0 and R[0, 0] == R[1, 1]


def _project_np(cam8, xv, y):
    w, x, yy, z = xv[3:7]
    Rwr = np.array([[1 - 2 * (yy * yy + z * z), 2 * (x * yy - z * w), 2 * (x * z + yy * w)],
                    [2 * (x * yy + z * w), 1 - 2 * (x * x + z * z), 2 * (yy * z - x * w)],
                    [2 * (x * z - yy * w), 2 * (yy * z + x * w), 1 - 2 * (x * x + yy * yy)]])
    return synth.project(cam8, Rwr.T @ (y - xv[:3]))


def test_visibility(oracle):
    cam8 = synth.camera_params(320, 240)
    xp = np.array([0, 0, -0.6, 1, 0, 0, 0.0])
    y = np.array([0.05, 0.02, 0.0])
    h = oracle.predict_feature(cam8, np.concatenate([xp, np.zeros(5), [0.01]]), y, np.eye(13),
                               np.zeros((13, 3)), np.eye(3))[0]
    assert oracle.visibility_test(cam8, xp, y, xp, h) == 0
    assert oracle.visibility_test(cam8, xp, y, xp, np.array([5.0, 100.0])) & 1
    assert oracle.visibility_test(cam8, xp, y, xp, np.array([100.0, 230.0])) & 2
    far = xp.copy()
    far[2] = -2.0
    assert oracle.visibility_test(cam8, far, y, xp, h) & 4            # distance ratio > 2
    behind = xp.copy()
    behind[2] = 0.5
    assert oracle.visibility_test(cam8, behind, y, xp, h) & 16        # behind camera
    side = np.array([0.6, 0, -0.3, 1, 0, 0, 0.0])
    assert oracle.visibility_test(cam8, side, y, xp, h) & 8           # angle > 45 deg


def test_puinv_from_S(oracle):
    rng = np.random.default_rng(4)
    for _ in range(20):
        a = rng.standard_normal((2, 2))
        S = a @ a.T + 0.1 * np.eye(2)
        p = oracle.puinv_from_S(S)
        Si = np.linalg.inv(S)
        assert np.allclose(p, [Si[0, 0], Si[0, 1], Si[1, 1]], rtol=1e-12)


def test_elliptical_search_vs_python(oracle):
    """Independent (slow, pure Python) restatement of monoslam.cpp:401-477 on small cases,
    including border clamping on all four sides, plateaus (ties -> last wins) and gating."""
    rng = np.random.default_rng(6)
    img = synth.make_texture(rng, 60, 80)
    img[5:30, 5:40] = 200      # flat: sigma < 10 -> never accepted
    B = 11
    patch = img[35:46, 50:61].copy()
    cases = [([55.2, 40.4], [0.09, 0.0, 0.09]), ([3.0, 3.0], [0.05, 0.01, 0.07]),
             ([78.6, 58.9], [0.04, -0.01, 0.06]), ([20.0, 15.0], [0.1, 0.0, 0.1]),
             ([55.0, 2.0], [0.02, 0.0, 0.3]), ([1.0, 40.0], [0.3, 0.0, 0.02])]
    for c, p in cases:
        u, v, f, best = oracle.elliptical_search(img, patch[None], np.array([c]), np.array([p]))
        eu, ev, ef, eb = _py_search(img, patch, c, p, B)
        assert (int(f[0]), best[0]) == (ef, eb)
        if eb < 1e6:
            assert (u[0], v[0]) == (eu, ev)


def _py_search(img, patch, c, p, B):
    H, W = img.shape
    P00, P01, P11 = p
    hw = int(3.0 / np.sqrt(P00 - P01 * P01 / P11))
    hh = int(3.0 / np.sqrt(P11 - P01 * P01 / P00))
    uc, vc = int(c[0] + 0.5), int(c[1] + 0.5)
    us, uf, vs, vf = -hw, hw, -hh, hh
    half = (B - 1) // 2
    if uc + us - half < 0:
        us = half - uc
    if uc + uf - half > W - B:
        uf = W - B - uc + half
    if vc + vs - half < 0:
        vs = half - vc
    if vc + vf - half > H - B:
        vf = H - B - vc + half
    best, bu, bv = 1000000.0, -1, -1
    g0 = patch.astype(np.int64)
    n = float(B * B)
    for ur in range(us, uf + 1):
        for vr in range(vs, vf + 1):
            if P00 * ur * ur + 2 * P01 * ur * vr + P11 * vr * vr < 9.0:
                x, y = uc + ur - half, vc + vr - half
                g1 = img[y:y + B, x:x + B].astype(np.int64)
                s0, s1, s01, s00, s11 = map(float, (g0.sum(), g1.sum(), (g0 * g1).sum(),
                                                    (g0 * g0).sum(), (g1 * g1).sum()))
                m0, m1 = s0 / n, s1 / n
                v0, v1 = s00 / n - m0 * m0, s11 / n - m1 * m1
                sd0, sd1 = np.sqrt(v0), np.sqrt(v1)
                if sd0 == 0 or sd1 == 0:
                    corr = 0.0 if (sd0 == 0 and sd1 == 0) else 1.0
                else:
                    k = m0 / sd0 - m1 / sd1
                    C = (s00 / v0 + s11 / v1 + n * (k * k) - s01 * 2.0 / (sd0 * sd1)
                         - s0 * 2.0 * k / sd0 + s1 * 2.0 * k / sd1)
                    corr = C / n
                if corr <= best and sd0 >= 10 and sd1 >= 10:
                    best, bu, bv = corr, ur + uc, vr + vc
    return bu, bv, int(best <= 0.40), best


def test_find_best_patch_vs_numpy(oracle):
    """N3 (monoslam.cpp:1070-1205): the oracle's running-sum restatement equals a brute-force
    evaluation (all partial sums are exact multiples of 0.25, so the order cannot matter)."""
    rng = np.random.default_rng(9)
    img = synth.make_texture(rng, 80, 100)
    I = img.astype(np.int64)
    dx = np.zeros_like(I)
    dy = np.zeros_like(I)
    dx[:, 1:-1] = I[:, 2:] - I[:, :-2]
    dy[1:-1, :] = I[2:, :] - I[:-2, :]
    for B, reg in ((11, (20, 15, 70, 60)), (15, (0, 0, 100, 80)), (11, (90, 70, 200, 200))):
        half = (B - 1) // 2
        us, vs = max(reg[0], half + 1), max(reg[1], half + 1)
        uf, vf = min(reg[2], 100 - half - 1), min(reg[3], 80 - half - 1)
        best = (0.0, -1, -1)
        for v in range(vs, vf):
            for u in range(us, uf):
                w = (slice(v - half, v + half + 1), slice(u - half, u + half + 1))
                A, C, Bm = (dx[w] ** 2).sum() / 4.0, (dy[w] ** 2).sum() / 4.0, (dx[w] * dy[w]).sum() / 4.0
                BB = np.sqrt((A + C) * (A + C) - 4 * (A * C - Bm * Bm))
                e2 = (A + C - BB) / 2.0
                if e2 > best[0]:
                    best = (e2, u, v)
        u, v, ev = oracle.find_best_patch(img, B, reg)
        assert (ev, u, v) == best
    assert oracle.find_best_patch(img, 11, (50, 50, 50, 60), ubest=3, vbest=4) == (50, 50, 0.0)


def test_score_equals_two_minus_two_rho_within_1e9(oracle):
    """Basis of the CUDA pre-filter (search.cu): for patches passing the sigma >= 10 gates the
    reference score C/n (improc.cpp:127-133) equals 2 - 2*rho of the exact integer moments up to
    FP64 rounding.  The kernel's acceptance window (1e-5) assumes |error| <= 1e-9."""
    rng = np.random.default_rng(10)
    worst = 0.0
    for B in (11, 15):
        n = B * B
        for trial in range(6):
            img = synth.make_texture(rng, 70, 90)
            if trial % 3 == 1:                       # bright, low-contrast: large mean/sigma ratio
                img = (200 + (img.astype(np.int64) - 128) // 5).clip(0, 255).astype(np.uint8)
            if trial % 3 == 2:
                img = (img.astype(np.int64) // 4 + 180).clip(0, 255).astype(np.uint8)
            py, px = int(rng.integers(0, 70 - B)), int(rng.integers(0, 90 - B))
            patch = img[py:py + B, px:px + B].copy()
            c = [px + B // 2 + 1.2, py + B // 2 - 0.7]
            box, corr, sd, inside = oracle.score_map(img, patch, c, [0.02, 0.0, 0.02])
            g0 = patch.astype(np.int64)
            s0, s00 = int(g0.sum()), int((g0 * g0).sum())
            v0 = n * s00 - s0 * s0
            if np.sqrt(v0) / n < 10:
                continue
            half = (B - 1) // 2
            for iu in range(corr.shape[0]):
                for iv in range(corr.shape[1]):
                    if sd[iu, iv] < 10:
                        continue
                    x, y = box[4] + box[0] + iu - half, box[5] + box[2] + iv - half
                    g1 = img[y:y + B, x:x + B].astype(np.int64)
                    s1, s11, s01 = int(g1.sum()), int((g1 * g1).sum()), int((g0 * g1).sum())
                    rho = np.longdouble(n * s01 - s0 * s1) / np.sqrt(np.longdouble(v0) * np.longdouble(n * s11 - s1 * s1))
                    worst = max(worst, abs(float(np.longdouble(2) - 2 * rho) - corr[iu, iv]))
    assert 0 < worst < 1e-9, worst


def test_particle_update_against_numpy(oracle):
    """N2 (monoslam.cpp:1447-1493, feature_init_info.cpp:95-172): Bayes re-weighting, normalisation,
    pruning below thr/K, re-normalisation, mean / variance of lambda — against a NumPy restatement."""
    rng = np.random.default_rng(3)
    for K in (1, 7, 100):
        h = rng.uniform(50, 150, (K, 2))
        a, b, c = rng.uniform(0.01, 0.05, K), rng.uniform(-0.005, 0.005, K), rng.uniform(0.01, 0.05, K)
        Sinv3 = np.column_stack([a, b, c])
        det = 1.0 / (a * c - b * b)
        z = np.rint(h + rng.normal(0, 4, (K, 2))).astype(np.int32)
        found = (rng.uniform(size=K) < 0.8).astype(np.uint8)
        if K == 1:
            found[:] = 1
        lam = np.linspace(0.5, 4.5, K)
        p0 = rng.uniform(0.1, 1, K)
        p0 /= p0.sum()
        left, prob, keep, cum, mv = oracle.particle_update(h, Sinv3, det, lam, z, found, 0.05, p0)
        nu = z - h
        q = a * nu[:, 0] ** 2 + 2 * b * nu[:, 0] * nu[:, 1] + c * nu[:, 1] ** 2
        w = p0 * np.where(found > 0, np.exp(-0.5 * q) / np.sqrt(2 * np.pi * det), 0.0)
        w /= w.sum()
        k2 = w >= 0.05 / K
        w2 = np.where(k2, w, 0.0)
        w2 /= w2.sum()
        assert left == int(k2.sum()) and (keep.astype(bool) == k2).all()
        np.testing.assert_allclose(prob[k2], w2[k2], rtol=1e-12)
        np.testing.assert_allclose(cum[k2], np.cumsum(w2[k2]), rtol=1e-12)
        assert (cum[~k2] == 0).all()
        mean = (w2 * lam).sum()
        np.testing.assert_allclose(mv, [mean, (w2 * lam * lam).sum() - mean * mean], rtol=1e-10, atol=1e-14)
    # all matches failed -> 0 survivors, probabilities zero (the reference deletes the feature)
    left, prob, keep, cum, mv = oracle.particle_update(h, Sinv3, det, lam, z, np.zeros(K, np.uint8), 0.05, p0)
    assert left == 0 and not keep.any() and (prob == 0).all()
