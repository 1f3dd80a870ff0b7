"""GPU parity, patch search: CUDA path through the C ABI vs the CPU oracle.  Bit-exact (integer
match positions, found flags AND the FP64 score bits)."""
import os

import numpy as np
import pytest

from gpu_util import ctx_for_image, ctx_from_scenes, random_puinv, sl2, synth

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _compare_search(oracle, ctx, image, patches, feat, centres, pu):
    u, v, f, best = ctx.patch_search(0, 0, feat, centres, pu)
    ou, ov, of, obest = oracle.elliptical_search(image, patches[feat], centres, pu)
    assert (f == of).all()
    assert best.tobytes() == obest.tobytes()          # corrmax bits, incl. the 1e6 sentinel
    acc = obest < 1e6
    assert (u[acc] == ou[acc]).all() and (v[acc] == ov[acc]).all()
    assert (u[~acc] == -1).all() and (v[~acc] == -1).all()
    return f


def test_c2_search_bit_exact(oracle):
    sc = synth.make_scene("C2", n_frames=3)
    ctx = ctx_from_scenes([sc])
    rng = np.random.default_rng(1)
    for t in range(3):
        ctx.set_frame(0, 0, sc.frames[t])
        n = sc.n_features
        centres = sc.pix + rng.uniform(-4, 4, (n, 2))
        pu = np.tile([9 / 400.0, 0.0, 9 / 400.0], (n, 1))
        pu[::5] = random_puinv(rng, len(pu[::5]), 6, 20, iso_fraction=0.2)
        f = _compare_search(oracle, ctx, sc.frames[t], sc.patches, np.arange(n, dtype=np.int32),
                            centres, pu)
        assert f.mean() > 0.9
    ctx.close()


def test_c3_full_size_search_bit_exact(oracle):
    sc = synth.make_scene("C3", n_frames=1)
    assert sc.patches.shape == (100, 15, 15) and sc.frames.shape[1:] == (480, 640)
    ctx = ctx_from_scenes([sc])
    ctx.set_frame(0, 0, sc.frames[0])
    n = sc.n_features
    centres = sc.pix + np.random.default_rng(2).uniform(-10, 10, (n, 2))
    pu = np.tile([9 / 1600.0, 0.0, 9 / 1600.0], (n, 1))
    f = _compare_search(oracle, ctx, sc.frames[0], sc.patches, np.arange(n, dtype=np.int32),
                        centres, pu)
    assert f.all()
    ctx.close()


@pytest.mark.parametrize("B", [11, 15])
def test_score_map_bits_borders_ties_gating(oracle, B):
    rng = np.random.default_rng(3)
    img = synth.make_texture(rng, 120, 160)
    img[30:75, 40:100] = 200          # flat plateau: sigma < 10 (gated) and exact ties
    img[90:120, 0:30] = (np.arange(30)[None, :] * 8).astype(np.uint8)   # columns-only gradient: ties along v
    half = B // 2
    patches = np.stack([img[80 - half:80 + half + 1, 120 - half:120 + half + 1],
                        img[100 - half:100 + half + 1, 12 - half:12 + half + 1],
                        np.full((B, B), 77, np.uint8),                      # sigma0 = 0
                        rng.integers(0, 256, (B, B), dtype=np.uint8)])
    ctx = ctx_for_image(img, patches, radius=20)
    cases = [(0, [120.3, 80.6], [0.02, 0.0, 0.02]), (0, [2.0, 3.0], [0.01, 0.002, 0.02]),
             (0, [158.9, 118.2], [0.03, -0.01, 0.02]), (1, [14.0, 104.0], [0.05, 0.0, 0.01]),
             (1, [70.0, 50.0], [0.04, 0.0, 0.04]), (2, [120.0, 80.0], [0.1, 0.0, 0.1]),
             (3, [80.0, 5.0], [0.009, 0.0, 0.5]), (3, [3.0, 60.0], [0.5, 0.0, 0.009]),
             (0, [120.0, 80.0], [0.3, 0.29, 0.3])]
    for feat, c, p in cases:
        box, corr, sd, inside = ctx.score_map(0, 0, feat, c, p)
        obox, ocorr, osd, oinside = oracle.score_map(img, patches[feat], c, p)
        assert (box == obox).all()
        assert (inside == oinside).all()
        assert corr.tobytes() == ocorr.tobytes()      # every candidate, inside or not: bit-exact
        assert sd.tobytes() == osd.tobytes()
        feat_i = np.array([feat], np.int32)
        u, v, f, best = ctx.patch_search(0, 0, feat_i, np.array([c]), np.array([p]))
        ou, ov, of, obest = oracle.elliptical_search(img, patches[feat_i], np.array([c]), np.array([p]))
        assert (f[0], best[0]) == (of[0], obest[0])
        if obest[0] < 1e6:
            assert (u[0], v[0]) == (ou[0], ov[0])
    ctx.close()


def test_large_ellipse_walks_several_tiles(oracle):
    sc = synth.make_scene("C2", n_frames=1, n_features=10)
    ctx = ctx_from_scenes([sc], search_tile_radius=12)    # ellipses below need up to 4x4 tiles
    ctx.set_frame(0, 0, sc.frames[0])
    rng = np.random.default_rng(4)
    n = sc.n_features
    pu = random_puinv(rng, n, 25, 60, iso_fraction=0.3)
    centres = sc.pix + rng.uniform(-15, 15, (n, 2))
    _compare_search(oracle, ctx, sc.frames[0], sc.patches, np.arange(n, dtype=np.int32), centres, pu)
    ctx.close()


def test_smoe_search_matches_oracle_and_reference_golden(oracle):
    k = np.load(os.path.join(G, "a11_ref_kat.npz"))
    ctx = ctx_for_image(k["image"], k["patch"][None], radius=20)
    ru, rv, rf = ctx.smoe_search(0, 0, 0, k["puinv3"], k["centres"])
    assert (ru == k["res_u"]).all() and (rv == k["res_v"]).all() and (rf == k["res_flag"]).all()
    rng = np.random.default_rng(5)
    pu = random_puinv(rng, 16, 4, 18, iso_fraction=0.3)
    centres = np.column_stack([rng.uniform(0, 160, 16), rng.uniform(0, 120, 16)])
    ru, rv, rf = ctx.smoe_search(0, 0, 0, pu, centres)
    ou, ov, of, _ = oracle.smoe_search(k["image"], k["patch"], pu, centres)
    assert (ru == ou).all() and (rv == ov).all() and (rf == of).all()
    ctx.close()


def test_search_arguments_are_checked():
    sc = synth.make_scene("C2", n_frames=1, n_features=4)
    ctx = ctx_from_scenes([sc])
    with pytest.raises(sl2.Sl2Error):
        ctx.patch_search(0, 0, np.array([7], np.int32), np.zeros((1, 2)), np.ones((1, 3)))
    with pytest.raises(sl2.Sl2Error):
        ctx.patch_search(3, 0, np.array([0], np.int32), np.zeros((1, 2)), np.ones((1, 3)))
    ctx.close()


@pytest.mark.parametrize("B", [11, 15])
def test_shi_tomasi_best_patch_bit_exact(oracle, B):
    """N3: MonoSLAM::find_best_patch_inside_region (monoslam.cpp:1070-1205) — position and the
    FP64 eigenvalue bits, incl. border clamps, an empty region and a flat image."""
    rng = np.random.default_rng(8)
    img = synth.make_texture(rng, 120, 160)
    img[0:50, 0:70] = 93                                   # flat corner: score 0 there
    patches = np.zeros((1, B, B), np.uint8)
    ctx = ctx_for_image(img, patches)
    regions = np.array([[40, 30, 120, 90], [-5, -7, 60, 40], [100, 70, 400, 300], [0, 0, 160, 120],
                        [50, 50, 50, 80], [3, 3, 40, 30], [90, 20, 91, 21], [10, 10, 30, 25]], np.int32)
    u, v, ev = ctx.find_best_patch(0, 0, regions, ubest=-7, vbest=-9)
    for i, reg in enumerate(regions):
        ou, ov, oev = oracle.find_best_patch(img, B, reg, ubest=-7, vbest=-9)
        assert (u[i], v[i]) == (ou, ov), (i, reg)
        assert np.float64(ev[i]).tobytes() == np.float64(oev).tobytes(), (i, reg)
    assert ev[0] > 0 and ev[4] == 0.0
    ctx.close()


@pytest.mark.parametrize("B", [11, 15])
def test_shi_tomasi_many_regions_and_ties_across_tiles(oracle, B):
    """N3, batched: 48 regions of a 320x240 frame in one call (the kernel works on 32x32 tiles of positions and
    reduces across tiles), and a periodic image whose maxima repeat every 32 / 40 pixels, so that equal eigenvalues
    meet in different tiles: the first in (v, u) scan order must win, like the reference's strict `>`."""
    rng = np.random.default_rng(80 + B)
    img = synth.make_texture(rng, 240, 320)
    patches = np.zeros((1, B, B), np.uint8)
    ctx = ctx_for_image(img, patches)
    x0 = rng.integers(-10, 300, 48)
    y0 = rng.integers(-10, 220, 48)
    regions = np.column_stack([x0, y0, x0 + rng.integers(1, 130, 48), y0 + rng.integers(1, 100, 48)]).astype(np.int32)
    regions[0] = [0, 0, 320, 240]
    u, v, ev = ctx.find_best_patch(0, 0, regions, ubest=-1, vbest=-1)
    for i, reg in enumerate(regions):
        ou, ov, oev = oracle.find_best_patch(img, B, reg, ubest=-1, vbest=-1)
        assert (u[i], v[i]) == (ou, ov), (i, reg)
        assert np.float64(ev[i]).tobytes() == np.float64(oev).tobytes(), (i, reg)
    for period in (32, 40):
        cell = synth.make_texture(rng, period, period)
        per = np.tile(cell, (240 // period + 1, 320 // period + 1))[:240, :320].copy()
        ctx.set_frame(0, 0, per)
        regs = np.array([[20, 20, 300, 220], [0, 0, 320, 240], [33, 41, 200, 150]], np.int32)
        u, v, ev = ctx.find_best_patch(0, 0, regs, ubest=-1, vbest=-1)
        for i, reg in enumerate(regs):
            ou, ov, oev = oracle.find_best_patch(per, B, reg, ubest=-1, vbest=-1)
            assert (u[i], v[i]) == (ou, ov), (period, i)
            assert np.float64(ev[i]).tobytes() == np.float64(oev).tobytes()
    ctx.close()


def test_odd_frame_size_pitch_padding(oracle):
    """Frame width not a multiple of 16 (device pitch != width): search, score map and detector."""
    rng = np.random.default_rng(12)
    img = synth.make_texture(rng, 101, 150)
    B = 11
    pts = [(30, 40), (140, 90), (75, 8), (6, 95), (144, 6)]
    patches = np.stack([img[y - 5:y + 6, x - 5:x + 6] for x, y in pts])
    ctx = ctx_for_image(img, patches)
    n = len(pts)
    centres = np.array(pts, float) + rng.uniform(-3, 3, (n, 2))
    pu = random_puinv(rng, n, 6, 25, iso_fraction=0.4)
    f = _compare_search(oracle, ctx, img, patches, np.arange(n, dtype=np.int32), centres, pu)
    assert f.all()
    box, corr, sd, inside = ctx.score_map(0, 0, 1, [141.0, 91.0], [0.02, 0.0, 0.02])
    obox, ocorr, osd, oinside = oracle.score_map(img, patches[1], [141.0, 91.0], [0.02, 0.0, 0.02])
    assert (box == obox).all() and corr.tobytes() == ocorr.tobytes() and (inside == oinside).all()
    u, v, ev = ctx.find_best_patch(0, 0, [[0, 0, 150, 101], [100, 50, 150, 101]])
    for i, reg in enumerate(([0, 0, 150, 101], [100, 50, 150, 101])):
        assert (u[i], v[i], ev[i]) == oracle.find_best_patch(img, B, reg)
    ctx.close()


def test_particle_measurement_and_reweighting_matches_oracle(oracle):
    """N2: sl2_measure_particles = measure_feature_with_multiple_priors (monoslam.cpp:1408-1438) + the
    particle branch of update_partially_initialised_feature_probabilities (monoslam.cpp:1447-1493,
    feature_init_info.cpp:95-172).  Match positions / flags / surviving set bit-exact against the oracle
    (reference-pinned SMOE search), probabilities within 1e-13 (device exp is not glibc's)."""
    k = np.load(os.path.join(G, "a11_ref_kat.npz"))
    img, patch = k["image"], k["patch"]
    ctx = ctx_for_image(img, patch[None], radius=20)
    rng = np.random.default_rng(21)
    # where the template really is: a successful match of the reference's own golden run
    hit = int(np.flatnonzero(k["res_flag"])[0])
    cu, cv = [float(k["res_u"][hit])], [float(k["res_v"][hit])]
    for K, spread in ((40, 6.0), (100, 25.0), (1, 0.5)):
        h = np.column_stack([cu[0] + rng.normal(0, spread, K), cv[0] + rng.normal(0, spread, K)])
        pu = random_puinv(rng, K, 4, 12, iso_fraction=0.3)
        det = 1.0 / (pu[:, 0] * pu[:, 2] - pu[:, 1] ** 2)          # det S = 1 / det Sinv
        lam = np.linspace(0.5, 4.5, K)
        p0 = rng.uniform(0.2, 1.0, K)
        p0 /= p0.sum()
        left, prob, z, found, keep, cum, mv = ctx.measure_particles(0, 0, 0, h, pu, det, lam, 0.05, p0)
        ou, ov, of, _ = oracle.smoe_search(img, patch, pu, h)
        oleft, oprob, okeep, ocum, omv = oracle.particle_update(h, pu, det, lam, np.column_stack([ou, ov]),
                                                                of, 0.05, p0)
        assert (found == of).all()
        assert (z[of > 0, 0] == ou[of > 0]).all() and (z[of > 0, 1] == ov[of > 0]).all()
        assert left == oleft and (keep == okeep).all()
        np.testing.assert_allclose(prob, oprob, rtol=1e-13, atol=1e-300)
        np.testing.assert_allclose(cum, ocum, rtol=1e-13, atol=1e-300)
        np.testing.assert_allclose(mv, omv, rtol=1e-12, atol=1e-15)
        if left:
            assert abs(prob[keep > 0].sum() - 1.0) < 1e-12 and (prob[keep == 0] * 0 == 0).all()
    # every match fails (ellipses on a flat, textureless area): the reference deletes the feature
    flat = img.copy()
    flat[:] = 127
    ctx.set_frame(0, 0, flat)
    K = 8
    h = np.column_stack([rng.uniform(40, 120, K), rng.uniform(40, 80, K)])
    pu = random_puinv(rng, K, 4, 8, iso_fraction=1.0)
    det = 1.0 / (pu[:, 0] * pu[:, 2] - pu[:, 1] ** 2)
    left, prob, z, found, keep, cum, mv = ctx.measure_particles(0, 0, 0, h, pu, det, np.ones(K), 0.05,
                                                                np.full(K, 1.0 / K))
    assert left == 0 and not found.any() and not keep.any() and (prob == 0).all()
    ctx.close()


def test_partial_features_batched_prediction_search_reweighting(oracle):
    """N2 in one call (sl2_measure_partial_features): device prediction of every particle's ellipse
    (monoslam.cpp:1347-1400; bit-exact against the reference-pinned oracle: only IEEE + - * / sqrt are involved),
    the shared-score-map SMOE search (positions / flags exact) and the re-weighting (1e-13), for three features
    with different particle counts on one frame."""
    rng = np.random.default_rng(77)
    img = synth.make_texture(rng, 240, 320)
    B = 11
    ctx = ctx_for_image(img, np.zeros((1, B, B), np.uint8))
    cfg = ctx.cfg
    cam8 = np.array([cfg.width, cfg.height, cfg.fku, cfg.fkv, cfg.u0, cfg.v0, cfg.kd1, cfg.sd], float)
    xv = np.zeros(13)
    xv[:3] = [0.06, -0.03, 0.01]
    q = np.array([1.0, 0.01, -0.02, 0.015])
    xv[3:7] = q / np.linalg.norm(q)
    A = rng.normal(0, 1, (16, 16))
    P = A @ A.T * 2e-6 + 1e-8 * np.eye(16)
    x = np.concatenate([xv, [0.1, 0.1, 2.0]])
    ctx.set_state(0, x, P)
    F, Kmax = 3, 60
    K = np.array([Kmax, Kmax - 7, 1], np.int32)
    pix = [(120.0, 100.0), (215.0, 150.0), (160.0, 60.0)]
    ypi = np.zeros((F, 6))
    Pxy = np.zeros((F, 13, 6))
    Pyy = np.zeros((F, 6, 6))
    lam = np.zeros((F, Kmax))
    prob = np.zeros((F, Kmax))
    patches = np.zeros((F, B, B), np.uint8)
    for f, (u, v) in enumerate(pix):
        hh = np.array([-(u - cfg.u0) / cfg.fku, -(v - cfg.v0) / cfg.fkv, 1.0])
        ypi[f] = np.concatenate([[0.0, 0.0, 0.0], hh / np.linalg.norm(hh)])
        Af = rng.normal(0, 1, (19, 19))
        Pf = Af @ Af.T * 2e-6 + 1e-8 * np.eye(19)
        Pf[:13, :13] = P[:13, :13]
        Pxy[f], Pyy[f] = Pf[:13, 13:], Pf[13:, 13:]
        lam[f] = np.linspace(0.4, 6.0, Kmax)
        p0 = rng.uniform(0.2, 1.0, Kmax)
        p0[K[f]:] = 0.0
        prob[f] = p0 / p0.sum()
    # templates: what the image shows where the middle particle of each feature projects
    for f in range(F):
        h_mid = oracle.predict_particles(cam8, xv, ypi[f], lam[f, :1] * 0 + lam[f, K[f] // 2], P[:13, :13], Pxy[f],
                                         Pyy[f])[0][0]
        cu, cv = int(round(h_mid[0])), int(round(h_mid[1]))
        patches[f] = img[cv - 5:cv + 6, cu - 5:cu + 6]
    out = ctx.measure_partial_features(0, 0, patches, ypi, Pxy, Pyy, lam, 0.05, prob, K=K)
    any_found = False
    for f in range(F):
        k = K[f]
        oh, oS, osi, odet = oracle.predict_particles(cam8, xv, ypi[f], lam[f, :k], P[:13, :13], Pxy[f], Pyy[f])
        assert out["h"][f, :k].tobytes() == oh.tobytes(), f
        assert out["Sinv3"][f, :k].tobytes() == osi.tobytes(), f
        assert out["detS"][f, :k].tobytes() == odet.tobytes(), f
        ou, ov, of, _ = oracle.smoe_search(img, patches[f], osi, oh)
        assert (out["found"][f, :k] == of).all(), f
        z = out["z"][f, :k]
        assert (z[of > 0, 0] == ou[of > 0]).all() and (z[of > 0, 1] == ov[of > 0]).all(), f
        any_found |= bool(of.any())
        oleft, oprob, okeep, ocum, omv = oracle.particle_update(oh, osi, odet, lam[f, :k], np.column_stack([ou, ov]),
                                                                of, 0.05, prob[f, :k])
        assert out["left"][f] == oleft and (out["keep"][f, :k] == okeep).all(), f
        np.testing.assert_allclose(out["prob"][f, :k], oprob, rtol=1e-13, atol=1e-300)
        np.testing.assert_allclose(out["cumulative"][f, :k], ocum, rtol=1e-13, atol=1e-300)
        np.testing.assert_allclose(out["mean_var"][f], omv, rtol=1e-12, atol=1e-15)
    assert any_found
    # argument checks
    with pytest.raises(sl2.Sl2Error):
        ctx.measure_partial_features(0, 0, patches, ypi, Pxy, Pyy, lam, 0.05, prob, K=np.array([Kmax + 1, 1, 1], np.int32))
    ctx.close()


def test_partial_features_at_the_limits_640x480_box15(oracle):
    """sl2_measure_partial_features at its limits: SL2_MAX_PARTIAL = 16 features x SL2_MAX_PARTICLES = 256 particle
    slots on a 640x480 frame with 15x15 templates, particle counts from 0 to 256, rays that leave the image (ellipse
    boxes clipped at the border, ellipses with no valid location)."""
    rng = np.random.default_rng(91)
    img = synth.make_texture(rng, 480, 640)
    B = 15
    ctx = ctx_for_image(img, np.zeros((1, B, B), np.uint8))
    cfg = ctx.cfg
    cam8 = np.array([cfg.width, cfg.height, cfg.fku, cfg.fkv, cfg.u0, cfg.v0, cfg.kd1, cfg.sd], float)
    xv = np.zeros(13)
    xv[:3] = [0.12, 0.05, -0.02]
    q = np.array([1.0, -0.02, 0.03, 0.01])
    xv[3:7] = q / np.linalg.norm(q)
    A = rng.normal(0, 1, (16, 16))
    P = A @ A.T * 4e-6 + 1e-8 * np.eye(16)
    ctx.set_state(0, np.concatenate([xv, [0.1, 0.1, 2.0]]), P)
    F, Kmax = 16, 256
    K = rng.integers(1, Kmax + 1, F).astype(np.int32)
    K[0], K[1], K[2] = Kmax, 0, 1
    ypi = np.zeros((F, 6))
    Pxy = np.zeros((F, 13, 6))
    Pyy = np.zeros((F, 6, 6))
    lam = np.tile(np.linspace(0.3, 8.0, Kmax), (F, 1))
    prob = np.zeros((F, Kmax))
    patches = rng.integers(0, 256, (F, B, B), dtype=np.uint8)
    for f in range(F):
        # pixel the ray points at from the origin; some of them close to / beyond the image border
        u, v = rng.uniform(-20, cfg.width + 20), rng.uniform(-20, cfg.height + 20)
        hh = np.array([-(u - cfg.u0) / cfg.fku, -(v - cfg.v0) / cfg.fkv, 1.0])
        ypi[f, 3:] = hh / np.linalg.norm(hh)
        Af = rng.normal(0, 1, (19, 19))
        Pf = Af @ Af.T * 4e-6 + 1e-8 * np.eye(19)
        Pxy[f], Pyy[f] = Pf[:13, 13:], Pf[13:, 13:]
        if K[f]:
            p0 = rng.uniform(0.2, 1.0, K[f])
            prob[f, :K[f]] = p0 / p0.sum()
    for f in (0, 3, 7):   # a real template where the ray's middle particle lands, when that is inside the image
        hm = oracle.predict_particles(cam8, xv, ypi[f], [lam[f, K[f] // 2]], P[:13, :13], Pxy[f], Pyy[f])[0][0]
        cu, cv = int(round(hm[0])), int(round(hm[1]))
        if 7 <= cu < cfg.width - 8 and 7 <= cv < cfg.height - 8:
            patches[f] = img[cv - 7:cv + 8, cu - 7:cu + 8]
    before = prob.copy()
    out = ctx.measure_partial_features(0, 0, patches, ypi, Pxy, Pyy, lam, 0.05, prob, K=K)
    assert out["left"][1] == 0 and (out["prob"][1] == 0).all()          # no particles: nothing to do
    for f in range(F):
        k = K[f]
        if k == 0:
            continue
        oh, oS, osi, odet = oracle.predict_particles(cam8, xv, ypi[f], lam[f, :k], P[:13, :13], Pxy[f], Pyy[f])
        assert out["h"][f, :k].tobytes() == oh.tobytes() and out["Sinv3"][f, :k].tobytes() == osi.tobytes(), f
        if f % 3 == 0 or k < 40:   # the CPU side of the search is the slow part of this test
            ou, ov, of, _ = oracle.smoe_search(img, patches[f], osi, oh)
            assert (out["found"][f, :k] == of).all(), f
            assert (out["z"][f, :k, 0] == ou).all() and (out["z"][f, :k, 1] == ov).all(), f
            oleft, oprob, okeep, ocum, omv = oracle.particle_update(oh, osi, odet, lam[f, :k],
                                                                    np.column_stack([ou, ov]), of, 0.05, before[f, :k])
            assert out["left"][f] == oleft and (out["keep"][f, :k] == okeep).all(), f
            np.testing.assert_allclose(out["prob"][f, :k], oprob, rtol=1e-13, atol=1e-300)
    with pytest.raises(sl2.Sl2Error):   # one feature too many
        ctx.measure_partial_features(0, 0, np.zeros((17, B, B), np.uint8), np.zeros((17, 6)), np.zeros((17, 13, 6)),
                                     np.zeros((17, 6, 6)), np.ones((17, 4)), 0.05, np.ones((17, 4)))
    ctx.close()


def test_raw_template_variants_of_smoe_and_particles(oracle):
    """sl2_smoe_search_patch / sl2_measure_particles_patch: the template of a partially-initialised feature is not
    a map feature (the reference hands Feature::patch_ to the SMOE search, monoslam.cpp:1413).  Same results as
    the oracle, and as the feat_index variants when the same template is registered as a map feature; the map
    templates are untouched by the scratch slot."""
    k = np.load(os.path.join(G, "a11_ref_kat.npz"))
    img, patch = k["image"], k["patch"]
    other = np.ascontiguousarray(img[20:31, 30:41])                       # the only map template: a different patch
    ctx = ctx_for_image(img, other[None], radius=20)
    ru, rv, rf = ctx.smoe_search_patch(0, 0, patch, k["puinv3"], k["centres"])
    assert (ru == k["res_u"]).all() and (rv == k["res_v"]).all() and (rf == k["res_flag"]).all()
    rng = np.random.default_rng(77)
    hit = int(np.flatnonzero(k["res_flag"])[0])
    K = 60
    h = np.column_stack([k["res_u"][hit] + rng.normal(0, 6, K), k["res_v"][hit] + rng.normal(0, 6, K)])
    pu = random_puinv(rng, K, 4, 12, iso_fraction=0.3)
    det = 1.0 / (pu[:, 0] * pu[:, 2] - pu[:, 1] ** 2)
    lam = np.linspace(0.5, 4.5, K)
    p0 = np.full(K, 1.0 / K)
    left, prob, z, found, keep, cum, mv = ctx.measure_particles(0, 0, -1, h, pu, det, lam, 0.05, p0, patch=patch)
    ou, ov, of, _ = oracle.smoe_search(img, patch, pu, h)
    oleft, oprob, okeep, ocum, omv = oracle.particle_update(h, pu, det, lam, np.column_stack([ou, ov]), of, 0.05, p0)
    assert (found == of).all() and (z[of > 0, 0] == ou[of > 0]).all() and (z[of > 0, 1] == ov[of > 0]).all()
    assert left == oleft and left > 0 and (keep == okeep).all()
    np.testing.assert_allclose(prob, oprob, rtol=1e-13, atol=1e-300)
    np.testing.assert_allclose(mv, omv, rtol=1e-12, atol=1e-15)
    # the map feature still carries its own template
    ou2, ov2, of2, _ = oracle.smoe_search(img, other, pu[:8], np.tile([35.3, 25.4], (8, 1)))
    ru2, rv2, rf2 = ctx.smoe_search(0, 0, 0, pu[:8], np.tile([35.3, 25.4], (8, 1)))
    assert (rf2 == of2).all() and (ru2[of2 > 0] == ou2[of2 > 0]).all() and (rv2[of2 > 0] == ov2[of2 > 0]).all()
    ctx.close()
