"""Whole-step oracle behaviour (monoslam.cpp:108-180) on the synthetic scenes."""
import os

import numpy as np

from scenelib2_b200 import synth

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def make_oracle_slam(oracle, sc):
    cfg = oracle.make_config(width=sc.width, height=sc.height, fku=sc.cam8[2], fkv=sc.cam8[3],
                             u0=sc.cam8[4], v0=sc.cam8[5], kd1=sc.cam8[6], sd=sc.cam8[7],
                             delta_t=sc.delta_t, n_select=sc.n_select, boxsize=sc.boxsize,
                             search_override=sc.search_override)
    s = oracle.Slam(cfg)
    for i in range(sc.n_features):
        s.add_feature(sc.x0[13 + 3 * i:16 + 3 * i], sc.xp_org[i], sc.patches[i])
    s.set_state(sc.x0, sc.P0)
    return s


def test_c2_tracks_true_matches(oracle):
    sc = synth.make_scene("C2", n_frames=5, override=False)
    s = make_oracle_slam(oracle, sc)
    x, P = s.get_state()
    assert (P == sc.P0).all() and (x == sc.x0).all()       # gather/scatter round trip
    tr_prev = np.trace(sc.P0[13:, 13:])
    for t in range(5):
        s.step(sc.frames[t])
        f = s.features()
        assert ((f["flags"] & 2) > 0).all()
        assert (f["z"] == sc.pix + sc.shifts[t]).all()      # bit-exact match positions
        x, P = s.get_state()
        assert np.abs(P - P.T).max() == 0.0                  # symmetrised (monoslam.cpp:143-150)
        assert np.trace(P[13:, 13:]) <= tr_prev                # map uncertainty never grows
        tr_prev = np.trace(P[13:, 13:])
    assert np.linalg.eigvalsh(P).min() > -1e-12


def test_staged_equals_step(oracle):
    sc = synth.make_scene("C2", n_frames=2, n_features=20)
    a = make_oracle_slam(oracle, sc)
    b = make_oracle_slam(oracle, sc)
    for t in range(2):
        a.step(sc.frames[t])
        b.predict()
        b.select()
        if b.measure(sc.frames[t]):
            b.update()
            b.normalise()
        b.finish()
        xa, Pa = a.get_state()
        xb, Pb = b.get_state()
        assert (xa == xb).all() and (Pa == Pb).all()


def test_c1_selection_and_known_patches(oracle):
    kp = np.load(os.path.join(G, "known_patches.npy"))
    sc = synth.make_scene("C1", n_frames=3, known_patches=kp)
    assert (sc.patches[:4] == kp).all()
    s = make_oracle_slam(oracle, sc)
    s.step(sc.frames[0])
    f = s.features()
    sel = f["select_rank"] >= 0
    assert sel.sum() == 10                                   # data/SceneLib2.cfg:60
    score = f["S"][:, 0] + f["S"][:, 3]
    order = np.argsort(f["select_rank"][sel])
    ranked = score[sel][order]
    assert (np.diff(ranked) <= 0).all()                      # descending trace(S)
    assert score[~sel].max() <= ranked.min()
    assert ((f["flags"][sel] & 2) > 0).all()
    assert (f["attempted"][sel] == 1).all() and (f["attempted"][~sel] == 0).all()


def test_bad_features_are_deleted(oracle):
    sc = synth.make_scene("C2", n_frames=2, n_features=12)
    bad = sc.patches.copy()
    rng = np.random.default_rng(0)
    bad[3] = rng.integers(0, 256, bad[3].shape, dtype=np.uint8)   # template that never matches
    sc.patches = bad
    s = make_oracle_slam(oracle, sc)
    n0 = s.n
    for t in range(12):
        s.step(sc.frames[t % 2])
    assert s.num_features == 11 and s.n == n0 - 3            # monoslam.cpp:644-660: >=10 attempts, <50 %
    x, P = s.get_state()
    assert P.shape == (n0 - 3, n0 - 3) and np.isfinite(P).all()


def test_c1_trajectory_1000_steps_matches_fixture(oracle):
    """1 000 GoOneStep calls on C1 against the fixture produced by the REFERENCE'S OWN code
    (tests/golden/make_c1_trajectory.py): integer results (selection ranks, flags, match positions of every
    step) by hash, camera state / covariance at every 100th step numerically."""
    import sys
    sys.path.insert(0, G)
    import make_c1_trajectory as gen
    k = np.load(os.path.join(G, "c1_trajectory_1000.npz"))
    out = gen.run(oracle)
    assert (out["integer_hash"] == k["integer_hash"]).all()
    assert (out["nfeat"] == k["nfeat"]).all()
    assert (out["attempted"] == k["attempted"]).all() and (out["successful"] == k["successful"]).all()
    for name in ("xv", "Pxx_diag", "x_final", "P_diag_final"):
        np.testing.assert_allclose(out[name], k[name], rtol=1e-9, atol=1e-15)
