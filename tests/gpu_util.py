"""Helpers shared by the GPU parity tests (CUDA path through the C ABI vs the CPU oracle)."""
import numpy as np

import scenelib2_b200 as sl2
from scenelib2_b200 import synth

# north star: "FP state/covariance within 1e-5 relative".  The tests hold the CUDA path to a
# much tighter bound (different summation order only) so that real bugs cannot hide.
RTOL_NORTH_STAR = 1e-5
RTOL_TEST = 1e-8


def oracle_slam_from_scene(oracle, sc):
    cfg = oracle.make_config(width=sc.width, height=sc.height, fku=sc.cam8[2], fkv=sc.cam8[3],
                             u0=sc.cam8[4], v0=sc.cam8[5], kd1=sc.cam8[6], sd=sc.cam8[7],
                             delta_t=sc.delta_t, n_select=sc.n_select, boxsize=sc.boxsize,
                             search_override=sc.search_override)
    s = oracle.Slam(cfg)
    for i in range(sc.n_features):
        s.add_feature(sc.x0[13 + 3 * i:16 + 3 * i], sc.xp_org[i], sc.patches[i])
    s.set_state(sc.x0, sc.P0)
    return s


def ctx_from_scenes(scenes, frame_slots=1, **kw):
    cfg = sl2.config_for_scene(scenes[0], num_streams=len(scenes), frame_slots=frame_slots, **kw)
    ctx = sl2.Context(cfg)
    for s, sc in enumerate(scenes):
        sl2.load_scene(ctx, s, sc)
    return ctx


def ctx_for_image(image, patches, radius=20, boxsize=None):
    """Context with one stream whose frame is `image` and whose templates are `patches`."""
    patches = np.ascontiguousarray(patches, np.uint8)
    n, B = patches.shape[0], patches.shape[1]
    cfg = sl2.default_config()
    cfg.width, cfg.height = image.shape[1], image.shape[0]
    cfg.boxsize = B
    cfg.max_features = max(n, 1)
    cfg.search_tile_radius = radius
    ctx = sl2.Context(cfg)
    ctx.set_features(0, np.zeros((n, 3)), np.tile([0, 0, 0, 1, 0, 0, 0.0], (n, 1)), patches)
    ctx.set_frame(0, 0, image)
    return ctx


def state_err(xg, Pg, xo, Po):
    """Largest error relative to the natural scale sqrt(P_ii P_jj) (covariance) / sigma_i (state)."""
    d = np.sqrt(np.abs(np.diag(Po))) + 1e-300
    eP = np.abs(Pg - Po) / (d[:, None] * d[None, :])
    ex = np.abs(xg - xo) / np.maximum(np.abs(xo), d)
    return float(ex.max()), float(eP.max())


def assert_state_close(xg, Pg, xo, Po, rtol=RTOL_TEST):
    ex, eP = state_err(xg, Pg, xo, Po)
    assert ex <= rtol, "state error %.3e" % ex
    assert eP <= rtol, "covariance error %.3e" % eP
    return ex, eP


def random_puinv(rng, n, lo, hi, iso_fraction=0.5):
    out = np.zeros((n, 3))
    for i in range(n):
        a, b = rng.uniform(lo, hi, 2)
        r = 0.0 if rng.random() < iso_fraction else rng.uniform(-0.8, 0.8)
        if rng.random() < iso_fraction:
            b = a
        Si = np.linalg.inv(np.array([[a * a / 9, r * a * b / 9], [r * a * b / 9, b * b / 9]]))
        out[i] = [Si[0, 0], Si[0, 1], Si[1, 1]]
    return out


__all__ = ["sl2", "synth", "np"]
