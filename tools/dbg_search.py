"""Debug helper: one patch_search call against the oracle (run under compute-sanitizer)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import scenelib2_b200 as sl2
from scenelib2_b200 import synth
from oracle import pyoracle as po
from gpu_util import ctx_from_scenes

sc = synth.make_scene("C2", n_frames=1, n_features=8)
ctx = ctx_from_scenes([sc])
ctx.set_frame(0, 0, sc.frames[0])
n = sc.n_features
centres = sc.pix.astype(float) + 0.3
pu = np.tile([9 / 400.0, 0.0, 9 / 400.0], (n, 1))
u, v, f, best = ctx.patch_search(0, 0, np.arange(n, dtype=np.int32), centres, pu)
ou, ov, of, obest = po.elliptical_search(sc.frames[0], sc.patches, centres, pu)
print(u, ou); print(v, ov); print(f, of); print(best - obest)
