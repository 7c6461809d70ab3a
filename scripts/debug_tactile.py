import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from mujoco_contact_surfaces_b200 import scenes
from parity_utils import make_engine, make_oracle, oracle_env
for S in (4, 8):
    sc = scenes.myrmex("box", sampling_resolution=S)
    eng, orc = make_engine(sc, 2), make_oracle(sc)
    xp, xm, ve = sc.poses(2, seed=7)
    eng.step(xp, xm, ve, with_sensors=True)
    img = eng.sensor_image(0)
    for e in range(2):
        _, ref = oracle_env(orc, sc, xp[e], xm[e], ve[e])
        ref = ref[0]; g = img[e]
        bad = np.nonzero(np.abs(g - ref) > 1e-6 * np.maximum(np.abs(ref), 1e-3 * ref.max()))[0]
        print("S", S, "env", e, "nonzero ref", (ref != 0).sum(), "nonzero gpu", (g != 0).sum(), "bad", len(bad), "sum ref", ref.sum(), "sum gpu", g.sum())
        for b in bad[:6]:
            print("   taxel", b % 16, b // 16, "ref", ref[b], "gpu", g[b], "ratio", g[b] / ref[b] if ref[b] else None)
