"""developer probe: where do neighbour counts / density sums differ from the oracle?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from ti_sph_b200 import _capi as K
from util import make_pair, small_scene, jitter, rel_err

for state in ("lattice", "jitter"):
    for mode in ("reference", "summed"):
        scene = small_scene()
        x = None
        if state == "jitter":
            o0, e0 = make_pair(scene); x = jitter(o0.x, 0.01); e0.close()
        ora, eng = make_pair(scene, density_mode=mode, x=x)
        t = ora.step(trace=True)
        eng.set_param(K.P_DIAGNOSTICS, 1)
        eng.stage(K.STAGE_UPDATE)
        eng.stage(K.STAGE_DENSITY)
        nc = eng.download(K.F_NEIGHBOR_COUNT)
        S = eng.download(K.F_DENSITY_SUM)
        bad = np.nonzero(nc != t["neighbor_count"])[0]
        print(state, mode, "n", ora.n, "count mismatches", len(bad), "S relerr", rel_err(S, t["S"], floor=1.0),
              "items", eng.get_param(K.P_STAT_ITEMS), "fb", eng.get_param(K.P_STAT_FALLBACK_DENSITY), eng.get_param(K.P_STAT_FALLBACK_FORCE))
        pr, pref = eng.download(K.F_PRESSURE).astype(np.float64), t["pressure"].astype(np.float64)
        x7 = (t["density"].astype(np.float64) / 1000.0) ** 7
        print("  pressure: worst |dp| / allowed", np.max(np.abs(pr - pref) / (1e-5 * np.abs(pref) + 400 * 1.1920929e-07 * x7)))
        if len(bad):
            d = nc[bad] - t["neighbor_count"][bad]
            print("  diff hist", np.unique(d, return_counts=True))
            scan = t["scan"]; keys = t["keys"]
            cell = keys[bad]
            start = np.where(cell > 0, scan[np.maximum(cell - 1, 0)], 0)
            print("  first bad", bad[:10], "offset in cell", (bad - start)[:10], "cell count", (scan[cell] - start)[:10])
        eng.stage(K.STAGE_FORCE_ADVECT)
        a = eng.download(K.F_D_VELOCITY)
        print("  dvel err", np.abs(a - t["d_velocity"]).max(), "scale", np.abs(t["d_velocity"]).max())
        eng.close()
