import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from ti_sph_b200 import _capi as K, scene as sc
from util import make_pair, jitter, rel_err
scene = sc.bench_scene("C3")
o0, e0 = make_pair(scene); x = jitter(o0.x, 0.01); e0.close()
ora, eng = make_pair(scene, x=x)
t = ora.step(trace=True)
eng.stage(K.STAGE_UPDATE); eng.stage(K.STAGE_DENSITY)
nc = eng.download(K.F_NEIGHBOR_COUNT); S = eng.download(K.F_DENSITY_SUM)
bad = np.nonzero(nc != t["neighbor_count"])[0]
print("mismatches", len(bad), "S relerr", rel_err(S, t["S"], floor=1.0), "items", eng.get_param(K.P_STAT_ITEMS),
      eng.get_param(K.P_STAT_FALLBACK_DENSITY), eng.get_param(K.P_STAT_FALLBACK_FORCE))
if len(bad):
    d = nc[bad] - t["neighbor_count"][bad]
    print("diff hist", np.unique(d, return_counts=True))
    scan, keys = t["scan"], t["keys"]
    cell = keys[bad]; start = np.where(cell > 0, scan[np.maximum(cell - 1, 0)], 0)
    print("bad idx", bad[:20]); print("offset in cell", (bad - start)[:20]); print("cell count", (scan[cell] - start)[:20])
    print("S err at bad", (S[bad] - t["S"][bad])[:20] / t["S"][bad][:20])
    # tile totals of bad cells
    g = ora.grid_num
    cnts = t["counts"].reshape(g[0], g[1], g[2])
    from scipy import ndimage
    tot = np.rint(ndimage.uniform_filter(cnts.astype(np.float64), 3, mode="constant") * 27).astype(int).ravel()
    print("tile totals of bad cells", np.unique(tot[cell], return_counts=True))
    print("max tile overall", tot[t["counts"] > 0].max())
