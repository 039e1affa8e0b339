"""Diagnostic: LocalCluster (all ranks on one GPU) on a bench scene; reports the first failing step."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ti_sph_b200 import _capi as K, scene as sc
from ti_sph_b200.sharded import LocalCluster

name, world, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
cl = LocalCluster(sc.bench_scene(name), world)
print("edges", cl.sims[0].edges, "owned", [s.engine.particle_num for s in cl.sims], flush=True)
for st in range(steps):
    try:
        cl.step(1)
    except Exception as e:
        print("step", st, "failed:", e)
        for r, s in enumerate(cl.sims):
            try:
                x = s.engine.download(K.F_X)
                cx = (x[:, 0] / np.float32(s.parts.h)).astype(np.int32)
                bad = (cx < s.plane_lo - 1) | (cx >= s.plane_hi + 1)
                print(" rank", r, "planes", s.plane_lo, s.plane_hi, "n", len(x), "cx range", cx.min(), cx.max(), "bad", int(bad.sum()))
                if bad.any():
                    print(x[bad][:8], np.nonzero(bad)[0][:8])
            except Exception as e2:
                print(" rank", r, "download failed:", e2)
        break
    if st % 10 == 9:
        print("step", st + 1, "owned", [s.engine.particle_num for s in cl.sims], flush=True)
