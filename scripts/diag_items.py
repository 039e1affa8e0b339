"""Diagnostic: per-step work-item / fallback statistics and neighbour-count range on the GPU."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ti_sph_b200 import _capi as K, scene as sc
from core.partice_system.partice_systemv4 import ParticleSystemV4
from core.sph.wcsphv2 import WCSPHV2

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
ps = ParticleSystemV4(sc.bench_scene(name))
solver = WCSPHV2(ps)
eng = ps.engine
for s in range(steps):
    solver.step()
    nc = eng.download(K.F_NEIGHBOR_COUNT)
    cc = eng.download(K.F_CELL_COUNT)
    print(s, "items", int(eng.get_param(K.P_STAT_ITEMS)), "fb_d", int(eng.get_param(K.P_STAT_FALLBACK_DENSITY)),
          "fb_f", int(eng.get_param(K.P_STAT_FALLBACK_FORCE)), "nbr max/mean", nc.max(), round(float(nc.mean()), 1),
          "cell max", cc.max(), "cells>64", int((cc > 64).sum()), flush=True)
