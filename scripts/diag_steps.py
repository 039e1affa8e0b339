"""Diagnostic: step time and work-item statistics along a run (reference or summed density mode)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ti_sph_b200 import _capi as K, scene as sc
from core.partice_system.partice_systemv4 import ParticleSystemV4
from core.sph.wcsphv2 import WCSPHV2

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
mode = sys.argv[3] if len(sys.argv) > 3 else "reference"
ps = ParticleSystemV4(sc.bench_scene(name), density_mode=mode)
solver = WCSPHV2(ps)
eng = ps.engine
n = eng.particle_num
for s0 in range(0, steps, 10):
    eng.sync(); t0 = time.perf_counter()
    for _ in range(10):
        solver.step()
    eng.sync(); dt = (time.perf_counter() - t0) / 10
    nc = eng.download(K.F_NEIGHBOR_COUNT); cc = eng.download(K.F_CELL_COUNT)
    print(f"steps {s0+10:4d}: {dt*1e3:7.3f} ms/step {n/dt/1e6:7.1f} M/s  items {int(eng.get_param(K.P_STAT_ITEMS))} "
          f"fb {int(eng.get_param(K.P_STAT_FALLBACK_DENSITY))}/{int(eng.get_param(K.P_STAT_FALLBACK_FORCE))} "
          f"nbr mean/max {nc.mean():.1f}/{nc.max()} cell max {cc.max()}", flush=True)
