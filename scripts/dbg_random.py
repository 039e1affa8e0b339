import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from oracle.oracle import Gen2Oracle
from ti_sph_b200 import _capi as K, scene as sc
from ti_sph_b200.engine import Engine
from util import small_scene
from test_gpu_random_states import random_state
for variant in (0, 1):
    scene = small_scene(domain_end=(1.0, 0.8, 0.6))
    x, v, density, material = random_state(2, 6000, (1.0, 0.8, 0.6), 0.04)
    ora = Gen2Oracle(scene); ora.set_state(x, v, density, material)
    eng = Engine(sc.gen2_config(scene["configuration"], len(x)))
    eng.add_particles(ora.x, ora.v, ora.density, ora.pressure, ora.material, ora.color)
    eng.set_param(K.P_DIAGNOSTICS, 1); eng.set_param(K.P_KERNEL_VARIANT, variant)
    t = ora.step(trace=True)
    eng.step(1)
    ap = eng.download(K.F_A_PRESSURE).astype(np.float64)
    d = np.linalg.norm(ap - t["a_pressure"], axis=1)
    mp = t["mag_pressure"].astype(np.float64); pf = t["mag_pressure_floor"]
    r = np.where(mp > 0, np.maximum(d - pf, 0) / np.maximum(mp, 1e-300), 0)
    w = np.argsort(-r)[:5]
    print("variant", variant, "fallback items", eng.get_param(K.P_STAT_FALLBACK_FORCE), eng.get_param(K.P_STAT_ITEMS))
    for i in w:
        print(i, "ratio", r[i], "d", d[i], "mag", mp[i], "floor", pf[i], "a_gpu", ap[i], "a_ora", t["a_pressure"][i], "ncount", t["neighbor_count"][i],
              "p", t["pressure"][i], "rho", t["density"][i], "mat", t["material"][i])
    eng.close()
