import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ti_sph_b200 import scene as sc, _capi as K
from core.partice_system.partice_systemv4 import ParticleSystemV4
from core.sph.wcsphv2 import WCSPHV2
name = sys.argv[1]
ps = ParticleSystemV4(sc.bench_scene(name)); solver = WCSPHV2(ps); eng = ps.engine
def T(label, f):
    torch.cuda.synchronize(); t0 = time.time(); r = f(); eng.sync(); torch.cuda.synchronize()
    print(f"{label}: {(time.time()-t0)*1e3:.2f} ms", flush=True); return r
T("step x3", lambda: eng.step(3))
host = T("dump pageable", lambda: ps.dump())
pin = {k: torch.empty(v.shape, dtype=torch.float32 if v.dtype == np.float32 else torch.int32).pin_memory() for k, v in host.items()}
outs = {k: v.numpy() for k, v in pin.items()}
hx, hv = outs["position"], outs["velocity"]
hx[:] = host["position"]; hv[:] = host["velocity"]
for it in range(3):
    T("upload_xv", lambda: eng.upload_xv(hx, hv))
    T("step", lambda: solver.step())
    T("dump x", lambda: eng.download(K.F_X, outs["position"]))
    T("dump v", lambda: eng.download(K.F_V, outs["velocity"]))
    T("dump mat", lambda: eng.download(K.F_MATERIAL, outs["material"]))
    T("dump color", lambda: eng.download(K.F_COLOR, outs["color"]))
    print(eng.stage_times(True))
