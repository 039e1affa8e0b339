"""Developer timing probe (not the contract bench): stage times of a few scenes."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ti_sph_b200 import scene as sc, _capi as K
from ti_sph_b200.engine import Engine

def run(name, mode=0, warm=20, steps=20, variant=0):
    s = sc.bench_scene(name)
    cfgd = s["configuration"]; blk = s["fluidBlocks"][0]
    r = cfgd["particleRadius"]
    size = [blk["end"][i] - blk["start"][i] for i in range(3)]
    x = sc.cube_positions(blk["start"], size, r, 3)
    n = len(x)
    eng = Engine(sc.gen2_config(cfgd, n, density_mode=mode))
    eng.set_param(K.P_KERNEL_VARIANT, variant)
    eng.add_particles(x, np.full(x.shape, blk["velocity"], np.float32), np.full(n, 1000.0, np.float32),
                      np.zeros(n, np.float32), np.ones(n, np.int32), None)
    del x
    eng.step(warm); eng.sync()
    eng.stage_times(True)
    t0 = time.time(); eng.step(steps); eng.sync(); wall = (time.time() - t0) / steps
    st = eng.stage_times(False)
    tot = st["update_ms"] + st["density_ms"] + st["force_ms"]
    print(f"{name} mode={mode} variant={variant} n={n} wall_ms={wall*1e3:.3f} stage_ms={st} total={tot:.3f} "
          f"Mupd/s={n/tot/1e3:.1f}", flush=True)
    eng.close()

if __name__ == "__main__":
    names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["C2", "C3"]
    variants = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
    for nm in names:
        for var in variants:
            for mode in (0, 1):
                run(nm, mode, variant=var)
