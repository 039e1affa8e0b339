"""Target of the ncu captures under profiles/: C5 (or another workload), `steps` steps from the t=0 lattice in one
density mode.  The walk kernels of the LAST step are the ones captured:

    ncu --set full --import-source on --clock-control none -k regex:'k_(density|force)_list' \
        --launch-skip $((2 * (steps - 1))) --launch-count 2 -o gpurun_out/prof python scripts/ncu_target.py C5 reference 3
    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep r02_v20_C5 C5/reference 16000000
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from ti_sph_b200 import scene as sc
from ti_sph_b200.engine import Engine


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C5"
    mode = {"reference": 0, "summed": 1}[sys.argv[2] if len(sys.argv) > 2 else "reference"]
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    s = sc.bench_scene(name)
    cfg, blk = s["configuration"], s["fluidBlocks"][0]
    r = cfg["particleRadius"]
    x = sc.cube_positions(blk["start"], [blk["end"][i] - blk["start"][i] for i in range(3)], r, 3)
    n = len(x)
    eng = Engine(sc.gen2_config(cfg, n, density_mode=mode))
    eng.add_particles(x, np.full(x.shape, blk["velocity"], np.float32), np.full(n, 1000.0, np.float32),
                      np.zeros(n, np.float32), np.ones(n, np.int32), None)
    eng.step(steps)
    eng.sync()
    print(f"{name} mode={mode}: {n} particles, {steps} steps", flush=True)
    eng.close()


if __name__ == "__main__":
    main()
