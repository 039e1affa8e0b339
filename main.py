"""2D dam break from a scene file -- counterpart of the reference's main.py (ParticleSystemV2 + WCSPH).

The reference opens ./data/scenes/demo.json, which it does not ship; demo_2d.json is its 2D scene.
With Taichi installed and no --headless flag the reference's ti.GUI loop is used, otherwise the
loop runs headless.  `python main.py --headless --frames 20`
"""
import argparse
import json
import time

from core.partice_system import partice_systemv2
from core.sph.wcsph import WCSPH


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="./data/scenes/demo_2d.json")
    ap.add_argument("--headless", action="store_true")
    ap.add_argument("--frames", type=int, default=20)
    args = ap.parse_args()
    with open(args.scene, "r") as f:
        simulation_config = json.load(f)
    ps = partice_systemv2.ParticleSystemV2((512, 512), simulation_config)
    ps.add_fluid_and_rigid()
    wcsph = WCSPH(ps)
    run(ps, wcsph, args)


def run(ps, wcsph, args):
    ti = None
    if not args.headless:
        try:
            import taichi as ti
            ti.init(arch=ti.gpu)
        except ImportError:
            print("taichi not installed: running headless")
    if ti is None:
        t0 = time.time()
        for _ in range(args.frames):
            for _ in range(5):
                wcsph.step()
            particle_info = ps.dump()
        n = ps.particle_num[None]
        print(f"{args.frames} frames x 5 steps, {n} particles in {time.time() - t0:.2f} s; "
              f"y range {particle_info['position'][:, 1].min():.3f}..{particle_info['position'][:, 1].max():.3f}")
        return
    gui = ti.GUI(background_color=0xFFFFFF)
    while gui.running:
        for _ in range(5):
            wcsph.step()
        particle_info = ps.dump()
        gui.circles(particle_info['position'] * ps.screen_to_world_ratio / 512,
                    radius=ps.particle_radius / 1.5 * ps.screen_to_world_ratio, color=0x111113)
        gui.show()


if __name__ == "__main__":
    main()
