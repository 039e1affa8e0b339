"""3D dam break -- counterpart of the reference's main_3d.py on the B200-native engine.

With Taichi installed and no --headless flag the ggui window of the reference is used
(scene.particles gets a Taichi mirror of ps.x); otherwise the loop runs headless and prints
throughput.  `python main_3d.py --headless --frames 20`
"""
import argparse
import json
import time

from core.partice_system import partice_systemv4
from core.sph.wcsphv2 import WCSPHV2
from utils.lines import getlines


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="./data/scenes/demo_3d.json")
    ap.add_argument("--headless", action="store_true")
    ap.add_argument("--frames", type=int, default=20)
    args = ap.parse_args()
    with open(args.scene, "r") as f:
        simulation_config = json.load(f)
    ti = None
    if not args.headless:
        try:
            import taichi as ti
            ti.init(arch=ti.cuda)
        except ImportError:
            print("taichi not installed: running headless")
    points, indices = getlines(simulation_config['configuration'])
    ps = partice_systemv4.ParticleSystemV4(simulation_config)
    wcsph = WCSPHV2(ps)
    if ti is None:
        t0 = time.time()
        for frame in range(args.frames):
            for _ in range(5):
                wcsph.step()
            particle_info = ps.dump()
        dt = time.time() - t0
        n = ps.particle_num[None]
        print(f"{args.frames} frames x 5 steps, {n} particles: {5 * args.frames * n / dt / 1e6:.1f} M particle-updates/s "
              f"(incl. dump); y range {particle_info['position'][:, 1].min():.3f}..{particle_info['position'][:, 1].max():.3f}")
        return
    window = ti.ui.Window('SPH', (1024, 1024), show_window=True, vsync=False)
    canvas, scene, camera = window.get_canvas(), window.get_scene(), ti.ui.Camera()
    camera.position(5.5, 2.5, 4.0); camera.up(0.0, 1.0, 0.0); camera.lookat(-1.0, 0.0, 0.0); camera.fov(70)
    while window.running:
        for _ in range(5):
            wcsph.step()
        camera.track_user_inputs(window, movement_speed=0.03, hold_key=ti.ui.LMB)
        scene.set_camera(camera)
        scene.point_light(pos=(2, 2, 2), color=(1, 1, 1))
        scene.particles(ps.x.to_taichi(), color=(0.68, 0.26, 0.19), radius=0.01)
        scene.lines(points, width=1.0, indices=indices, color=(0.99, 0.68, 0.28))
        canvas.scene(scene)
        window.show()


if __name__ == "__main__":
    main()
