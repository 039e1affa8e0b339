"""2D dam break from add_cube -- counterpart of the reference's demo.py (ParticleSystem + WCSPH)."""
import argparse

from core.partice_system.partice_system import ParticleSystem
from core.sph.wcsph import WCSPH
from main import run

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--headless", action="store_true")
    ap.add_argument("--frames", type=int, default=20)
    args = ap.parse_args()
    ps = ParticleSystem((512, 512))
    ps.add_cube(lower_corner=[3, 1], cube_size=[3.0, 5.0], color=0x111111, velocity=[0, -20],
                density=1000.0, material=1)
    wcsph_solver = WCSPH(ps)
    run(ps, wcsph_solver, args)
