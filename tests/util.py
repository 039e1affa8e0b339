"""Shared helpers of the parity tests: scenes, the oracle/engine pair, tolerances."""
import copy

import numpy as np

from oracle.oracle import Gen2Oracle
from ti_sph_b200 import _capi, scene as sc
from ti_sph_b200.engine import Engine

RTOL = 1e-5    # BASELINE.json north_star: density, pressure, acceleration within 1e-5 relative (fp32)


def small_scene(start=(0.3, 0.1, 0.7), end=(0.6, 0.4, 1.0), velocity=(0.0, -1.0, 10.0), radius=0.01,
                domain_end=(5.0, 3.0, 2.0)):
    s = copy.deepcopy(sc.DEMO_3D)
    s["configuration"]["particleRadius"] = radius
    s["configuration"]["domainEnd"] = list(domain_end)
    blk = s["fluidBlocks"][0]
    blk["start"], blk["end"], blk["velocity"] = list(start), list(end), list(velocity)
    return s


def jitter(x, radius, seed=1234, amp=0.2):
    """lattice + uniform jitter in [-amp r, amp r] (SURVEY 8(d) parity state ii)."""
    rng = np.random.default_rng(seed)
    return (x + rng.uniform(-amp * radius, amp * radius, size=x.shape)).astype(np.float32)


def make_pair(scene, density_mode="reference", volume_mode="reference", x=None, v=None,
              boundary_points=None):
    """Oracle and CUDA engine holding the same initial state."""
    ora = Gen2Oracle(scene, density_mode=density_mode, volume_mode=volume_mode,
                     boundary_points=boundary_points)
    if x is not None:
        ora.set_state(x, ora.v if v is None else v, ora.density, ora.material)
    cfg = sc.gen2_config(scene["configuration"], ora.n,
                         density_mode={"reference": 0, "summed": 1}[density_mode],
                         volume_mode={"reference": 0, "akinci": 1}[volume_mode])
    eng = Engine(cfg)
    eng.add_particles(ora.x, ora.v, ora.density, ora.pressure, ora.material, ora.color)
    return ora, eng


def rel_err(a, b, floor=0.0):
    """max |a-b| / max(|b|, floor), elementwise for scalars"""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0


def vec_rel_err(a, b, floor):
    """max over particles of |a-b| / max(|b|, floor) with vector norms (hard part 2)"""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    d = np.linalg.norm(a - b, axis=1)
    s = np.maximum(np.linalg.norm(b, axis=1), floor)
    return float(np.max(d / s)) if len(a) else 0.0
