"""Randomised single-step parity: clouds of particles with random positions (edge cells of the
grid and cell 0 included), random velocities, two fluid densities and scattered boundary
particles, in all four density/volume mode combinations, against the CPU oracle."""
import numpy as np
import pytest

from oracle.oracle import Gen2Oracle
from ti_sph_b200 import _capi as K
from ti_sph_b200 import scene as sc
from ti_sph_b200.engine import Engine
from util import RTOL, check_force_stage, rel_err, small_scene

pytestmark = pytest.mark.gpu


def random_state(seed, n, domain, h):
    rng = np.random.default_rng(seed)
    kind = seed % 3
    if kind == 0:        # a dense blob (30-120 per cell) straddling the -x / -y / -z edge cells, cell 0 included
        x = rng.uniform(0.0005, 2.6 * h, size=(n, 3))
    elif kind == 1:      # the +x +y +z corner of the grid
        x = np.array(domain) - rng.uniform(0.0005, 2.2 * h, size=(n, 3))
    else:                # clusters of very different density in the interior
        centres = rng.uniform(0.3, 0.6, size=(6, 3))
        x = centres[rng.integers(0, 6, n)] + rng.normal(scale=rng.choice([0.004, 0.015, 0.04], size=(n, 1)), size=(n, 3))
        x = np.clip(x, 0.01, np.array(domain) - 0.01)
    v = rng.normal(scale=3.0, size=(n, 3))
    material = (rng.uniform(size=n) > 0.25).astype(np.int32)             # 25 % boundary particles
    density = np.where(rng.uniform(size=n) > 0.5, 1000.0, 4500.0)        # 4500: mass W(0) > rho0, p > 0 in reference mode
    return x.astype(np.float32), v.astype(np.float32), density.astype(np.float32), material


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("dmode,vmode", [("reference", "reference"), ("summed", "akinci"), ("summed", "reference")])
def test_random_cloud(seed, dmode, vmode):
    _random_cloud(seed, dmode, vmode, diagnostics=True)


@pytest.mark.parametrize("seed", [0, 2, 5])
@pytest.mark.parametrize("dmode,vmode", [("summed", "akinci"), ("reference", "reference")])
def test_random_cloud_plain_step(seed, dmode, vmode):
    """without TISPH_P_DIAGNOSTICS (a plain step): one accumulator for both force sums, boundary neighbours included"""
    _random_cloud(seed, dmode, vmode, diagnostics=False)


def _random_cloud(seed, dmode, vmode, diagnostics):
    scene = small_scene(domain_end=(1.0, 0.8, 0.6))
    h = 0.04
    x, v, density, material = random_state(seed, 6000, (1.0, 0.8, 0.6), h)
    ora = Gen2Oracle(scene, density_mode=dmode, volume_mode=vmode)
    ora.set_state(x, v, density, material)
    eng = Engine(sc.gen2_config(scene["configuration"], len(x), density_mode={"reference": 0, "summed": 1}[dmode],
                                volume_mode={"reference": 0, "akinci": 1}[vmode]))
    eng.add_particles(ora.x, ora.v, ora.density, ora.pressure, ora.material, ora.color)
    eng.set_param(K.P_DIAGNOSTICS, 1 if diagnostics else 0)
    t = ora.step(trace=True)
    eng.stage(K.STAGE_UPDATE)
    assert np.array_equal(eng.download(K.F_GRID_PARTICLES_NUM), t["scan"])
    assert np.array_equal(eng.download(K.F_ORIG_ID), t["orig"])
    eng.stage(K.STAGE_DENSITY)
    fl = t["material"] == 1
    # the oracle counts neighbours for every particle; the engine's count of a boundary particle is
    # only defined where the reference walks it (akinci volumes), so fluid rows are compared
    assert np.array_equal(eng.download(K.F_NEIGHBOR_COUNT)[fl], t["neighbor_count"][fl])
    assert rel_err(eng.download(K.F_DENSITY_SUM)[fl], t["S"][fl], floor=10.0) < RTOL
    assert rel_err(eng.download(K.F_VOLUME), t["volume"]) < RTOL
    assert rel_err(eng.download(K.F_DENSITY), t["density"]) < RTOL
    eng.stage(K.STAGE_FORCE_ADVECT)
    check_force_stage(eng, t, split=diagnostics)      # 1e-5 relative to the magnitude sums of the terms (util.accel_err)
    assert np.array_equal(eng.download(K.F_MATERIAL), t["material"])
    bd = ~fl
    assert np.array_equal(eng.download(K.F_X)[bd], t["x"][bd]) and np.array_equal(eng.download(K.F_V)[bd], t["v"][bd])
    eng.sync()
    eng.close()
