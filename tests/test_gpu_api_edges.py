"""Edge cases of the C ABI / drop-in classes on the GPU: empty and over-full systems, particles
outside the grid, stage order, dt changes, checkpoint replay, the zero-copy device view."""
import copy

import numpy as np
import pytest

from ti_sph_b200 import _capi as K
from ti_sph_b200 import scene as sc
from ti_sph_b200._capi import TisphError
from ti_sph_b200.engine import Engine
from util import make_pair, small_scene

pytestmark = pytest.mark.gpu


def _engine(cap=1000):
    return Engine(sc.gen2_config(small_scene()["configuration"], cap))


def test_empty_system_refuses_to_step():
    eng = _engine()
    assert eng.particle_num == 0
    with pytest.raises(TisphError, match="no particles"):
        eng.step(1)
    assert eng.download(K.F_X).shape == (0, 3)
    eng.close()


def test_capacity_is_enforced_and_reported():
    eng = _engine(cap=10)
    x = np.full((8, 3), 0.5, np.float32); v = np.zeros((8, 3), np.float32)
    one = np.ones(8, np.float32); mat = np.ones(8, np.int32)
    eng.add_particles(x, v, 1000 * one, 0 * one, mat)
    with pytest.raises(TisphError) as e:
        eng.add_particles(x, v, 1000 * one, 0 * one, mat)
    assert e.value.code == -3 and "exceeds particle_max_num" in str(e.value)
    assert eng.particle_num == 8                                   # nothing was added
    eng.close()


def test_particle_outside_the_grid_is_flagged_not_undefined():
    """the reference indexes out of bounds here; the engine clamps the key and reports"""
    eng = _engine()
    x = np.array([[0.5, 0.5, 0.5], [7.0, 0.5, 0.5]], np.float32)   # domain is 5 x 3 x 2
    eng.add_particles(x, np.zeros((2, 3), np.float32), np.full(2, 1000, np.float32), np.zeros(2, np.float32),
                      np.array([1, 0], np.int32))
    eng.step(1)
    with pytest.raises(TisphError) as e:
        eng.sync()
    assert e.value.code == -4 and "left the grid" in str(e.value)
    eng.sync()                                                     # the flag is cleared once reported
    eng.close()


def test_stages_must_be_issued_in_order():
    ora, eng = make_pair(small_scene(end=(0.4, 0.2, 0.8)))
    with pytest.raises(TisphError, match="out of order"):
        eng.stage(K.STAGE_DENSITY)
    eng.stage(K.STAGE_UPDATE)
    with pytest.raises(TisphError, match="out of order"):
        eng.stage(K.STAGE_FORCE_ADVECT)
    with pytest.raises(TisphError, match="middle"):
        eng.step(1)
    eng.stage(K.STAGE_DENSITY); eng.stage(K.STAGE_FORCE_ADVECT)
    eng.step(1)
    eng.sync(); eng.close()


def test_dt_is_a_settable_zero_d_field_like_the_reference():
    from core.partice_system.partice_systemv4 import ParticleSystemV4
    from core.sph.wcsphv2 import WCSPHV2
    from oracle.oracle import Gen2Oracle
    scene = small_scene(end=(0.4, 0.2, 0.8))
    ps = ParticleSystemV4(copy.deepcopy(scene)); solver = WCSPHV2(ps)
    assert solver.dt[None] == pytest.approx(2e-4)
    solver.dt[None] = 1e-4
    ora = Gen2Oracle(scene); ora.cfg.dt = 1e-4
    solver.step(); ora.step()
    assert np.abs(ps.x.to_numpy() - ora.x).max() < 1e-6
    assert np.array_equal(ps.dump()["color"], ora.dump()["color"])
    ps.engine.close()


def test_checkpoint_replay_is_bit_identical():
    ora, eng = make_pair(small_scene(end=(0.45, 0.25, 0.85)))
    eng.step(2); eng.save_state()
    eng.step(3); a = (eng.download(K.F_X), eng.download(K.F_V), eng.download(K.F_ORIG_ID))
    eng.restore_state()
    eng.step(3); b = (eng.download(K.F_X), eng.download(K.F_V), eng.download(K.F_ORIG_ID))
    for p, q in zip(a, b):
        assert np.array_equal(p, q)
    eng.close()


def test_zero_copy_view_of_ps_x_matches_dump():
    import torch
    from core.partice_system.partice_systemv4 import ParticleSystemV4
    from core.sph.wcsphv2 import WCSPHV2
    ps = ParticleSystemV4(small_scene(end=(0.4, 0.2, 0.8))); solver = WCSPHV2(ps)
    solver.step()
    ps.engine.sync()
    t = ps.x.to_torch()                                            # strided view of the float4 records
    assert t.is_cuda and t.shape == (ps.particle_num[None], 3)
    assert np.array_equal(t.cpu().numpy(), ps.dump()["position"])
    ps.engine.close()


def test_gen1_capacity_is_two_to_the_fifteen():
    from core.partice_system.partice_system import ParticleSystem
    ps = ParticleSystem((512, 512))
    assert ps.particle_max_num == 2 ** 15
    ps.add_cube(lower_corner=[1, 1], cube_size=[8.0, 8.0], material=1)          # 160 x 160 = 25,600
    with pytest.raises(AssertionError):                                          # partice_system.py:150
        ps.add_cube(lower_corner=[1, 1], cube_size=[8.0, 8.0], material=1)
    ps.engine.close()


def test_cfl_time_step_extension():
    """TISPH_P_CFL (extension): dt = min(dt_max, cfl h / (c_s + max|v|)) before every step"""
    from oracle.oracle import Gen2Oracle
    scene = small_scene(end=(0.4, 0.2, 0.8))                        # v0 = (0,-1,10): |v| = 10.05
    ora, eng = make_pair(scene)
    eng.set_param(K.P_CFL, 0.4)
    eng.step(1)
    dt = eng.get_param(K.P_DT)
    expect = 0.4 * np.float32(0.04) / (np.float32(88.5) + np.sqrt(np.float32(101.0)))
    assert dt == pytest.approx(float(expect), rel=1e-6) and dt < 2e-4
    ora.cfg.dt = dt
    ora.step()
    assert np.abs(eng.download(K.F_X) - ora.x).max() < 1e-6
    eng.set_param(K.P_CFL, 0.0)                                     # back to the reference's fixed step
    assert eng.get_param(K.P_DT) == pytest.approx(2e-4)
    eng.close()


def test_asynchronous_upload_and_dump_match_the_blocking_calls():
    """tisph_upload_xv_async / tisph_dump_async / tisph_dump_wait (pinned host arrays, copy streams) against
    upload_xv / dump(), pipelined over a few steps the way bench.py's e2e loop uses them"""
    import torch
    from core.partice_system.partice_systemv4 import ParticleSystemV4
    from core.sph.wcsphv2 import WCSPHV2
    scene = small_scene(end=(0.5, 0.3, 0.9))
    ps_a, ps_b = ParticleSystemV4(copy.deepcopy(scene)), ParticleSystemV4(copy.deepcopy(scene))
    sa, sb = WCSPHV2(ps_a), WCSPHV2(ps_b)
    n = ps_a.particle_num[None]
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
    hx, hv = pin((n, 3), torch.float32), pin((n, 3), torch.float32)
    outs = {"position": pin((n, 3), torch.float32), "velocity": pin((n, 3), torch.float32),
            "material": pin((n,), torch.int32), "color": pin((n, 3), torch.int32)}
    d = ps_b.dump()
    for step in range(4):
        hx[:] = d["position"]; hv[:] = d["velocity"] * np.float32(1.0 + 0.01 * step)
        ps_a.engine.upload_xv_async(hx, hv)
        sa.step()
        ps_a.dump_async(outs)
        ps_b.engine.upload_xv(hx.copy(), hv.copy())
        sb.step()
        d = ps_b.dump()
        ps_a.dump_wait()
        for k in ("position", "velocity", "material", "color"):
            assert np.array_equal(outs[k], d[k]), (step, k)
    ids = pin((n,), torch.int32)
    ps_a.engine.dump_async(orig_id=ids)
    ps_a.engine.sync()                                       # sync() completes a pending dump too
    assert np.array_equal(ids, ps_b.engine.download(K.F_ORIG_ID))
    ps_a.engine.close(); ps_b.engine.close()


def test_skipping_the_discarded_sum_changes_no_field_of_the_reference():
    """TISPH_P_SKIP_DISCARDED_SUM (opt-in): in the reference density mode the neighbour sum is overwritten
    (wcsphv2.py:32-34), so a density walk that only builds the lists must give bit-identical x, v, density,
    pressure and d_velocity; in summed mode the switch has no effect"""
    for mode in ("reference", "summed"):
        scene = small_scene(end=(0.5, 0.3, 0.9))
        ora, a = make_pair(scene, density_mode=mode)
        _, b = make_pair(scene, density_mode=mode)
        b.set_param(K.P_SKIP_DISCARDED_SUM, 1)
        a.step(3); b.step(3)
        for f in (K.F_X, K.F_V, K.F_DENSITY, K.F_PRESSURE, K.F_D_VELOCITY, K.F_ORIG_ID):
            assert np.array_equal(a.download(f), b.download(f)), (mode, f)
        same_sum = np.array_equal(a.download(K.F_DENSITY_SUM), b.download(K.F_DENSITY_SUM))
        assert same_sum == (mode == "summed")
        a.close(); b.close()
