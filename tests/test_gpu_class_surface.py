"""The drop-in classes driven kernel by kernel, exactly as tests/golden/make_golden.py:run_gen2 drives
the reference's ParticleSystemV4 / WCSPHV2 (update_gird_id, prefix_sum_executor.run, resort,
compute_volume_of_boundary_particle, compute_densities, compute_non_pressure_force,
compute_pressure_force, advert, enforce_boundary), and every snapshot compared with what the
reference's own sources recorded in tests/golden/*.npz.  Plus the extension points of the classes:
subclass hooks, assignable solver attributes, ps.update() + solver.step().
"""
import copy
import json
import os

import numpy as np
import pytest

from core.partice_system.partice_systemv4 import ParticleSystemV4
from core.sph.wcsphv2 import WCSPHV2
from ti_sph_b200 import _capi as K
from util import RTOL, accel_err, golden_gen2_force_reference, rel_err, small_scene

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def build(case, z, tmp_path):
    scene = copy.deepcopy(case["scene"])
    for k, rb in enumerate(scene["rigidBodies"]):
        path = os.path.join(tmp_path, f"rigid{k}.npy")
        np.save(path, z["points." + rb["geometryFile"]])
        rb["geometryFile"] = path
    ps = ParticleSystemV4(scene)
    return ps, WCSPHV2(ps)


@pytest.mark.parametrize("name", ["gen2_block", "gen2_walls", "gen2_two_blocks", "gen2_boundary"])
def test_kernel_by_kernel_driver_matches_the_reference_vectors(name, tmp_path):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    case = json.loads(str(z["case_json"]))
    ps, solver = build(case, z, str(tmp_path))
    n = int(ps.particle_num[None])
    assert n == int(z["n"]) == ps.particle_max_num
    ps.engine.set_param(K.P_DIAGNOSTICS, 1)          # ps.paritcle_index_temp is a diagnostic here

    def same(tag, *names, exact=True, floor=0.0):
        for nm in names:
            got = (solver.d_velocity if nm == "d_velocity" else getattr(ps, nm)).to_numpy()
            want = z[f"{tag}.{nm}"]
            if exact:
                assert np.array_equal(got, want), f"{tag}.{nm}"
            else:
                assert rel_err(got, want, floor=floor) < RTOL, f"{tag}.{nm}"

    same("init", "x", "v", "density", "pressure", "material", "color", "mass", "volume")
    for s in range(1):        # (the second golden step starts from the reference's f32 trajectory: test_gpu_golden.py)
        t = f"s{s}"
        # ---- SPHBaseV2.step(), sph_basev2.py:210-214, call by call
        ps.update_gird_id()
        assert np.array_equal(ps.grid_particles_num.to_numpy(), z[f"{t}.counts"])
        assert np.array_equal(ps.grid_ids.to_numpy(), z[f"{t}.sorted.grid_ids"][z[f"{t}.sorted.paritcle_index_temp"]])
        ps.prefix_sum_executor.run(ps.grid_particles_num)
        ps.resort()
        same(t + ".sorted", "grid_ids", "grid_particles_num", "paritcle_index_temp", "x", "v", "density", "pressure",
             "material", "color", "mass", "volume")
        solver.compute_volume_of_boundary_particle()
        same(t + ".volume", "volume", exact=False)
        solver.compute_densities()
        same(t + ".density", "density", exact=False)
        assert np.array_equal(ps.pressure.to_numpy(), z[f"{t}.sorted.pressure"])     # not written yet
        solver.compute_non_pressure_force()
        fl = z[f"{t}.sorted.material"] == 1
        ref = golden_gen2_force_reference(case, z, s)       # golden arrays + the magnitude sums that scale 1e-5
        assert accel_err(solver.d_velocity.to_numpy(), ref["a_nonpressure"], ref["mag_nonpressure"]) < RTOL
        assert np.array_equal(ps.x.to_numpy(), z[f"{t}.sorted.x"])                   # not advected yet
        solver.compute_pressure_force()
        same(t + ".pressure", "density", exact=False)
        p, p_ref = ps.pressure.to_numpy().astype(np.float64), z[f"{t}.pressure.pressure"].astype(np.float64)
        x7 = (z[f"{t}.pressure.density"].astype(np.float64) / 1000.0) ** 7
        assert np.all(np.abs(p - p_ref)[fl] <= (RTOL * np.abs(p_ref) + 50 * 8 * np.finfo(np.float32).eps * x7)[fl])
        mag = ref["mag_nonpressure"].astype(np.float64) + ref["mag_pressure"]
        assert accel_err(solver.d_velocity.to_numpy(), ref["d_velocity"], mag, ref["mag_pressure_floor"]) < RTOL
        scale = float((mag + ref["mag_pressure_floor"] / RTOL).max())
        assert np.array_equal(ps.x.to_numpy(), z[f"{t}.sorted.x"])
        solver.advert()
        same(t + ".advert", "x", exact=False, floor=0.04)       # advected, walls not applied
        assert rel_err(ps.v.to_numpy(), z[f"{t}.advert.v"], floor=1.0) < RTOL + 2e-4 * scale * RTOL
        solver.enforce_boundary()
        same(t + ".end", "x", exact=False, floor=0.04)
        assert rel_err(ps.v.to_numpy(), z[f"{t}.end.v"], floor=1.0) < RTOL + 2e-4 * scale * RTOL
        same(t + ".end", "material", "color", "mass")
        d = ps.dump()
        assert np.array_equal(d["material"], z[f"{t}.dump.material"]) and np.array_equal(d["color"], z[f"{t}.dump.color"])
        assert rel_err(d["position"], z[f"{t}.dump.position"], floor=0.04) < RTOL
        np_x = np.zeros((n, 3), np.float32)
        ps.copy_to_numpy_nd(np_x, ps.x)
        np_m = np.zeros(n, np.int32)
        ps.copy_to_numpy(np_m, ps.material)
        assert np.array_equal(np_x, d["position"]) and np.array_equal(np_m, d["material"])
    if name == "gen2_walls":     # the walls moved something in this case, so advert() and the end state differ
        assert not np.array_equal(z["s0.advert.x"], z["s0.end.x"])
    # the unmodified step() on a second system reaches the same state.  Not bit for bit: driven kernel by kernel the
    # force walk keeps the non-pressure and the pressure sums apart, the fused step adds both terms of a pair into
    # one accumulator -- two f32 summation orders of the same terms, each within the tolerance above.
    ps2, solver2 = build(case, z, str(tmp_path))
    solver2.step()
    assert rel_err(ps2.x.to_numpy(), ps.x.to_numpy(), floor=0.04) < 1e-6
    assert rel_err(ps2.v.to_numpy(), ps.v.to_numpy(), floor=1.0) < 2 * (RTOL + 2e-4 * scale * RTOL)
    assert rel_err(ps2.v.to_numpy(), z[f"{t}.end.v"], floor=1.0) < RTOL + 2e-4 * scale * RTOL     # ... and the reference's
    ps.engine.close(); ps2.engine.close()


def test_subclass_hooks_are_called_in_the_reference_order():
    calls = []

    class Probe(WCSPHV2):
        def substep(self):
            calls.append("substep")
            super().substep()

        def enforce_boundary(self):
            calls.append("enforce_boundary")
            super().enforce_boundary()

    scene = small_scene(end=(0.4, 0.2, 0.8))
    ps, ps_ref = ParticleSystemV4(copy.deepcopy(scene)), ParticleSystemV4(copy.deepcopy(scene))
    probe, stock = Probe(ps), WCSPHV2(ps_ref)
    for _ in range(2):
        probe.step()
        stock.step()
    assert calls == ["substep", "enforce_boundary"] * 2
    # (same terms, two f32 summation orders: the hooked solver runs kernel by kernel with the force sums kept apart,
    #  the stock one fused with one accumulator -- hence a tolerance, not bit equality)
    assert rel_err(ps.x.to_numpy(), ps_ref.x.to_numpy(), floor=0.04) < 1e-6
    assert rel_err(ps.v.to_numpy(), ps_ref.v.to_numpy(), floor=1.0) < 4 * RTOL
    # ps.update() by the script, then solver.step(): the step continues from the sorted state
    ps.update()
    probe.step()
    ps_ref.update()
    stock.step()
    assert rel_err(ps.x.to_numpy(), ps_ref.x.to_numpy(), floor=0.04) < 1e-6
    ps.engine.close(); ps_ref.engine.close()


def test_solver_attributes_are_assignable_after_construction():
    from oracle.oracle import Gen2Oracle
    scene = small_scene(end=(0.4, 0.2, 0.8))
    scene["fluidBlocks"][0]["density"] = 4000.0          # mass W(0) > rho0: the Tait pressure works
    ps = ParticleSystemV4(copy.deepcopy(scene))
    solver = WCSPHV2(ps)
    solver.stiffness, solver.exponent, solver.viscosity = 80.0, 5.0, 0.02
    solver.g[1] = -3.0
    solver.dt[None] = 1e-4
    assert (solver.stiffness, solver.exponent) == (80.0, 5.0) and solver.viscosity == pytest.approx(0.02)
    ora = Gen2Oracle(scene)
    ora.cfg.stiffness, ora.cfg.exponent, ora.cfg.dt = 80.0, 5.0, 1e-4
    ora.cfg.visc_fluid_c = 2 * 0.02 * ora.support_length * scene["configuration"]["c_s"]
    ora.cfg.g[1] = -3.0
    solver.step()
    t = ora.step(trace=True)
    assert np.array_equal(ps.engine.download(K.F_ORIG_ID), t["orig"])
    assert float(np.abs(t["pressure"]).max()) > 0
    assert rel_err(ps.pressure.to_numpy(), t["pressure"], floor=1.0) < RTOL
    assert rel_err(ps.x.to_numpy(), t["x"], floor=0.04) < RTOL
    assert rel_err(ps.v.to_numpy(), t["v"], floor=1.0) < RTOL
    ps.engine.close()


def test_zero_copy_view_right_after_step_needs_no_manual_sync():
    """ps.x handed to torch straight after step(): the view orders itself after the engine's stream"""
    import torch
    ps = ParticleSystemV4(small_scene(end=(0.5, 0.3, 0.9)))
    solver = WCSPHV2(ps)
    for _ in range(3):
        solver.step()
        x_dev = ps.x.to_torch()                    # no engine.sync() in between
        x_host = ps.x.to_numpy()
        assert np.array_equal(x_dev.cpu().numpy(), x_host)
    ps.engine.close()


@pytest.mark.parametrize("name", ["gen1_cube", "gen1_scene"])
def test_gen1_kernel_by_kernel_driver_matches_the_reference_vectors(name):
    """the gen-1 classes driven like tests/golden/make_golden.py:run_gen1 drives the reference's"""
    from core.partice_system.partice_system import ParticleSystem
    from core.partice_system.partice_systemv2 import ParticleSystemV2
    from core.sph.wcsph import WCSPH
    from oracle.oracle import Gen1Oracle
    z = np.load(os.path.join(GOLD, name + ".npz"))
    case = json.loads(str(z["case_json"]))
    if case["kind"] == "v1":
        ps = ParticleSystem(tuple(case["res"]))
        ps.add_cube(**case["cube"])
    else:
        ps = ParticleSystemV2(tuple(case["res"]), case["scene"])
        ps.add_fluid_and_rigid()
    solver = WCSPH(ps)
    g = lambda k: z[f"s0.{k}"]
    ps.init()
    assert np.array_equal(ps.particle_neighbors.to_numpy(), g("init.particle_neighbors"))
    assert np.array_equal(ps.particle_neighbors_num.to_numpy(), g("init.particle_neighbors_num"))
    solver.compute_volume_of_boundary_particle()
    solver.compute_densities()
    assert rel_err(ps.density.to_numpy(), g("density.density"), floor=1.0) < RTOL
    ora = Gen1Oracle(tuple(case["res"]))
    ora.material = z["init.material"]
    mags = ora.force_magnitudes(z["init.x"], z["init.v"], g("density.density"), g("pressure.density"),
                                g("pressure.pressure"), g("init.particle_neighbors"), g("init.particle_neighbors_num"))
    solver.compute_non_pressure_force()
    assert accel_err(solver.d_velocity.to_numpy(), g("nonpressure.d_velocity"), mags["mag_nonpressure"]) < RTOL
    assert np.array_equal(ps.x.to_numpy(), z["init.x"])                      # not advected yet
    solver.compute_pressure_force()
    assert rel_err(ps.density.to_numpy(), g("pressure.density")) < RTOL
    mag = mags["mag_nonpressure"].astype(np.float64) + mags["mag_pressure"]
    assert accel_err(solver.d_velocity.to_numpy(), g("pressure.d_velocity"), mag,
                     mags["mag_pressure_floor"] + RTOL * mags["mag_pressure"]) < RTOL
    solver.advert()
    solver.enforce_boundary()
    assert rel_err(ps.x.to_numpy(), g("end.x"), floor=0.2) < RTOL
    # the unmodified step() continues from here like a fresh system that took one fused step
    solver.step()
    assert ps._kernel_stage == 0
    ps.engine.close()


def test_gen1_grid_and_neighbour_kernels_called_on_their_own():
    """partice_system.py:211-215: init() is fill + allocate_particles_to_grid() + search_neighbors(); a script may call
    the two kernels itself, and copy_to_numpy(_nd) (:167-176) instead of dump()"""
    from core.partice_system.partice_system import ParticleSystem
    from core.sph.wcsph import WCSPH
    z = np.load(os.path.join(GOLD, "gen1_cube.npz"))
    case = json.loads(str(z["case_json"]))
    ps = ParticleSystem(tuple(case["res"]))
    ps.add_cube(**case["cube"])
    solver = WCSPH(ps)
    ps.allocate_particles_to_grid()
    ps.search_neighbors()
    assert np.array_equal(ps.particle_neighbors.to_numpy(), z["s0.init.particle_neighbors"])
    assert np.array_equal(ps.particle_neighbors_num.to_numpy(), z["s0.init.particle_neighbors_num"])
    n = ps.particle_num[None]
    x = np.zeros((n + 3, 2), np.float32)
    mat = np.zeros(n + 3, np.int32)
    ps.copy_to_numpy_nd(x, ps.x)
    ps.copy_to_numpy(mat, ps.material)
    assert np.array_equal(x[:n], z["init.x"]) and not x[n:].any()
    assert np.array_equal(mat[:n], z["init.material"]) and not mat[n:].any()
    solver.step()                                   # continues from the grid the two calls built
    assert rel_err(ps.x.to_numpy(), z["s0.end.x"], floor=0.2) < RTOL
    from ti_sph_b200 import _capi as K
    assert int(ps.engine.get_param(K.P_PHASE)) == 0
    ps.search_neighbors()                           # a new state: the kernel rebuilds grid and table
    assert int(ps.engine.get_param(K.P_PHASE)) == 1
    cnt = ps.particle_neighbors_num.to_numpy()
    want = z["s1.init.particle_neighbors_num"]      # (the reference's own next state: knife-edge pairs of the lattice may flip)
    assert cnt.shape == want.shape and np.abs(cnt.astype(np.int64) - want).max() <= 4 and (cnt == want).mean() > 0.9
    ps.engine.close()
