"""Full BASELINE sizes on the GPU.

C3 (1 M particles) and C5 (16 M particles, the headline configuration): the whole step against the
CPU oracle, stage by stage, jittered lattice, both density modes (the oracle needs ~1 s / ~20 s
per step on the box's host cores).
C5 again: size-independent properties -- keys recomputed on the host, sortedness,
histogram / scan consistency, permutation, neighbour counts and density sums of a random sample
against a float64 KD-tree evaluation, momentum balance of the pair forces, bit-identical replay.
"""
import numpy as np
import pytest

from ti_sph_b200 import _capi as K
from ti_sph_b200 import scene as sc
from util import RTOL, check_force_stage, jitter, make_pair, rel_err, vec_rel_err

pytestmark = pytest.mark.gpu


def stage_by_stage(workload, radius, n_expected, mode, diagnostics=True):
    """one step of a BASELINE workload (jittered lattice) against the oracle, stage by stage:
    bit-exact histogram scan / order / neighbour counts, 1e-5 fields.  diagnostics=False: the force walk of a
    plain step (one accumulator for both sums -- what the benchmark times)"""
    scene = sc.bench_scene(workload)
    ora, eng = make_pair(scene, density_mode=mode)
    assert ora.n == n_expected
    x = jitter(ora.x, radius)
    ora.set_state(x, ora.v, ora.density, ora.material)
    eng.upload_xv(x, ora.v)
    t = ora.step(trace=True)
    eng.set_param(K.P_DIAGNOSTICS, 1 if diagnostics else 0)
    eng.stage(K.STAGE_UPDATE)
    assert np.array_equal(eng.download(K.F_CELL_COUNT), t["counts"])
    assert np.array_equal(eng.download(K.F_GRID_PARTICLES_NUM), t["scan"])
    assert np.array_equal(eng.download(K.F_GRID_IDS), t["keys"])
    assert np.array_equal(eng.download(K.F_ORIG_ID), t["orig"])
    assert np.array_equal(eng.download(K.F_X), t["x_sorted"])
    eng.stage(K.STAGE_DENSITY)
    assert np.array_equal(eng.download(K.F_NEIGHBOR_COUNT), t["neighbor_count"])
    assert rel_err(eng.download(K.F_DENSITY_SUM), t["S"], floor=1.0) < RTOL
    assert rel_err(eng.download(K.F_DENSITY_RAW), t["density_pre"]) < RTOL
    assert rel_err(eng.download(K.F_DENSITY), t["density"]) < RTOL
    p, p_ref = eng.download(K.F_PRESSURE).astype(np.float64), t["pressure"].astype(np.float64)
    x7 = (t["density"].astype(np.float64) / 1000.0) ** 7          # p = 50 (x^7 - 1) cancels near x = 1
    assert np.all(np.abs(p - p_ref) <= RTOL * np.abs(p_ref) + 50 * 8 * np.finfo(np.float32).eps * x7)
    eng.stage(K.STAGE_FORCE_ADVECT)
    check_force_stage(eng, t, split=diagnostics)      # 1e-5 relative to the magnitude sums of the terms (util.accel_err)
    assert np.array_equal(eng.download(K.F_MATERIAL), t["material"])
    assert int(eng.get_param(K.P_STAT_FALLBACK_FORCE)) == 0          # the fast path took everything
    eng.sync(); eng.close()


@pytest.mark.parametrize("mode", ["reference", "summed"])
def test_c3_one_million_particles_against_the_oracle(mode):
    stage_by_stage("C3", 0.01, 1_000_000, mode, diagnostics=(mode == "reference"))


@pytest.mark.parametrize("mode", ["reference", "summed"])
def test_c5_sixteen_million_particles_against_the_oracle(mode):
    """the headline configuration: 400 x 200 x 200 particles, r = 0.005"""
    # reference mode as the benchmark runs it (plain step), summed mode with the sums kept apart
    stage_by_stage("C5", 0.005, 16_000_000, mode, diagnostics=(mode == "summed"))


def test_c5_sixteen_million_particles_properties():
    from scipy.spatial import cKDTree
    scene = sc.bench_scene("C5")
    from core.partice_system.partice_systemv4 import ParticleSystemV4
    ps = ParticleSystemV4(scene, density_mode="summed")
    eng = ps.engine
    n = eng.particle_num
    assert n == 16_000_000
    x0 = eng.download(K.F_X)
    rng = np.random.default_rng(5)
    x0 = (x0 + rng.uniform(-0.2 * 0.005, 0.2 * 0.005, size=x0.shape)).astype(np.float32)   # off the knife edges
    eng.upload_xv(x0, eng.download(K.F_V))
    eng.save_state()
    eng.set_param(K.P_DIAGNOSTICS, 1)
    # ---- ps.update()
    eng.stage(K.STAGE_UPDATE)
    keys, xs = eng.download(K.F_GRID_IDS), eng.download(K.F_X)
    h = np.float32(0.02)
    cell = (xs / h).astype(np.int32)
    g = ps.grid_num.astype(np.int64)
    assert np.array_equal(keys, (cell[:, 0] * g[1] * g[2] + cell[:, 1] * g[2] + cell[:, 2]).astype(np.int32))
    assert np.all(np.diff(keys) >= 0)                                             # sorted by cell key
    scan, hist = eng.download(K.F_GRID_PARTICLES_NUM), eng.download(K.F_CELL_COUNT)
    assert scan[-1] == n and hist.sum() == n and np.array_equal(np.cumsum(hist, dtype=np.int64), scan)
    ids = eng.download(K.F_ORIG_ID)
    assert np.array_equal(np.sort(ids), np.arange(n, dtype=np.int32))             # a permutation
    assert np.array_equal(xs, x0[ids])                                            # records moved intact
    seg = np.nonzero(np.diff(keys) == 0)[0]
    assert np.all(ids[seg + 1] > ids[seg])                                        # stable inside every cell
    # ---- density: a random sample against float64
    eng.stage(K.STAGE_DENSITY)
    nc, S = eng.download(K.F_NEIGHBOR_COUNT), eng.download(K.F_DENSITY_SUM)
    assert nc.sum() % 2 == 0                                                      # neighbourhood is symmetric
    tree = cKDTree(xs.astype(np.float64))
    sample = rng.choice(n, 3000, replace=False)
    kw = 8 / np.pi / 0.02 ** 3
    mass = np.float64(np.float32(0.8 * 0.01 ** 3) * np.float32(1000.0))
    for i, nb in zip(sample, tree.query_ball_point(xs[sample].astype(np.float64), 0.02 * (1 - 1e-6))):
        nb = np.array([j for j in nb if j != i])
        q = np.linalg.norm(xs[nb].astype(np.float64) - xs[i], axis=1) / 0.02
        w = np.where(q <= 0.5, 6 * (q ** 3 - q ** 2) + 1, 2 * (1 - q) ** 3)
        assert abs(nc[i] - len(nb)) <= 1                                          # f32 vs f64 at the cutoff
        assert abs(S[i] - mass * kw * w.sum()) <= 2e-5 * max(S[i], 1.0)
    # ---- forces: the pair forces are antisymmetric, so the total momentum change vanishes
    eng.stage(K.STAGE_FORCE_ADVECT)
    a_p = eng.download(K.F_A_PRESSURE).astype(np.float64)
    assert np.linalg.norm(a_p.sum(axis=0)) < 1e-6 * np.abs(a_p).sum()
    out1 = (eng.download(K.F_X), eng.download(K.F_V))
    # ---- replay from the checkpoint: bit-identical
    eng.restore_state()
    eng.step(1)
    out2 = (eng.download(K.F_X), eng.download(K.F_V))
    assert np.array_equal(out1[0], out2[0]) and np.array_equal(out1[1], out2[1])
    assert int(eng.get_param(K.P_STAT_FALLBACK_FORCE)) == 0
    eng.sync(); eng.close()


def test_c2_shipped_demo_3d_scene_against_the_oracle_and_survey_values():
    """BASELINE config C2 = data/scenes/demo_3d.json as shipped (195,300 particles) through the
    drop-in classes; SURVEY section 4 known-answer values on the GPU result"""
    import json
    import os
    from core.partice_system.partice_systemv4 import ParticleSystemV4
    from core.sph.wcsphv2 import WCSPHV2
    from oracle.oracle import Gen2Oracle
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "data", "scenes", "demo_3d.json")) as f:
        scene = json.load(f)
    ps = ParticleSystemV4(scene)
    solver = WCSPHV2(ps)
    eng = ps.engine
    assert ps.particle_num[None] == 195300 and list(ps.grid_num) == [125, 75, 50]
    ora = Gen2Oracle(scene)
    t = ora.step(trace=True)
    solver.step()
    ids = eng.download(K.F_ORIG_ID)
    assert np.array_equal(ids, t["orig"])
    nc = eng.download(K.F_NEIGHBOR_COUNT)
    assert np.array_equal(nc, t["neighbor_count"]) and (nc.min(), nc.max(), int(nc.sum())) == (50, 255, 45273868)
    inv = np.empty(len(ids), np.int64); inv[ids] = np.arange(len(ids))
    S, a = eng.download(K.F_DENSITY_SUM), eng.download(K.F_D_VELOCITY)
    for idx, nn, s_ref, dv in [(0, 51, 1703.690, (19.0562, 9.2462, 19.0562)),
                               (99060, 253, 6145.846, (-1.0e-4, -9.80986, 1.0e-4)),
                               (1410, 150, 4065.557, (41.2480, -9.80990, 7.0e-5))]:
        s = inv[idx]
        assert nc[s] == nn and S[s] == pytest.approx(s_ref, rel=5e-6)
        assert np.allclose(a[s], dv, rtol=2e-5, atol=3e-5)
    d = ps.dump()
    assert rel_err(d["position"], ora.x, floor=0.04) < RTOL
    assert np.array_equal(d["material"], ora.material) and np.array_equal(d["color"], ora.color)
    eng.close()


def test_c4_four_million_particles_with_a_mesh_sampled_boundary():
    """BASELINE config C4: 200 x 200 x 100 fluid block (r = 0.005) over the voxelised Dragon_50k.obj
    (the reference's asset, data/models/), one step against the oracle fed with the same boundary points"""
    import sys
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import workload_scene
    from core.partice_system.partice_systemv4 import ParticleSystemV4
    from core.sph.wcsphv2 import WCSPHV2
    from oracle.oracle import Gen2Oracle
    scene = workload_scene("C4")
    ps = ParticleSystemV4(scene, volume_mode="akinci")
    solver = WCSPHV2(ps)
    eng = ps.engine
    pts = ps.rigidBodiesConfig[0]["voxelized_points"]
    n = ps.particle_num[None]
    assert scene["rigidBodies"][0]["geometryFile"].endswith("Dragon_50k.obj")
    assert n == 4_000_000 + len(pts) and abs(len(pts) - 129_815) <= 40          # tests/golden/dragon_c4_voxels.npz
    bare = dict(scene, rigidBodies=[dict(scene["rigidBodies"][0])])
    ora = Gen2Oracle(bare, volume_mode="akinci", boundary_points=pts)
    assert ora.n == n
    ora.step()
    solver.step()
    d = ps.dump()
    assert np.array_equal(eng.download(K.F_ORIG_ID), ora.orig)
    assert np.array_equal(d["material"], ora.material)
    assert rel_err(d["position"], ora.x, floor=0.02) < RTOL
    assert rel_err(ps.volume.to_numpy(), ora.volume) < RTOL
    assert vec_rel_err(d["velocity"], ora.v, floor=1.0) < 5 * RTOL
    assert int(eng.get_param(K.P_STAT_FALLBACK_FORCE)) == 0
    eng.close()
