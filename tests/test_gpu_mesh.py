"""GPU voxeliser (tisph_voxelize_mesh) and a rigid body through the drop-in ParticleSystemV4."""
import copy

import numpy as np
import pytest

from ti_sph_b200 import _capi as K
from ti_sph_b200 import mesh, scene as sc

pytestmark = pytest.mark.gpu


def brute_force_surface(v, f, pitch, centres):
    """float64 triangle/box separating-axis test, vectorised over voxel centres"""
    hw = pitch / 2
    hit = np.zeros(len(centres), bool)
    for tri in f:
        t = v[tri]                                           # [3,3]
        lo, hi = t.min(0) - hw, t.max(0) + hw
        cand = np.nonzero(np.all((centres >= lo) & (centres <= hi), axis=1) & ~hit)[0]
        if not len(cand):
            continue
        p = t[None, :, :] - centres[cand][:, None, :]        # [m,3,3]
        e = np.stack([p[:, 1] - p[:, 0], p[:, 2] - p[:, 1], p[:, 0] - p[:, 2]], 1)
        ok = np.ones(len(cand), bool)
        axes = [np.cross(e[:, 0], e[:, 1])]
        for k in range(3):
            for u in np.eye(3):
                axes.append(np.cross(np.broadcast_to(u, e[:, k].shape), e[:, k]))
        for a in axes:
            proj = np.einsum("mij,mj->mi", p, a)
            r = hw * np.abs(a).sum(1)
            ok &= ~((proj.min(1) > r) | (proj.max(1) < -r))
        hit[cand[ok]] = True
    return hit


def test_sphere_surface_and_fill():
    R, pitch = 0.2, 0.02
    v, f = mesh.icosphere(R, (0.5, 0.6, 0.7), subdivisions=3)
    surf = mesh.voxelize(v, f, pitch, fill=False)
    full = mesh.voxelize(v, f, pitch, fill=True)
    c = np.array([0.5, 0.6, 0.7])
    ds = np.linalg.norm(surf - c, axis=1)
    assert ds.max() < R + pitch * 0.87 and ds.min() > R * 0.97 - pitch * 0.87      # a shell one voxel thick
    df = np.linalg.norm(full - c, axis=1)
    assert df.max() < R + pitch * 0.87
    assert len(full) > len(surf)
    vol = len(full) * pitch ** 3
    assert abs(vol / (4 / 3 * np.pi * R ** 3) - 1) < 0.25                           # volume + half a shell
    # every lattice point well inside the sphere is filled
    k = np.round(full / pitch).astype(np.int64)
    assert np.allclose(k * pitch, full, atol=1e-6)                                  # centres on the world lattice
    have = set(map(tuple, k))
    g = np.arange(-12, 13)
    pts = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3) + np.round(c / pitch).astype(int)
    inside = np.linalg.norm(pts * pitch - c, axis=1) < R * 0.97 - pitch
    assert all(tuple(p) in have for p in pts[inside])


def test_surface_matches_a_float64_separating_axis_reference():
    ctr = np.array([0.4037, 0.3071, 0.4113])          # off-lattice: no triangle exactly touches a voxel face
    v, f = mesh.torus(0.15, 0.05, ctr, nu=24, nv=12)
    pitch = 0.02
    got = mesh.voxelize(v, f, pitch, fill=False)
    lo = np.floor(v.min(0) / pitch + 0.5).astype(int) - 1
    hi = np.floor(v.max(0) / pitch + 0.5).astype(int) + 1
    grid = np.stack(np.meshgrid(*[np.arange(lo[k], hi[k] + 1) for k in range(3)], indexing="ij"), -1).reshape(-1, 3)
    ref = brute_force_surface(v, f, pitch, grid * pitch)
    got_set = set(map(tuple, np.round(got / pitch).astype(int)))
    ref_set = set(map(tuple, grid[ref]))
    # f32 vs f64 may disagree only for voxels that a triangle touches within rounding error
    assert len(got_set ^ ref_set) <= max(2, len(ref_set) // 200)
    # a torus encloses its tube but not its hole
    full = mesh.voxelize(v, f, pitch, fill=True)
    centre_hole = np.linalg.norm(full - ctr, axis=1).min()
    assert centre_hole > 0.05 and len(full) > len(got)


def test_rigid_body_scene_through_the_drop_in_class(tmp_path):
    from core.partice_system.partice_systemv4 import ParticleSystemV4
    from core.sph.wcsphv2 import WCSPHV2
    from oracle.oracle import Gen2Oracle
    from util import RTOL, rel_err
    v, f = mesh.icosphere(0.06, (0.0, 0.0, 0.0), subdivisions=2)
    obj = tmp_path / "ball.obj"
    mesh.write_obj(obj, v, f)
    scene = copy.deepcopy(sc.DEMO_3D)
    scene["configuration"]["domainEnd"] = [1.0, 1.0, 1.0]
    scene["rigidBodies"] = [{"geometryFile": str(obj), "scale": [1, 1, 1], "translation": [0.4, 0.2, 0.4],
                             "rotationAngle": 30, "rotationAxis": [0, 1, 0], "color": [255, 255, 255],
                             "velocity": [0.0, 0.0, 0.0], "density": 1000.0}]
    scene["fluidBlocks"][0].update(start=[0.32, 0.27, 0.32], end=[0.48, 0.36, 0.48], velocity=[0.0, -2.0, 0.0])
    for vmode in ("reference", "akinci"):
        ps = ParticleSystemV4(copy.deepcopy(scene), volume_mode=vmode, density_mode="summed")
        solver = WCSPHV2(ps)
        pts = ps.rigidBodiesConfig[0]["voxelized_points"]
        assert len(pts) > 100 and np.all(ps.material.to_numpy()[:len(pts)] == 0)
        ora = Gen2Oracle(scene, density_mode="summed", volume_mode=vmode, boundary_points=pts)
        assert ora.n == ps.particle_num[None]
        solver.step()
        ora.step()
        d = ps.dump()
        assert np.array_equal(ps.engine.download(K.F_ORIG_ID), ora.orig)
        assert rel_err(d["position"], ora.x, floor=0.04) < 5 * RTOL
        assert rel_err(ps.volume.to_numpy(), ora.volume) < RTOL
        ps.engine.close()


def test_dragon_50k_voxelised_in_the_c4_placement_matches_the_float64_golden():
    """The reference's only mesh asset (data/models/Dragon_50k.obj, partice_systemv4.py:259-277) at the
    C4 placement and pitch: the GPU sampler against tests/golden/dragon_c4_voxels.npz (float64
    separating-axis surface + scipy binary_fill_holes, written by tests/golden/make_dragon_golden.py).
    The mesh is closed but not 2-manifold (74,968 edges with two faces, 16 with four, none with one),
    so the outside flood fill of the sampler is well defined: fill = everything the flood cannot reach."""
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from bench import workload_scene
    gold = np.load(os.path.join(root, "tests", "golden", "dragon_c4_voxels.npz"))
    scene = workload_scene("C4")
    body = scene["rigidBodies"][0]
    assert body["geometryFile"].endswith(os.path.join("data", "models", "Dragon_50k.obj"))
    assert np.allclose(body["translation"], gold["translation"])
    v, f = mesh.load_obj(body["geometryFile"])
    assert (len(v), len(f)) == (int(gold["n_vertices"]), int(gold["n_faces"])) == (25007, 50000)
    edges = np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), axis=1)
    _, cnt = np.unique(edges, axis=0, return_counts=True)
    assert np.bincount(cnt).tolist() == gold["edge_face_histogram"].tolist() == [0, 0, 74968, 0, 16]   # closed surface
    pitch = float(gold["pitch"])
    dims, lo = tuple(int(d) for d in gold["dims"]), gold["lo"].astype(np.int64)
    n = int(np.prod(dims))
    want_surface = np.unpackbits(gold["surface"])[:n].reshape(dims).astype(bool)
    want_full = np.unpackbits(gold["filled"])[:n].reshape(dims).astype(bool)
    assert (int(want_surface.sum()), int(want_full.sum())) == (39759, 129815)
    vt = mesh.transform_vertices(v, body)

    def occupancy(points):
        k = np.round(points.astype(np.float64) / pitch).astype(np.int64) - lo
        assert np.all(k >= 0) and np.all(k < np.array(dims))
        occ = np.zeros(dims, bool)
        occ[k[:, 0], k[:, 1], k[:, 2]] = True
        assert occ.sum() == len(points)                       # no duplicates
        return occ

    surf = occupancy(mesh.voxelize(vt, f, pitch, fill=False))
    full = occupancy(mesh.sample_rigid_body(dict(body), pitch))
    # f32 (GPU) vs f64 (golden) may disagree only where a triangle touches a voxel face within rounding
    assert int((surf ^ want_surface).sum()) <= 40
    diff = full ^ want_full
    assert int(diff.sum()) <= 40 and not np.any(diff & ~(surf | want_surface))    # the interior is identical
    pts = np.argwhere(full) + lo
    assert np.array_equal(pts.min(0), lo + 1) and np.array_equal(pts.max(0), lo + np.array(dims) - 2)   # bbox
    interior = full.sum() - surf.sum()
    assert 0.68 < interior / full.sum() < 0.71            # 129,815 points, 90,056 of them strictly inside
