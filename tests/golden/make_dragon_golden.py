"""Golden voxel set of the reference's only mesh asset, data/models/Dragon_50k.obj, in the C4
placement (bench.workload_scene('C4')): float64 triangle/box separating-axis test per voxel of
every triangle's bounding box, interior by scipy.ndimage.binary_fill_holes (the routine behind
trimesh's VoxelGrid.fill(), partice_systemv4.py:276-277).  Independent of the CUDA voxeliser.

    python tests/golden/make_dragon_golden.py        # writes tests/golden/dragon_c4_voxels.npz

The file holds the occupancy bit-packed (np.packbits), the lattice origin and dims, the pitch, and
the mesh statistics the GPU test pins.
"""
import os
import sys

import numpy as np
from scipy import ndimage

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from bench import workload_scene          # noqa: E402
from ti_sph_b200 import mesh              # noqa: E402


def surface_voxels(v, f, pitch, lo, dims):
    """occupancy [dims] of the voxels (cubes of edge `pitch` centred on k * pitch) a triangle touches"""
    hw = pitch / 2
    occ = np.zeros(dims, bool)
    eye = np.eye(3)
    for tri in f:
        t = v[tri]
        klo = np.floor((t.min(0) - hw) / pitch + 0.5).astype(int)
        khi = np.floor((t.max(0) + hw) / pitch + 0.5).astype(int)
        g = np.stack(np.meshgrid(*[np.arange(klo[k], khi[k] + 1) for k in range(3)], indexing="ij"), -1).reshape(-1, 3)
        p = t[None, :, :] - (g * pitch)[:, None, :]
        e = np.stack([p[:, 1] - p[:, 0], p[:, 2] - p[:, 1], p[:, 0] - p[:, 2]], 1)
        ok = np.all((p.min(1) <= hw) & (p.max(1) >= -hw), axis=1)
        axes = [np.cross(e[:, 0], e[:, 1])]
        for k in range(3):
            for u in eye:
                axes.append(np.cross(np.broadcast_to(u, e[:, k].shape), e[:, k]))
        for a in axes:
            proj = np.einsum("mij,mj->mi", p, a)
            r = hw * np.abs(a).sum(1)
            ok &= ~((proj.min(1) > r) | (proj.max(1) < -r))
        idx = g[ok] - lo
        occ[idx[:, 0], idx[:, 1], idx[:, 2]] = True
    return occ


def main():
    scene = workload_scene("C4")
    body = scene["rigidBodies"][0]
    assert body["geometryFile"].endswith("Dragon_50k.obj"), "data/models/Dragon_50k.obj is missing"
    pitch = 2 * scene["configuration"]["particleRadius"]
    v, f = mesh.load_obj(body["geometryFile"])
    edges = np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), axis=1)
    _, cnt = np.unique(edges, axis=0, return_counts=True)
    vt = mesh.transform_vertices(v, body)
    vt32 = vt.astype(np.float32).astype(np.float64)          # the sampler receives f32 vertices
    lo = np.floor(vt32.min(0) / pitch + 0.5).astype(int) - 1
    hi = np.floor(vt32.max(0) / pitch + 0.5).astype(int) + 1
    dims = tuple(int(d) for d in hi - lo + 1)
    surf = surface_voxels(vt32, f, pitch, lo, dims)
    full = ndimage.binary_fill_holes(surf)                   # 6-connectivity, like trimesh's fill()
    out = os.path.join(ROOT, "tests", "golden", "dragon_c4_voxels.npz")
    np.savez_compressed(out, surface=np.packbits(surf), filled=np.packbits(full), lo=lo, dims=np.array(dims),
                        pitch=pitch, n_vertices=len(v), n_faces=len(f),
                        edge_face_histogram=np.bincount(cnt), translation=np.array(body["translation"]))
    print(f"{out}: {len(v)} vertices, {len(f)} faces, edge/face histogram {np.bincount(cnt).tolist()}, "
          f"grid {dims}, surface {int(surf.sum())}, filled {int(full.sum())}")


if __name__ == "__main__":
    main()
