"""Rigid bodies -> boundary particles: the host side of ParticleSystemV4.load_rigid_body
(partice_systemv4.py:259-277) without trimesh.

    mesh = trim.load(geometryFile); mesh.apply_scale(scale)
    rotate by rotationAngle (degrees) about rotationAxis through the vertex mean; translate
    points = mesh.voxelized(pitch=particle_diameter).fill().points   (f32)

The OBJ reader handles what the reference ships (data/models/Dragon_50k.obj: plain `v x y z` /
`f a b c`, Meshlab export) plus `a/b/c` index forms and polygons (fan triangulation).  The
voxelisation runs on the GPU (tisph_voxelize_mesh, csrc/tisph_voxel.cuh); its conventions are
stated there.  `geometryFile` may also name a .npy file holding voxel-centre points directly
(sampler bypass, used by parity tests that share a point set with the oracle).
"""
import ctypes as C

import numpy as np

from . import _capi


def load_obj(path):
    """(vertices f64 [nv,3], faces i32 [nf,3]) of a Wavefront OBJ file."""
    verts, faces = [], []
    with open(path, "r") as fh:
        for line in fh:
            if line.startswith("v "):
                p = line.split()
                verts.append((float(p[1]), float(p[2]), float(p[3])))
            elif line.startswith("f "):
                idx = [int(tok.split("/")[0]) for tok in line.split()[1:]]
                idx = [i - 1 if i > 0 else len(verts) + i for i in idx]      # OBJ is 1-based; negatives are relative
                for k in range(1, len(idx) - 1):
                    faces.append((idx[0], idx[k], idx[k + 1]))
    if not verts or not faces:
        raise ValueError(f"{path}: no vertices / faces found")
    return np.array(verts, np.float64), np.array(faces, np.int32)


def rotation_matrix(angle, axis, point):
    """4x4 rotation by `angle` (radians) about `axis` through `point` (trimesh.transformations)."""
    d = np.asarray(axis, np.float64)
    d = d / np.linalg.norm(d)
    c, s = np.cos(angle), np.sin(angle)
    K = np.array([[0, -d[2], d[1]], [d[2], 0, -d[0]], [-d[1], d[0], 0]])
    R = c * np.eye(3) + s * K + (1 - c) * np.outer(d, d)
    M = np.eye(4)
    M[:3, :3] = R
    p = np.asarray(point, np.float64)
    M[:3, 3] = p - R @ p
    return M


def transform_vertices(vertices, rigid_body):
    """scale, rotate about the vertex mean, translate (partice_systemv4.py:266-273)"""
    v = np.asarray(vertices, np.float64) * np.asarray(rigid_body.get("scale", 1.0), np.float64)
    angle = rigid_body.get("rotationAngle", 0) * np.pi / 180
    M = rotation_matrix(angle, rigid_body.get("rotationAxis", [0, 1, 0]), v.mean(axis=0))
    v = v @ M[:3, :3].T + M[:3, 3]
    return v + np.asarray(rigid_body.get("translation", [0, 0, 0]), np.float64)


def voxelize(vertices, faces, pitch, fill=True, device=0):
    """Voxel-centre points (f32 [n,3], x-major order) of a triangle mesh: surface voxels plus,
    with fill=True, the enclosed interior."""
    lib = _capi.load()
    v = np.ascontiguousarray(vertices, np.float32)
    f = np.ascontiguousarray(faces, np.int32)
    lo = np.floor(v.min(axis=0).astype(np.float64) / pitch + 0.5).astype(np.int32) - 1       # one empty layer
    hi = np.floor(v.max(axis=0).astype(np.float64) / pitch + 0.5).astype(np.int32) + 1
    dims = (hi - lo + 1).astype(np.int32)
    occ = np.zeros(tuple(int(d) for d in dims), np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    _capi.check(lib.tisph_voxelize_mesh(int(device), p(v), len(v), p(f), len(f), float(pitch), int(bool(fill)),
                                        p(lo), p(dims), p(occ)))
    idx = np.argwhere(occ != 0)
    return ((idx + lo[None, :]).astype(np.float64) * pitch).astype(np.float32)


def sample_rigid_body(rigid_body, pitch, device=0):
    path = rigid_body["geometryFile"]
    if str(path).endswith(".npy"):
        return np.ascontiguousarray(np.load(path), np.float32)
    vertices, faces = load_obj(path)
    vertices = transform_vertices(vertices, rigid_body)
    rigid_body["mesh"] = {"vertices": vertices, "faces": faces}              # the reference keeps mesh.copy()
    return voxelize(vertices, faces, pitch, fill=True, device=device)


# ---- procedural meshes (tests, synthetic benchmark scenes) --------------------------------------
def icosphere(radius=1.0, center=(0, 0, 0), subdivisions=3):
    t = (1.0 + 5 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2),
         (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11),
         (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    v = [np.array(p, np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(subdivisions):
        cache, nf = {}, []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = v[a] + v[b]
                v.append(m / np.linalg.norm(m))
                cache[key] = len(v) - 1
            return cache[key]
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    return np.array(v) * radius + np.asarray(center, np.float64), np.array(f, np.int32)


def torus(R=1.0, r=0.3, center=(0, 0, 0), nu=64, nv=32):
    u = np.linspace(0, 2 * np.pi, nu, endpoint=False)
    w = np.linspace(0, 2 * np.pi, nv, endpoint=False)
    uu, ww = np.meshgrid(u, w, indexing="ij")
    v = np.stack([(R + r * np.cos(ww)) * np.cos(uu), r * np.sin(ww), (R + r * np.cos(ww)) * np.sin(uu)], -1).reshape(-1, 3)
    f = []
    for i in range(nu):
        for j in range(nv):
            a, b = i * nv + j, i * nv + (j + 1) % nv
            c, d = ((i + 1) % nu) * nv + j, ((i + 1) % nu) * nv + (j + 1) % nv
            f += [(a, c, b), (b, c, d)]
    return v + np.asarray(center, np.float64), np.array(f, np.int32)


def write_obj(path, vertices, faces):
    with open(path, "w") as fh:
        for p in vertices:
            fh.write(f"v {p[0]:.9g} {p[1]:.9g} {p[2]:.9g}\n")
        for t in faces:
            fh.write(f"f {t[0] + 1} {t[1] + 1} {t[2] + 1}\n")
