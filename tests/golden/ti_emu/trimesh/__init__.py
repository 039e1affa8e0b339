"""Stand-in for `trimesh` (TEST INFRASTRUCTURE ONLY, used by tests/golden/make_golden.py).

The reference samples rigid bodies with trimesh.load(...).voxelized(pitch).fill().points
(core/partice_system/partice_systemv4.py:259-277).  trimesh is not installable in the build
image, and the sampler is an *input* of the hot path, not part of it, so this stub hands back a
preset point set: `geometryFile` names a .npy file of voxel-centre points that already are in
mesh coordinates; scale / rotation / translation are applied to those points the way the
reference applies them to the mesh vertices.
"""
import numpy as np


class _Voxels:
    def __init__(self, points):
        self.points = points

    def fill(self):
        return self


class _Mesh:
    def __init__(self, vertices):
        self.vertices = np.array(vertices, dtype=np.float64)

    def apply_scale(self, s):
        self.vertices = self.vertices * s

    def apply_transform(self, m):
        v = np.c_[self.vertices, np.ones(len(self.vertices))] @ np.asarray(m).T
        self.vertices = v[:, :3]

    def copy(self):
        return _Mesh(self.vertices.copy())

    def voxelized(self, pitch):
        return _Voxels(self.vertices.copy())


def load(path):
    return _Mesh(np.load(path))


class transformations:
    @staticmethod
    def rotation_matrix(angle, direction, point=None):
        d = np.asarray(direction, dtype=np.float64)
        d = d / np.linalg.norm(d)
        c, s = np.cos(angle), np.sin(angle)
        K = np.array([[0, -d[2], d[1]], [d[2], 0, -d[0]], [-d[1], d[0], 0]])
        R = c * np.eye(3) + s * K + (1 - c) * np.outer(d, d)
        M = np.eye(4)
        M[:3, :3] = R
        if point is not None:
            p = np.asarray(point, dtype=np.float64)
            M[:3, 3] = p - R @ p
        return M
