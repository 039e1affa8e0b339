"""Drop-in for the reference's utils/lines.py: the 8 corners and 12 edges of the domain box
that main_3d.py hands to ggui `scene.lines`.  Returns Taichi fields when Taichi is importable
(what ggui needs), numpy arrays otherwise."""
import numpy as np

_EDGES = [(0, 1), (0, 2), (1, 3), (2, 3), (4, 5), (4, 6), (5, 7), (6, 7),
          (0, 4), (1, 5), (2, 6), (3, 7)]


def getlines(config):
    lo, hi = config['domainStart'], config['domainEnd']
    corners = np.array([[(hi if (i >> 1) & 1 else lo)[0], (hi if i & 1 else lo)[1],
                         (hi if (i >> 2) & 1 else lo)[2]] for i in range(8)], dtype=np.float32)
    indices = np.array(_EDGES, dtype=np.int32).reshape(-1)
    try:
        import taichi as ti
    except ImportError:
        return corners, indices
    points = ti.Vector.field(3, dtype=ti.f32, shape=8)
    points.from_numpy(corners)
    box_lines_indices = ti.field(int, shape=24)
    box_lines_indices.from_numpy(indices)
    return points, box_lines_indices
