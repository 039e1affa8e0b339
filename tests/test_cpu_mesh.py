"""Host side of the mesh sampler: OBJ reader, rigid-body transform (partice_systemv4.py:259-273)."""
import numpy as np

from ti_sph_b200 import mesh


def test_obj_round_trip_and_index_forms(tmp_path):
    v, f = mesh.icosphere(0.5, (1, 2, 3), subdivisions=2)
    p = tmp_path / "s.obj"
    mesh.write_obj(p, v, f)
    v2, f2 = mesh.load_obj(p)
    assert np.allclose(v, v2, atol=1e-8) and np.array_equal(f, f2)
    q = tmp_path / "q.obj"
    q.write_text("# quad with texture/normal indices and a relative face\n"
                 "v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvt 0 0\nvn 0 0 1\nf 1/1/1 2/1/1 3/1/1 4/1/1\nf -4 -3 -2\n")
    v3, f3 = mesh.load_obj(q)
    assert len(v3) == 4 and f3.tolist() == [[0, 1, 2], [0, 2, 3], [0, 1, 2]]


def test_transform_matches_the_reference_order_of_operations():
    v, _ = mesh.icosphere(1.0, (0, 0, 0), 1)
    v = v + np.array([2.0, 0.0, 0.0])
    body = {"scale": [2, 2, 2], "rotationAngle": 90, "rotationAxis": [0, 0, 1], "translation": [0.0, 1.0, 0.0]}
    w = mesh.transform_vertices(v, body)
    # scale first (about the origin), rotate about the scaled vertex mean, then translate
    c = (v * 2).mean(axis=0)
    rel = v * 2 - c
    expect = c + np.stack([-rel[:, 1], rel[:, 0], rel[:, 2]], -1) + np.array([0.0, 1.0, 0.0])
    assert np.allclose(w, expect, atol=1e-12)
