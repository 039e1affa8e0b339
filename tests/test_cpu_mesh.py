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


def test_dragon_asset_is_closed_but_not_two_manifold():
    """data/models/Dragon_50k.obj, the reference's only mesh (partice_systemv4.py:259-277): every edge carries two
    faces except 16 that carry four, none is a boundary edge -- not "watertight" in trimesh's sense, but without
    holes, which is all the outside-flood fill of the sampler needs."""
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    v, f = mesh.load_obj(os.path.join(root, "data", "models", "Dragon_50k.obj"))
    assert v.shape == (25007, 3) and f.shape == (50000, 3) and len(np.unique(f)) == len(v)
    e = np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), axis=1)
    _, cnt = np.unique(e, axis=0, return_counts=True)
    assert np.bincount(cnt).tolist() == [0, 0, 74968, 0, 16]
    gold = np.load(os.path.join(root, "tests", "golden", "dragon_c4_voxels.npz"))
    assert gold["edge_face_histogram"].tolist() == [0, 0, 74968, 0, 16]
    surface = np.unpackbits(gold["surface"])[:int(np.prod(gold["dims"]))]
    filled = np.unpackbits(gold["filled"])[:int(np.prod(gold["dims"]))]
    assert (int(surface.sum()), int(filled.sum())) == (39759, 129815) and not np.any(surface & ~filled)
