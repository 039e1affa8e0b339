"""CPU oracle against the known-answer values of SURVEY.md section 4 (float64 numpy values the
surveyor derived from the reference's formulas on the shipped scenes) and against an independent
float64 brute-force evaluation written here from the formulas of SURVEY.md 2.3."""
import numpy as np
import pytest

from oracle.oracle import Gen1Oracle, Gen2Oracle, kernel_constants
from ti_sph_b200 import scene as sc
from util import small_scene


@pytest.fixture(scope="module")
def demo3d_trace():
    out = {}
    for mode in ("reference", "summed"):
        o = Gen2Oracle(sc.DEMO_3D, density_mode=mode)
        t = o.step(trace=True)
        inv = np.empty(o.n, np.int64)
        inv[t["orig"]] = np.arange(o.n)
        out[mode] = (o, t, inv)
    return out


def test_demo3d_sizes(demo3d_trace):
    o, t, _ = demo3d_trace["reference"]
    assert o.n == 195300 and list(o.grid_num) == [125, 75, 50] and o.ncell == 468750
    nc = t["neighbor_count"]
    assert (nc.min(), nc.max(), int(nc.sum())) == (50, 255, 45273868)
    assert abs(nc.mean() - 231.8) < 0.05
    assert np.count_nonzero(t["counts"]) == 3726 and t["counts"].max() == 64


KAT = [  # original index, n_nbr, S_i, rho/p summed, d_velocity (reference mode)
    (0, 51, 1703.690, 1958.338, 5473.106, (19.0562, 9.2462, 19.0562)),
    (99060, 253, 6145.846, 6400.494, 2.20021e7, (-1.0e-4, -9.80986, 1.0e-4)),
    (1410, 150, 4065.557, 4320.205, 1.404380e6, (41.2480, -9.80990, 7.0e-5)),
]


@pytest.mark.parametrize("idx,nn,S,rho_s,p_s,dv", KAT)
def test_demo3d_kat(demo3d_trace, idx, nn, S, rho_s, p_s, dv):
    o, t, inv = demo3d_trace["reference"]
    s = inv[idx]
    assert t["neighbor_count"][s] == nn
    assert t["S"][s] == pytest.approx(S, rel=2e-6)
    assert t["density_pre"][s] == pytest.approx(254.6479, rel=1e-6)
    assert t["density"][s] == 1000.0 and t["pressure"][s] == 0.0
    assert np.allclose(t["d_velocity"][s], dv, rtol=1e-5, atol=2e-5)
    o2, t2, inv2 = demo3d_trace["summed"]
    s2 = inv2[idx]
    assert t2["density"][s2] == pytest.approx(rho_s, rel=2e-6)
    assert t2["pressure"][s2] == pytest.approx(p_s, rel=2e-5)


def test_gen1_kat():
    scene = {"fluidBlocks": [{"start": [3, 1], "end": [6, 6], "velocity": [0, -20], "density": 1000.0}]}
    g = Gen1Oracle((512, 512), scene)
    assert g.n == 6000
    t = g.step(trace=True)
    assert t["neighbor_count"][0] == 14 and t["neighbor_count"][3050] == 46
    assert t["density_pre"][0] == pytest.approx(1072.804, rel=2e-6)
    assert t["pressure"][0] == pytest.approx(31.7736, rel=2e-5)
    assert t["density_pre"][3050] == pytest.approx(2836.081, rel=2e-6)
    assert t["pressure"][3050] == pytest.approx(73740.48, rel=2e-5)


def _brute_force_f64(x, v, mass, h, c_s, g):
    """float64 O(N^2) evaluation of S_i and the reference-mode acceleration (p = 0)."""
    from scipy.spatial import cKDTree
    kw, kdw = kernel_constants(3, h)
    n = len(x)
    x = x.astype(np.float64); v = v.astype(np.float64)
    tree = cKDTree(x)
    pairs = tree.query_pairs(h * 1.0000001, output_type="ndarray")
    i = np.concatenate([pairs[:, 0], pairs[:, 1]]); j = np.concatenate([pairs[:, 1], pairs[:, 0]])
    r = x[i] - x[j]
    rn = np.linalg.norm(r, axis=1)
    keep = rn < np.float32(h)
    i, j, r, rn = i[keep], j[keep], r[keep], rn[keep]
    q = rn / h
    W = np.where(q <= 0.5, kw * (6 * (q ** 3 - q ** 2) + 1), kw * 2 * (1 - q) ** 3)
    dW = np.where(q <= 0.5, kdw * q * (3 * q - 2), -kdw * (1 - q) ** 2)
    gW = (dW / (rn * h))[:, None] * r
    S = np.bincount(i, weights=mass[i] * W, minlength=n)
    rho = mass * kw
    nu = 2 * 0.05 * h * c_s / (rho[i] + rho[j])
    vx = np.einsum("ij,ij->i", v[i] - v[j], r)
    pi = -nu * np.minimum(0, vx) / (rn ** 2 + 0.01 * h ** 2)
    term = (0.01 / mass[i] * mass[j] * W)[:, None] * r + (mass[j] * pi)[:, None] * gW
    a = np.tile(np.array(g, np.float64), (n, 1))
    for k in range(3):
        a[:, k] -= np.bincount(i, weights=term[:, k], minlength=n)
    return S, a


def test_oracle_vs_float64_bruteforce():
    scene = small_scene(end=(0.5, 0.3, 0.9))
    o = Gen2Oracle(scene)
    rng = np.random.default_rng(7)
    x = (o.x + rng.uniform(-0.002, 0.002, o.x.shape)).astype(np.float32)
    v = (o.v + rng.normal(0, 0.5, o.v.shape)).astype(np.float32)
    o.set_state(x, v, o.density, o.material)
    t = o.step(trace=True)
    S64, a64 = _brute_force_f64(t["x_sorted"], t["v_sorted"], o.mass.astype(np.float64), 0.04, 88.5,
                                (0.0, -9.81, 0.0))
    assert np.max(np.abs(t["S"] - S64) / np.maximum(S64, 1.0)) < 2e-6
    err = np.linalg.norm(t["a_nonpressure"] - a64, axis=1) / np.maximum(np.linalg.norm(a64, axis=1), 9.81)
    assert err.max() < 1e-5


def test_sort_is_stable_and_a_permutation():
    scene = small_scene(end=(0.5, 0.3, 0.9))
    o = Gen2Oracle(scene)
    rng = np.random.default_rng(3)
    o.set_state(o.x[rng.permutation(o.n)], o.v, o.density, o.material)
    x0 = o.x.copy()
    o.update()
    assert sorted(o.orig.tolist()) == list(range(o.n))
    assert np.array_equal(o.x, x0[o.orig])
    assert np.all(np.diff(o.keys) >= 0)
    same = o.keys[1:] == o.keys[:-1]
    assert np.all(o.orig[1:][same] > o.orig[:-1][same])          # ascending original index in a cell
    assert o.scan[-1] == o.n and np.array_equal(np.diff(np.concatenate([[0], o.scan])), o.counts)


def test_wall_reflection_corner():
    """a particle pushed past two walls is clamped to both and reflected about the diagonal"""
    scene = small_scene(start=(4.90, 0.05, 0.7), end=(4.96, 0.11, 0.76), velocity=(60.0, -60.0, 0.0))
    o = Gen2Oracle(scene)
    t = o.step(trace=True)
    assert np.all(t["x"][:, 0] <= np.float32(5.0 - 0.04)) and np.all(t["x"][:, 1] >= np.float32(0.04))
    hit = (t["x_advected"][:, 0] > np.float32(4.96)) & (t["x_advected"][:, 1] <= np.float32(0.04))
    assert hit.any()
    va, vb = t["v_advected"][hit].astype(np.float64), t["v"][hit].astype(np.float64)
    nrm = np.array([1.0, -1.0, 0.0]) / np.sqrt(2.0)
    assert np.allclose(vb, va - 1.5 * (va @ nrm)[:, None] * nrm, rtol=1e-5, atol=1e-4)
