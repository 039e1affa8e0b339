"""Shared helpers of the parity tests: scenes, the oracle/engine pair, tolerances."""
import copy

import numpy as np

from oracle.oracle import Gen2Oracle
from ti_sph_b200 import _capi, scene as sc
from ti_sph_b200.engine import Engine

RTOL = 1e-5    # BASELINE.json north_star: density, pressure, acceleration within 1e-5 relative (fp32)


def small_scene(start=(0.3, 0.1, 0.7), end=(0.6, 0.4, 1.0), velocity=(0.0, -1.0, 10.0), radius=0.01,
                domain_end=(5.0, 3.0, 2.0)):
    s = copy.deepcopy(sc.DEMO_3D)
    s["configuration"]["particleRadius"] = radius
    s["configuration"]["domainEnd"] = list(domain_end)
    blk = s["fluidBlocks"][0]
    blk["start"], blk["end"], blk["velocity"] = list(start), list(end), list(velocity)
    return s


def jitter(x, radius, seed=1234, amp=0.2):
    """lattice + uniform jitter in [-amp r, amp r] (SURVEY 8(d) parity state ii)."""
    rng = np.random.default_rng(seed)
    return (x + rng.uniform(-amp * radius, amp * radius, size=x.shape)).astype(np.float32)


def make_pair(scene, density_mode="reference", volume_mode="reference", x=None, v=None,
              boundary_points=None):
    """Oracle and CUDA engine holding the same initial state."""
    ora = Gen2Oracle(scene, density_mode=density_mode, volume_mode=volume_mode,
                     boundary_points=boundary_points)
    if x is not None:
        ora.set_state(x, ora.v if v is None else v, ora.density, ora.material)
    cfg = sc.gen2_config(scene["configuration"], ora.n,
                         density_mode={"reference": 0, "summed": 1}[density_mode],
                         volume_mode={"reference": 0, "akinci": 1}[volume_mode])
    eng = Engine(cfg)
    eng.add_particles(ora.x, ora.v, ora.density, ora.pressure, ora.material, ora.color)
    return ora, eng


def rel_err(a, b, floor=0.0):
    """max |a-b| / max(|b|, floor), elementwise for scalars"""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0


def vec_rel_err(a, b, floor):
    """max over particles of |a-b| / max(|b|, floor) with vector norms (hard part 2)"""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    d = np.linalg.norm(a - b, axis=1)
    s = np.maximum(np.linalg.norm(b, axis=1), floor)
    return float(np.max(d / s)) if len(a) else 0.0


def accel_err(a, b, mag, floor=None):
    """max over particles of (|a - b| - floor) / mag: accelerations are sums of ~250 pair terms that
    largely cancel, and an f32 sum carries a rounding error proportional to the sum of the
    MAGNITUDES of its terms (`mag`, exported by the oracle: Gen2Oracle.force_magnitudes), not to what
    is left of them.  `floor` is an absolute allowance per particle (the pressure sum inherits the
    cancellation floor of p = B (x^7 - 1)).  Particles without terms (mag = 0: boundary particles,
    zero pressure) must agree exactly."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    d = np.linalg.norm(a - b, axis=1)
    if floor is not None:
        d = np.maximum(d - np.asarray(floor, np.float64), 0.0)
    mag = np.asarray(mag, np.float64)
    if np.any(d[mag == 0] != 0):
        return np.inf
    ok = mag > 0
    return float(np.max(d[ok] / mag[ok])) if ok.any() else 0.0


def check_force_stage(eng, t, dt=2e-4, rtol=RTOL, p_rtol=RTOL, split=True):
    """the force + advect + walls stage of a traced oracle step `t` against the engine (after
    STAGE_FORCE_ADVECT with diagnostics on).  Every sum is held to `rtol` (1e-5) relative to the sum of
    the magnitudes of its terms.  The pressure terms are linear in the pressures the density stage
    produced, which the tests accept within |dp| <= p_rtol |p| + floor (p = B (x^7 - 1)): the pressure
    sum inherits exactly that, p_rtol x its magnitude sum plus the floor carried through the sum
    (t["mag_pressure_floor"]).  split=False: the step ran without TISPH_P_DIAGNOSTICS -- the force walk then keeps
    one accumulator for both sums, and only the reference's own fields (d_velocity, v, x) exist."""
    from ti_sph_b200 import _capi as K
    mnp, mp = t["mag_nonpressure"].astype(np.float64), t["mag_pressure"].astype(np.float64)
    pf = t.get("mag_pressure_floor")
    pf = (np.zeros_like(mp) if pf is None else pf.astype(np.float64)) + p_rtol * mp
    # d_velocity after compute_non_pressure_force and after compute_pressure_force are the reference's fields
    # (wcsphv2.py:93, :53).  The pressure sum on its own is a diagnostic of this library: it is held to the
    # scale of the field it is added to (a lone neighbour at the very edge of the support has a pressure
    # term of 1e-7 m/s^2 whose own relative accuracy means nothing).
    worst = {"d_velocity": accel_err(eng.download(K.F_D_VELOCITY), t["d_velocity"], mnp + mp, pf)}
    if split:
        worst["a_nonpressure"] = accel_err(eng.download(K.F_A_NONPRESSURE), t["a_nonpressure"], mnp)
        worst["a_pressure"] = accel_err(eng.download(K.F_A_PRESSURE), t["a_pressure"], mnp + mp, pf)
    fl = t["material"] == 1
    mag, pf = (mnp + mp)[fl], pf[fl]
    # v' = v + dt a ; x' = x + dt v' (then the wall clamp): both inherit dt (dt^2) times the acceleration tolerance
    dv = np.linalg.norm(eng.download(K.F_V).astype(np.float64)[fl] - t["v"][fl], axis=1)
    vn = np.linalg.norm(t["v"][fl].astype(np.float64), axis=1)
    worst["v"] = float(np.max(np.maximum(dv - dt * pf, 0.0) / (vn + dt * mag))) if fl.any() else 0.0
    dx = np.linalg.norm(eng.download(K.F_X).astype(np.float64)[fl] - t["x"][fl], axis=1)
    xn = np.linalg.norm(t["x"][fl].astype(np.float64), axis=1)
    worst["x"] = float(np.max(np.maximum(dx - dt * dt * pf, 0.0) / (xn + dt * vn + dt * dt * mag))) if fl.any() else 0.0
    bad = {k: v for k, v in worst.items() if not v < rtol}
    assert not bad, f"beyond {rtol:g} relative: {bad} (all: {worst})"
    return worst


def golden_gen2_force_reference(case, z, s=0):
    """the force-stage arrays of golden step `s` (tests/golden/gen2_*.npz: what the reference's own
    sources computed) in the layout check_force_stage takes; the magnitude sums that scale the
    tolerance are evaluated by the oracle on the golden state"""
    from oracle.oracle import Gen2Oracle
    g = lambda k: z[f"s{s}.{k}"]
    ora = Gen2Oracle(case["scene"], boundary_points=z["init.x"][z["init.material"] == 0])
    assert ora.n == int(z["n"])
    ora.set_state(g("sorted.x"), g("sorted.v"), g("pressure.density"), g("sorted.material"),
                  pressure=g("pressure.pressure"), volume=g("volume.volume"), mass=g("sorted.mass"))
    ora.scan = np.ascontiguousarray(g("sorted.grid_particles_num"), np.int32)
    mnp, mp = ora.force_magnitudes(g("density.density"))
    p_floor = 50 * 8 * np.finfo(np.float32).eps * (g("pressure.density").astype(np.float64) / 1000.0) ** 7
    mpf = ora.force_magnitudes(g("density.density"), pressure=p_floor)[1]
    a_np_ref, a_ref = g("nonpressure.d_velocity"), g("pressure.d_velocity")
    return {"mag_nonpressure": mnp, "mag_pressure": mp, "mag_pressure_floor": mpf, "a_nonpressure": a_np_ref,
            "a_pressure": a_ref - a_np_ref, "d_velocity": a_ref, "material": g("sorted.material"),
            "v": g("end.v"), "x": g("end.x")}


def check_end_state(x, v, t, dt=2e-4, rtol=RTOL):
    """x, v after a whole step against a traced oracle step `t` (rows in the oracle's order): the same
    bounds check_force_stage puts on v' and x'"""
    mnp, mp = t["mag_nonpressure"].astype(np.float64), t["mag_pressure"].astype(np.float64)
    pf = t["mag_pressure_floor"].astype(np.float64) + rtol * mp
    fl = t["material"] == 1
    mag, pf = (mnp + mp)[fl], pf[fl]
    dv = np.linalg.norm(np.asarray(v, np.float64)[fl] - t["v"][fl], axis=1)
    vn = np.linalg.norm(t["v"][fl].astype(np.float64), axis=1)
    dx = np.linalg.norm(np.asarray(x, np.float64)[fl] - t["x"][fl], axis=1)
    xn = np.linalg.norm(t["x"][fl].astype(np.float64), axis=1)
    worst = {"v": float(np.max(np.maximum(dv - dt * pf, 0.0) / (vn + dt * mag))),
             "x": float(np.max(np.maximum(dx - dt * dt * pf, 0.0) / (xn + dt * vn + dt * dt * mag)))}
    bad = {k: w for k, w in worst.items() if not w < rtol}
    assert not bad, f"beyond {rtol:g} relative: {bad}"
    assert np.array_equal(np.asarray(x)[~fl], t["x"][~fl]) and np.array_equal(np.asarray(v)[~fl], t["v"][~fl])
    return worst
