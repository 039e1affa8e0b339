"""ctypes/numpy front end of the CPU oracle (oracle/sph_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of sph_oracle.c.  Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never by
the product path (ti_sph_b200/, core/, utils/).

The classes restate the *host side* of the reference constructors so that the oracle
is independent of the product's host code:
  Gen2Oracle  <- core/partice_system/partice_systemv4.py:8-78,148-168,347-373
                 + core/sph/sph_basev2.py:10-16 + core/sph/wcsphv2.py:8-16
  Gen1Oracle  <- core/partice_system/partice_system.py:8-34,134-164
                 (+ partice_systemv2.py:124-136) + core/sph/sph_base.py:10-16
"""
import ctypes as C
import os
import subprocess
from functools import reduce

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OraConfig(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("grid_num", C.c_int32 * 3), ("h", C.c_float),
        ("domain_size", C.c_float * 3), ("padding", C.c_float), ("wall_hi", C.c_float * 3),
        ("dt", C.c_float), ("g", C.c_float * 3), ("c_s", C.c_float), ("rho0", C.c_float),
        ("ps_density0", C.c_float), ("stiffness", C.c_float), ("exponent", C.c_float),
        ("k_w", C.c_float), ("k_dw", C.c_float), ("visc_fluid_c", C.c_float),
        ("visc_bound_c", C.c_float), ("eps_h2", C.c_float), ("m_V", C.c_float),
        ("g1_visc_c", C.c_float), ("g1_mass", C.c_float), ("g1_press_c", C.c_float),
        ("density_mode", C.c_int32), ("volume_mode", C.c_int32),
        ("max_per_cell", C.c_int32), ("max_neighbors", C.c_int32),
    ]


def build(force=False):
    """Compile oracle/_build/liboracle.so with the committed Makefile."""
    so = os.path.join(_HERE, "_build", "liboracle.so")
    src = os.path.join(_HERE, "sph_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        assert _LIB.ora_config_size() == C.sizeof(OraConfig)
        _LIB.ora_num_threads.restype = C.c_int
    return _LIB


def set_pow_mode(mode):
    """0: q**3 by repeated multiplication (Taichi's lowering; default).  1: libm powf, the way
    numpy evaluates `**` under the Taichi stand-in that produced tests/golden/ (bit-comparable)."""
    lib().ora_set_pow_mode(int(mode))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def kernel_constants(dim, h):
    """(k/h^dim, 6k/h^dim) evaluated in float64 like the Python-scope constants of
    sph_basev2.py:22-30,42-50, then rounded to f32 by the caller."""
    k = {1: 4 / 3, 2: 40 / (7 * np.pi), 3: 8 / np.pi}[dim]
    k6 = {1: 4 / 3, 2: 40 / 7 / np.pi, 3: 8 / np.pi}[dim]
    return k / h ** dim, 6. * k6 / h ** dim


def cube_positions(lower_corner, cube_size, spacing, dim):
    """partice_systemv4.py:356-366 / partice_system.py:143-156 (np.arange per axis,
    meshgrid ij, cast f32, x slowest)."""
    num_dim = [np.arange(lower_corner[i], lower_corner[i] + cube_size[i], spacing)
               for i in range(dim)]
    n = reduce(lambda x, y: x * y, [len(a) for a in num_dim])
    pos = np.array(np.meshgrid(*num_dim, indexing='ij'), dtype=np.float32)
    return np.ascontiguousarray(pos.reshape(dim, n).T)


class Gen2Oracle:
    """ParticleSystemV4 + WCSPHV2 on the CPU."""

    def __init__(self, scene, density_mode="reference", volume_mode="reference",
                 boundary_points=None, threads=None, boundary_color=(255, 255, 255)):
        cfg = scene["configuration"]
        self.dim = cfg["dim"]
        assert self.dim == 3
        domain_start = np.array(cfg["domainStart"])
        domain_end = np.array(cfg["domainEnd"])
        self.domain_size = domain_end - domain_start                   # :22
        self.particle_radius = cfg["particleRadius"]
        self.support_length = 4.0 * self.particle_radius               # :34
        self.padding = self.support_length                             # :35
        self.particle_diameter = 2 * self.particle_radius              # :47
        self.m_V0 = 0.8 * self.particle_diameter ** self.dim           # :48
        self.grid_num = np.ceil(self.domain_size / self.support_length).astype(np.int32)  # :59
        self.ncell = int(np.prod(self.grid_num.astype(np.int64)))
        h = self.support_length
        kw, kdw = kernel_constants(3, h)
        c = OraConfig()
        c.dim = 3
        c.grid_num[:] = [int(g) for g in self.grid_num]
        c.h = h
        c.domain_size[:] = [float(s) for s in self.domain_size]
        c.padding = self.padding
        c.wall_hi[:] = [float(s - self.padding) for s in self.domain_size]
        c.dt = 2e-4                                                    # sph_basev2.py:15
        c.g[:] = [float(g) for g in cfg["gravitation"]]                 # sph_basev2.py:16
        c.c_s = cfg["c_s"]                                             # wcsphv2.py:16
        c.rho0 = 1000.0                                                # sph_basev2.py:13
        c.ps_density0 = cfg["density0"]
        c.stiffness = 50.0
        c.exponent = 7.0
        c.k_w, c.k_dw = kw, kdw
        c.visc_fluid_c = 2 * 0.05 * h * cfg["c_s"]                     # wcsphv2.py:69
        c.visc_bound_c = 0.08 * h * cfg["c_s"]                         # wcsphv2.py:76
        c.eps_h2 = 0.01 * h ** 2
        c.density_mode = {"reference": 0, "summed": 1}[density_mode]
        c.volume_mode = {"reference": 0, "akinci": 1}[volume_mode]
        self.cfg = c
        if threads:
            lib().ora_set_num_threads(int(threads))
        # initial state: rigid (boundary) particles first, then fluid blocks (:102-146)
        xs, vs, ds, mats = [], [], [], []
        if boundary_points is not None and len(boundary_points):
            bp = _f32(boundary_points)
            xs.append(bp); vs.append(np.zeros_like(bp))
            ds.append(np.full(len(bp), 1000.0, np.float32)); mats.append(np.zeros(len(bp), np.int32))
        for fluid in scene["fluidBlocks"]:
            start, end = fluid["start"], fluid["end"]
            size = [end[i] - start[i] for i in range(3)]
            pos = cube_positions(start, size, self.particle_radius, 3)
            pre = reduce(lambda a, b: a * b, [len(np.arange(start[i], end[i], self.particle_radius))
                                              for i in range(3)])       # :160-168
            assert pre == len(pos), "particle_max_num pre-pass disagrees with add_cube (Q10)"
            xs.append(pos)
            vs.append(np.full(pos.shape, fluid["velocity"], dtype=np.float32))
            dens = fluid["density"]
            ds.append(np.full(len(pos), dens if dens is not None else 1000.0, np.float32))
            mats.append(np.ones(len(pos), np.int32))
        self.set_state(np.concatenate(xs), np.concatenate(vs), np.concatenate(ds),
                       np.concatenate(mats))
        # rigid colours: ints are divided by 255.0 and the f32 result is cast back into the i32
        # colour field (:111-114,190), so [255,255,255] is stored as (1,1,1)
        bc = [c / 255.0 if type(boundary_color[0]) == int else c for c in boundary_color]
        self.color[self.material == 0] = np.array(bc, np.float32).astype(np.int32)

    def set_state(self, x, v, density, material, pressure=None, volume=None, mass=None, color=None):
        n = len(x)
        self.n = n
        self.x = _f32(x).copy(); self.v = _f32(v).copy()
        self.density = _f32(density).copy()
        self.pressure = np.zeros(n, np.float32) if pressure is None else _f32(pressure).copy()
        self.material = np.ascontiguousarray(material, np.int32).copy()
        self.volume = (np.full(n, self.m_V0, np.float32) if volume is None
                       else _f32(volume).copy())                        # :203
        self.mass = (self.volume * self.density if mass is None else _f32(mass).copy())  # :204
        if color is None:
            self.color = np.zeros((n, 3), np.int32)
            self.color[self.material == 1] = 0x111111                  # :144 (scalar -> every lane)
        else:
            self.color = np.ascontiguousarray(color, np.int32).reshape(n, 3).copy()
        self.m = np.zeros(n, np.float32)
        self.orig = np.arange(n, dtype=np.int32)      # bookkeeping only (not a reference field)
        self.keys = np.zeros(n, np.int32)
        self.new_index = np.zeros(n, np.int32)
        self.counts = np.zeros(self.ncell, np.int32)
        self.scan = np.zeros(self.ncell, np.int32)
        self.dvel = np.zeros((n, 3), np.float32)
        self.S = np.zeros(n, np.float32)

    # -- stages ---------------------------------------------------------
    def update(self):
        """ps.update(): keys, histogram, inclusive scan, stable rank, reorder (:251-256)."""
        L = lib(); c = C.byref(self.cfg)
        bad = L.ora_bin_count(c, self.n, _p(self.x), _p(self.keys), _p(self.counts), _p(self.scan))
        if bad:
            raise RuntimeError(f"{bad} particles outside the grid (reference UB)")
        L.ora_sort_rank(c, self.n, _p(self.keys), _p(self.scan), _p(self.new_index))
        for name in ("x", "v", "mass", "volume", "density", "pressure", "material", "color", "m",
                     "keys", "orig"):
            a = getattr(self, name)
            out = np.empty_like(a)
            words = a.shape[1] if a.ndim == 2 else 1
            L.ora_reorder(self.n, words, _p(self.new_index), _p(a), _p(out))
            setattr(self, name, out)
        return self.new_index

    def neighbor_count(self):
        out = np.zeros(self.n, np.int32)
        lib().ora_neighbor_count(C.byref(self.cfg), self.n, _p(self.x), _p(self.scan), _p(out))
        return out

    def boundary_volume(self):
        lib().ora_boundary_volume(C.byref(self.cfg), self.n, _p(self.x), _p(self.material),
                                  _p(self.scan), _p(self.volume))

    def compute_densities(self):
        lib().ora_density(C.byref(self.cfg), self.n, _p(self.x), _p(self.mass), _p(self.material),
                          _p(self.scan), _p(self.density), _p(self.S))

    def compute_non_pressure_force(self):
        lib().ora_non_pressure(C.byref(self.cfg), self.n, _p(self.x), _p(self.v), _p(self.mass),
                               _p(self.volume), _p(self.density), _p(self.material),
                               _p(self.scan), _p(self.dvel))

    def compute_pressure_force(self):
        L = lib(); c = C.byref(self.cfg)
        L.ora_eos(c, self.n, _p(self.density), _p(self.pressure))
        self.a_pressure = np.zeros((self.n, 3), np.float32)
        L.ora_pressure_force(c, self.n, _p(self.x), _p(self.mass), _p(self.volume),
                             _p(self.density), _p(self.pressure), _p(self.material),
                             _p(self.scan), _p(self.dvel), _p(self.a_pressure))

    def force_magnitudes(self, density_pre, pressure=None):
        """(sum |non-pressure term|, sum |pressure term|) per particle: the scale of the rounding error of
        the two acceleration sums (test support; call after compute_pressure_force, before advert).
        `pressure` replaces the pressure field (e.g. by its tolerance floor)."""
        mnp, mp = np.zeros(self.n, np.float32), np.zeros(self.n, np.float32)
        pr = self.pressure if pressure is None else _f32(pressure)
        lib().ora_force_magnitudes(C.byref(self.cfg), self.n, _p(self.x), _p(self.v), _p(self.mass),
                                   _p(self.volume), _p(_f32(density_pre)), _p(self.density), _p(pr),
                                   _p(self.material), _p(self.scan), _p(mnp), _p(mp))
        return mnp, mp

    def advert(self):
        lib().ora_advect(C.byref(self.cfg), self.n, _p(self.x), _p(self.v), _p(self.dvel),
                         _p(self.material))

    def enforce_boundary(self):
        lib().ora_enforce_boundary(C.byref(self.cfg), self.n, _p(self.x), _p(self.v),
                                   _p(self.material))

    def step(self, trace=False):
        """sph_basev2.py:210-214.  With trace=True returns every intermediate."""
        t = {}
        self.update()
        if trace:
            t.update(keys=self.keys.copy(), scan=self.scan.copy(), counts=self.counts.copy(),
                     new_index=self.new_index.copy(), orig=self.orig.copy(),
                     x_sorted=self.x.copy(), v_sorted=self.v.copy(),
                     neighbor_count=self.neighbor_count())
        self.boundary_volume()
        self.compute_densities()
        if trace:
            t.update(volume=self.volume.copy(), S=self.S.copy(), density_pre=self.density.copy())
        self.compute_non_pressure_force()
        if trace:
            t.update(a_nonpressure=self.dvel.copy())
        self.compute_pressure_force()
        if trace:
            t.update(density=self.density.copy(), pressure=self.pressure.copy(),
                     a_pressure=self.a_pressure.copy(), d_velocity=self.dvel.copy())
            t["mag_nonpressure"], t["mag_pressure"] = self.force_magnitudes(t["density_pre"])
            # p = B (x^7 - 1) cancels near x = 1: the tests allow |dp| <= rtol |p| + p_floor; the pressure sum
            # inherits the same sum with p_floor in the place of p
            t["p_floor"] = (50 * 8 * np.finfo(np.float32).eps * (self.density.astype(np.float64) / 1000.0) ** 7).astype(np.float32)
            t["mag_pressure_floor"] = self.force_magnitudes(t["density_pre"], pressure=t["p_floor"])[1]
        self.advert()
        if trace:
            t.update(x_advected=self.x.copy(), v_advected=self.v.copy())
        self.enforce_boundary()
        if trace:
            t.update(x=self.x.copy(), v=self.v.copy(), material=self.material.copy())
        return t

    def step_fast(self, nsteps=1):
        """Whole steps inside C (used for CPU-baseline timing)."""
        L = lib()
        for _ in range(nsteps):
            rc = L.ora_step_gen2(C.byref(self.cfg), self.n, _p(self.x), _p(self.v), _p(self.mass),
                                 _p(self.volume), _p(self.density), _p(self.pressure),
                                 _p(self.material), _p(self.color), _p(self.m), _p(self.keys),
                                 _p(self.new_index), _p(self.counts), _p(self.scan),
                                 _p(self.dvel), _p(self.S))
            if rc:
                raise RuntimeError(f"oracle step failed: {rc}")

    def dump(self):
        return {"position": self.x.copy(), "velocity": self.v.copy(),
                "material": self.material.copy(), "color": self.color.copy()}


class Gen1Oracle:
    """ParticleSystem / ParticleSystemV2 + WCSPH (2D) on the CPU."""

    def __init__(self, res=(512, 512), scene=None, threads=None):
        self.dim = len(res)
        assert self.dim == 2
        self.res = res
        self.screen_to_world_ratio = 50
        self.particle_radius = 0.05                                    # partice_system.py:20
        self.particle_diameter = 2 * self.particle_radius
        self.support_radius = self.particle_radius * 4.0
        self.m_V = 0.8 * self.particle_diameter ** self.dim            # :23
        self.particle_max_num = 2 ** 15
        h = self.support_radius
        self.grid_num = np.ceil(np.array(res) / h).astype(int)         # :30
        kw, kdw = kernel_constants(2, h)
        c = OraConfig()
        c.dim = 2
        c.grid_num[:] = [int(self.grid_num[0]), int(self.grid_num[1]), 1]
        c.h = h
        c.padding = h
        c.dt = 2e-4
        c.g[:] = [0.0, -9.80, 0.0]                                     # const.py:2, wcsph.py:59
        c.rho0 = 1000.0
        c.stiffness = 50.0
        c.exponent = 7.0
        c.k_w, c.k_dw = kw, kdw
        c.eps_h2 = 0.01 * h ** 2
        c.m_V = self.m_V
        c.g1_visc_c = 2 * (self.dim + 2) * 0.05                        # sph_base.py:81
        c.g1_mass = self.m_V * 1000.0                                  # sph_base.py:16
        c.g1_press_c = -1000.0 * self.m_V                              # sph_base.py:68
        c.max_per_cell = 100
        c.max_neighbors = 100
        self.cfg = c
        if threads:
            lib().ora_set_num_threads(int(threads))
        self.n = 0
        self.x = np.zeros((0, 2), np.float32); self.v = np.zeros((0, 2), np.float32)
        self.density = np.zeros(0, np.float32); self.pressure = np.zeros(0, np.float32)
        self.material = np.zeros(0, np.int32); self.color = np.zeros(0, np.int32)
        if scene is not None:                       # partice_systemv2.py:124-136
            for fluid in scene["fluidBlocks"]:
                start, end = fluid["start"], fluid["end"]
                self.add_cube(start, [end[0] - start[0], end[1] - start[1]], material=1,
                              color=0x111111, density=fluid["density"],
                              velocity=fluid["velocity"])

    def add_cube(self, lower_corner, cube_size, material, color=0xFFFFFF, density=None,
                 pressure=None, velocity=None):
        pos = cube_positions(lower_corner, cube_size, self.particle_radius, 2)
        k = len(pos)
        assert self.n + k <= self.particle_max_num                     # :150
        vel = np.full(pos.shape, 0 if velocity is None else velocity, dtype=np.float32)
        self.x = np.concatenate([self.x, pos]); self.v = np.concatenate([self.v, vel])
        self.density = np.concatenate([self.density, np.full(k, density if density is not None else 1000., np.float32)])
        self.pressure = np.concatenate([self.pressure, np.full(k, pressure if pressure is not None else 0., np.float32)])
        self.material = np.concatenate([self.material, np.full(k, material, np.int32)])
        self.color = np.concatenate([self.color, np.full(k, color, np.int32)])
        self.n += k
        self.neighbors = np.zeros((self.n, 100), np.int32)
        self.neighbors_num = np.zeros(self.n, np.int32)
        self.dvel = np.zeros((self.n, 2), np.float32)

    def set_state(self, x, v):
        self.x = _f32(x).copy(); self.v = _f32(v).copy()

    def step(self, trace=False):
        L = lib(); c = C.byref(self.cfg); n = self.n
        t = {}
        ov = L.ora_g1_search_neighbors(c, n, _p(self.x), _p(self.material), _p(self.neighbors),
                                       _p(self.neighbors_num))
        if ov:
            raise RuntimeError(f"{ov} cell/neighbour list overflows (reference UB)")
        L.ora_g1_density(c, n, _p(self.x), _p(self.material), _p(self.neighbors),
                         _p(self.neighbors_num), _p(self.density))
        if trace:
            t.update(neighbors=self.neighbors.copy(), neighbor_count=self.neighbors_num.copy(),
                     density_pre=self.density.copy())
        L.ora_g1_non_pressure(c, n, _p(self.x), _p(self.v), _p(self.density), _p(self.material),
                              _p(self.neighbors), _p(self.neighbors_num), _p(self.dvel))
        if trace:
            t.update(a_nonpressure=self.dvel.copy())
        L.ora_eos(c, n, _p(self.density), _p(self.pressure))
        L.ora_g1_pressure_force(c, n, _p(self.x), _p(self.density), _p(self.pressure),
                                _p(self.material), _p(self.neighbors), _p(self.neighbors_num),
                                _p(self.dvel))
        if trace:
            t.update(density=self.density.copy(), pressure=self.pressure.copy(),
                     d_velocity=self.dvel.copy())
            t.update(self.force_magnitudes(self.x, self.v, t["density_pre"], self.density, self.pressure,
                                           self.neighbors, self.neighbors_num))
            t.update(a_pressure=(t["d_velocity"].astype(np.float64) - t["a_nonpressure"]).astype(np.float32),
                     material=self.material.copy())
        L.ora_advect(c, n, _p(self.x), _p(self.v), _p(self.dvel), _p(self.material))
        if trace:
            t.update(x=self.x.copy(), v=self.v.copy())
        return t

    def dump(self):
        return {"position": self.x.copy(), "velocity": self.v.copy(),
                "material": self.material.copy(), "color": self.color.copy()}


def _g1_force_magnitudes(self, x, v, density_pre, density, pressure, neighbors, neighbors_num):
    """test support (see Gen2Oracle.force_magnitudes): magnitude sums of the gen-1 acceleration terms for
    the given pre-advection state, and the same pressure sum with the tolerance floor of
    p = B (x^7 - 1) in the place of p"""
    n = len(x)
    c = C.byref(self.cfg)
    mnp, mp, mpf, scratch = (np.zeros(n, np.float32) for _ in range(4))
    args = (_p(_f32(x)), _p(_f32(v)), _p(_f32(density_pre)), _p(_f32(density)))
    tail = (_p(np.ascontiguousarray(self.material, np.int32)), _p(np.ascontiguousarray(neighbors, np.int32)),
            _p(np.ascontiguousarray(neighbors_num, np.int32)))
    lib().ora_g1_force_magnitudes(c, n, *args, _p(_f32(pressure)), *tail, _p(mnp), _p(mp))
    p_floor = (50 * 8 * np.finfo(np.float32).eps * (np.asarray(density, np.float64) / 1000.0) ** 7).astype(np.float32)
    lib().ora_g1_force_magnitudes(c, n, *args, _p(p_floor), *tail, _p(scratch), _p(mpf))
    return {"mag_nonpressure": mnp, "mag_pressure": mp, "mag_pressure_floor": mpf}


Gen1Oracle.force_magnitudes = _g1_force_magnitudes
