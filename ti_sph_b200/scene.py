"""Scene helpers: the reference's JSON schema, its derived constants and the benchmark scenes.

Host-side restatement of the constructor arithmetic of
core/partice_system/partice_systemv4.py:8-78 and core/sph/sph_basev2.py / wcsphv2.py
(`__init__`), evaluated in float64 exactly where the reference's Python scope does.
"""
import copy
from functools import reduce

import numpy as np

from . import _capi

MATERIAL_BOUNDARY = 0     # partice_systemv4.py:24
MATERIAL_FLUID = 1        # partice_systemv4.py:25

DEMO_3D = {               # data/scenes/demo_3d.json of the reference (consumed keys only)
    "configuration": {
        "dim": 3, "domainStart": [0.0, 0.0, 0.0], "domainEnd": [5.0, 3.0, 2.0],
        "particleRadius": 0.01, "density0": 1000, "gravitation": [0.0, -9.81, 0.0], "c_s": 88.5,
    },
    "rigidBodies": [],
    "fluidBlocks": [{"objectId": 0, "start": [0.3, 0.1, 0.7], "end": [1.0, 1.0, 1.0],
                     "velocity": [0.0, -1.0, 10.0], "density": 1000.0, "color": [50, 100, 200]}],
}


def bench_scene(name):
    """The BASELINE.md configurations C2..C5 (C1 is the 2D gen-1 demo)."""
    s = copy.deepcopy(DEMO_3D)
    blk = s["fluidBlocks"][0]
    if name == "C2":      # demo_3d.json as shipped: 70x90x31 = 195,300
        pass
    elif name == "C3":    # 100^3 = 1,000,000
        blk["start"], blk["end"] = [0.3, 0.1, 0.3], [1.3, 1.1, 1.3]
    elif name == "C4":    # 200x200x100 = 4,000,000 (+ mesh boundary added by the caller)
        s["configuration"]["particleRadius"] = 0.005
        blk["start"], blk["end"] = [0.5, 0.5, 0.5], [1.5, 1.5, 1.0]
    elif name == "C5":    # 400x200x200 = 16,000,000
        s["configuration"]["particleRadius"] = 0.005
        blk["start"], blk["end"] = [0.3, 0.1, 0.3], [2.3, 1.1, 1.3]
    else:
        raise ValueError(f"unknown bench scene {name!r}")
    return s


def kernel_constants(dim, h):
    """k/h^dim and 6k/h^dim (sph_basev2.py:22-30, 42-50), float64."""
    k = {1: 4 / 3, 2: 40 / (7 * np.pi), 3: 8 / np.pi}[dim]
    k6 = {1: 4 / 3, 2: 40 / 7 / np.pi, 3: 8 / np.pi}[dim]
    return k / h ** dim, 6. * k6 / h ** dim


def cube_positions(lower_corner, cube_size, spacing, dim):
    """Particle lattice of add_cube (partice_systemv4.py:356-366): np.arange per axis with
    step `spacing`, meshgrid(indexing='ij'), f32, first axis slowest."""
    num_dim = [np.arange(lower_corner[i], lower_corner[i] + cube_size[i], spacing)
               for i in range(dim)]
    n = reduce(lambda x, y: x * y, [len(a) for a in num_dim])
    positions = np.array(np.meshgrid(*num_dim, sparse=False, indexing='ij'), dtype=np.float32)
    return np.ascontiguousarray(positions.reshape(dim, n).T)


def cube_particle_num(start, end, spacing, dim):
    """compute_cube_particles_num (partice_systemv4.py:160-168)."""
    return reduce(lambda x, y: x * y, [len(np.arange(start[i], end[i], spacing)) for i in range(dim)])


def gen2_config(configuration, capacity, device=0, density_mode=0, volume_mode=0):
    """tisph_config for ParticleSystemV4 + WCSPHV2."""
    dim = configuration["dim"]
    domain_size = np.array(configuration["domainEnd"]) - np.array(configuration["domainStart"])
    r = configuration["particleRadius"]
    h = 4.0 * r                                                # :34
    grid_num = np.ceil(domain_size / h).astype(np.int32)      # :59
    kw, kdw = kernel_constants(dim, h)
    c = _capi.Config()
    c.struct_size = _capi.C.sizeof(_capi.Config)
    c.generation, c.dim, c.device, c.capacity = 2, dim, device, int(capacity)
    c.grid_num[:] = [int(g) for g in grid_num] + [1] * (3 - dim)
    c.support = h
    c.padding = h                                              # :35
    c.domain_size[:] = [float(s) for s in domain_size] + [0.0] * (3 - dim)
    c.wall_hi[:] = [float(s - h) for s in domain_size] + [0.0] * (3 - dim)
    c.m_V0 = 0.8 * (2 * r) ** dim                              # :47-48
    c.dt = 2e-4                                                # sph_basev2.py:15
    c.gravity[:] = [float(g) for g in configuration["gravitation"]]
    c.c_s = configuration["c_s"]
    c.rho0 = 1000.0                                            # sph_basev2.py:13
    c.ps_density0 = configuration["density0"]
    c.stiffness, c.exponent = 50.0, 7.0                        # wcsphv2.py:10-11
    c.k_w, c.k_dw = kw, kdw
    c.visc_fluid_c = 2 * 0.05 * h * configuration["c_s"]       # wcsphv2.py:69
    c.visc_bound_c = 0.08 * h * configuration["c_s"]           # wcsphv2.py:75-76
    c.eps_h2 = 0.01 * h ** 2
    c.density_mode, c.volume_mode = int(density_mode), int(volume_mode)
    return c


DEMO_2D = {               # data/scenes/demo_2d.json of the reference (the keys gen-1 consumes)
    "configuration": {"domainStart": [0.0, 0.0, 0.0], "domainEnd": [5.0, 3.0, 2.0], "particleRadius": 0.01,
                      "density0": 1000, "viscosity": 0.01, "gravitation": [0.0, -9.81, 0.0]},
    "rigidBodies": [],
    "fluidBlocks": [{"objectId": 1, "start": [3, 1], "end": [6, 6], "velocity": [0, -20], "density": 1000.0,
                     "color": [50, 100, 200]}],
}

GEN1_GRAVITY = -9.80      # core/const.py:2


def gen1_config(res, device=0):
    """tisph_config for ParticleSystem / ParticleSystemV2 + WCSPH (2D; partice_system.py:8-34,
    sph_base.py:10-16, wcsph.py:8-15).  Everything but `res` is hard-coded in the reference."""
    dim = len(res)
    if dim != 2:
        raise ValueError("the gen-1 path is 2D (the reference's 3D variant of it is not runnable)")
    r = 0.05                                                   # partice_system.py:20
    h = r * 4.0                                                # :22
    m_V = 0.8 * (2 * r) ** dim                                 # :23
    grid_num = np.ceil(np.array(res) / h).astype(int)          # :30 (pixel resolution / support radius)
    kw, kdw = kernel_constants(dim, h)
    c = _capi.Config()
    c.struct_size = _capi.C.sizeof(_capi.Config)
    c.generation, c.dim, c.device, c.capacity = 1, dim, device, 2 ** 15      # :24
    c.grid_num[:] = [int(grid_num[0]), int(grid_num[1]), 1]
    c.support = h
    c.padding = h                                              # :34
    c.m_V0 = m_V
    c.dt = 2e-4                                                # sph_base.py:14-15
    c.gravity[:] = [0.0, GEN1_GRAVITY, 0.0]                    # wcsph.py:59: d_v[dim-1] = const.g
    c.rho0 = 1000.0                                            # sph_base.py:13
    c.ps_density0 = 1000.0
    c.stiffness, c.exponent = 50.0, 7.0                        # wcsph.py:10-11
    c.k_w, c.k_dw = kw, kdw
    c.eps_h2 = 0.01 * h ** 2                                   # sph_base.py:82
    c.g1_visc_c = 2 * (dim + 2) * 0.05                         # sph_base.py:81
    c.g1_mass = m_V * 1000.0                                   # sph_base.py:16
    c.g1_press_c = -1000.0 * m_V                               # sph_base.py:68
    return c
