"""Drop-in for the reference's core/partice_system/partice_system.py (ParticleSystem, gen-1, 2D).

Same constructor, attributes and methods; the Taichi kernels are calls into libtisph.so
(include/tisph.h, generation 1):

  reference (partice_system.py)                              here
  --------------------------------------------------------   --------------------------------
  fields x v density pressure material color ... (:36-59)    device records, FieldView
  add_particles kernel (:70-89)                              tisph_add_particles
  init() = fill(0) x2 + allocate_particles_to_grid
           + search_neighbors (:102-132, :211-215)           tisph_stage_run(TISPH_STAGE_UPDATE)
  particle_neighbors / particle_neighbors_num                TISPH_F_NEIGHBORS / TISPH_F_NEIGHBOR_COUNT
  dump(), copy_to_numpy(_nd) (:166-209)                      tisph_download
"""
from functools import reduce

import numpy as np

from ti_sph_b200 import _capi as K
from ti_sph_b200 import scene as _scene
from ti_sph_b200.engine import Engine
from ti_sph_b200.fields import FieldView, ScalarView


class ParticleSystem:
    def __init__(self, res, device=0):
        self.res = res
        self.dim = len(res)
        assert self.dim > 1
        self.screen_to_world_ratio = 50
        self.bound = np.array(res) / self.screen_to_world_ratio
        self.material_boundary = _scene.MATERIAL_BOUNDARY
        self.material_fluid = _scene.MATERIAL_FLUID
        self.particle_radius = 0.05
        self.particle_diameter = 2 * self.particle_radius
        self.support_radius = self.particle_radius * 4.0
        self.m_V = 0.8 * self.particle_diameter ** self.dim
        self.particle_max_num = 2 ** 15
        self.particle_max_num_per_cell = 100
        self.particle_max_num_neighbor = 100
        self.grid_size = self.support_radius
        self.grid_num = np.ceil(np.array(res) / self.grid_size).astype(int)
        self.padding = self.grid_size
        self.engine = Engine(_scene.gen1_config(res, device=device))
        self.particle_num = ScalarView(lambda: self.engine.particle_num)
        for name, fid in (("x", K.F_X), ("v", K.F_V), ("density", K.F_DENSITY), ("pressure", K.F_PRESSURE),
                          ("material", K.F_MATERIAL), ("color", K.F_COLOR), ("volume", K.F_VOLUME),
                          ("particle_neighbors", K.F_NEIGHBORS), ("particle_neighbors_num", K.F_NEIGHBOR_COUNT),
                          ("grid_particles_num", K.F_GRID_PARTICLES_NUM)):
            setattr(self, name, FieldView(self, fid, name))
        self._overrides = {}         # field name -> library field, while a step is driven kernel by kernel
        self._kernel_stage = 0       # ... and how far that step got (core/sph/sph_base.py)

    def _field_override(self, name, field):
        return self._overrides.get(name, field)

    # ---- particles -------------------------------------------------------------------------
    def add_particles(self, num, particle_position, particle_velocity, particle_density,
                      particle_pressure, particle_material, particle_color):
        self.engine.add_particles(particle_position[:num], particle_velocity[:num], particle_density[:num],
                                  particle_pressure[:num],
                                  np.asarray(particle_material[:num]).astype(np.int32),
                                  np.asarray(particle_color[:num]).astype(np.int64).astype(np.int32))

    def add_cube(self, lower_corner, cube_size, material, color=0xFFFFFF, density=None, pressure=None,
                 velocity=None):
        positions = _scene.cube_positions(lower_corner, cube_size, self.particle_radius, self.dim)
        num = positions.shape[0]
        assert self.particle_num[None] + num <= self.particle_max_num              # :150
        velocity = np.full(positions.shape, fill_value=0 if velocity is None else velocity, dtype=np.float32)
        self.add_particles(num, positions, velocity,
                           np.full(num, density if density is not None else 1000.0),
                           np.full(num, pressure if pressure is not None else 0.0),
                           np.full(num, material), np.full(num, color))

    def compute_fluid_particle_num(self, start, end):
        return reduce(lambda a, b: a * b,
                      [len(np.arange(start[i], end[i], self.particle_diameter)) for i in range(self.dim)])

    # ---- per step ----------------------------------------------------------------------------
    def init(self):
        """clear the grid and the neighbour table, rebuild both (partice_system.py:211-215)"""
        self._overrides.clear()
        self._kernel_stage = 0
        self.engine.stage(K.STAGE_UPDATE)

    def allocate_particles_to_grid(self):
        """partice_system.py:128-132 (dense cell lists).  The library builds the cell lists and the neighbour table
        in ONE stage (TISPH_STAGE_UPDATE), so either of the two kernels of init() runs that stage for the current
        state if it has not run yet, and the other one finds its result in place."""
        if int(self.engine.get_param(K.P_PHASE)) == 0:
            self.init()

    def search_neighbors(self):
        """partice_system.py:103-121 (neighbour table particle_neighbors / particle_neighbors_num)"""
        if int(self.engine.get_param(K.P_PHASE)) == 0:
            self.init()

    def copy_to_numpy(self, np_arr, src_arr):
        """partice_system.py:167-170: np_arr[i] = src_arr[i] for the particles in use"""
        n = self.engine.particle_num
        np_arr[:n] = src_arr.to_numpy()[:n]

    def copy_to_numpy_nd(self, np_arr, src_arr):
        """partice_system.py:172-176"""
        n = self.engine.particle_num
        np_arr[:n, :self.dim] = src_arr.to_numpy()[:n]

    def pos_to_index(self, pos):
        return (np.asarray(pos, np.float32) / np.float32(self.grid_size)).astype(np.int32)

    def is_valid_cell(self, cell):
        return all(0 <= cell[i] < self.grid_num[i] for i in range(self.dim))

    def dump(self, out=None):
        e, out = self.engine, out or {}
        return {'position': e.download(K.F_X, out.get('position')),
                'velocity': e.download(K.F_V, out.get('velocity')),
                'material': e.download(K.F_MATERIAL, out.get('material')),
                'color': e.download(K.F_COLOR, out.get('color'))}
