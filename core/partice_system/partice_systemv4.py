"""Drop-in for the reference's core/partice_system/partice_systemv4.py (ParticleSystemV4).

Same module path, class name, constructor signature, attributes and methods as the reference
class, but every Taichi kernel is replaced by a call into libtisph.so (hand-written sm_100a
CUDA, see include/tisph.h):

  reference (partice_systemv4.py)                    here
  ------------------------------------------------   ---------------------------------------
  fields x v mass volume density ... (:36-49,64-78)  device-resident float4 records, FieldView
  add_particles kernel (:171-204)                    tisph_add_particles
  update() = update_gird_id + scan + resort (:251)   tisph_stage_run(TISPH_STAGE_UPDATE)
  for_all_neighbors (:331-345)                       inside the density / force kernels
  update_gird_id / prefix_sum_executor.run / resort  TISPH_STAGE_UPDATE_BIN / _SCAN / _SORT
  dump(), copy_to_numpy(_nd) (:279-307)              tisph_download
  load_rigid_body via trimesh (:259-277)             ti_sph_b200.mesh (OBJ + voxeliser)

A script may drive a step kernel by kernel, as the reference allows (update_gird_id(),
prefix_sum_executor.run(...), resort(), then the solver's compute_* / advert / enforce_boundary):
between those calls every field view shows what the reference's field holds at that point.
"""
from functools import reduce

import numpy as np

from ti_sph_b200 import _capi as K
from ti_sph_b200 import scene as _scene
from ti_sph_b200.engine import Engine
from ti_sph_b200.fields import ConstantField, FieldView, ScalarView


class ParticleSystemV4:
    def __init__(self, simulation_config, device=0, density_mode="reference",
                 volume_mode="reference"):
        self.simulation_config = simulation_config
        self.configuration = simulation_config['configuration']
        self.rigidBodiesConfig = simulation_config['rigidBodies']
        self.fluidBlocksConfig = simulation_config['fluidBlocks']
        self.fluid = self.fluidBlocksConfig
        self.rigid = self.rigidBodiesConfig
        cfg = self.configuration
        self.density0 = cfg['density0']
        self.dim = cfg['dim']
        if self.dim != 3:
            raise ValueError("ParticleSystemV4 is the 3D system; use ParticleSystem(V2) for 2D")
        self.domain_start = np.array(cfg['domainStart'])
        self.domain_end = np.array(cfg['domainEnd'])
        self.domain_size = self.domain_end - self.domain_start
        self.material_boundary = _scene.MATERIAL_BOUNDARY
        self.material_fluid = _scene.MATERIAL_FLUID
        self.particle_radius = cfg['particleRadius']
        self.support_length = 4.0 * self.particle_radius
        self.padding = self.support_length
        self.particle_diameter = 2 * self.particle_radius
        self.m_V0 = 0.8 * self.particle_diameter ** self.dim
        self.grid_size = self.support_length
        self.grid_num = np.ceil(self.domain_size / self.grid_size).astype(np.int32)

        self._rigid_points = {}
        self.particle_max_num = 0
        self.compute_particle_num()

        self.engine = Engine(_scene.gen2_config(
            cfg, max(self.particle_max_num, 1), device=device,
            density_mode={"reference": 0, "summed": 1}[density_mode],
            volume_mode={"reference": 0, "akinci": 1}[volume_mode]))
        self.particle_num = ScalarView(lambda: self.engine.particle_num)
        for name, fid in (("x", K.F_X), ("v", K.F_V), ("mass", K.F_MASS), ("volume", K.F_VOLUME),
                          ("density", K.F_DENSITY), ("pressure", K.F_PRESSURE),
                          ("material", K.F_MATERIAL), ("color", K.F_COLOR),
                          ("grid_ids", K.F_GRID_IDS),
                          ("grid_particles_num", K.F_GRID_PARTICLES_NUM)):
            setattr(self, name, FieldView(self, fid, name))
        self.grid_particles_num_temp = FieldView(self, K.F_CELL_COUNT, "grid_particles_num_temp")
        self.paritcle_index_temp = FieldView(self, K.F_PARTICLE_INDEX, "paritcle_index_temp")   # needs P_DIAGNOSTICS
        self.prefix_sum_executor = _PrefixSumExecutor(self)          # ti.algorithms.PrefixSumExecutor, :62
        self._overrides = {}         # field name -> library field, while a step is driven kernel by kernel
        self._kernel_stage = 0       # ... and how far that step got (core/sph/sph_basev2.py)
        self.m = ConstantField(self, "m")            # allocated, sorted along and never written by the reference (:39)
        self.add_fluid_and_rigid()
        if self.engine.particle_num != self.particle_max_num:
            # reference quirk Q10: its counting pre-pass and add_cube can disagree by round-off;
            # it then simulates zero-initialised phantom particles. Refuse instead.
            raise RuntimeError("particle count pre-pass (%d) != particles added (%d)"
                               % (self.particle_max_num, self.engine.particle_num))

    # ---- scene -> particles ------------------------------------------------------------
    def compute_cube_particles_num(self, start, end):
        return _scene.cube_particle_num(start, end, self.particle_radius, self.dim)

    def compute_particle_num(self):
        for fluid in self.fluidBlocksConfig:
            self.particle_max_num += self.compute_cube_particles_num(fluid['start'], fluid['end'])
        for k, rigid in enumerate(self.rigidBodiesConfig):
            self.particle_max_num += self._rigid_body_points(k, rigid).shape[0]

    def _rigid_body_points(self, k, rigid):
        if k not in self._rigid_points:       # the reference voxelises every body twice
            self._rigid_points[k] = self.load_rigid_body(rigid)
        return self._rigid_points[k]

    def load_rigid_body(self, rigid_body):
        from ti_sph_b200 import mesh
        return mesh.sample_rigid_body(rigid_body, pitch=self.particle_diameter)

    def add_fluid_and_rigid(self):
        for k, rigid in enumerate(self.rigidBodiesConfig):
            points = self._rigid_body_points(k, rigid)
            num = points.shape[0]
            rigid['partice_num'] = num
            rigid['voxelized_points'] = points
            density = rigid.get('density')
            # the reference divides integer RGB by 255.0 and hands the float (num,3) array to the
            # i32 colour field (:111-114,190), which truncates: [255,255,255] is stored as (1,1,1)
            color = rigid.get('color', [0, 0, 0])
            if type(color[0]) == int:
                color = [c / 255.0 for c in color]
            color = np.tile(np.array(color, dtype=np.float32).astype(np.int32), (num, 1))
            self.add_particles(num, points,
                               np.tile(np.array(rigid['velocity'], dtype=np.float32), (num, 1)),
                               np.full(num, density if density is not None else 1000.0),
                               np.zeros(num),
                               np.full((num,), self.material_boundary, dtype=np.int32), color)
        for fluid in self.fluidBlocksConfig:
            start, end = fluid['start'], fluid['end']
            self.add_cube(lower_corner=start,
                          cube_size=[end[i] - start[i] for i in range(self.dim)],
                          material=self.material_fluid, color=0x111111,
                          density=fluid['density'], velocity=fluid['velocity'])

    def add_cube(self, lower_corner, cube_size, material, color=0xFFFFFF, density=None,
                 pressure=None, velocity=None):
        positions = _scene.cube_positions(lower_corner, cube_size, self.particle_radius, self.dim)
        num = positions.shape[0]
        velocity = np.full(positions.shape, fill_value=0 if velocity is None else velocity,
                           dtype=np.float32)
        self.add_particles(num, positions, velocity,
                           np.full(num, density if density is not None else 1000.0),
                           np.full(num, pressure if pressure is not None else 0.0),
                           np.full(num, material), np.full(num, color))

    def add_particles(self, num, particle_position, particle_velocity, particle_density,
                      particle_pressure, particle_material, particle_color):
        color = np.asarray(particle_color)
        if color.ndim == 1:                      # a scalar colour fills the 3-vector (:202)
            color = np.repeat(color.astype(np.int64).astype(np.int32)[:, None], 3, axis=1)
        self.engine.add_particles(particle_position[:num], particle_velocity[:num],
                                  particle_density[:num], particle_pressure[:num],
                                  np.asarray(particle_material[:num]).astype(np.int32),
                                  color[:num])

    # ---- per-step --------------------------------------------------------------------
    def _field_override(self, name, field):
        return self._overrides.get(name, field)

    def update_gird_id(self):
        """cell key per particle + histogram (partice_systemv4.py:206-215); ps.grid_ids then holds the keys
        in the particles' current order and ps.grid_particles_num the counts"""
        self._overrides.clear()
        self._kernel_stage = 0
        self.engine.stage(K.STAGE_UPDATE_BIN)

    def resort(self):
        """stable counting-sort rank + reorder of every per-particle array (partice_systemv4.py:217-249)"""
        self.engine.stage(K.STAGE_UPDATE_SORT)

    def update(self):
        """bin + prefix scan + stable counting sort + reorder (partice_systemv4.py:251-256)"""
        self.update_gird_id()
        self.prefix_sum_executor.run(self.grid_particles_num)
        self.resort()

    def search_neighbors(self):
        """partice_systemv4.py:310-330 is a leftover of the gen-1 class: it reads self.grid_particles,
        self.particle_neighbors(_num) and self.support_radius, none of which ParticleSystemV4 defines, so the
        reference raises as soon as Taichi compiles the kernel.  Same error here; gen-2 walks cells, it keeps no table."""
        raise AttributeError("'ParticleSystemV4' object has no attribute 'grid_particles' "
                             "(search_neighbors is dead code in the reference: partice_systemv4.py:310-330)")

    def copy_to_numpy(self, np_arr, src_arr):
        """partice_systemv4.py:298-301: np_arr[i] = src_arr[i] for the particles in use"""
        n = self.engine.particle_num
        np_arr[:n] = src_arr.to_numpy()[:n]

    def copy_to_numpy_nd(self, np_arr, src_arr):
        """partice_systemv4.py:303-307"""
        n = self.engine.particle_num
        np_arr[:n, :self.dim] = src_arr.to_numpy()[:n]

    def dump(self, out=None):
        """dict of host copies like the reference (:279-296). `out` (extension) may hold
        preallocated -- e.g. pinned -- arrays to fill instead of fresh ones."""
        e, out = self.engine, out or {}
        return {'position': e.download(K.F_X, out.get('position')),
                'velocity': e.download(K.F_V, out.get('velocity')),
                'material': e.download(K.F_MATERIAL, out.get('material')),
                'color': e.download(K.F_COLOR, out.get('color'))}

    def dump_async(self, out):
        """dump() that does not wait (extension): starts filling the preallocated -- ideally page-locked --
        arrays out['position'|'velocity'|'material'|'color'] and returns; dump_wait() completes them.  The copies
        run on the engine's copy stream, next to the following steps."""
        self.engine.dump_async(out.get('position'), out.get('velocity'), out.get('material'), out.get('color'))
        return out

    def dump_wait(self):
        self.engine.dump_wait()

    def is_valid_cell(self, cell):
        return all(0 <= cell[i] < self.grid_num[i] for i in range(self.dim))

    def pos_to_index(self, pos):
        h = np.float32(self.grid_size)
        return (np.asarray(pos, np.float32) / h).astype(np.int32)

    def flatten_grid_index(self, grid_index):
        return int(grid_index[0]) * int(self.grid_num[1]) * int(self.grid_num[2]) + \
            int(grid_index[1]) * int(self.grid_num[2]) + int(grid_index[2])

    def get_flatten_grid_index(self, pos):
        return self.flatten_grid_index(self.pos_to_index(pos))


class _PrefixSumExecutor:
    """ti.algorithms.PrefixSumExecutor as ParticleSystemV4 uses it (partice_systemv4.py:62,255): an
    in-place inclusive scan of ps.grid_particles_num."""

    def __init__(self, ps):
        self._ps = ps

    def run(self, field):
        if field is not self._ps.grid_particles_num:
            raise ValueError("the executor of a ParticleSystemV4 scans ps.grid_particles_num")
        self._ps.engine.stage(K.STAGE_UPDATE_SCAN)
