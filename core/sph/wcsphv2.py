"""Drop-in for the reference's core/sph/wcsphv2.py (WCSPHV2).

The reference's kernels map onto libtisph.so stages:
  compute_densities (:28-34) + clamp/EOS (:45-47)          TISPH_STAGE_DENSITY
  compute_non_pressure_force (:83-93), compute_pressure_force launch B (:49-54),
  advert (:95-100), enforce_boundary (sph_basev2.py:204)   TISPH_STAGE_FORCE_ADVECT (fused)

step() of an unmodified solver is one fused call.  Called one by one (a script that inspects the
fields between the kernels, a subclass that hooks into substep()), every kernel method runs the
stage that contains it the first time one of its kernels is asked for, and the field views
(ps.x, ps.density, ps.pressure, solver.d_velocity ...) are redirected to what the reference's
fields hold at that point of the step:

  after compute_densities          ps.density = unclamped density, ps.pressure = carried-over pressure
  after compute_non_pressure_force solver.d_velocity = gravity + cohesion + viscosity only
  after compute_pressure_force     ps.density clamped, ps.pressure = Tait EOS, d_velocity = total
  after advert                     ps.x, ps.v advected, walls not applied yet
  after enforce_boundary           the end-of-step state
"""
from core.sph.sph_basev2 import SPHBaseV2, _engine_attr
from ti_sph_b200 import _capi as K
from ti_sph_b200.fields import FieldView


class WCSPHV2(SPHBaseV2):
    exponent = _engine_attr(K.P_EXPONENT, "wcsphv2.py:10; assignable")
    stiffness = _engine_attr(K.P_STIFFNESS, "wcsphv2.py:11; assignable")

    def __init__(self, particle_system):
        super().__init__(particle_system)
        self.exponent = 7.0
        self.stiffness = 50.0
        self.d_velocity = FieldView(self, K.F_D_VELOCITY, "d_velocity")
        self.c_s = self.ps.configuration['c_s']

    def compute_densities(self):
        self._ensure_density()

    def _ensure_forces(self):
        self._ensure_density()
        if self.ps._kernel_stage >= 2:
            return
        # the fused stage, stopped after advert(): the walls are enforce_boundary()'s; the diagnostics
        # keep the non-pressure and the pressure sums apart
        self.engine.set_param(K.P_DIAGNOSTICS, 1)
        self.engine.set_param(K.P_SPLIT_WALLS, 1)
        self.engine.stage(K.STAGE_FORCE_ADVECT)
        self.ps._kernel_stage = 2

    def compute_non_pressure_force(self):
        self._ensure_forces()
        self.ps._overrides.update(x=K.F_X_IN, v=K.F_V_IN, density=K.F_DENSITY_RAW, pressure=K.F_PRESSURE_STORED,
                                  d_velocity=K.F_A_NONPRESSURE)

    def compute_pressure_force(self):
        self._ensure_forces()
        self.ps._overrides.clear()
        self.ps._overrides.update(x=K.F_X_IN, v=K.F_V_IN)

    def advert(self):
        self._ensure_forces()
        self.ps._overrides.clear()

    def substep(self):
        """wcsphv2.py:102-106"""
        self.compute_densities()
        self.compute_non_pressure_force()
        self.compute_pressure_force()
        self.advert()
