"""Drop-in for the reference's core/sph/wcsph.py (WCSPH, gen-1 2D).

  compute_densities (:18-32) + clamp/EOS (:37-40)                  TISPH_STAGE_DENSITY
  compute_non_pressure_force (:52-65), compute_pressure_force
  launch B (:42-49), advert (:67-72)                               TISPH_STAGE_FORCE_ADVECT

Called one by one, every kernel method runs the stage that contains it the first time one of its
kernels is asked for, and the field views show what the reference's fields hold at that point
(the same scheme as core/sph/wcsphv2.py).
"""
import core.const as const
from core.sph.sph_base import SPHBase
from core.sph.sph_basev2 import _engine_attr
from ti_sph_b200 import _capi as K
from ti_sph_b200.fields import FieldView


class WCSPH(SPHBase):
    exponent = _engine_attr(K.P_EXPONENT, "wcsph.py:10; assignable")
    stiffness = _engine_attr(K.P_STIFFNESS, "wcsph.py:11; assignable")
    g = _engine_attr(K.P_GRAVITY_Y, "wcsph.py:59: d_v[dim-1] = const.g; assignable")

    def __init__(self, particle_system):
        super().__init__(particle_system)
        self.exponent = 7.0
        self.stiffness = 50.0
        self.g = const.g
        self.d_velocity = FieldView(self, K.F_D_VELOCITY, "d_velocity")

    def compute_densities(self):
        self._ensure_density()

    def _ensure_forces(self):
        self._ensure_density()
        if self.ps._kernel_stage >= 2:
            return
        self.engine.set_param(K.P_DIAGNOSTICS, 1)     # keeps the non-pressure and the pressure sums apart
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
        """wcsph.py:74-78"""
        self.compute_densities()
        self.compute_non_pressure_force()
        self.compute_pressure_force()
        self.advert()
