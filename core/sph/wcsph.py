"""Drop-in for the reference's core/sph/wcsph.py (WCSPH, gen-1 2D).

  compute_densities (:18-32) + clamp/EOS (:37-40)                  TISPH_STAGE_DENSITY
  compute_non_pressure_force (:52-65), compute_pressure_force
  launch B (:42-49), advert (:67-72)                               TISPH_STAGE_FORCE_ADVECT
"""
import core.const as const
from core.sph.sph_base import SPHBase
from ti_sph_b200 import _capi as K
from ti_sph_b200.fields import FieldView


class WCSPH(SPHBase):
    def __init__(self, particle_system):
        super().__init__(particle_system)
        self.exponent = 7.0
        self.stiffness = 50.0
        self.g = const.g
        self.d_velocity = FieldView(self, K.F_D_VELOCITY, "d_velocity")

    def compute_densities(self):
        self.engine.stage(K.STAGE_DENSITY)

    def substep(self):
        self.engine.stage(K.STAGE_DENSITY)
        self.engine.stage(K.STAGE_FORCE_ADVECT)
