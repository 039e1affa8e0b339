"""Drop-in for the reference's core/sph/wcsphv2.py (WCSPHV2).

The reference's kernels map onto libtisph.so stages:
  compute_densities (:28-34) + clamp/EOS (:45-47)          TISPH_STAGE_DENSITY
  compute_non_pressure_force (:83-93), compute_pressure_force launch B (:49-54),
  advert (:95-100), enforce_boundary (sph_basev2.py:204)   TISPH_STAGE_FORCE_ADVECT (fused)
"""
from core.sph.sph_basev2 import SPHBaseV2
from ti_sph_b200 import _capi as K
from ti_sph_b200.fields import FieldView


class WCSPHV2(SPHBaseV2):
    def __init__(self, particle_system):
        super().__init__(particle_system)
        self.exponent = 7.0
        self.stiffness = 50.0
        self.d_velocity = FieldView(self, K.F_D_VELOCITY, "d_velocity")
        self.c_s = self.ps.configuration['c_s']

    def compute_densities(self):
        self.engine.stage(K.STAGE_DENSITY)

    def substep(self):
        """densities, forces, advection -- the fused CUDA stage also applies the walls, which
        the reference runs right after substep() (sph_basev2.py:213-214)."""
        self.engine.stage(K.STAGE_DENSITY)
        self.engine.stage(K.STAGE_FORCE_ADVECT)
