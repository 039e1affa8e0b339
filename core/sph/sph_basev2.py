"""Drop-in for the reference's core/sph/sph_basev2.py (SPHBaseV2).

step() keeps the reference order (sph_basev2.py:210-214): ps.update(), boundary volumes,
substep(), enforce_boundary() -- issued as the three stages of libtisph.so.
"""
from ti_sph_b200 import _capi as K
from ti_sph_b200.fields import ScalarView


class SPHBaseV2:
    def __init__(self, particle_system):
        self.ps = particle_system
        self.engine = particle_system.engine
        self.viscosity = 0.05
        self.density_0 = 1000.0
        self.dt = ScalarView(lambda: self.engine.get_param(K.P_DT),
                             lambda v: self.engine.set_param(K.P_DT, v))
        self.dt[None] = 2e-4
        self.g = self.ps.configuration['gravitation']

    def substep(self):
        pass

    def step(self):
        self.engine.step(1)
