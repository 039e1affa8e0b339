"""Drop-in for the reference's core/sph/sph_base.py (SPHBase, gen-1).  step() keeps the reference
order (sph_base.py:168-172): ps.init(), boundary volumes (no boundary particles can exist in
gen-1), substep(), enforce_boundary() (a no-op in the reference, :161-166)."""
from ti_sph_b200 import _capi as K
from ti_sph_b200.fields import ScalarView


class SPHBase:
    def __init__(self, particle_system):
        self.ps = particle_system
        self.engine = particle_system.engine
        self.viscosity = 0.05
        self.density_0 = 1000.0
        self.dt = ScalarView(lambda: self.engine.get_param(K.P_DT),
                             lambda v: self.engine.set_param(K.P_DT, v))
        self.dt[None] = 2e-4
        self.mass = self.ps.m_V * self.density_0

    def substep(self):
        pass

    def enforce_boundary(self):
        pass

    def step(self):
        self.engine.step(1)
