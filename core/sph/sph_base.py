"""Drop-in for the reference's core/sph/sph_base.py (SPHBase, gen-1).  step() keeps the reference
order (sph_base.py:168-172): ps.init(), boundary volumes (no boundary particles can exist in
gen-1), substep(), enforce_boundary() (a no-op in the reference, :161-166).  An unmodified solver
takes the fused path; a subclass that overrides a hook gets the hooks called in that order, and the
kernel methods may be called one by one (see core/sph/wcsph.py)."""
from core.sph.sph_basev2 import _engine_attr
from ti_sph_b200 import _capi as K
from ti_sph_b200.fields import ScalarView


class SPHBase:
    viscosity = _engine_attr(K.P_VISCOSITY, "sph_base.py:12; assignable, the kernels' coefficient follows")
    density_0 = _engine_attr(K.P_DENSITY0, "sph_base.py:13")

    def __init__(self, particle_system):
        self.ps = particle_system
        self.engine = particle_system.engine
        self.viscosity = 0.05
        self.density_0 = 1000.0
        self.dt = ScalarView(lambda: self.engine.get_param(K.P_DT),
                             lambda v: self.engine.set_param(K.P_DT, v))
        self.dt[None] = 2e-4
        self.mass = self.ps.m_V * self.density_0

    def _field_override(self, name, field):
        return self.ps._field_override(name, field)

    def _phase(self):
        return int(self.engine.get_param(K.P_PHASE))

    def _ensure_density(self):
        if self.ps._kernel_stage >= 1:
            return
        if self._phase() == 0:
            self.ps.init()
        self.engine.stage(K.STAGE_DENSITY)
        self.ps._kernel_stage = 1
        self.ps._overrides.update(density=K.F_DENSITY_RAW, pressure=K.F_PRESSURE_STORED)

    def compute_volume_of_boundary_particle(self):
        """sph_base.py:143-153: walks boundary particles only, and gen-1 scenes have none"""

    def substep(self):
        pass

    def enforce_boundary(self):
        """sph_base.py:161-166 is a no-op in the reference; it ends a step driven kernel by kernel"""
        self.ps._overrides.clear()
        self.ps._kernel_stage = 0

    def _defined_by_library(self, name):
        for klass in type(self).__mro__:
            if name in vars(klass):
                return klass.__module__.startswith("core.sph.")
        return True

    def step(self):
        hooks = ("substep", "enforce_boundary", "compute_volume_of_boundary_particle", "compute_densities",
                 "compute_non_pressure_force", "compute_pressure_force", "advert")
        phase = self._phase()
        if all(self._defined_by_library(n) for n in hooks) and phase in (0, 1) and self.ps._kernel_stage == 0:
            self.ps._overrides.clear()
            if phase == 0:
                self.engine.step(1)
            else:                                 # ps.init() was called by the script already
                self.engine.stage(K.STAGE_DENSITY)
                self.engine.stage(K.STAGE_FORCE_ADVECT)
            return
        if phase == 0 and self.ps._kernel_stage == 0:
            self.ps.init()
        self.compute_volume_of_boundary_particle()
        self.substep()
        self.enforce_boundary()
