"""Drop-in for the reference's core/sph/sph_basev2.py (SPHBaseV2).

step() is the reference's (sph_basev2.py:210-214): ps.update(), boundary volumes, substep(),
enforce_boundary().  A solver whose hooks are the stock ones takes the fused path (three stages of
libtisph.so, one call); a subclass that overrides substep() / enforce_boundary() / a compute_*
kernel gets them called in the reference's order, each mapped onto the stage that contains it:

  compute_volume_of_boundary_particle (:195-201)   TISPH_STAGE_DENSITY (with compute_densities, clamp + EOS)
  enforce_boundary (:204-208)                      TISPH_STAGE_WALLS after a split TISPH_STAGE_FORCE_ADVECT
"""
from ti_sph_b200 import _capi as K
from ti_sph_b200.fields import ScalarView


class _Gravity(list):
    """solver.g: a list whose item assignment reaches the kernels (sph_basev2.py:16)"""

    def __init__(self, values, push):
        super().__init__(values)
        self._push = push

    def __setitem__(self, k, v):
        super().__setitem__(k, v)
        self._push(self)


def _engine_attr(param, doc):
    def get(self):
        return self.engine.get_param(param)

    def set_(self, value):
        cur = self.engine.get_param(param)
        if abs(cur - value) > 1e-6 * abs(value):      # (re-deriving a coefficient from an unchanged value could move its last bit)
            self.engine.set_param(param, value)
    return property(get, set_, doc=doc)


class SPHBaseV2:
    viscosity = _engine_attr(K.P_VISCOSITY, "sph_basev2.py:12; assignable, the kernels' coefficients follow")
    density_0 = _engine_attr(K.P_DENSITY0, "sph_basev2.py:13")

    def __init__(self, particle_system):
        self.ps = particle_system
        self.engine = particle_system.engine
        self.viscosity = 0.05
        self.density_0 = 1000.0
        self.dt = ScalarView(lambda: self.engine.get_param(K.P_DT),
                             lambda v: self.engine.set_param(K.P_DT, v))
        self.dt[None] = 2e-4
        self.g = self.ps.configuration['gravitation']

    @property
    def g(self):
        return self._g

    @g.setter
    def g(self, values):
        self._g = _Gravity(values, self._push_gravity)
        self._push_gravity(self._g)

    def _push_gravity(self, g):
        for k, p in enumerate((K.P_GRAVITY_X, K.P_GRAVITY_Y, K.P_GRAVITY_Z)[:len(g)]):
            self.engine.set_param(p, g[k])

    def _field_override(self, name, field):
        return self.ps._field_override(name, field)

    # ---- the kernels of a step, callable one by one like the reference's ----------------------
    def _phase(self):
        return int(self.engine.get_param(K.P_PHASE))

    def _ensure_density(self):
        """run the density stage unless this step's already ran (ps._kernel_stage: 0 between steps,
        1 after the density stage, 2 after the force stage of a step driven kernel by kernel)"""
        if self.ps._kernel_stage >= 1:
            return
        if self._phase() == 0:                    # nobody sorted yet: ps.update() is part of the step
            self.ps.update()
        self.engine.stage(K.STAGE_DENSITY)
        self.ps._kernel_stage = 1
        # until compute_pressure_force: density is the unclamped one, pressure the carried-over one
        self.ps._overrides.update(density=K.F_DENSITY_RAW, pressure=K.F_PRESSURE_STORED)

    def compute_volume_of_boundary_particle(self):
        self._ensure_density()

    def enforce_boundary(self):
        if int(self.engine.get_param(K.P_SPLIT_WALLS)):
            self.engine.stage(K.STAGE_WALLS)
            self.engine.set_param(K.P_SPLIT_WALLS, 0)
        self.ps._overrides.clear()
        self.ps._kernel_stage = 0

    def substep(self):
        pass

    def step(self):
        hooks = ("substep", "enforce_boundary", "compute_volume_of_boundary_particle", "compute_densities",
                 "compute_non_pressure_force", "compute_pressure_force", "advert")
        fused = all(self._defined_by_library(n) for n in hooks)
        phase = self._phase()
        if fused and phase in (0, 1):
            self.ps._overrides.clear()
            if phase == 0:
                self.engine.step(1)
            else:                                 # ps.update() was called by the script already
                self.engine.stage(K.STAGE_DENSITY)
                self.engine.stage(K.STAGE_FORCE_ADVECT)
            return
        if phase == 0:
            self.ps.update()
        self.compute_volume_of_boundary_particle()
        self.substep()
        self.enforce_boundary()

    def _defined_by_library(self, name):
        """True when `name` is not overridden by a subclass outside this package"""
        for klass in type(self).__mro__:
            if name in vars(klass):
                return klass.__module__.startswith("core.sph.")
        return True
