"""Constants module of the gen-1 (2D) solver, at the reference's module path (core/const.py).
`g` is the gravity that WCSPH puts on the last axis (wcsph.py:59); `limit` is unused by the
reference as well and kept only so that `from core.const import limit` keeps working."""
GRAVITY_LAST_AXIS = -9.80
g = GRAVITY_LAST_AXIS
limit = 1.0e-5
