"""Field views: what the reference exposes as Taichi fields (`ps.x`, `ps.v`, `solver.dt`, ...).

A FieldView never copies on its own: `.to_numpy()` is a device->host copy of the field in the
reference's layout, `__cuda_array_interface__` / `.to_torch()` are zero-copy views of the packed
float4 records on the device (what `scene.particles(ps.x)` needs, main_3d.py:41), and
`.to_taichi()` mirrors into a Taichi field when Taichi is importable.
"""
import numpy as np

from . import _capi as K

_ZERO_COPY = (K.F_X, K.F_V, K.F_D_VELOCITY)


class FieldView:
    def __init__(self, owner, field, name):
        self._owner = owner          # object with an `.engine`
        self._field0 = field
        self.name = name

    @property
    def _field(self):
        """the library field behind this view: while a step is driven kernel by kernel the owner
        redirects some names to what the reference's field holds between those kernels"""
        override = getattr(self._owner, "_field_override", None)
        return self._field0 if override is None else override(self.name, self._field0)

    @property
    def _eng(self):
        return self._owner.engine

    @property
    def dtype(self):
        return np.dtype(self._eng.field_shape(self._field)[0])

    @property
    def shape(self):
        # Taichi fields are sized particle_max_num / ncell; vector width is not part of shape
        shp = self._eng.field_shape(self._field)[1]
        return shp[:1]

    def to_numpy(self):
        return self._eng.download(self._field)

    def __array__(self, dtype=None, copy=None):
        a = self.to_numpy()
        return a if dtype is None else a.astype(dtype)

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, idx):
        return self.to_numpy()[idx]

    @property
    def __cuda_array_interface__(self):
        if self._field not in _ZERO_COPY:
            raise AttributeError(f"{self.name} has no zero-copy device view; use to_numpy()")
        # The engine runs on its own non-blocking stream and step() is asynchronous: a consumer on another
        # stream (torch's default, ggui) must not read records the force kernel is still writing.  The
        # step is therefore completed before the pointer is handed out (CAI v3 "stream": None = the data is
        # ready, no further synchronisation is needed).
        self._eng.sync()
        ptr, stride = self._eng.device_ptr(self._field)
        n = self._eng.particle_num
        return {"shape": (n, self._eng.dim), "typestr": "<f4", "data": (ptr, False),
                "strides": (stride, 4), "version": 3, "stream": None}

    def to_torch(self, device=None):
        import torch
        return torch.as_tensor(self, device=device or "cuda")

    def to_taichi(self):
        """mirror into a Taichi field (kept and refilled on later calls, so that a GUI loop such as
        main_3d.py:41 `scene.particles(ps.x, ...)` does not allocate a field per frame).  Fields with a
        device view (x, v, d_velocity) are refilled device to device through `field.from_torch` -- no host
        round trip; the others, and Taichi builds without torch interop, go through numpy."""
        import taichi as ti     # optional: only for the ggui hand-off
        shape = self._eng.field_shape(self._field)[1]
        f = getattr(self, "_ti_field", None)
        if f is None or tuple(f.shape) != tuple(shape[:1]):
            dt = ti.f32 if self.dtype == np.float32 else ti.i32
            f = ti.Vector.field(shape[1], dtype=dt, shape=shape[0]) if len(shape) == 2 else ti.field(dtype=dt, shape=shape[0])
            self._ti_field = f
        if self._field in _ZERO_COPY and hasattr(f, "from_torch"):
            try:
                f.from_torch(self.to_torch().contiguous())       # strided float4 view -> packed [n][dim], on the device
                return f
            except Exception:
                pass
        f.from_numpy(self.to_numpy())
        return f


class ScalarView:
    """0-d field (`ps.particle_num[None]`, `solver.dt[None]`)."""

    def __init__(self, getter, setter=None):
        self._get, self._set = getter, setter

    def __getitem__(self, idx):
        assert idx is None
        return self._get()

    def __setitem__(self, idx, value):
        assert idx is None
        if self._set is None:
            raise TypeError("read-only field")
        self._set(value)

    def to_numpy(self):
        return np.asarray(self._get())


class ConstantField:
    """A per-particle field that the reference allocates but never writes (ParticleSystemV4.m,
    partice_systemv4.py:39): all zeros, nothing is stored for it."""

    def __init__(self, owner, name, dtype=np.float32):
        self._owner, self.name, self.dtype = owner, name, np.dtype(dtype)

    @property
    def shape(self):
        return (self._owner.engine.particle_num,)

    def to_numpy(self):
        return np.zeros(self.shape, self.dtype)

    def __array__(self, dtype=None, copy=None):
        return self.to_numpy() if dtype is None else self.to_numpy().astype(dtype)

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, idx):
        return self.to_numpy()[idx]
