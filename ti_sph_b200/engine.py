"""Engine: thin object wrapper over the C ABI (include/tisph.h). All compute is CUDA."""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import check

_FIELD_SPEC = {   # field -> (dtype, components: 'dim' | int | 'color' | 'cell')
    _capi.F_X: (np.float32, "dim"), _capi.F_V: (np.float32, "dim"),
    _capi.F_MASS: (np.float32, 1), _capi.F_VOLUME: (np.float32, 1),
    _capi.F_DENSITY: (np.float32, 1), _capi.F_PRESSURE: (np.float32, 1),
    _capi.F_MATERIAL: (np.int32, 1), _capi.F_COLOR: (np.int32, "color"),
    _capi.F_GRID_IDS: (np.int32, 1), _capi.F_GRID_PARTICLES_NUM: (np.int32, "cell"),
    _capi.F_D_VELOCITY: (np.float32, "dim"), _capi.F_DENSITY_SUM: (np.float32, 1),
    _capi.F_DENSITY_RAW: (np.float32, 1), _capi.F_NEIGHBOR_COUNT: (np.int32, 1),
    _capi.F_ORIG_ID: (np.int32, 1), _capi.F_A_NONPRESSURE: (np.float32, "dim"),
    _capi.F_A_PRESSURE: (np.float32, "dim"), _capi.F_CELL_COUNT: (np.int32, "cell"),
    _capi.F_NEIGHBORS: (np.int32, 100),
    _capi.F_X_IN: (np.float32, "dim"), _capi.F_V_IN: (np.float32, "dim"),
    _capi.F_PRESSURE_STORED: (np.float32, 1), _capi.F_PARTICLE_INDEX: (np.int32, 1),
}


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Engine:
    """One tisph_ctx: device-resident particle state + the step kernels."""

    def __init__(self, config):
        self._lib = _capi.load()
        self.config = config
        self.dim = config.dim
        self.generation = config.generation
        self.ncell = int(config.grid_num[0]) * int(config.grid_num[1]) * int(config.grid_num[2])
        self._ctx = C.c_void_p()
        check(self._lib.tisph_create(C.byref(config), C.byref(self._ctx)))

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._lib.tisph_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- particles ---------------------------------------------------------------------
    @property
    def particle_num(self):
        n = C.c_int32()
        check(self._lib.tisph_particle_num(self._ctx, C.byref(n)))
        return n.value

    def add_particles(self, pos, vel, density, pressure, material, color=None):
        n = len(pos)
        pos = np.ascontiguousarray(pos, np.float32).reshape(n, self.dim)
        vel = np.ascontiguousarray(vel, np.float32).reshape(n, self.dim)
        density = np.ascontiguousarray(density, np.float32).reshape(n)
        pressure = np.ascontiguousarray(pressure, np.float32).reshape(n)
        material = np.ascontiguousarray(material, np.int32).reshape(n)
        if color is not None:
            color = np.ascontiguousarray(color, np.int32)
        check(self._lib.tisph_add_particles(self._ctx, n, _ptr(pos), _ptr(vel), _ptr(density),
                                            _ptr(pressure), _ptr(material), _ptr(color)))

    def reset(self):
        check(self._lib.tisph_reset(self._ctx))

    def save_state(self):
        check(self._lib.tisph_state_save(self._ctx))

    def restore_state(self):
        check(self._lib.tisph_state_restore(self._ctx))

    def upload_xv(self, pos, vel):
        """overwrite x and v of the (owned) particles in their current order"""
        n = self.particle_num
        pos = np.ascontiguousarray(pos, np.float32).reshape(n, self.dim)
        vel = np.ascontiguousarray(vel, np.float32).reshape(n, self.dim)
        check(self._lib.tisph_upload_xv(self._ctx, _ptr(pos), _ptr(vel)))

    def upload_xv_async(self, pos, vel):
        """upload_xv without blocking: `pos` / `vel` must be C-contiguous f32 [n][dim] (page-locked for a truly
        asynchronous copy) and stay untouched until the engine is synchronised"""
        n = self.particle_num
        assert pos.dtype == np.float32 and vel.dtype == np.float32 and pos.flags.c_contiguous and vel.flags.c_contiguous
        assert pos.size == n * self.dim and vel.size == n * self.dim
        check(self._lib.tisph_upload_xv_async(self._ctx, _ptr(pos), _ptr(vel)))

    def upload_xv_stage(self, pos, vel):
        """first half of upload_xv_async: start the host->device copy; allowed while a step is running"""
        assert pos.dtype == np.float32 and vel.dtype == np.float32 and pos.flags.c_contiguous and vel.flags.c_contiguous
        assert pos.size == vel.size and pos.size % self.dim == 0
        check(self._lib.tisph_upload_xv_stage(self._ctx, _ptr(pos), _ptr(vel), pos.size // self.dim))

    def upload_xv_commit(self):
        """second half: between steps, the staged arrays become x, v of the owned particles"""
        check(self._lib.tisph_upload_xv_commit(self._ctx))

    def dump_async(self, position=None, velocity=None, material=None, color=None, orig_id=None):
        """start filling the given preallocated host arrays (any subset) with the current state; they are
        complete after dump_wait()"""
        n = self.particle_num
        for a, dt, cols in ((position, np.float32, self.dim), (velocity, np.float32, self.dim), (material, np.int32, 1),
                            (color, np.int32, 3 if self.generation == 2 else 1), (orig_id, np.int32, 1)):
            assert a is None or (a.dtype == dt and a.flags.c_contiguous and a.size == n * cols)
        check(self._lib.tisph_dump_async(self._ctx, _ptr(position), _ptr(velocity), _ptr(material), _ptr(color),
                                         _ptr(orig_id)))

    def dump_wait(self):
        check(self._lib.tisph_dump_wait(self._ctx))

    # -- stepping ----------------------------------------------------------------------
    def step(self, nsteps=1):
        check(self._lib.tisph_step(self._ctx, int(nsteps)))

    def stage(self, stage):
        check(self._lib.tisph_stage_run(self._ctx, int(stage)))

    def sync(self):
        check(self._lib.tisph_sync(self._ctx))

    # -- data --------------------------------------------------------------------------
    def field_shape(self, field):
        dtype, comp = _FIELD_SPEC[field]
        n = self.particle_num
        if comp == "cell":
            return dtype, (self.ncell,)
        if comp == "dim":
            return dtype, (n, self.dim)
        if comp == "color":
            return dtype, ((n, 3) if self.generation == 2 else (n,))
        if isinstance(comp, int) and comp > 1:
            return dtype, (n, comp)
        return dtype, (n,)

    def download(self, field, out=None):
        dtype, shape = self.field_shape(field)
        if out is None:
            out = np.empty(shape, dtype)
        assert out.dtype == dtype and out.flags.c_contiguous and out.size == int(np.prod(shape))
        check(self._lib.tisph_download(self._ctx, int(field), _ptr(out), out.nbytes))
        return out

    def device_ptr(self, field):
        p, stride = C.c_void_p(), C.c_int32()
        check(self._lib.tisph_device_ptr(self._ctx, int(field), C.byref(p), C.byref(stride)))
        return p.value, stride.value

    # -- parameters ----------------------------------------------------------------------
    def set_param(self, param, value):
        check(self._lib.tisph_set_param(self._ctx, int(param), float(value)))

    def get_param(self, param):
        v = C.c_double()
        check(self._lib.tisph_get_param(self._ctx, int(param), C.byref(v)))
        return v.value

    def set_stream(self, cuda_stream):
        check(self._lib.tisph_set_stream(self._ctx, C.c_void_p(cuda_stream or 0)))

    @property
    def launch_count(self):
        v = C.c_int64()
        check(self._lib.tisph_launch_count(self._ctx, C.byref(v)))
        return v.value

    def stage_times(self, enable=True):
        a, b, c, s = C.c_float(), C.c_float(), C.c_float(), C.c_int32()
        check(self._lib.tisph_stage_times(self._ctx, int(bool(enable)), C.byref(a), C.byref(b),
                                          C.byref(c), C.byref(s)))
        return {"update_ms": a.value, "density_ms": b.value, "force_ms": c.value, "steps": s.value}

    # -- slab sharding (include/tisph.h, "spatial-slab sharding") ------------------------------
    def shard_config(self, plane_lo, plane_hi, ghost_planes, left_lo, right_hi, message_capacity):
        """left_lo / right_hi: far edges of the neighbouring slabs, -1 where there is no neighbour"""
        check(self._lib.tisph_shard_config(self._ctx, int(plane_lo), int(plane_hi), int(ghost_planes),
                                           int(left_lo), int(right_hi), int(message_capacity)))
        self._msg_cap = int(message_capacity)

    def shard_config_rows(self, row_lo, row_hi, ghost_planes, left_row_lo, right_row_hi, message_capacity):
        """slab of the cell rows cx * gy + cy in [row_lo, row_hi); far edges of the neighbours, -1 = no neighbour"""
        check(self._lib.tisph_shard_config_rows(self._ctx, int(row_lo), int(row_hi), int(ghost_planes),
                                                int(left_row_lo), int(right_row_hi), int(message_capacity)))
        self._msg_cap = int(message_capacity)

    def row_counts(self):
        """particles per cell row (cx, cy) at the last sort (ghosts included), gx * gy values"""
        out = np.zeros(int(self.config.grid_num[0]) * int(self.config.grid_num[1]), np.int32)
        check(self._lib.tisph_row_counts(self._ctx, _ptr(out)))
        return out

    def plane_counts(self):
        """particles per x-plane at the last sort (ghosts included)"""
        out = np.zeros(int(self.config.grid_num[0]), np.int32)
        check(self._lib.tisph_plane_counts(self._ctx, _ptr(out)))
        return out

    def shard_pack(self):
        nl, nr = C.c_int32(), C.c_int32()
        check(self._lib.tisph_shard_pack(self._ctx, C.byref(nl), C.byref(nr)))
        return nl.value, nr.value

    def message_tensor(self, which, n):
        """torch view (no copy) of the first n records of a message buffer:
        0 send-left, 1 send-right, 2 recv-left, 3 recv-right; a record is 12 f32."""
        import torch
        p, cap = C.c_void_p(), C.c_int32()
        check(self._lib.tisph_shard_buffer(self._ctx, int(which), C.byref(p), C.byref(cap)))
        if n > cap.value:
            raise _capi.TisphError(-3, f"{n} halo records exceed the message capacity {cap.value}")
        if n == 0:
            return torch.empty((0, 12), dtype=torch.float32, device=f"cuda:{self.config.device}")
        return torch.as_tensor(_DeviceArray(p.value, (int(n), 12)), device=f"cuda:{self.config.device}")

    def shard_ipc_export(self):
        """CUDA IPC handles of this context's four receive buffers (bytes)"""
        buf = C.create_string_buffer(256)
        check(self._lib.tisph_shard_ipc_export(self._ctx, buf, 256))
        return buf.raw

    def shard_ipc_connect(self, side, handles):
        """map the receive buffers of the neighbour on `side` (0 left, 1 right): packs then write there"""
        check(self._lib.tisph_shard_ipc_connect(self._ctx, int(side), C.create_string_buffer(handles, 256), 256))

    def shard_ipc_disconnect(self):
        check(self._lib.tisph_shard_ipc_disconnect(self._ctx))

    def shard_append(self, n_from_left, n_from_right):
        check(self._lib.tisph_shard_append(self._ctx, int(n_from_left), int(n_from_right)))


class _DeviceArray:
    """__cuda_array_interface__ carrier for a raw device pointer (f32, C-contiguous)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 3, "strides": None}
