"""ctypes binding of include/tisph.h (the C ABI of libtisph.so).

Fails loudly when the CUDA library is missing: this package has no CPU fallback.
"""
import ctypes as C
import os

from . import build as _build

_lib = None


class TisphError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libtisph error {code}: {message}")
        self.code = code


class Config(C.Structure):
    """struct tisph_config (include/tisph.h)."""
    _fields_ = [
        ("struct_size", C.c_int32), ("generation", C.c_int32), ("dim", C.c_int32),
        ("device", C.c_int32), ("capacity", C.c_int32), ("grid_num", C.c_int32 * 3),
        ("support", C.c_float), ("padding", C.c_float), ("domain_size", C.c_float * 3),
        ("wall_hi", C.c_float * 3), ("m_V0", C.c_float), ("dt", C.c_float),
        ("gravity", C.c_float * 3), ("c_s", C.c_float), ("rho0", C.c_float),
        ("ps_density0", C.c_float), ("stiffness", C.c_float), ("exponent", C.c_float),
        ("k_w", C.c_float), ("k_dw", C.c_float), ("visc_fluid_c", C.c_float),
        ("visc_bound_c", C.c_float), ("eps_h2", C.c_float), ("g1_visc_c", C.c_float),
        ("g1_mass", C.c_float), ("g1_press_c", C.c_float), ("density_mode", C.c_int32),
        ("volume_mode", C.c_int32), ("reserved", C.c_int32 * 8),
    ]


# enum tisph_field
F_X, F_V, F_MASS, F_VOLUME, F_DENSITY, F_PRESSURE, F_MATERIAL, F_COLOR, F_GRID_IDS, \
    F_GRID_PARTICLES_NUM, F_D_VELOCITY, F_DENSITY_SUM, F_DENSITY_RAW, F_NEIGHBOR_COUNT, \
    F_ORIG_ID, F_A_NONPRESSURE, F_A_PRESSURE, F_CELL_COUNT, F_NEIGHBORS, F_X_IN, F_V_IN, \
    F_PRESSURE_STORED, F_PARTICLE_INDEX = range(23)
# enum tisph_stage
STAGE_UPDATE, STAGE_DENSITY, STAGE_FORCE_ADVECT, STAGE_UPDATE_BIN, STAGE_UPDATE_SCAN, STAGE_UPDATE_SORT, \
    STAGE_WALLS = range(7)
# enum tisph_param
P_DT, P_DENSITY_MODE, P_VOLUME_MODE, P_DIAGNOSTICS, P_KERNEL_VARIANT, P_ID_BASE, P_HAS_BOUNDARY, \
    P_STAT_ITEMS, P_STAT_FALLBACK_DENSITY, P_STAT_FALLBACK_FORCE, P_CFL, P_STAT_CHECK_FAILURES, \
    P_SPLIT_WALLS, P_PHASE, P_STIFFNESS, P_EXPONENT, P_VISCOSITY, P_DENSITY0, P_GRAVITY_X, P_GRAVITY_Y, \
    P_GRAVITY_Z, P_MAX_SPEED, P_SKIP_DISCARDED_SUM = range(23)
ABI_VERSION = 2

ERR_NO_DEVICE = -5

# every symbol include/tisph.h declares: (name, restype, argtypes)
_vp, _i32, _fp, _ip = C.c_void_p, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_int32)
SYMBOLS = {
    "tisph_last_error": (C.c_char_p, []),
    "tisph_abi_version": (C.c_int, []),
    "tisph_device_count": (C.c_int, []),
    "tisph_create": (C.c_int, [C.POINTER(Config), C.POINTER(_vp)]),
    "tisph_destroy": (C.c_int, [_vp]),
    "tisph_add_particles": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tisph_reset": (C.c_int, [_vp]),
    "tisph_particle_num": (C.c_int, [_vp, _ip]),
    "tisph_state_save": (C.c_int, [_vp]),
    "tisph_state_restore": (C.c_int, [_vp]),
    "tisph_step": (C.c_int, [_vp, _i32]),
    "tisph_stage_run": (C.c_int, [_vp, _i32]),
    "tisph_download": (C.c_int, [_vp, _i32, _vp, C.c_size_t]),
    "tisph_upload_xv": (C.c_int, [_vp, _vp, _vp]),
    "tisph_upload_xv_async": (C.c_int, [_vp, _vp, _vp]),
    "tisph_upload_xv_stage": (C.c_int, [_vp, _vp, _vp, _i32]),
    "tisph_upload_xv_commit": (C.c_int, [_vp]),
    "tisph_dump_async": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "tisph_dump_wait": (C.c_int, [_vp]),
    "tisph_device_ptr": (C.c_int, [_vp, _i32, C.POINTER(_vp), _ip]),
    "tisph_set_param": (C.c_int, [_vp, _i32, C.c_double]),
    "tisph_get_param": (C.c_int, [_vp, _i32, C.POINTER(C.c_double)]),
    "tisph_set_stream": (C.c_int, [_vp, _vp]),
    "tisph_sync": (C.c_int, [_vp]),
    "tisph_launch_count": (C.c_int, [_vp, C.POINTER(C.c_int64)]),
    "tisph_stage_times": (C.c_int, [_vp, _i32, _fp, _fp, _fp, _ip]),
    "tisph_shard_config": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i32]),
    "tisph_shard_pack": (C.c_int, [_vp, _ip, _ip]),
    "tisph_shard_buffer": (C.c_int, [_vp, _i32, C.POINTER(_vp), _ip]),
    "tisph_shard_append": (C.c_int, [_vp, _i32, _i32]),
    "tisph_plane_counts": (C.c_int, [_vp, _vp]),
    "tisph_shard_config_rows": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i32]),
    "tisph_row_counts": (C.c_int, [_vp, _vp]),
    "tisph_shard_ipc_export": (C.c_int, [_vp, _vp, C.c_size_t]),
    "tisph_shard_ipc_connect": (C.c_int, [_vp, _i32, _vp, C.c_size_t]),
    "tisph_shard_ipc_disconnect": (C.c_int, [_vp]),
    "tisph_voxelize_mesh": (C.c_int, [_i32, _vp, _i32, _vp, _i32, C.c_float, _i32, _vp, _vp, _vp]),
}


def library_path():
    return _build.LIB


def load():
    """dlopen libtisph.so and bind every declared symbol. No fallback of any kind."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if os.path.exists(path) and not _build.is_fresh() and _build.have_nvcc():
        _build.build()                       # stale against csrc/ or include/: rebuild rather than run old kernels
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `python -m ti_sph_b200.build` (needs nvcc). "
            "ti_sph_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    if lib.tisph_abi_version() != ABI_VERSION:
        raise ImportError(f"{path} has ABI version {lib.tisph_abi_version()}, this package binds {ABI_VERSION}: "
                          "rebuild it with `python -m ti_sph_b200.build --force`")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().tisph_last_error()
        raise TisphError(rc, msg.decode() if msg else "unknown error")
