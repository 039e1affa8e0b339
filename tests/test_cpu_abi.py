"""Host-side checks that need no GPU: the C-ABI library loads and exports every symbol that
include/tisph.h declares, struct layouts agree, and there is no CPU fallback."""
import ctypes
import os
import re

import numpy as np
import pytest

import ti_sph_b200
from ti_sph_b200 import _capi, scene as sc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "tisph.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tisph_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ti_sph_b200.load()
    names = header_symbols()
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), f"{name} declared in tisph.h but not exported"
    assert sorted(_capi.SYMBOLS) == names, "ctypes binding and header disagree"
    header = open(os.path.join(ROOT, "include", "tisph.h")).read()
    assert lib.tisph_abi_version() == _capi.ABI_VERSION == int(re.search(r"#define TISPH_ABI_VERSION (\d+)", header).group(1))


def test_config_struct_matches_header_layout():
    # the library refuses a struct of the wrong size: probe with a deliberately wrong one
    lib = ti_sph_b200.load()
    cfg = sc.gen2_config(sc.DEMO_3D["configuration"], 100)
    assert cfg.struct_size == ctypes.sizeof(_capi.Config) == 176
    bad = sc.gen2_config(sc.DEMO_3D["configuration"], 100)
    bad.struct_size = 7
    ctx = ctypes.c_void_p()
    assert lib.tisph_create(ctypes.byref(bad), ctypes.byref(ctx)) == -1
    assert b"size mismatch" in lib.tisph_last_error()


def test_no_cpu_fallback_without_device():
    lib = ti_sph_b200.load()
    if lib.tisph_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(ti_sph_b200.TisphError) as e:
        ti_sph_b200.Engine(sc.gen2_config(sc.DEMO_3D["configuration"], 100))
    assert e.value.code == _capi.ERR_NO_DEVICE and "no CPU fallback" in str(e.value)
    from core.partice_system.partice_systemv4 import ParticleSystemV4
    with pytest.raises(ti_sph_b200.TisphError):
        ParticleSystemV4(sc.DEMO_3D)


def test_product_path_never_imports_the_oracle():
    for sub in ("ti_sph_b200", "core", "utils"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, sub)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    text = open(os.path.join(dirpath, f)).read()
                    assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), (dirpath, f)
                    assert "sph_oracle" not in text and "liboracle" not in text, (dirpath, f)


def test_scene_constants_and_counts():
    cfg = sc.gen2_config(sc.DEMO_3D["configuration"], 195300)
    assert list(cfg.grid_num) == [125, 75, 50]
    assert cfg.support == np.float32(0.04) and cfg.m_V0 == np.float32(0.8 * 0.02 ** 3)
    assert cfg.k_w == pytest.approx(39788.734, rel=1e-6)
    for name, n in (("C2", 195300), ("C3", 1000000), ("C4", 4000000), ("C5", 16000000)):
        s = sc.bench_scene(name)
        blk, r = s["fluidBlocks"][0], s["configuration"]["particleRadius"]
        assert sc.cube_particle_num(blk["start"], blk["end"], r, 3) == n
    x = sc.cube_positions([0.3, 0.1, 0.7], [0.7, 0.9, 0.3], 0.01, 3)
    assert x.shape == (195300, 3) and x.dtype == np.float32
    assert np.allclose(x[0], [0.3, 0.1, 0.7]) and np.allclose(x[1], [0.3, 0.1, 0.71])


def test_lines_helper():
    from utils.lines import getlines
    pts, idx = getlines(sc.DEMO_3D["configuration"])
    pts = np.asarray(pts.to_numpy() if hasattr(pts, "to_numpy") else pts)
    assert pts.shape == (8, 3) and len(np.asarray(idx.to_numpy() if hasattr(idx, "to_numpy") else idx)) == 24


def test_enum_values_of_the_header_match_the_ctypes_binding():
    """tisph_field / tisph_stage / tisph_param / tisph_status: the numbers in include/tisph.h are the
    numbers ti_sph_b200/_capi.py uses"""
    text = open(os.path.join(ROOT, "include", "tisph.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    values = {m.group(1): int(m.group(2)) for m in re.finditer(r"\b(TISPH_[A-Z0-9_]+)\s*=\s*(-?\d+)", text)}
    assert len(values) >= 40
    for name, value in values.items():
        for prefix, strip in (("TISPH_F_", "F_"), ("TISPH_STAGE_", "STAGE_"), ("TISPH_P_", "P_")):
            if name.startswith(prefix):
                assert getattr(_capi, strip + name[len(prefix):]) == value, name
    assert values["TISPH_ERR_NO_DEVICE"] == _capi.ERR_NO_DEVICE
    fields = sorted(v for k, v in values.items() if k.startswith("TISPH_F_"))
    assert fields == list(range(len(fields)))                       # dense numbering, no duplicates
    params = sorted(v for k, v in values.items() if k.startswith("TISPH_P_"))
    assert params == list(range(len(params)))


def test_lines_helper_returns_taichi_fields_when_taichi_is_importable():
    """main_3d.py:21,43 hands the result to ggui scene.lines; with a `taichi` module on the path
    (here: the stand-in of tests/golden/ti_emu) the helper must return fields, not arrays"""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "from utils.lines import getlines\n"
            "from ti_sph_b200 import scene as sc\n"
            "p, i = getlines(sc.DEMO_3D['configuration'])\n"
            "import taichi\n"
            "assert isinstance(p, taichi.Field) and isinstance(i, taichi.Field)\n"
            "assert p.to_numpy().shape == (8, 3) and i.to_numpy().shape == (24,)\n"
            "assert p.to_numpy().max() == 5.0 and sorted(set(i.to_numpy())) == list(range(8))\n"
            "print('ok')\n") % (ROOT, os.path.join(ROOT, "tests", "golden", "ti_emu"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-1500:]
