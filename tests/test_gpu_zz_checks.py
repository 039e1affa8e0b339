"""Runs last (file name): in a debug build of libtisph.so (TISPH_CHECKS=1, device-side bounds checks
of the shared-memory tiles, pending lists and list-pool rows -- compute-sanitizer is not available
on the B200 pool) no check may have failed during the whole GPU test session; also pushes the
largest and the most crowded cases through the checked kernels once more."""
import numpy as np
import pytest

from ti_sph_b200 import _capi as K
from ti_sph_b200 import scene as sc
from util import jitter, make_pair, small_scene

pytestmark = pytest.mark.gpu


def test_no_device_side_bounds_check_failed():
    scene = sc.bench_scene("C3")
    ora, eng = make_pair(scene, density_mode="summed")
    eng.step(3)                                             # 1 M particles, explosive summed mode: sparse + dense cells
    eng.sync()
    g = np.arange(20, dtype=np.float64) * 0.004
    x = np.stack(np.meshgrid(0.4 + g, 0.4 + g, 0.4 + g, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    x = jitter(x, 0.004, seed=9)
    from test_gpu_gen2_parity import _custom_pair
    ora2, eng2 = _custom_pair(x)                            # 1000 particles per cell: many tiles per item
    eng2.step(1)
    eng2.sync()
    v = eng.get_param(K.P_STAT_CHECK_FAILURES)
    eng.close(); eng2.close()
    if v < 0:
        pytest.skip("libtisph.so was built without TISPH_CHECKS")
    assert v == 0, f"{int(v)} bounds check(s) failed, first at tisph source line {round((v - int(v)) * 1e6)}"
