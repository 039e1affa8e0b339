"""Slab sharding on the GPU.

* LocalCluster: all ranks' engines inside this process on cuda:0, messages as device-to-device
  copies -- exercises the shard kernels (pack / append / owned range / ghost-aware walks) under
  the single-GPU test run.  Compared with the CPU oracle: bit-exact order and positions equal to
  1e-5 after one step, fp32 round-off afterwards.
* torchrun + NCCL over 2 GPUs (skipped on a 1-GPU box): the same through torch.distributed.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle.oracle import Gen2Oracle
from ti_sph_b200 import _capi as K
from ti_sph_b200.sharded import LocalCluster
from test_cpu_sharded import _scene
from util import RTOL, check_end_state, rel_err, vec_rel_err

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def check_against_oracle(step_fn, dump_fn, mode, steps=4):
    """every step is a single-step comparison at 1e-5: the oracle is re-synced from the sharded state (in
    original-id order, which is also the sharded runs' intra-cell order) before the next one"""
    ora = Gen2Oracle(_scene(), density_mode=mode)
    n = ora.n
    dens0, mat0 = ora.density.copy(), ora.material.copy()
    for s in range(steps):
        t = ora.step(trace=True)
        step_fn()
        d = dump_fn()
        ids = d["orig_id"]
        assert len(ids) == n and np.array_equal(np.sort(ids), np.arange(n))     # nobody lost or duplicated
        assert np.array_equal(ids, ora.orig)                                   # global cell-sorted order, by id inside a cell
        check_end_state(d["position"], d["velocity"], t)
        assert np.array_equal(d["material"], ora.material)
        back = np.argsort(ids)
        ora.set_state(d["position"][back], d["velocity"][back], dens0, mat0)


@pytest.mark.parametrize("world,mode", [(2, "reference"), (2, "summed"), (3, "reference"), (3, "summed")])
def test_local_cluster_matches_the_oracle(world, mode):
    cl = LocalCluster(_scene(), world, density_mode=mode)
    owned0 = [s.engine.particle_num for s in cl.sims]
    assert sum(owned0) == cl.sims[0].global_particle_num
    check_against_oracle(lambda: cl.step(1), cl.dump, mode)
    owned1 = [s.engine.particle_num for s in cl.sims]
    assert sum(owned1) == sum(owned0) and owned1 != owned0      # particles migrated between slabs
    for s in cl.sims:
        s.engine.sync()
        s.engine.close()


def test_local_cluster_save_restore_replays_the_same_steps():
    cl = LocalCluster(_scene(), 2)
    cl.step(2)
    for s in cl.sims:
        s.save_state()
    cl.step(2)
    a = cl.dump()
    for s in cl.sims:
        s.restore_state()
    cl.step(2)
    b = cl.dump()
    assert np.array_equal(a["orig_id"], b["orig_id"]) and np.array_equal(a["position"], b["position"])


@pytest.mark.parametrize("p2p", ["1", "0"], ids=["p2p-halo", "nccl-halo"])
def test_two_gpus_over_torchrun(p2p):
    """one process per GPU: records either written by the pack kernel straight into the neighbour's
    receive buffer (CUDA IPC over NVLink) or sent with NCCL send/recv"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(HERE, "dist_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, TISPH_P2P_HALO=p2p))
    if res.returncode != 0 or "SHARDED-NCCL-OK" not in res.stdout:
        os.makedirs(os.path.join(os.path.dirname(HERE), "gpurun_out"), exist_ok=True)
        with open(os.path.join(os.path.dirname(HERE), "gpurun_out", f"dist_worker_failure_p2p{p2p}.log"), "w") as fh:
            fh.write(res.stdout + "\n----- stderr -----\n" + res.stderr)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "SHARDED-NCCL-OK" in res.stdout
    assert f"halo path: {'p2p' if p2p == '1' else 'nccl'}" in res.stdout


@pytest.mark.parametrize("vmode,dmode", [("reference", "reference"), ("akinci", "summed")])
def test_local_cluster_with_boundary_particles(vmode, dmode):
    """a slab of boundary particles under the moving block, cut by the slab faces: rigid points are
    distributed by x-plane, ghosts carry boundary particles (TISPH_P_HAS_BOUNDARY), akinci volumes
    and summed densities need two ghost planes"""
    import copy
    from ti_sph_b200.sharded import x_plane
    scene = copy.deepcopy(_scene())
    g = np.arange(0.28, 0.46, 0.02)
    slab = np.stack(np.meshgrid(g, [0.26, 0.28], np.arange(0.28, 0.40, 0.02), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    # ids follow the cell rows cx * gy + cy (SceneParts); gy = 25 cells of 0.04 in a unit domain
    slab = slab[np.argsort(x_plane(slab[:, 0], 0.04) * 25 + x_plane(slab[:, 1], 0.04), kind="stable")]
    scene["rigidBodies"] = [{"geometryFile": "(points)", "scale": [1, 1, 1], "translation": [0, 0, 0], "rotationAngle": 0,
                             "rotationAxis": [0, 1, 0], "color": [255, 255, 255], "velocity": [0.0, 0.0, 0.0], "density": 1000.0}]
    cl = LocalCluster(scene, 3, density_mode=dmode, volume_mode=vmode, rigid_points=[slab])
    ora = Gen2Oracle(scene, density_mode=dmode, volume_mode=vmode, boundary_points=slab)
    n = ora.n
    assert sum(s.engine.particle_num for s in cl.sims) == n
    dens0, mat0 = ora.density.copy(), ora.material.copy()
    for s in range(3):          # single-step parity, the oracle re-synced from the sharded state every step
        t = ora.step(trace=True); cl.step(1)
        d = cl.dump()
        ids = d["orig_id"]
        assert np.array_equal(ids, ora.orig)
        assert np.array_equal(d["material"], ora.material)
        check_end_state(d["position"], d["velocity"], t)
        back = np.argsort(ids)
        ora.set_state(d["position"][back], d["velocity"][back], dens0, mat0)
    for s in cl.sims:
        s.engine.sync(); s.engine.close()


def test_local_cluster_one_million_particles():
    """C3 (1 M particles) cut into 4 slabs at cell-row granularity: one step bit-identical in order, 1e-5 in fields"""
    from ti_sph_b200 import scene as sc
    scene = sc.bench_scene("C3")
    cl = LocalCluster(scene, 4)
    owned = [s.engine.particle_num for s in cl.sims]
    assert max(owned) - min(owned) <= 2 * 16 * 100 and any(s.row_lo % s.gy for s in cl.sims[1:])    # faces inside a plane
    ora = Gen2Oracle(scene)
    t = ora.step(trace=True); cl.step(1)
    d = cl.dump()
    assert np.array_equal(d["orig_id"], ora.orig)
    check_end_state(d["position"], d["velocity"], t)
    assert all(int(s.engine.get_param(K.P_STAT_FALLBACK_FORCE)) == 0 for s in cl.sims)
    for s in cl.sims:
        s.engine.sync(); s.engine.close()


def test_local_cluster_rebalancing():
    """lopsided slabs, re-balanced after two steps (tisph_row_counts + re-issued
    tisph_shard_config_rows): faces move, nobody is lost, results still follow the oracle"""
    bad_edges = [0, 8, 11, 25]
    cl = LocalCluster(_scene(), 3, edges=bad_edges)
    ora = Gen2Oracle(_scene())
    n = ora.n
    loads = []
    dens0, mat0 = ora.density.copy(), ora.material.copy()
    for s in range(5):
        if s == 2:
            new = cl.rebalance()               # faces in cell rows (25 per plane here), not necessarily whole planes
            assert new != [25 * e for e in bad_edges] and all(b - a >= 3 * 25 for a, b in zip(new, new[1:]))
        t = ora.step(trace=True); cl.step(1)
        d = cl.dump()
        ids = d["orig_id"]
        assert np.array_equal(ids, ora.orig)
        check_end_state(d["position"], d["velocity"], t)
        back = np.argsort(ids)
        ora.set_state(d["position"][back], d["velocity"][back], dens0, mat0)
        loads.append(max(sim.engine.particle_num for sim in cl.sims))
    assert loads[-1] < loads[1]
    hist = sum(sim.owned_plane_counts() for sim in cl.sims)
    assert hist.sum() == n
    for sim in cl.sims:
        sim.engine.sync(); sim.engine.close()


def test_cfl_step_of_a_sharded_run_is_agreed_by_the_ranks():
    """TISPH_P_CFL is per context: refused on a sharded one; TISPH_P_MAX_SPEED is what ShardedSim.set_cfl reduces"""
    from ti_sph_b200._capi import TisphError
    cl = LocalCluster(_scene(), 2)
    cl.step(1)
    with pytest.raises(TisphError):
        cl.sims[0].engine.set_param(K.P_CFL, 0.4)
    d = cl.dump()
    per_rank = [s.engine.get_param(K.P_MAX_SPEED) for s in cl.sims]
    assert max(per_rank) == pytest.approx(float(np.linalg.norm(d["velocity"], axis=1).max()), rel=1e-6)
    dt = cl.sims[0].set_cfl(0.4)              # (comm-less sim: its own maximum)
    assert 0 < dt <= 2e-4
    for s in cl.sims:
        s.engine.sync(); s.engine.close()
