"""Host side of the multi-GPU path on CPU: slab planning and a world_size-2 (and 3) run of
ShardedSim over torch.distributed/gloo, with tests/fake_shard_engine.py (the CPU oracle behind the
shard protocol) in place of the CUDA engine.  The sharded result must equal a single global
oracle: bit-exact after one step (same neighbour order), within fp32 round-off afterwards."""
import copy
import os
import socket
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.oracle import Gen2Oracle
from ti_sph_b200 import scene as sc
from ti_sph_b200.sharded import SceneParts, ShardedSim, TorchDistComm, plan_slabs, x_plane
from util import small_scene

HERE = os.path.dirname(os.path.abspath(__file__))


def test_plan_slabs_balances_and_respects_min_thickness():
    counts = np.zeros(125, np.int64)
    counts[7:33] = 6400
    for world in (1, 2, 3, 4, 8):
        e = plan_slabs(counts, world)
        assert e[0] == 0 and e[-1] == 125 and len(e) == world + 1
        assert all(b - a >= 3 for a, b in zip(e, e[1:]))
        loads = [counts[a:b].sum() for a, b in zip(e, e[1:])]
        assert sum(loads) == counts.sum()
        assert max(loads) <= counts.sum() / world + 6400
    with pytest.raises(ValueError):
        plan_slabs(np.ones(5), 2)


def test_scene_parts_cover_the_scene_exactly_once():
    scene = sc.bench_scene("C2")
    parts = SceneParts(scene)
    assert parts.total == 195300
    counts = parts.plane_counts()
    assert counts.sum() == 195300
    full = sc.cube_positions([0.3, 0.1, 0.7], [0.7, 0.9, 0.3], 0.01, 3)
    assert np.array_equal(np.bincount(x_plane(full[:, 0], 0.04), minlength=125), counts)
    edges = plan_slabs(counts, 3)
    seen = np.zeros(parts.total, bool)
    for a, b in zip(edges, edges[1:]):
        for id0, pos, vel, dens, mat in parts.slab_particles(a, b):
            assert np.array_equal(pos, full[id0:id0 + len(pos)])
            assert not seen[id0:id0 + len(pos)].any()
            seen[id0:id0 + len(pos)] = True
            assert np.all((x_plane(pos[:, 0], 0.04) >= a) & (x_plane(pos[:, 0], 0.04) < b))
    assert seen.all()


def test_row_granular_slabs_cover_the_scene_exactly_once_and_balance_to_a_row():
    """slabs cut at (x-plane, y-row) granularity: contiguous id chunks, every particle once, loads equal to
    within one cell row of particles"""
    scene = sc.bench_scene("C2")
    parts = SceneParts(scene)
    gy = int(parts.grid_num[1])
    rows = parts.row_counts()
    assert rows.sum() == parts.total and np.array_equal(rows.reshape(-1, gy).sum(axis=1), parts.plane_counts())
    full = sc.cube_positions([0.3, 0.1, 0.7], [0.7, 0.9, 0.3], 0.01, 3)
    assert np.array_equal(np.bincount(parts.row_of(full), minlength=len(rows)), rows)
    for world in (2, 3, 5):
        edges = plan_slabs(rows, world, min_planes=3 * gy)
        seen = np.zeros(parts.total, bool)
        loads = []
        for a, b in zip(edges, edges[1:]):
            n = 0
            for id0, pos, vel, dens, mat in parts.slab_particles_rows(a, b):
                assert np.array_equal(pos, full[id0:id0 + len(pos)])
                assert not seen[id0:id0 + len(pos)].any()
                seen[id0:id0 + len(pos)] = True
                r = parts.row_of(pos)
                assert np.all((r >= a) & (r < b))
                n += len(pos)
            loads.append(n)
        assert seen.all()
        assert max(loads) - min(loads) <= 2 * rows.max()          # whole planes: up to a plane (7 440 here) apart
        assert any(e % gy for e in edges[1:-1])                   # at least one face inside a plane


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _scene():
    # 13 x 8 x 8 block, moving in +x fast enough that particles cross the slab face within 3 steps
    return small_scene(start=(0.30, 0.30, 0.30), end=(0.43, 0.38, 0.38), velocity=(25.0, -1.0, 3.0),
                       domain_end=(1.0, 1.0, 1.0))


def _worker(rank, world, port, mode, steps, outdir, edges=None, rebalance_at=None):
    sys.path.insert(0, HERE)
    from fake_shard_engine import FakeShardEngine
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    scene = _scene()
    sim = ShardedSim(scene, rank, world, comm=TorchDistComm(device="cpu"), density_mode=mode,
                     engine_factory=lambda cfg: FakeShardEngine(cfg, scene), edges=edges)
    owned = [sim.engine.particle_num]
    dumps = []
    for st in range(steps):
        if rebalance_at is not None and st == rebalance_at:
            sim.rebalance()
        sim.step()
        owned.append(sim.engine.particle_num)
        dumps.append(sim.dump())
    if rank == 0:
        np.savez(os.path.join(outdir, "out.npz"), edges=np.array(sim.edges),
                 **{f"{k}{s}": v for s, d in enumerate(dumps) for k, v in d.items()})
    np.save(os.path.join(outdir, f"owned{rank}.npy"), np.array(owned))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,mode", [(2, "reference"), (2, "summed"), (3, "reference")])
def test_sharded_gloo_run_matches_the_global_oracle(world, mode):
    steps = 4
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_worker, args=(world, _free_port(), mode, steps, tmp), nprocs=world, join=True)
        out = np.load(os.path.join(tmp, "out.npz"))
        owned = [np.load(os.path.join(tmp, f"owned{r}.npy")) for r in range(world)]
    ora = Gen2Oracle(_scene(), density_mode=mode)
    n = ora.n
    assert all(sum(o[s] for o in owned) == n for s in range(steps + 1))      # nobody lost or duplicated
    assert any(o[0] != o[-1] for o in owned), "the scene is meant to make particles migrate"
    for s in range(steps):
        ora.step()
        ids = out[f"orig_id{s}"]
        assert np.array_equal(np.sort(ids), np.arange(n))
        assert np.array_equal(out[f"material{s}"], ora.material) and np.all(out[f"color{s}"] == 0x111111)
        if s == 0:     # identical neighbour order: bit-exact, and the same (cell-sorted) dump order
            assert np.array_equal(ids, ora.orig)
            assert np.array_equal(out[f"position{s}"], ora.x)
            assert np.array_equal(out[f"velocity{s}"], ora.v)
        else:          # intra-cell order differs (by id vs by previous position): fp32 round-off
            inv = np.empty(n, np.int64); inv[ora.orig] = np.arange(n)
            sel = inv[ids]
            assert np.allclose(out[f"position{s}"], ora.x[sel], rtol=0, atol=2e-6)
            assert np.allclose(out[f"velocity{s}"], ora.v[sel], rtol=2e-5, atol=2e-4)


def test_rebalancing_moves_the_faces_and_loses_nobody():
    """start from deliberately lopsided slabs, re-balance after two steps: the faces move towards
    equal loads (at most to a neighbouring old face per call), results still match the oracle"""
    steps, world = 5, 3
    bad_edges = [0, 8, 11, 25]                 # the block spans planes 7..10: rank 0 gets 1/4, rank 2 nothing
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_worker, args=(world, _free_port(), "reference", steps, tmp, bad_edges, 2), nprocs=world, join=True)
        out = np.load(os.path.join(tmp, "out.npz"))
        owned = [np.load(os.path.join(tmp, f"owned{r}.npy")) for r in range(world)]
    assert list(out["edges"]) != [25 * e for e in bad_edges]           # (faces are kept in cell rows: 25 per plane here)
    ora = Gen2Oracle(_scene())
    n = ora.n
    assert all(sum(o[s] for o in owned) == n for s in range(steps + 1))
    before = max(o[2] for o in owned) / (n / world)
    after = max(o[-1] for o in owned) / (n / world)
    assert after < before                                         # better balanced than before
    for s in range(steps):
        ora.step()
        ids = out[f"orig_id{s}"]
        assert np.array_equal(np.sort(ids), np.arange(n))
        inv = np.empty(n, np.int64); inv[ora.orig] = np.arange(n)
        assert np.allclose(out[f"position{s}"], ora.x[inv[ids]], rtol=0, atol=2e-6)
