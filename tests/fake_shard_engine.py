"""TEST INFRASTRUCTURE ONLY: a CPU stand-in for ti_sph_b200.engine.Engine that speaks the shard
protocol of include/tisph.h (shard_config / shard_pack / message_tensor / shard_append / step) with
the CPU oracle as its compute.  It lets the world_size-2 gloo tests exercise the host side of the
multi-GPU path (slab planning, count and record exchange, migration, ghost layers, dump order) in
a container without GPUs.  It is never imported by the product path.
"""
import ctypes as C

import numpy as np
import torch

from oracle.oracle import Gen2Oracle
from ti_sph_b200 import _capi as K

_FIELDS = ("x", "v", "mass", "volume", "density", "pressure", "material", "orig")


class FakeShardEngine:
    def __init__(self, config, scene):
        self.config = config
        self.h = np.float32(config.support)
        mode = {0: "reference", 1: "summed"}[config.density_mode]
        vmode = {0: "reference", 1: "akinci"}[config.volume_mode]
        self._proto = (scene, mode, vmode)
        self.x = np.zeros((0, 3), np.float32); self.v = np.zeros((0, 3), np.float32)
        for f in ("mass", "volume", "density", "pressure"):
            setattr(self, f, np.zeros(0, np.float32))
        self.material = np.zeros(0, np.int32); self.orig = np.zeros(0, np.int32)
        self.id_base = 0
        self.m_V0 = np.float32(config.m_V0)
        self.msg = {}
        self.steps = 0

    # ---- Engine surface used by ShardedSim
    def shard_config(self, plane_lo, plane_hi, ghost, left_lo, right_hi, cap):
        gy = int(self.config.grid_num[1])
        self.shard_config_rows(plane_lo * gy, plane_hi * gy, ghost, left_lo * gy if left_lo >= 0 else -1,
                               right_hi * gy if right_hi >= 0 else -1, cap)

    def shard_config_rows(self, row_lo, row_hi, ghost, left_row_lo, right_row_hi, cap):
        self.row_lo, self.row_hi, self.ghost = row_lo, row_hi, ghost
        self.has_left, self.has_right, self.cap = left_row_lo >= 0, right_row_hi >= 0, cap

    def _cells(self):
        return (self.x[:, 0] / self.h).astype(np.int32), (self.x[:, 1] / self.h).astype(np.int32)

    def set_param(self, param, value):
        if param == K.P_ID_BASE:
            self.id_base = int(value)

    def add_particles(self, pos, vel, density, pressure, material, color=None):
        n = len(pos)
        vol = np.full(n, self.m_V0, np.float32)
        new = dict(x=np.asarray(pos, np.float32), v=np.asarray(vel, np.float32), mass=vol * np.asarray(density, np.float32),
                   volume=vol, density=np.asarray(density, np.float32), pressure=np.asarray(pressure, np.float32),
                   material=np.asarray(material, np.int32), orig=self.id_base + np.arange(n, dtype=np.int32))
        for f in _FIELDS:
            setattr(self, f, np.concatenate([getattr(self, f), new[f]]))
        self.id_base += n

    @property
    def particle_num(self):
        return len(self.x)

    def _records(self, mask):
        rec = np.zeros((int(mask.sum()), 12), np.float32)
        rec[:, 0:3] = self.x[mask]; rec[:, 3] = self.mass[mask]
        rec[:, 4:7] = self.v[mask]; rec[:, 7] = self.volume[mask]
        rec[:, 8] = self.density[mask]; rec[:, 9] = self.pressure[mask]
        rec[:, 10] = self.material[mask].view(np.float32); rec[:, 11] = self.orig[mask].view(np.float32)
        return torch.from_numpy(rec)

    def shard_pack(self):
        # the neighbour needs every particle that has one of ITS cell rows within `ghost` layers in x and y
        cx, cy = self._cells()
        gx, gy, g = int(self.config.grid_num[0]), int(self.config.grid_num[1]), self.ghost
        lowest = np.maximum(cx - g, 0) * gy + np.maximum(cy - g, 0)
        highest = np.minimum(cx + g, gx - 1) * gy + np.minimum(cy + g, gy - 1)
        left = (lowest < self.row_lo) if self.has_left else np.zeros(len(cx), bool)
        right = (highest >= self.row_hi) if self.has_right else np.zeros(len(cx), bool)
        self.msg[0], self.msg[1] = self._records(left), self._records(right)
        assert len(self.msg[0]) <= self.cap and len(self.msg[1]) <= self.cap
        return len(self.msg[0]), len(self.msg[1])

    def message_tensor(self, which, n):
        if which >= 2:
            self.msg[which] = torch.zeros((n, 12), dtype=torch.float32)
        return self.msg[which][:n]

    def shard_append(self, nl, nr):
        for which, n in ((2, nl), (3, nr)):
            if not n:
                continue
            rec = self.msg[which][:n].numpy()
            new = dict(x=rec[:, 0:3], mass=rec[:, 3], v=rec[:, 4:7], volume=rec[:, 7], density=rec[:, 8],
                       pressure=rec[:, 9], material=rec[:, 10].copy().view(np.int32),
                       orig=rec[:, 11].copy().view(np.int32))
            for f in _FIELDS:
                setattr(self, f, np.concatenate([getattr(self, f), np.ascontiguousarray(new[f])]))

    def step(self, nsteps=1):
        assert nsteps == 1
        scene, mode, vmode = self._proto
        o = self._blank_oracle(scene, mode, vmode)
        order = np.argsort(self.orig, kind="stable")       # intra-cell order = by original id
        o.set_state(self.x[order], self.v[order], self.density[order], self.material[order],
                    pressure=self.pressure[order], volume=self.volume[order], mass=self.mass[order])
        o.orig = self.orig[order].copy()
        o.step()
        # owned = cell row at sort time in [row_lo, row_hi); everything else was a ghost
        own = o.keys // int(self.config.grid_num[2])                     # cell row cx * gy + cy
        own = (own >= self.row_lo) & (own < self.row_hi)
        self.x, self.v = o.x[own], o.v[own]
        self.mass, self.volume = o.mass[own], o.volume[own]
        self.density, self.pressure = o.density[own], o.pressure[own]
        self.material, self.orig = o.material[own], o.orig[own]
        self.steps += 1

    def _blank_oracle(self, scene, mode, vmode):
        """an oracle with the scene's constants and no particles of its own"""
        blank = dict(scene)
        blank["fluidBlocks"] = [dict(scene["fluidBlocks"][0], start=[0.5, 0.5, 0.5], end=[0.505, 0.505, 0.505])]
        blank["rigidBodies"] = []
        return Gen2Oracle(blank, density_mode=mode, volume_mode=vmode)

    def plane_counts(self):
        cx = (self.x[:, 0] / self.h).astype(np.int32)
        return np.bincount(cx, minlength=int(self.config.grid_num[0])).astype(np.int32)

    def row_counts(self):
        cx, cy = self._cells()
        gx, gy = int(self.config.grid_num[0]), int(self.config.grid_num[1])
        return np.bincount(cx * gy + cy, minlength=gx * gy).astype(np.int32)

    def download(self, field, out=None):
        return {K.F_X: self.x, K.F_V: self.v, K.F_MATERIAL: self.material, K.F_ORIG_ID: self.orig,
                K.F_DENSITY: self.density, K.F_PRESSURE: self.pressure}[field].copy()

    def save_state(self):
        self._snap = {f: getattr(self, f).copy() for f in _FIELDS}

    def restore_state(self):
        for f in _FIELDS:
            setattr(self, f, self._snap[f].copy())
