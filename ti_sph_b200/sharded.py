"""Slab-sharded WCSPH step: one process (and one Engine) per GPU, x-slabs, halo exchange + migration.

The reference is single-device; this is the multi-GPU row of SURVEY.md 8(e).
The cell key is x-major (partice_systemv4.py:98-100), so a rank that owns the cell ROWS
row = cx * gy + cy in [row_lo, row_hi) -- whole x-planes plus a partial plane at either end -- owns one
contiguous range of the sorted particle arrays.  A halo message is a plain run of 48-byte records, packed
by one kernel (csrc/tisph_shard.cuh) and written peer-to-peer over NVLink (or sent with NCCL send/recv)
-- there is no all-reduce or any other collective on the data path.  Per step and rank:

    pack  (owned particles with a cell of the neighbour within `ghost` cell layers, or inside the
           neighbour's rows = migrants)
    -> exchange counts with both neighbours -> exchange records
    -> append -> bin/scan/sort everything -> density on the owned rows + one layer -> forces/advect on the owned rows

`ghost` is 1 when a particle's density needs no neighbour data (density_mode "reference" with
volume_mode "reference": rho = mass W(0), wcsphv2.py:32-34), else 2 -- the densities of the first
ghost plane are then recomputed locally instead of being exchanged a second time.

Slab edges are chosen once from the per-row particle histogram so that every rank starts with the
same number of particles to within one cell row (`plan_slabs`; round 1 cut at whole planes: 4 % imbalance
at 8 ranks of the 16 M-particle scene).  Intra-cell order is by original particle id
(the single-GPU engine orders by array position like the serial reference; array positions are
rank-local here).  Concatenating the ranks' owned particles in rank order gives the global
cell-sorted order that ParticleSystemV4.dump() of a single engine returns.
"""
from functools import reduce

import numpy as np

from . import _capi as K
from . import scene as sc

LEFT, RIGHT = 0, 1
SEND_LEFT, SEND_RIGHT, RECV_LEFT, RECV_RIGHT = 0, 1, 2, 3
FLUID_COLOR = 0x111111            # partice_systemv4.py:144


# ------------------------------------------------------------------------------ slab planning
def x_plane(x, h):
    """cell plane of x-coordinates: IEEE f32 division, truncation (partice_systemv4.py:86-92)."""
    return (np.asarray(x, np.float32) / np.float32(h)).astype(np.int32)


def plan_slabs(plane_counts, world, min_planes=3):
    """Edges e[0]=0 < e[1] < ... < e[world]=len(plane_counts): slab k owns planes [e[k], e[k+1]).
    Cuts are put where the cumulative particle count crosses k/world of the total; every slab is
    at least `min_planes` thick (the ghost-layer argument of the module docstring needs 3).
    Works the same on a per-ROW histogram (units = cell rows, min_planes = 3 * gy)."""
    counts = np.asarray(plane_counts, np.int64)
    gx = len(counts)
    if world * min_planes > gx:
        raise ValueError(f"{gx} cell planes cannot be split into {world} slabs of >= {min_planes} planes")
    cum = np.concatenate([[0], np.cumsum(counts)])
    total = cum[-1]
    edges = [0]
    for k in range(1, world):
        target = total * k / world
        e = int(np.searchsorted(cum, target, side="left"))
        # the cut that leaves the cumulative count closest to the target
        if e > 0 and abs(cum[e - 1] - target) <= abs(cum[min(e, gx)] - target):
            e -= 1
        e = max(e, edges[-1] + min_planes)
        e = min(e, gx - (world - k) * min_planes)
        edges.append(e)
    edges.append(gx)
    return edges


class SceneParts:
    """The scene's particles in the reference's insertion order (rigid bodies first, then the fluid
    blocks: partice_systemv4.py:102-146) without materialising them: per-plane histogram and
    per-slab generation."""

    def __init__(self, scene, rigid_points=()):
        cfg = scene["configuration"]
        self.cfg = cfg
        self.dim = cfg["dim"]
        assert self.dim == 3
        self.r = cfg["particleRadius"]
        self.h = 4.0 * self.r
        size = np.array(cfg["domainEnd"]) - np.array(cfg["domainStart"])
        self.grid_num = np.ceil(size / self.h).astype(np.int32)
        self.parts = []          # (kind, id0, data)
        nid = 0
        for body, pts in zip(scene["rigidBodies"], rigid_points):
            pts = np.asarray(pts, np.float32)
            order = np.argsort(self.row_of(pts), kind="stable")             # ids follow the cell rows
            self.parts.append(("rigid", nid, (pts[order], body)))
            nid += len(pts)
        for blk in scene["fluidBlocks"]:
            start, end = blk["start"], blk["end"]
            axes = [np.arange(start[i], start[i] + (end[i] - start[i]), self.r) for i in range(3)]
            pre = sc.cube_particle_num(start, end, self.r, 3)
            n = reduce(lambda a, b: a * b, [len(a) for a in axes])
            if pre != n:
                raise RuntimeError("particle count pre-pass != add_cube count (reference quirk Q10)")
            self.parts.append(("fluid", nid, (axes, blk)))
            nid += n
        self.total = nid

    def row_of(self, pts):
        """cell row cx * gy + cy of positions (n, 3)"""
        return x_plane(pts[:, 0], self.h).astype(np.int64) * int(self.grid_num[1]) + x_plane(pts[:, 1], self.h)

    def row_counts(self):
        """particles per cell row (cx, cy): gx * gy values"""
        gx, gy = int(self.grid_num[0]), int(self.grid_num[1])
        counts = np.zeros(gx * gy, np.int64)
        for kind, _, data in self.parts:
            if kind == "rigid":
                np.add.at(counts, self.row_of(data[0]), 1)
            else:
                axes, _ = data
                px, py = x_plane(axes[0].astype(np.float32), self.h), x_plane(axes[1].astype(np.float32), self.h)
                np.add.at(counts, (px[:, None].astype(np.int64) * gy + py[None, :]).ravel(), len(axes[2]))
        return counts

    def plane_counts(self):
        counts = np.zeros(int(self.grid_num[0]), np.int64)
        for kind, _, data in self.parts:
            if kind == "rigid":
                np.add.at(counts, x_plane(data[0][:, 0], self.h), 1)
            else:
                axes, _ = data
                np.add.at(counts, x_plane(axes[0].astype(np.float32), self.h), len(axes[1]) * len(axes[2]))
        return counts

    def slab_particles(self, lo, hi):
        """[(id0, pos, vel, density, material)] of the particles whose x-plane is in [lo, hi);
        every chunk has contiguous original ids starting at id0."""
        gy = int(self.grid_num[1])
        return self.slab_particles_rows(lo * gy, hi * gy)

    def slab_particles_rows(self, row_lo, row_hi):
        """the same for the cell rows cx * gy + cy in [row_lo, row_hi).  A fluid block is a lattice with ids
        x-major, then y, then z: the x-layers of a partial plane contribute one contiguous id range each."""
        gy = int(self.grid_num[1])
        out = []
        for kind, id0, data in self.parts:
            if kind == "rigid":
                pts, body = data
                rows = self.row_of(pts)                                # non-decreasing (sorted in __init__)
                a, b = int(np.searchsorted(rows, row_lo, "left")), int(np.searchsorted(rows, row_hi, "left"))
                if b > a:
                    dens = body.get("density")
                    out.append((id0 + a, pts[a:b], np.tile(np.array(body["velocity"], np.float32), (b - a, 1)),
                                np.full(b - a, dens if dens is not None else 1000.0, np.float32),
                                np.zeros(b - a, np.int32)))
                continue
            axes, blk = data
            px = x_plane(axes[0].astype(np.float32), self.h).astype(np.int64)       # non-decreasing in x
            py = x_plane(axes[1].astype(np.float32), self.h).astype(np.int64)       # ... and in y
            ny, nz = len(axes[1]), len(axes[2])
            dens = blk["density"]

            def chunk(ia, ib, ja, jb):
                """x-layers [ia, ib) x y-layers [ja, jb) x all z: contiguous ids when ja, jb cover all y or ib = ia + 1"""
                pos = np.array(np.meshgrid(axes[0][ia:ib], axes[1][ja:jb], axes[2], indexing="ij"), dtype=np.float32)
                pos = np.ascontiguousarray(pos.reshape(3, -1).T)
                out.append((id0 + (ia * ny + ja) * nz, pos, np.full(pos.shape, blk["velocity"], dtype=np.float32),
                            np.full(len(pos), dens if dens is not None else 1000.0, np.float32),
                            np.ones(len(pos), np.int32)))

            # x-layers whose whole plane is inside: one chunk; the others layer by layer with their y-range
            full = (px * gy >= row_lo) & ((px + 1) * gy <= row_hi)
            ia = 0
            while ia < len(px):
                if full[ia]:
                    ib = ia
                    while ib < len(px) and full[ib]:
                        ib += 1
                    chunk(ia, ib, 0, ny)
                    ia = ib
                    continue
                rows = px[ia] * gy + py
                ja, jb = int(np.searchsorted(rows, row_lo, "left")), int(np.searchsorted(rows, row_hi, "left"))
                if jb > ja:
                    chunk(ia, ia + 1, ja, jb)
                ia += 1
        return out

    def color_of(self, ids):
        """ps.color of the particles with these original ids (fluid: 0x111111 in every lane, :144)."""
        ids = np.asarray(ids)
        col = np.zeros((len(ids), 3), np.int32)
        for kind, id0, data in self.parts:
            n = len(data[0][0]) if kind == "rigid" else reduce(lambda a, b: a * b, [len(a) for a in data[0]])
            m = (ids >= id0) & (ids < id0 + n)
            if kind == "fluid":
                col[m] = FLUID_COLOR
            else:
                col[m] = np.array(data[1].get("color", [0, 0, 0]), np.int32)
        return col


# ------------------------------------------------------------------------------ communication
class TorchDistComm:
    """Neighbour exchange over torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self, group=None, device=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = device

    def exchange(self, rank, send_left, send_right, recv_left, recv_right):
        """send_* / recv_* are tensors or None; posts everything at once and waits."""
        dist = self.dist
        ops = []
        if recv_left is not None:
            ops.append(dist.P2POp(dist.irecv, recv_left, rank - 1, self.group))
        if recv_right is not None:
            ops.append(dist.P2POp(dist.irecv, recv_right, rank + 1, self.group))
        if send_left is not None:
            ops.append(dist.P2POp(dist.isend, send_left, rank - 1, self.group))
        if send_right is not None:
            ops.append(dist.P2POp(dist.isend, send_right, rank + 1, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    def all_gather_objects(self, obj):
        out = [None] * self.world
        self.dist.all_gather_object(out, obj, group=self.group)
        return out


class ShmCounts:
    """Halo record counts exchanged through a POSIX shared-memory segment instead of a device
    round trip: every rank of ONE node publishes (step, to_left, to_right) and polls its
    neighbours' slots.  Costs microseconds and no stream synchronisation; used when all ranks
    live on this node (LOCAL_WORLD_SIZE == WORLD_SIZE), the NCCL path otherwise."""

    def __init__(self, comm, rank, world):
        import mmap
        import os
        import tempfile
        self.rank, self.world = rank, world
        size = world * 64
        # A plain file in /dev/shm mapped by every rank (multiprocessing.shared_memory would hand the segment to
        # Python's resource tracker, which unlinks attached segments at exit and warns about "leaks").  Rank 0
        # removes the name as soon as everybody has mapped it, so nothing is left behind even after a crash.
        path = None
        if rank == 0:
            fd, path = tempfile.mkstemp(prefix="tisph_counts_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
            os.ftruncate(fd, size)
        path = comm.all_gather_objects(path)[0]
        if rank != 0:
            fd = os.open(path, os.O_RDWR)
        self.mm = mmap.mmap(fd, size)
        os.close(fd)
        # two slots per rank, used alternately: a rank can be at most one step ahead of a neighbour
        # (it needs the neighbour's counts of step k+1 before it can publish step k+2)
        self.arr = np.ndarray((world, 2, 4), dtype=np.int64, buffer=self.mm)
        self.step = 0
        comm.all_gather_objects(None)          # everybody attached before anybody publishes ...
        if rank == 0:
            os.unlink(path)                    # ... and before the name goes away

    def exchange(self, nl, nr, has_left, has_right):
        import time
        self.step += 1
        slot = self.step & 1
        me = self.arr[self.rank, slot]
        me[1], me[2] = nl, nr
        me[0] = self.step                      # publish last (x86: stores are not reordered)
        ml = mr = 0
        t0 = time.perf_counter()
        for nb, col, flag in ((self.rank - 1, 2, has_left), (self.rank + 1, 1, has_right)):
            if not flag:
                continue
            while self.arr[nb, slot, 0] != self.step:
                if time.perf_counter() - t0 > 120.0:
                    raise RuntimeError("halo count exchange timed out (a neighbour rank died?)")
            if nb < self.rank:
                ml = int(self.arr[nb, slot, col])
            else:
                mr = int(self.arr[nb, slot, col])
        return ml, mr

    def close(self):
        self.arr = None
        try:
            self.mm.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------ one rank
class ShardedSim:
    """One rank's slab of a gen-2 (ParticleSystemV4 + WCSPHV2) simulation."""

    def __init__(self, scene, rank, world, comm=None, density_mode="reference", volume_mode="reference",
                 device=0, rigid_points=(), engine_factory=None, capacity_factor=1.6, edges=None,
                 message_capacity=None, row_edges=None):
        """edges: slab faces in x-planes (whole planes); row_edges: in cell rows cx * gy + cy.  Default: equal
        particle counts to within one cell row."""
        import torch
        self.torch = torch
        self.rank, self.world, self.comm = rank, world, comm
        self.scene = scene
        if scene["rigidBodies"] and len(rigid_points) != len(scene["rigidBodies"]):
            if len(rigid_points):
                raise ValueError(f"{len(scene['rigidBodies'])} rigid bodies but {len(rigid_points)} point sets")
            # like ParticleSystemV4.load_rigid_body: every rank voxelises the bodies itself (same result on every
            # rank).  Within a body the original ids follow the x-planes (SceneParts), not the sampler's order.
            from . import mesh
            pitch = 2 * scene["configuration"]["particleRadius"]
            rigid_points = [mesh.sample_rigid_body(dict(body), pitch, device=device) for body in scene["rigidBodies"]]
        self.parts = SceneParts(scene, rigid_points)
        self.ghost = 1 if (density_mode == "reference" and volume_mode == "reference") else 2
        counts = self.parts.plane_counts()
        self.gy = gy = int(self.parts.grid_num[1])
        if row_edges is not None:
            self.edges = [int(e) for e in row_edges]
        elif edges is not None:
            self.edges = [int(e) * gy for e in edges]
        else:
            self.edges = plan_slabs(self.parts.row_counts(), world, min_planes=3 * gy)
        self.row_lo, self.row_hi = self.edges[rank], self.edges[rank + 1]
        self.plane_lo, self.plane_hi = self.row_lo // gy, -(-self.row_hi // gy)     # the planes the slab touches
        self.has_left, self.has_right = rank > 0, rank < world - 1
        chunks = self.parts.slab_particles_rows(self.row_lo, self.row_hi)
        n_own = sum(len(c[1]) for c in chunks)
        halo = int(counts[max(self.plane_lo - self.ghost, 0):self.plane_lo + 1].sum() +
                   counts[max(self.plane_hi - 1, 0):self.plane_hi + self.ghost].sum())
        peak_plane = int(counts.max())
        if message_capacity is None:
            message_capacity = max(1024, 2 * (self.ghost + 1) * peak_plane)
        capacity = int(capacity_factor * (n_own + halo)) + 2 * message_capacity + 1024
        cfg = sc.gen2_config(scene["configuration"], capacity, device=device,
                             density_mode={"reference": 0, "summed": 1}[density_mode],
                             volume_mode={"reference": 0, "akinci": 1}[volume_mode])
        if engine_factory is None:
            from .engine import Engine
            engine_factory = Engine
        self.engine = eng = engine_factory(cfg)
        eng.shard_config_rows(self.row_lo, self.row_hi, self.ghost,
                              self.edges[rank - 1] if self.has_left else -1,
                              self.edges[rank + 2] if self.has_right else -1, message_capacity)
        if scene["rigidBodies"]:
            eng.set_param(K.P_HAS_BOUNDARY, 1)
        for id0, pos, vel, dens, mat in chunks:
            eng.set_param(K.P_ID_BASE, id0)
            eng.add_particles(pos, vel, dens, np.zeros(len(pos), np.float32), mat)
        self._message_capacity = message_capacity
        self.global_particle_num = self.parts.total
        self.initial_owned = n_own
        self._counts_dev = None
        self._shm = None
        # The engine and the NCCL exchanges must be ordered on ONE stream: the exchange's wait() only
        # blocks the stream that is current when it is posted.  (Handing the engine torch's legacy
        # default stream would not do: its handle is 0, which tisph_set_stream reads as "own stream".)
        self.stream = None
        if comm is not None and str(getattr(comm, "device", "cpu")).startswith("cuda"):
            self.stream = torch.cuda.Stream(device=comm.device)
            eng.set_stream(self.stream.cuda_stream)
        import os
        if comm is not None and world > 1 and os.environ.get("TISPH_SHM_COUNTS", "1") != "0" and \
                os.environ.get("LOCAL_WORLD_SIZE", str(world)) == str(world):
            self._shm = ShmCounts(comm, rank, world)
        # Peer-to-peer halo: with every rank on this node and CUDA IPC available, the pack kernel writes
        # the records straight into the neighbour's receive buffer over NVLink; the host only passes the
        # two counts (ShmCounts).  All ranks must agree, otherwise everybody stays on NCCL send/recv.
        self.p2p = False
        if self._shm is not None and self.stream is not None and engine_factory.__name__ == "Engine" and \
                os.environ.get("TISPH_P2P_HALO", "1") != "0":
            handles = comm.all_gather_objects(eng.shard_ipc_export())
            ok = True
            try:
                if self.has_left:
                    eng.shard_ipc_connect(0, handles[rank - 1])
                if self.has_right:
                    eng.shard_ipc_connect(1, handles[rank + 1])
            except Exception:
                ok = False
            self.p2p = all(comm.all_gather_objects(ok))
            if not self.p2p:
                eng.shard_ipc_disconnect()
        self.profile = {} if os.environ.get("TISPH_SHARD_PROFILE") else None

    # -- the phases of one step (LocalCluster drives them for several ranks in one process) -----
    def pack(self):
        self._nsend = self.engine.shard_pack()
        return self._nsend

    def exchange(self):
        """counts, then records, with both neighbours (torch.distributed P2P)."""
        if self.stream is not None and self.torch.cuda.current_stream(self.stream.device) != self.stream:
            with self.torch.cuda.stream(self.stream):
                return self._exchange()
        return self._exchange()

    def _exchange(self):
        torch, eng, comm = self.torch, self.engine, self.comm
        nl, nr = self._nsend
        if self._counts_dev is None:          # persistent count buffers: [to_left, to_right], [from_left, from_right]
            self._counts_dev = (torch.zeros(2, dtype=torch.int32, device=comm.device),
                                torch.zeros(2, dtype=torch.int32, device=comm.device),
                                torch.zeros(2, dtype=torch.int32).pin_memory() if str(comm.device) != "cpu"
                                else torch.zeros(2, dtype=torch.int32))
        if self._shm is not None:
            ml, mr = self._shm.exchange(nl, nr, self.has_left, self.has_right)
        else:
            snd, rcv, host = self._counts_dev
            host[0], host[1] = nl, nr
            snd.copy_(host, non_blocking=True)
            comm.exchange(self.rank, snd[0:1] if self.has_left else None, snd[1:2] if self.has_right else None,
                          rcv[0:1] if self.has_left else None, rcv[1:2] if self.has_right else None)
            ml, mr = rcv.tolist()                  # one synchronisation for both counts
            ml = ml if self.has_left else 0
            mr = mr if self.has_right else 0
        if self.p2p:                           # the records are already in place (written by the neighbours' packs)
            self._nrecv = (ml, mr)
            return self._nrecv
        comm.exchange(self.rank,
                      eng.message_tensor(SEND_LEFT, nl) if self.has_left and nl else None,
                      eng.message_tensor(SEND_RIGHT, nr) if self.has_right and nr else None,
                      eng.message_tensor(RECV_LEFT, ml) if ml else None,
                      eng.message_tensor(RECV_RIGHT, mr) if mr else None)
        self._nrecv = (ml, mr)
        return self._nrecv

    def compute(self):
        self.engine.shard_append(*self._nrecv)
        self.engine.step(1)

    def step(self, nsteps=1):
        if self.profile is None:
            for _ in range(nsteps):
                self.pack()
                self.exchange()
                self.compute()
            return
        import time
        for _ in range(nsteps):            # host-side phase times (TISPH_SHARD_PROFILE=1); compute is asynchronous
            t0 = time.perf_counter(); self.pack()
            t1 = time.perf_counter(); self.exchange()
            t2 = time.perf_counter(); self.compute()
            t3 = time.perf_counter()
            for k, v in (("pack", t1 - t0), ("exchange", t2 - t1), ("compute_issue", t3 - t2), ("steps", 1)):
                self.profile[k] = self.profile.get(k, 0.0) + v

    def set_cfl(self, cfl):
        """One CFL step for all ranks (extension; TISPH_P_CFL itself is per context and refused on a sharded
        one): dt = min(dt0, cfl h / (c_s + max |v|)) with the maximum taken over every rank.  Collective;
        call between steps, as often as the step should follow the flow."""
        vmax = max(self.comm.all_gather_objects(self.engine.get_param(K.P_MAX_SPEED))) if self.comm else \
            self.engine.get_param(K.P_MAX_SPEED)
        cfg = self.scene["configuration"]
        dt0 = getattr(self, "_dt0", None)
        if dt0 is None:
            dt0 = self._dt0 = self.engine.get_param(K.P_DT)
        dt = min(dt0, cfl * 4.0 * cfg["particleRadius"] / (cfg["c_s"] + vmax))
        self.engine.set_param(K.P_DT, dt)
        return dt

    # -- re-balancing --------------------------------------------------------------------------
    def owned_row_counts(self):
        """this rank's contribution to the global per-row histogram (its own cell rows only)"""
        counts = np.asarray(self.engine.row_counts(), np.int64)
        mine = np.zeros_like(counts)
        mine[self.row_lo:self.row_hi] = counts[self.row_lo:self.row_hi]
        return mine

    def owned_plane_counts(self):
        """... summed over the rows of every x-plane"""
        return self.owned_row_counts().reshape(-1, self.gy).sum(axis=1)

    def apply_histogram(self, row_counts):
        """Move the slab faces to where `row_counts` (global per-row histogram, identical on all ranks; a per-PLANE
        histogram is accepted too and cuts at whole planes) says the load is balanced.  A face moves at most to
        the old position of a neighbouring face, so that every particle's new owner is the old owner or one of
        its neighbours and the next pack can carry it there.  Must be called by all ranks between the same two
        steps."""
        row_counts = np.asarray(row_counts, np.int64)
        gy = self.gy
        if len(row_counts) == int(self.parts.grid_num[0]):            # a per-plane histogram
            want = [e * gy for e in plan_slabs(row_counts, self.world)]
        else:
            want = plan_slabs(row_counts, self.world, min_planes=3 * gy)
        old = self.edges
        new = [0]
        for k in range(1, self.world):
            e = min(max(want[k], old[k - 1]), old[k + 1])
            e = max(e, new[-1] + 3 * gy)
            e = min(e, old[-1] - 3 * gy * (self.world - k))
            new.append(int(e))
        new.append(old[-1])
        self.edges = new
        r = self.rank
        self.row_lo, self.row_hi = new[r], new[r + 1]
        self.plane_lo, self.plane_hi = self.row_lo // gy, -(-self.row_hi // gy)
        self.engine.shard_config_rows(self.row_lo, self.row_hi, self.ghost,
                                      new[r - 1] if self.has_left else -1, new[r + 2] if self.has_right else -1,
                                      self._message_capacity)
        return new

    def rebalance(self):
        """collective: gather the histogram, move the faces (see apply_histogram)"""
        total = sum(self.comm.all_gather_objects(self.owned_row_counts()))
        return self.apply_histogram(total)

    def close(self):
        """release the engine, the peer mappings and (rank 0: unlink) the shared-memory count segment"""
        if self._shm is not None:
            self._shm.close()
            self._shm = None
        eng, self.engine = self.engine, None
        if eng is not None:
            try:
                eng.sync()
                if self.p2p:
                    eng.shard_ipc_disconnect()
            finally:
                eng.close()

    # -- state -------------------------------------------------------------------------------
    def save_state(self):
        self.engine.save_state()

    def restore_state(self):
        self.engine.restore_state()

    def upload_xv(self, pos, vel):
        """overwrite x, v of this rank's owned particles (order of dump_local) from host arrays"""
        self.engine.upload_xv(pos, vel)

    def upload_xv_async(self, pos, vel):
        self.engine.upload_xv_async(pos, vel)

    def dump_local_async(self, out):
        """dump_local without waiting: starts filling out['position'|'velocity'|'material'|'orig_id'] (preallocated,
        sized for this rank's owned particles); complete after dump_wait()"""
        self.engine.dump_async(out.get("position"), out.get("velocity"), out.get("material"), None, out.get("orig_id"))
        return out

    def dump_wait(self):
        self.engine.dump_wait()

    def dump_local(self, out=None, color=True):
        """this rank's owned particles, in sorted order (keys of dump() + 'orig_id'); `out` may hold
        preallocated (pinned) arrays for 'position', 'velocity', 'material', 'orig_id'.  The colour is a host-side
        function of the original id (a sharded engine does not carry it); color=False skips it."""
        e, out = self.engine, out or {}
        ids = e.download(K.F_ORIG_ID, out.get("orig_id"))
        d = {"position": e.download(K.F_X, out.get("position")), "velocity": e.download(K.F_V, out.get("velocity")),
             "material": e.download(K.F_MATERIAL, out.get("material")), "orig_id": ids}
        if color:
            d["color"] = self.parts.color_of(ids)
        return d

    def dump(self):
        """ParticleSystemV4.dump() of the whole simulation on every rank: the ranks' owned
        particles concatenated in rank order (= global cell-sorted order)."""
        parts = self.comm.all_gather_objects(self.dump_local())
        return {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}


# ------------------------------------------------------------------------------ in-process cluster
class LocalCluster:
    """All ranks of a sharded run inside one process (engines may share one GPU).  Used by the
    single-GPU parity tests of the shard kernels; messages are device-to-device copies."""

    def __init__(self, scene, world, **kw):
        self.world = world
        self.sims = [ShardedSim(scene, r, world, comm=None, **kw) for r in range(world)]

    def step(self, nsteps=1):
        sims = self.sims
        for _ in range(nsteps):
            counts = [s.pack() for s in sims]             # pack synchronises each engine's stream
            for r, s in enumerate(sims):
                ml = counts[r - 1][1] if r > 0 else 0
                mr = counts[r + 1][0] if r < self.world - 1 else 0
                if ml:
                    s.engine.message_tensor(RECV_LEFT, ml).copy_(sims[r - 1].engine.message_tensor(SEND_RIGHT, ml))
                if mr:
                    s.engine.message_tensor(RECV_RIGHT, mr).copy_(sims[r + 1].engine.message_tensor(SEND_LEFT, mr))
                s._nrecv = (ml, mr)
            sync = getattr(sims[0].torch.cuda, "synchronize", None)
            if sims[0].torch.cuda.is_available():
                sync()
            for s in sims:
                s.compute()

    def rebalance(self):
        total = sum(s.owned_row_counts() for s in self.sims)
        return [s.apply_histogram(total) for s in self.sims][0]

    def dump(self):
        parts = [s.dump_local() for s in self.sims]
        return {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
