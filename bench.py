#!/usr/bin/env python
"""bench.py -- particle-updates/s of one full WCSPH step (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # the CUDA engine (this repo)
  python bench.py --impl reference ...                      # CPU restatement of the reference

A "step" is SPHBaseV2.step(): bin/scan/sort/reorder, density+EOS, forces+advect+walls over the
whole particle set.  Workload at N=1: C5 of BASELINE.md, the 16 M-particle 3D dam break the
metric is quoted on (`--workload C3` gives the 1 M run).  `value` has the state resident in HBM;
`e2e` goes through the drop-in classes (ParticleSystemV4 / WCSPHV2) with host buffers: pinned
host x,v -> device, step(), dump() -> host, every step.  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle_updates_per_sec"
UNIT = "particle-updates/s"
# algorithmic HBM bytes per particle-step
BYTES_STEP = {"reference": 170.0, "summed": 186.0}
BYTES_FORCE = 68.0      # fused forces+advect+walls: R {x,v,mass,volume,material,rho,p} 44 + W {x,v} 24
BYTES_DENSITY = {"reference": 8.0, "summed": 24.0}     # R x 12 + mass 4, W rho 4 + p 4 (rho = mass W(0) needs no x)


COUNTERS = "profiles/r02_counters.json"
SM_COUNT, SMSP_PER_SM = 148, 4


def ncu_counters(workload, mode, kernel):
    """Per-PARTICLE counters of `kernel` from the committed `ncu --set full` capture of this workload
    (profiles/r02_counters.json, written by scripts/ncu_summary.py): DRAM bytes (read + write), warp
    instructions, FMA-pipe busy cycles.  None when the workload / mode was not captured."""
    try:
        with open(os.path.join(ROOT, COUNTERS)) as f:
            return json.load(f).get(f"{workload}/{mode}", {}).get(kernel)
    except Exception:
        return None


def log(*a):
    print(*a, file=sys.stderr, flush=True)


_RESULT_FD = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on
    stdout): from here on file descriptor 1 goes to stderr, and emit() writes the line to the real stdout."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def mark(self):
        """start of the timed region: samples taken before it (launched early because nvidia-smi
        needs ~100 ms to start) are dropped, unless nothing else is left"""
        import datetime
        self.t0 = datetime.datetime.now()

    def stop(self):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f")
                    rows.append((ts, float(f[1]), float(f[2]), f[5:9]))
                except ValueError:
                    continue
            os.unlink(self.path)
        except Exception:
            pass
        t0 = getattr(self, "t0", None)
        timed = [r for r in rows if t0 is None or r[0] >= t0]
        use = timed or rows[-3:]
        if use:
            reasons = set()
            for r in use:
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            out = {"sm_mhz": float(np.median([r[1] for r in use])), "sm_max_mhz": float(max(r[2] for r in use)),
                   "reasons": sorted(reasons), "samples": len(use), "samples_in_timed_region": len(timed)}
        return out


def workload_scene(name, with_mesh=True):
    """BASELINE.md configurations.  C4 adds the mesh-sampled boundary of the reference's own asset,
    data/models/Dragon_50k.obj (an input fixture of this repo), voxelised at pitch 2r = 0.01 and set
    down on the floor inside the padded domain."""
    from ti_sph_b200 import scene as sc
    s = sc.bench_scene(name)
    if name == "C4" and with_mesh:
        from ti_sph_b200 import mesh
        path = os.path.join(ROOT, "data", "models", "Dragon_50k.obj")
        v, _ = mesh.load_obj(path)
        s["rigidBodies"] = [{
            "geometryFile": path, "scale": [1, 1, 1], "rotationAngle": 0, "rotationAxis": [0, 1, 0],
            "color": [255, 255, 255], "velocity": [0.0, 0.0, 0.0], "density": 1000.0,
            "translation": [float(t) for t in np.array([1.0, 0.06, 0.75]) - [v[:, 0].mean(), v[:, 1].min(), v[:, 2].mean()]]}]
    return s


def sample_scene(name, target_particles):
    """A bounded sample of workload `name` for the CPU legs: same radius/grid/velocity, the
    fluid block cut down along x then y to ~target_particles."""
    s = workload_scene(name, with_mesh=False)           # the CPU legs time a cut of the fluid block only
    blk = s["fluidBlocks"][0]
    r = s["configuration"]["particleRadius"]
    dims = [int(round((blk["end"][i] - blk["start"][i]) / r)) for i in range(3)]
    n = dims[0] * dims[1] * dims[2]
    for ax in (0, 1, 2):
        while n > target_particles * 1.5 and dims[ax] > 40:
            dims[ax] //= 2
            n = dims[0] * dims[1] * dims[2]
    blk["end"] = [blk["start"][i] + dims[i] * r for i in range(3)]
    return s


# ----------------------------------------------------------------------------- CPU legs
def cpu_leg(workload, mode, steps, warmup, target_particles):
    """Times the CPU oracle (restatement of the reference's kernels, oracle/) on host cores."""
    from oracle.oracle import Gen2Oracle, lib
    scene = sample_scene(workload, target_particles)
    # torchrun exports OMP_NUM_THREADS=1: take every core this process may run on, explicitly
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    ora = Gen2Oracle(scene, density_mode=mode, threads=threads)
    cores = lib().ora_num_threads()
    for _ in range(warmup):
        ora.step_fast(1)
    t0 = time.perf_counter()
    for _ in range(steps):
        ora.step_fast(1)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    blk = scene["fluidBlocks"][0]
    sample = (f"{ora.n} particles: block {blk['start']}-{[round(e, 4) for e in blk['end']]} of {workload} "
              f"(r={scene['configuration']['particleRadius']}), {steps} step(s) after {warmup} warm-up")
    return {"value": ora.n / dt, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
            "ms_per_step": dt * 1e3, "n": ora.n}


def run_reference_impl(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    leg = cpu_leg(args.workload, args.mode, args.steps, args.warmup, args.ref_particles)
    try:                                        # re-probe in every run
        import taichi  # noqa: F401
        taichi_state = "importable, but the reference checkout does not travel to this box: CPU restatement timed"
    except Exception as e:
        taichi_state = f"not importable ({type(e).__name__})"
    line = {
        "impl": "reference", "taichi": taichi_state, "metric": METRIC, "value": leg["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": leg["ms_per_step"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload} (bounded sample)", "density_mode": args.mode,
                   "particles": leg["n"]},
        "cpu_baseline": {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": leg["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "Taichi is not installable here, so the reference's own kernels cannot run; this is "
                "the CPU restatement of them (oracle/sph_oracle.c, OpenMP over all host threads)",
    }
    emit(line)


# ----------------------------------------------------------------------------- GPU arm
def engine_at_spacing(workload, mode, factor, start, size):
    """NOT a reference configuration: a block filled at `factor` x the reference's lattice spacing (the reference
    puts particles one RADIUS apart, SURVEY Q7: h = 4 x spacing, 64 particles per cell; factor 2 is the conventional
    one particle diameter: h = 2 x spacing, 8 per cell).  Same solver constants, grid and kernels."""
    from ti_sph_b200 import scene as sc
    from ti_sph_b200.engine import Engine
    s = sc.bench_scene(workload)
    cfg, blk = s["configuration"], s["fluidBlocks"][0]
    x = sc.cube_positions(start, size, factor * cfg["particleRadius"], 3)
    n = len(x)
    eng = Engine(sc.gen2_config(cfg, n, density_mode={"reference": 0, "summed": 1}[mode]))
    eng.add_particles(x, np.full(x.shape, blk["velocity"], np.float32), np.full(n, 1000.0, np.float32),
                      np.zeros(n, np.float32), np.ones(n, np.int32), None)
    return eng


def also_measure(workload, mode, pre, chain, args, torch, stream, lists_only=False, spacing=None):
    """device-resident timing of a second workload / density mode with the rules of the main one (single GPU)"""
    if spacing is None:
        from core.partice_system.partice_systemv4 import ParticleSystemV4
        ps = ParticleSystemV4(workload_scene(workload), density_mode=mode)
        eng = ps.engine
    else:
        eng = engine_at_spacing(workload, mode, *spacing)
    eng.set_stream(stream.cuda_stream)
    if lists_only:
        eng.set_param(_K().P_SKIP_DISCARDED_SUM, 1)
    eng.step(pre)
    eng.save_state()

    def run(k):
        done = 0
        while done < k:
            m = min(chain, k - done)
            eng.restore_state()
            eng.step(m)
            done += m
    run(max(args.warmup, 3))
    eng.stage_times(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    run(args.steps)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / args.steps
    st = eng.stage_times(False)
    n = eng.particle_num
    pairs = int(eng.download(_K().F_NEIGHBOR_COUNT).astype(np.int64).sum())
    fb = int(eng.get_param(_K().P_STAT_FALLBACK_FORCE))
    eng.close()
    extra = {"note": "opt-in TISPH_P_SKIP_DISCARDED_SUM: the density walk builds the neighbour lists only; the sum that "
                     "wcsphv2.py:32-34 overwrites is not evaluated. NOT the headline: the default computes it."} if lists_only else {}
    if spacing is not None:
        extra = {"note": f"NOT a reference configuration (SURVEY 8(d), optional): block {spacing[1]} + {spacing[2]} filled at "
                         f"{spacing[0]} x the reference's lattice spacing, i.e. one particle diameter apart (h = 2 x spacing: 216 "
                         "candidates and ~30 neighbours per particle instead of 1728 and ~250). Shows how the HBM fraction moves "
                         "with the neighbour count; the walk kernels are laid out for the reference's 64 particles per cell."}
    label = f"{workload}: {n} particles" if spacing is None else f"{workload} grid, block at {spacing[0]} x the lattice spacing: {n} particles"
    return {**extra, "workload": label, "density_mode": mode, "value": n / (ms * 1e-3), "unit": UNIT,
            "ms_per_step": ms, "stage_ms": {k: st[k] for k in ("update_ms", "density_ms", "force_ms")},
            "step_hbm_frac": BYTES_STEP[mode] * n / (ms * 1e-3) / 1e9 / measured_peaks()[0],
            "pair_interactions_per_s": None if lists_only else 2.0 * pairs / (ms * 1e-3), "fallback_force_items": fb,
            "state": f"after {pre} steps from the t=0 lattice, replayed in chains of {chain}",
            "l2": "state fits L2 (small workload)" if n * 96 <= 126e6 else "state larger than L2"}


def _K():
    from ti_sph_b200 import _capi
    return _capi


def sharded_check(args, rank, world, local_rank, comm):
    """Correctness of the multi-process path, outside the timed region: two sharded steps of a
    <= 1 M-particle cut of the workload against the single engine on rank 0 -- the same particle ids
    after migration, x and v within 1e-5 (relative to the block size / the initial speed)."""
    from ti_sph_b200.sharded import ShardedSim
    scene = sample_scene(args.workload, 1_000_000)
    small = ShardedSim(scene, rank, world, comm=comm, density_mode=args.mode, device=local_rank)
    small.step(2)
    got = small.dump()                       # gathered on every rank, rank order = global cell order
    small.engine.sync()
    small.close()
    if rank != 0:
        return None
    from core.partice_system.partice_systemv4 import ParticleSystemV4
    ps = ParticleSystemV4(scene, device=local_rank, density_mode=args.mode)
    ps.engine.step(2)
    want = ps.dump()
    ids = ps.engine.download(_K().F_ORIG_ID)
    ps.engine.close()
    n = len(ids)
    if len(got["orig_id"]) != n or not np.array_equal(np.sort(got["orig_id"]), np.arange(n, dtype=np.int32)):
        return f"FAILED: {len(got['orig_id'])} particles / ids are not a permutation of 0..{n - 1}"
    a, b = np.argsort(got["orig_id"], kind="stable"), np.argsort(ids, kind="stable")
    dx = float(np.abs(got["position"][a].astype(np.float64) - want["position"][b]).max())
    dv = float(np.abs(got["velocity"][a].astype(np.float64) - want["velocity"][b]).max())
    vscale = float(np.abs(want["velocity"]).max())
    ok = dx <= 1e-5 * 1.0 and dv <= 1e-5 * max(vscale, 1.0)
    return ("ok" if ok else "FAILED") + f": {n} particles x{world} ranks, 2 steps, max |dx| {dx:.2e} m, max |dv| {dv:.2e} m/s"


def run_gpu(args):
    import torch
    from ti_sph_b200 import _capi as K

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    scene = workload_scene(args.workload)
    hbm_peak, peak_kind = measured_peaks()

    if world == 1:
        from core.partice_system.partice_systemv4 import ParticleSystemV4
        from core.sph.wcsphv2 import WCSPHV2
        t0 = time.time()
        ps = ParticleSystemV4(scene, device=local_rank, density_mode=args.mode)
        solver = WCSPHV2(ps)
        eng = ps.engine
        stream = torch.cuda.Stream()
        torch.cuda.set_stream(stream)      # torch events and the engine share this stream
        eng.set_stream(stream.cuda_stream)
        n_total = eng.particle_num
        log(f"[bench] {args.workload}: {n_total} particles set up in {time.time() - t0:.1f}s")
        step = lambda k=1: eng.step(k)
        sim = None
    else:
        from ti_sph_b200.sharded import ShardedSim, TorchDistComm
        sim = ShardedSim(scene, rank, world, comm=TorchDistComm(device=f"cuda:{local_rank}"),
                         density_mode=args.mode, device=local_rank)
        eng = sim.engine
        stream = sim.stream                # engine kernels and NCCL exchanges are ordered on this stream
        torch.cuda.set_stream(stream)      # ... and so are the timing events
        n_total = sim.global_particle_num
        log(f"[bench] rank {rank}: cell rows [{sim.row_lo},{sim.row_hi}) of {sim.gy} per plane = planes "
            f"[{sim.plane_lo},{sim.plane_hi}), {sim.initial_owned} particles")
        step = lambda k=1: sim.step(k)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ---------------------------------------------------------
    # The reference-mode dam break has no working pressure (wcsphv2.py:32-34 discards the density
    # sum) and clumps within a few dozen steps, after which the cost of a step explodes. So the
    # bench replays a saved state: PRE steps from the t=0 lattice, save, then chains of CHAIN steps
    # each restarted from the saved state (one device-to-device restore per chain, inside the timed
    # region). Every timed step is a full step on simulated steps PRE..PRE+CHAIN of the dam break.
    PRE, CHAIN = args.pre_steps, args.chain

    def run_steps(k):
        done = 0
        while done < k:
            m = min(CHAIN, k - done)
            if sim is None:
                eng.restore_state()
            else:
                sim.restore_state()
            step(m)
            done += m

    step(PRE)
    if sim is None:
        eng.save_state()
    else:
        sim.save_state()
    clocks = ClockSampler(local_rank)
    run_steps(args.warmup)
    barrier()
    eng.stage_times(True)
    l0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if sim is not None and sim.profile is not None:
        sim.profile.clear()
    clocks.mark()
    ev0.record()
    run_steps(args.steps)
    ev1.record()
    barrier()
    clk = clocks.stop()
    ms = ev0.elapsed_time(ev1) / args.steps
    launches = eng.launch_count - l0
    stage = eng.stage_times(False)
    from ti_sph_b200 import _capi as KK
    items = {"items": int(eng.get_param(KK.P_STAT_ITEMS)),
             "fallback_density": int(eng.get_param(KK.P_STAT_FALLBACK_DENSITY)),
             "fallback_force": int(eng.get_param(KK.P_STAT_FALLBACK_FORCE))}
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    value = n_total / (ms * 1e-3)

    # ---- end-to-end through the public API, host buffers ---------------------------------------
    # every step: this step's input state (x, v) goes host -> device from pinned memory, step(),
    # and the step's result (position / velocity / material / colour [+ ids when sharded]) comes back.
    dbg = os.environ.get("BENCH_DEBUG")

    def pinned(shape, dtype):
        return torch.empty(shape, dtype=torch.float32 if dtype == np.float32 else torch.int32).pin_memory().numpy()

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    if sim is None:
        eng.restore_state()
        host = ps.dump()
        hx, hv = pinned(host["position"].shape, np.float32), pinned(host["velocity"].shape, np.float32)
        hx[:] = host["position"]; hv[:] = host["velocity"]
        outs = {k: pinned(v.shape, v.dtype) for k, v in host.items()}
        h2d = hx.nbytes + hv.nbytes
        d2h = sum(v.nbytes for v in outs.values())

        def e2e_step():
            # host -> device: this step's input state (pinned x, v) on the engine's copy stream; step; device -> host:
            # the step's result, packed by one kernel and copied on a second copy stream.  Nothing here blocks,
            # so the copies of neighbouring steps overlap each other and the kernels (PCIe is full duplex);
            # dump_async waits for the previous dump before it reuses the host arrays.
            eng.upload_xv_async(hx, hv)
            solver.step()
            ps.dump_async(outs)

        def e2e_finish():
            ps.dump_wait()
        api = "Engine.upload_xv_async + WCSPHV2.step + ParticleSystemV4.dump_async / dump_wait"
    else:
        sim.restore_state()
        rows = int(1.25 * eng.particle_num) + 4096
        bufs = [{"position": pinned((rows, 3), np.float32), "velocity": pinned((rows, 3), np.float32),
                 "material": pinned((rows,), np.int32), "orig_id": pinned((rows,), np.int32)} for _ in range(2)]

        def views(b, k):
            return {name: a[:k] for name, a in b.items()}
        state = {"k": 0, "d": None}

        def dump_start():
            k = eng.particle_num           # the owned set changes as particles migrate (this waits for the step)
            state["k"] += 1
            state["d"] = sim.dump_local_async(views(bufs[state["k"] & 1], k))
        dump_start()
        sim.dump_wait()
        k0 = eng.particle_num
        hx0, hv0 = pinned((k0, 3), np.float32), pinned((k0, 3), np.float32)
        hx0[:] = state["d"]["position"]; hv0[:] = state["d"]["velocity"]
        bytes_io = [hx0.nbytes + hv0.nbytes, 0]
        sim.restore_state()
        eng.upload_xv_stage(hx0, hv0)

        def e2e_step():
            # every step starts from the saved state (device-side restore, so that the owned set matches the host
            # arrays), takes its input x, v from pinned host memory and returns its result to the host; the
            # device -> host copy of step k runs on a copy stream next to the upload and the kernels of step k+1
            sim.restore_state()
            eng.upload_xv_commit()
            sim.step(1)
            eng.upload_xv_stage(hx0, hv0)  # the next step's input, while this step runs
            dump_start()
            bytes_io[1] = sum(state["d"][k].nbytes for k in ("position", "velocity", "material", "orig_id"))

        def e2e_finish():
            sim.dump_wait()
        api = ("ShardedSim.restore_state + Engine.upload_xv_stage / _commit + step + dump_local_async / dump_wait "
               "(per rank; bytes summed over ranks)")
    e2e_step()
    e2e_finish()
    barrier()
    ev0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e2e_finish()                           # the last result is on the host: inside the timed region
    ev1.record()
    barrier()
    ems = ev0.elapsed_time(ev1) / e2e_steps
    if sim is not None:
        import torch.distributed as dist
        t = torch.tensor([ems, float(bytes_io[0]), float(bytes_io[1])], device="cuda", dtype=torch.float64)
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ems, h2d, d2h = float(tm[0].item()), int(t[1].item()), int(t[2].item())
    e2e = {"value": n_total / (ems * 1e-3), "unit": UNIT, "ms_per_step": ems, "steps": e2e_steps,
           "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "api": api}

    if sim is not None and sim.profile:
        k = max(sim.profile.get("steps", 1), 1)
        log(f"[bench] rank {rank} host phases per step (ms): " +
            ", ".join(f"{n} {1e3 * v / k:.3f}" for n, v in sim.profile.items() if n != "steps"))
    check = None
    if sim is not None:
        import torch.distributed as dist
        if not args.no_check:
            check = sharded_check(args, rank, world, local_rank, sim.comm)
            if rank == 0:
                log(f"[bench] sharded_check: {check}")
        if rank != 0:
            sim.close()
            dist.barrier()
            dist.destroy_process_group()
            return
    # ---- roofline of the dominant kernel, measured live with CUDA events on the engine's stream ----
    n_local = eng.particle_num
    cand = {"density": ("k_density_list (+k_density_fb): boundary volume, density summation, clamp, Tait EOS",
                        BYTES_DENSITY[args.mode], stage["density_ms"], "k_density_list"),
            "force": ("k_force_list (+k_force_fb): non-pressure + pressure forces, advect, walls",
                      BYTES_FORCE, stage["force_ms"], "k_force_list")}
    dom = max(cand, key=lambda k: cand[k][2])
    kname, kbytes, kms, ktag = cand[dom]
    achieved = kbytes * n_local / (kms * 1e-3) / 1e9 if kms > 0 else 0.0
    # The bound that binds is instruction issue / the FP32 pipe, not HBM: the committed ncu
    # capture gives warp instructions and FMA-pipe cycles per particle, the live kernel time turns them
    # into fractions of the issue and pipe peaks (SMs x 4 schedulers x SM clock).
    sm_hz = 1e6 * float((clk or {}).get("sm_mhz") or 1965.0)
    slots_per_s = SM_COUNT * SMSP_PER_SM * sm_hz
    per_kernel = {}
    for key, (_, _, kms_k, tag) in cand.items():
        c = ncu_counters(args.workload, args.mode, tag)
        if c and kms_k > 0:
            per_kernel[key] = {
                "issue_frac": c["warp_inst_per_particle"] * n_local / (kms_k * 1e-3) / slots_per_s,
                "fp32_pipe_frac": c["fma_pipe_cycles_per_particle"] * n_local / (kms_k * 1e-3) / slots_per_s,
                "traffic": c["dram_bytes_per_particle"] * n_local}
    cdom = per_kernel.get(dom, {})
    pairs = int(eng.download(K.F_NEIGHBOR_COUNT).astype(np.int64).sum())       # of the last step, this rank
    roofline = {
        "bound": "hbm", "kernel": kname, "achieved": achieved,
        "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "peak_kind": peak_kind,
        "traffic": cdom.get("traffic"), "bytes_per_particle": kbytes, "launch_ms": kms,
        "issue_frac": cdom.get("issue_frac"), "fp32_pipe_frac": cdom.get("fp32_pipe_frac"),
        "pairs_per_s": 2.0 * pairs / (ms * 1e-3),
        "per_kernel": per_kernel,
        "step_hbm_frac": BYTES_STEP[args.mode] * n_total / (ms * 1e-3) / 1e9 / hbm_peak / world,
        "stage_ms": {k: stage[k] for k in ("update_ms", "density_ms", "force_ms")},
        "stage_hbm_frac": {"update": (16 + 2 + 76) * n_local / max(stage["update_ms"], 1e-9) / 1e6 / hbm_peak,
                           "density": BYTES_DENSITY[args.mode] * n_local / max(stage["density_ms"], 1e-9) / 1e6 / hbm_peak,
                           "force": BYTES_FORCE * n_local / max(stage["force_ms"], 1e-9) / 1e6 / hbm_peak},
        "work_items_last_step": items,
        "definitions": "traffic = DRAM read+write bytes per launch on this rank: per-particle bytes of the committed "
                       f"ncu --set full capture ({COUNTERS}) x this rank's particles; issue_frac = warp instructions / "
                       "(kernel time x 148 SMs x 4 schedulers x SM clock); fp32_pipe_frac = FMA-pipe busy cycles over "
                       "the same denominator (a packed f32x2 instruction holds the pipe two cycles); pairs_per_s = "
                       "neighbour pairs evaluated per second by this rank (sum of the neighbour counts, two walks per step)",
        "note": "both walks are bound by instruction issue / the FP32 pipe and by shared-memory gathers, not by HBM: "
                "at the reference's h = 4 x spacing every particle tests ~1730-2030 candidates and evaluates ~250 "
                "pairs in each walk; see DESIGN.md section 5 and profiles/",
    }
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        leg = cpu_leg(args.workload, args.mode, 5, 1, args.cpu_particles)     # ~10 s of CPU work
        cpu = {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: 3D WCSPH dam break, {n_total} particles "
                               f"(BASELINE.md {args.workload})",
                   "particles": n_total, "density_mode": args.mode,
                   "state": f"dam-break block after {PRE} steps from the t=0 lattice; replayed in chains of "
                            f"{CHAIN} steps (device-side restore inside the timed region)",
                   "l2": "state (48 B/particle x 2 copies) larger than L2" if n_total * 96 > 126e6
                         else "state fits L2 (small workload)",
                   "parallelism": "single GPU" if world == 1 else
                   f"x-slabs x{world}, halo records " + ("written by the pack kernel into the neighbour's buffer over NVLink "
                                                          "(CUDA IPC), counts through shared memory" if sim.p2p
                                                          else "over NCCL send/recv")},
        "clocks": clk, "gpu_launches": int(launches), "roofline": roofline,
    }
    line["e2e"] = e2e
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if check is not None:
        line["sharded_check"] = check
    if world == 1 and args.workload == "C5" and not args.no_also:
        # the metric is quoted at 1 M and at 16 M particles, and `summed` is the mode in which the Tait
        # pressure works: both ride along in config.also (1 M: the state after 50 steps, BASELINE.md;
        # summed C5: steps 2..5 -- with a working pressure the reference's fixed dt = 2e-4 is beyond the CFL
        # limit at r = 0.005 and the block disintegrates within ~10 steps)
        eng.close()
        also = {}
        for key, (wl, mode, pre, chain, lo) in {"C3_1M": ("C3", args.mode, 50, 5, False),
                                                "C5_16M_summed": ("C5", "summed", 2, 3, False),
                                                "C5_16M_reference_lists_only": ("C5", "reference", 10, 5, True),
                                                **({"nonreference_spacing_2r_8M": ("C5", args.mode, 10, 5, False)}
                                                   if args.also_spacing else {})}.items():
            try:
                also[key] = also_measure(wl, mode, pre, chain, args, torch, stream, lists_only=lo,
                                         spacing=(2.0, [0.3, 0.1, 0.3], [4.0, 2.0, 1.0]) if key.startswith("nonreference") else None)
            except Exception as e:                  # never lose the main line over a side measurement
                also[key] = {"error": str(e)[:200]}
        line["config"]["also"] = also
    emit(line)
    if sim is not None:
        import torch.distributed as dist
        sim.close()
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="tisph", choices=["tisph", "reference"])
    ap.add_argument("--workload", default="C5", choices=["C2", "C3", "C4", "C5"])
    ap.add_argument("--mode", default="reference", choices=["reference", "summed"],
                    help="density mode: reference = bit-faithful to wcsphv2.py:32-34, summed = intent")
    ap.add_argument("--e2e-steps", type=int, default=20,
                    help="steps of the end-to-end leg (the first upload and the last read-back are not overlapped by "
                         "anything: over 5 steps they cost 3.7 ms per step at C5, over 20 less than 1)")
    ap.add_argument("--pre-steps", type=int, default=None,
                    help="steps from the lattice before the state is saved (default: 10; 2 in summed mode, where the "
                         "reference's fixed dt is beyond the CFL limit at r = 0.005 and the block disintegrates within ~10 steps)")
    ap.add_argument("--chain", type=int, default=None, help="steps per replay chain (default 5, as main_3d.py renders every 5; 3 in summed mode)")
    ap.add_argument("--cpu-particles", type=int, default=1000000, help="size of the cpu_baseline sample")
    ap.add_argument("--ref-particles", type=int, default=1000000, help="sample size of --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the side measurements of the C5 run (1 M particles; summed mode)")
    ap.add_argument("--also-spacing", action="store_true",
                    help="add the NON-reference side measurement at one-diameter lattice spacing (8 particles per cell) to config.also")
    ap.add_argument("--no-check", action="store_true", help="skip the sharded-vs-single-engine check of a multi-GPU run")
    args = ap.parse_args()
    if args.pre_steps is None:
        args.pre_steps = 2 if args.mode == "summed" else 10
    if args.chain is None:
        args.chain = 3 if args.mode == "summed" else 5
    if args.warmup < 3:
        log("[bench] warm-up raised to 3 (timing rules)")
        args.warmup = 3
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        claim_stdout()
        return run_reference_impl(args)
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-launch ourselves one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    claim_stdout()
    run_gpu(args)


if __name__ == "__main__":
    main()
