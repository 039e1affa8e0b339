/*
 * tisph.h -- C ABI of libtisph.so, the B200-native (sm_100a) WCSPH step engine.
 *
 * The reference (jiajun-c/Ti-SPH) has no FFI layer: its boundary is the Python class
 * surface (ParticleSystemV4 / WCSPHV2 and the gen-1 2D classes) whose methods are Taichi
 * kernels.  Each entry point below replaces one group of those Taichi kernels; the Python
 * classes in core/ and utils/ of this repo keep the reference's names and call these
 * functions through ctypes (see INTEGRATION.md for the binding).
 *
 * Plain C: pointers + sizes only, no torch / C++ types.  Every function returns 0 on
 * success or a negative tisph_status; the message is available from tisph_last_error().
 * A context owns all of its device memory and one CUDA stream; it is not thread-safe,
 * different contexts are independent.  Host arrays are copied, never retained.
 */
#ifndef TISPH_H
#define TISPH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TISPH_ABI_VERSION 2

typedef struct tisph_ctx tisph_ctx;

typedef enum tisph_status {
    TISPH_OK = 0,
    TISPH_ERR_INVALID = -1,      /* bad argument / wrong state */
    TISPH_ERR_CUDA = -2,         /* CUDA runtime error (message has the detail) */
    TISPH_ERR_CAPACITY = -3,     /* more particles than config.capacity */
    TISPH_ERR_DOMAIN = -4,       /* a particle left the grid (reference: undefined behaviour) */
    TISPH_ERR_NO_DEVICE = -5     /* no CUDA device: there is NO CPU fallback */
} tisph_status;

/* Scalars are the reference's Python-scope constants, evaluated by the host in float64
 * exactly as the reference's constructors do and then rounded to f32 (Taichi bakes them
 * into kernels the same way).  file:line = reference source. */
typedef struct tisph_config {
    int32_t struct_size;      /* = sizeof(tisph_config), ABI check */
    int32_t generation;       /* 2: ParticleSystemV4+WCSPHV2 (3D); 1: ParticleSystem(V2)+WCSPH (2D) */
    int32_t dim;              /* partice_systemv4.py:17 / partice_system.py:10 */
    int32_t device;           /* CUDA ordinal */
    int32_t capacity;         /* particle_max_num, partice_systemv4.py:37-38 / partice_system.py:24 */
    int32_t grid_num[3];      /* partice_systemv4.py:59 / partice_system.py:30 (z=1 in 2D) */
    float support;            /* support_length = 4r = grid_size, partice_systemv4.py:34,58 */
    float padding;            /* partice_systemv4.py:35 */
    float domain_size[3];     /* partice_systemv4.py:22 */
    float wall_hi[3];         /* domain_size - padding, sph_basev2.py:164-182 */
    float m_V0;               /* 0.8 d^dim, partice_systemv4.py:48 / partice_system.py:23 */
    float dt;                 /* sph_basev2.py:14-15 */
    float gravity[3];         /* sph_basev2.py:16 ; gen-1: const.py:2 on the last axis */
    float c_s;                /* wcsphv2.py:16 */
    float rho0;               /* solver density_0, sph_basev2.py:13 */
    float ps_density0;        /* ps.density0 (JSON), used at wcsphv2.py:80 */
    float stiffness;          /* wcsphv2.py:11 */
    float exponent;           /* wcsphv2.py:10 */
    float k_w;                /* k/h^dim,  sph_basev2.py:22-30 */
    float k_dw;               /* 6k/h^dim, sph_basev2.py:42-50 */
    float visc_fluid_c;       /* 2*viscosity*h*c_s, wcsphv2.py:69 */
    float visc_bound_c;       /* 0.08*h*c_s, wcsphv2.py:75-76 */
    float eps_h2;             /* 0.01*h^2, wcsphv2.py:72 / sph_base.py:82 */
    float g1_visc_c;          /* gen-1: 2*(dim+2)*viscosity, sph_base.py:81 */
    float g1_mass;            /* gen-1: m_V*density_0, sph_base.py:16 */
    float g1_press_c;         /* gen-1: -density_0*m_V, sph_base.py:68 */
    int32_t density_mode;     /* 0 = reference (wcsphv2.py:32-34: sum discarded), 1 = summed */
    int32_t volume_mode;      /* 0 = reference (sph_basev2.py:191 by-value arg), 1 = akinci */
    int32_t reserved[8];
} tisph_config;

/* Per-particle fields, named after the reference's ps.* / solver.* Taichi fields. */
typedef enum tisph_field {
    TISPH_F_X = 0,            /* ps.x        f32 [n][dim] */
    TISPH_F_V = 1,            /* ps.v        f32 [n][dim] */
    TISPH_F_MASS = 2,         /* ps.mass     f32 [n] */
    TISPH_F_VOLUME = 3,       /* ps.volume   f32 [n] */
    TISPH_F_DENSITY = 4,      /* ps.density  f32 [n] */
    TISPH_F_PRESSURE = 5,     /* ps.pressure f32 [n] */
    TISPH_F_MATERIAL = 6,     /* ps.material i32 [n] */
    TISPH_F_COLOR = 7,        /* ps.color    i32 [n][3] (gen-2) / i32 [n] (gen-1) */
    TISPH_F_GRID_IDS = 8,     /* ps.grid_ids i32 [n] (cell key of each sorted particle) */
    TISPH_F_GRID_PARTICLES_NUM = 9, /* ps.grid_particles_num i32 [ncell]: INCLUSIVE scan */
    TISPH_F_D_VELOCITY = 10,  /* solver.d_velocity f32 [n][dim] */
    /* diagnostics that the reference computes but does not keep */
    TISPH_F_DENSITY_SUM = 11, /* S_i = sum_j mass_i W(|x_ij|), wcsphv2.py:33   f32 [n] */
    TISPH_F_DENSITY_RAW = 12, /* density before the clamp of wcsphv2.py:46     f32 [n] */
    TISPH_F_NEIGHBOR_COUNT = 13, /* #j: j!=i and norm(x_ij) < h, partice_systemv4.py:344  i32 [n] */
    TISPH_F_ORIG_ID = 14,     /* index the particle had when it was added     i32 [n] */
    TISPH_F_A_NONPRESSURE = 15, /* d_velocity after wcsphv2.py:93 (needs TISPH_P_DIAGNOSTICS) */
    TISPH_F_A_PRESSURE = 16,  /* sum added at wcsphv2.py:53 (needs TISPH_P_DIAGNOSTICS) */
    TISPH_F_CELL_COUNT = 17,  /* histogram before the scan, partice_systemv4.py:213  i32 [ncell] */
    TISPH_F_NEIGHBORS = 18,   /* gen-1 only: ps.particle_neighbors i32 [n][100], zero-filled beyond the
                                 count (partice_system.py:102-121,213); counts = TISPH_F_NEIGHBOR_COUNT */
    /* what the reference's fields hold BETWEEN its kernels (the drop-in classes map ps.x, ps.pressure ...
       onto these while a step is driven kernel by kernel, as tests/golden/make_golden.py does) */
    TISPH_F_X_IN = 19,        /* ps.x between resort() and advert(): the sorted, not yet advected positions */
    TISPH_F_V_IN = 20,        /* ps.v, likewise */
    TISPH_F_PRESSURE_STORED = 21, /* ps.pressure before compute_pressure_force: the previous step's, sorted along */
    TISPH_F_PARTICLE_INDEX = 22   /* ps.paritcle_index_temp (partice_systemv4.py:219-224): sorted position of
                                 every pre-sort particle   i32 [n]  (needs TISPH_P_DIAGNOSTICS) */
} tisph_field;

/* Stages of one step, for stage-by-stage parity tests (tisph_step runs them in order). */
typedef enum tisph_stage {
    TISPH_STAGE_UPDATE = 0,   /* ps.update(): key, histogram, scan, stable sort, reorder
                                 (partice_systemv4.py:251-256) ; gen-1: ps.init() */
    TISPH_STAGE_DENSITY = 1,  /* compute_volume_of_boundary_particle + compute_densities +
                                 clamp/EOS (sph_basev2.py:195-201, wcsphv2.py:28-34,45-47) */
    TISPH_STAGE_FORCE_ADVECT = 2, /* compute_non_pressure_force + compute_pressure_force +
                                 advert + enforce_boundary (wcsphv2.py:83-100, sph_basev2.py:204) */
    /* TISPH_STAGE_UPDATE in the reference's three calls (gen-2), to be issued in this order: */
    TISPH_STAGE_UPDATE_BIN = 3,  /* ps.update_gird_id(): keys + histogram (partice_systemv4.py:206-215) */
    TISPH_STAGE_UPDATE_SCAN = 4, /* ps.prefix_sum_executor.run(ps.grid_particles_num) (:62,255) */
    TISPH_STAGE_UPDATE_SORT = 5, /* ps.resort() (:217-249) */
    /* enforce_boundary() on its own (sph_basev2.py:204-208): after a TISPH_STAGE_FORCE_ADVECT that ran with
       TISPH_P_SPLIT_WALLS = 1 and therefore stopped after advert() */
    TISPH_STAGE_WALLS = 6
} tisph_stage;

typedef enum tisph_param {
    TISPH_P_DT = 0,           /* solver.dt[None] */
    TISPH_P_DENSITY_MODE = 1,
    TISPH_P_VOLUME_MODE = 2,
    TISPH_P_DIAGNOSTICS = 3,  /* 1: also store a_nonpressure / a_pressure every step */
    TISPH_P_KERNEL_VARIANT = 4, /* implementation selector for A/B benchmarking (0 = default) */
    TISPH_P_ID_BASE = 5,      /* original id given to the next particle added (auto-increments);
                                 a sharded run sets it so that ids are global */
    TISPH_P_HAS_BOUNDARY = 6, /* 1: boundary (material 0) particles may reach this context as ghosts
                                 even though none was added to it (sharded runs with rigid bodies);
                                 set automatically when a non-fluid particle is added */
    /* read-only statistics of the last step (tisph_get_param synchronises the stream) */
    TISPH_P_STAT_ITEMS = 7,             /* work items (<= 64 targets of one occupied cell) */
    TISPH_P_STAT_FALLBACK_DENSITY = 8,  /* items whose candidate tile did not fit shared memory */
    TISPH_P_STAT_FALLBACK_FORCE = 9,    /* ... plus items whose neighbour lists overflowed */
    TISPH_P_CFL = 10          /* extension: > 0 turns on a CFL step,
                                 dt = min(TISPH_P_DT, cfl * h / (c_s + max|v|)), evaluated before every step of
                                 tisph_step; 0 (default) = the reference's fixed dt.  Rejected (TISPH_ERR_INVALID) on a
                                 sharded context: every rank would pick its own dt; ShardedSim.set_cfl max-reduces
                                 TISPH_P_MAX_SPEED over the ranks and sets TISPH_P_DT instead. */
    , TISPH_P_STAT_CHECK_FAILURES = 11, /* read-only: device-side bounds checks that failed so far, in libraries
                                 built with -DTISPH_CHECKS (count + first failing source line / 1e6);
                                 -1 when the checks are compiled out */
    TISPH_P_SPLIT_WALLS = 12, /* 1: TISPH_STAGE_FORCE_ADVECT stops after advert(); the walls are applied by
                                 TISPH_STAGE_WALLS (kernel-by-kernel drivers); 0 (default): fused */
    TISPH_P_PHASE = 13,       /* read-only: 0 between steps, 1 after UPDATE, 2 after DENSITY;
                                 + 10 (20) after UPDATE_BIN (UPDATE_SCAN) */
    /* the solver attributes of the reference, assignable after construction like there */
    TISPH_P_STIFFNESS = 14,   /* WCSPH(V2).stiffness, wcsphv2.py:11 */
    TISPH_P_EXPONENT = 15,    /* WCSPH(V2).exponent,  wcsphv2.py:10 */
    TISPH_P_VISCOSITY = 16,   /* SPHBase(V2).viscosity, sph_basev2.py:12 (the kernels' coefficients follow) */
    TISPH_P_DENSITY0 = 17,    /* SPHBase(V2).density_0, sph_basev2.py:13 */
    TISPH_P_GRAVITY_X = 18, TISPH_P_GRAVITY_Y = 19, TISPH_P_GRAVITY_Z = 20,   /* SPHBaseV2.g, sph_basev2.py:16 */
    TISPH_P_MAX_SPEED = 21,   /* read-only: max |v| over the fluid particles this context owns (synchronises);
                                 a sharded run max-reduces it over the ranks to agree on a CFL step */
    TISPH_P_SKIP_DISCARDED_SUM = 22 /* opt-in, default 0.  density_mode 0 reproduces wcsphv2.py:32-34, where the neighbour
                                 sum is accumulated and then overwritten; by default the sum is computed all the same
                                 (TISPH_F_DENSITY_SUM, TISPH_F_NEIGHBOR_COUNT).  1: a caller that reads neither lets the
                                 density walk build the neighbour lists only (no effect in the other modes, nor while
                                 TISPH_P_DIAGNOSTICS is on).  Every field of the reference comes out the same. */
} tisph_param;

const char *tisph_last_error(void);
int tisph_abi_version(void);
/* number of visible CUDA devices (0 => every other call fails with TISPH_ERR_NO_DEVICE) */
int tisph_device_count(void);

/* ParticleSystemV4.__init__ / ParticleSystem.__init__ + WCSPH(V2).__init__: allocate fields. */
int tisph_create(const tisph_config *cfg, tisph_ctx **out);
int tisph_destroy(tisph_ctx *ctx);

/* ParticleSystemV4.add_particles (partice_systemv4.py:171-204): append n particles.
 * pos/vel: [n][dim] f32; density/pressure: [n] f32; material: [n] i32;
 * color: [n][3] i32 (gen-2) or [n] i32 (gen-1).  mass = m_V0*density, volume = m_V0. */
int tisph_add_particles(tisph_ctx *ctx, int32_t n, const float *pos, const float *vel,
                        const float *density, const float *pressure, const int32_t *material,
                        const int32_t *color);
/* Forget all particles (particle_num = 0); used to restart from an identical state. */
int tisph_reset(tisph_ctx *ctx);
int tisph_particle_num(tisph_ctx *ctx, int32_t *n);
/* Device-resident checkpoint of the particle records (x, v, mass, volume, density, pressure,
 * material, ids) in their current order; restore makes it the current state again.  The
 * reference has no checkpointing (its JSON `outputInterval` is unused); bench.py uses this to
 * replay the same synthetic state. */
int tisph_state_save(tisph_ctx *ctx);
int tisph_state_restore(tisph_ctx *ctx);

/* SPHBaseV2.step / SPHBase.step (sph_basev2.py:210-214): nsteps whole steps, asynchronous. */
int tisph_step(tisph_ctx *ctx, int32_t nsteps);
/* One stage of a step (see tisph_stage); stages must be issued in order. */
int tisph_stage_run(tisph_ctx *ctx, int32_t stage);

/* ParticleSystemV4.dump / copy_to_numpy (partice_systemv4.py:279-307): synchronous copy of
 * one field to host memory in the reference's layout; bytes must equal the field size. */
int tisph_download(tisph_ctx *ctx, int32_t field, void *dst, size_t bytes);
/* Overwrite x and v of the current particles (current order) from host arrays [n][dim]. */
int tisph_upload_xv(tisph_ctx *ctx, const float *pos, const float *vel);
/* The same two transfers without blocking the host or the step: the host->device copy runs on a copy
 * stream of the context (it overlaps the kernels already queued), the dump is one packing kernel plus
 * device->host copies on a second copy stream.  Host arrays must be page-locked for the copies to be
 * asynchronous, and must stay untouched until consumed: `pos`/`vel` of an upload until the next call
 * that synchronises (tisph_sync, tisph_dump_wait after a later dump), the destinations of a dump until
 * tisph_dump_wait (or the next tisph_dump_async, which waits for the previous one first).
 * tisph_dump_async fills what ParticleSystemV4.dump() returns (partice_systemv4.py:279-296): position,
 * velocity [n][dim] f32, material [n] i32, colour [n][3] (gen-2) / [n] (gen-1) i32, plus the original ids
 * [n] i32; any destination may be NULL. */
int tisph_upload_xv_async(tisph_ctx *ctx, const float *pos, const float *vel);
/* tisph_upload_xv_async in its two halves: _stage starts the host->device copy of n particles and may be
 * called while a step is still running; _commit (between steps) makes them the x, v of the n owned particles. */
int tisph_upload_xv_stage(tisph_ctx *ctx, const float *pos, const float *vel, int32_t n);
int tisph_upload_xv_commit(tisph_ctx *ctx);
int tisph_dump_async(tisph_ctx *ctx, float *pos, float *vel, int32_t *material, int32_t *color, int32_t *orig_id);
int tisph_dump_wait(tisph_ctx *ctx);
/* Zero-copy hand-off (ggui scene.particles(ps.x), torch, cupy): device pointer of the packed
 * float4 arrays {x,y,z,mass} (TISPH_F_X) / {vx,vy,vz,volume} (TISPH_F_V) / {ax,ay,az,0}
 * (TISPH_F_D_VELOCITY); valid until the next step.  This call does NOT synchronise: tisph_step is
 * asynchronous on the context's own stream, so a consumer on another stream calls tisph_sync first
 * (the Python field views do) or shares its stream through tisph_set_stream. */
int tisph_device_ptr(tisph_ctx *ctx, int32_t field, void **ptr, int32_t *stride_bytes);

int tisph_set_param(tisph_ctx *ctx, int32_t param, double value);
int tisph_get_param(tisph_ctx *ctx, int32_t param, double *value);
/* Run on a caller-owned cudaStream_t (e.g. torch's current stream); NULL restores the own one. */
int tisph_set_stream(tisph_ctx *ctx, void *cuda_stream);
/* Block until the stream is idle and report deferred device-side errors (TISPH_ERR_DOMAIN). */
int tisph_sync(tisph_ctx *ctx);
/* Total kernel launches issued by this context so far. */
int tisph_launch_count(tisph_ctx *ctx, int64_t *launches);
/* Mean device time (ms, CUDA events on the context's stream) of each stage over the steps
 * (at most 64) issued since the previous call.  Events are recorded only while `enable`
 * (as set by the previous call) is non-zero; the call synchronises the stream. */
int tisph_stage_times(tisph_ctx *ctx, int32_t enable, float *ms_update, float *ms_density,
                      float *ms_force, int32_t *steps);

/* ---- spatial-slab sharding: one context per GPU, one process per GPU --------------------
 * The reference is single-device; there is no reference interface to cite.  The cell key is
 * x-major (partice_systemv4.py:98-100), so a rank that owns the x-planes [plane_lo, plane_hi) of
 * the grid owns one contiguous range of the sorted arrays.  Per step the host side does
 *     tisph_shard_pack -> exchange counts -> exchange records (NCCL send/recv between the
 *     message buffers) -> tisph_shard_append -> tisph_step(ctx, 1).
 * A record is 12 floats: {x,y,z,mass, vx,vy,vz,volume, density,pressure,material,orig_id}.
 * After tisph_shard_config every per-particle accessor (tisph_particle_num, tisph_download,
 * tisph_device_ptr, tisph_state_save/restore) covers the OWNED particles only. */
/* ghost_planes: 1 when a particle's density needs no neighbour data (density_mode 0 with
 * volume_mode 0), else 2.  left_lo / right_hi: the far edges of the neighbouring slabs
 * [left_lo, plane_lo) and [plane_hi, right_hi), -1 where there is no neighbour; a particle may
 * migrate anywhere into a neighbouring slab within one step.  message_capacity: records per
 * message buffer. */
int tisph_shard_config(tisph_ctx *ctx, int32_t plane_lo, int32_t plane_hi, int32_t ghost_planes,
                       int32_t left_lo, int32_t right_hi, int32_t message_capacity);
/* tisph_shard_config may be called again between steps to move the slab faces (re-balancing): all
 * ranks must switch at the same step; particles outside the new planes migrate with the next pack. */
/* Particles per x-plane (gx values) as of the last completed step, ghosts included: a rank reads
 * its own planes [plane_lo, plane_hi) from it to feed the global histogram that places the faces. */
int tisph_plane_counts(tisph_ctx *ctx, int32_t *counts);
/* The same at (x-plane, y-row) granularity: a rank owns the cell rows r = cx * gy + cy in [row_lo, row_hi) -- still
 * one contiguous key range of the sorted arrays -- so that equal-count slabs differ by one cell row (a few
 * hundred particles) instead of one plane.  The halo of a neighbour is every particle with one of the
 * neighbour's cells within ghost_planes cell layers in x and y.  tisph_shard_config(lo, hi, ...) is
 * tisph_shard_config_rows(lo * gy, hi * gy, ...).  tisph_row_counts: gx * gy values, like tisph_plane_counts. */
int tisph_shard_config_rows(tisph_ctx *ctx, int32_t row_lo, int32_t row_hi, int32_t ghost_planes,
                            int32_t left_row_lo, int32_t right_row_hi, int32_t message_capacity);
int tisph_row_counts(tisph_ctx *ctx, int32_t *counts);
/* Fill the two send buffers from the owned particles: every particle within ghost_planes of a
 * slab face, or beyond it (a migrant), goes to that neighbour.  Synchronous; returns counts. */
int tisph_shard_pack(tisph_ctx *ctx, int32_t *n_left, int32_t *n_right);
/* Device pointer of a message buffer: 0 send-left, 1 send-right, 2 recv-left, 3 recv-right. */
int tisph_shard_buffer(tisph_ctx *ctx, int32_t which, void **ptr, int32_t *capacity_records);
/* Peer-to-peer halo (optional, one node): every rank exports CUDA IPC handles of its four receive
 * buffers (two per side, used alternately) and connects to those of its neighbours; from then on
 * tisph_shard_pack writes the records straight into the neighbour's receive buffer over NVLink and
 * only the two counts have to be exchanged by the host.  handles: 4 cudaIpcMemHandle_t (256 bytes). */
int tisph_shard_ipc_export(tisph_ctx *ctx, void *handles, size_t bytes);
int tisph_shard_ipc_connect(tisph_ctx *ctx, int32_t side /* 0 left, 1 right */, const void *handles, size_t bytes);
/* Back to the local send buffers (all ranks must agree on which path they use). */
int tisph_shard_ipc_disconnect(tisph_ctx *ctx);
/* Make the received records part of the particle set of the coming step. */
int tisph_shard_append(tisph_ctx *ctx, int32_t n_from_left, int32_t n_from_right);

/* ---- mesh -> boundary particles: the sampling step of ParticleSystemV4.load_rigid_body
 * (partice_systemv4.py:276-277: mesh.voxelized(pitch=particle_diameter).fill().points, trimesh).
 * vertices [nv][3] f32 (already scaled / rotated / translated), faces [nf][3] i32.  Voxel centres
 * lie on the world lattice k * pitch; the grid covers lattice indices lo[k] .. lo[k]+dims[k]-1 and
 * must leave one empty layer around the mesh.  occupancy [dims0][dims1][dims2] u8 (host) receives 1
 * for surface voxels and, if `fill`, for every voxel not connected to the outside. */
int tisph_voxelize_mesh(int32_t device, const float *vertices, int32_t nv, const int32_t *faces,
                        int32_t nf, float pitch, int32_t fill, const int32_t *lo, const int32_t *dims,
                        uint8_t *occupancy);

#ifdef __cplusplus
}
#endif
#endif /* TISPH_H */
