// tisph.cu -- host side of libtisph.so: context, stream orchestration and the C ABI declared
// in include/tisph.h.  There is no CPU fallback: without a CUDA device every call fails.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <new>
#include <vector>

#include "tisph.h"
#include "tisph_kernels.cuh"
#include "tisph_walk.cuh"
#include "tisph_lists.cuh"
#include "tisph_shard.cuh"
#include "tisph_gen1.cuh"
#include "tisph_voxel.cuh"

using namespace tisph;

// ------------------------------------------------------------------------- error plumbing
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess)                                                            \
            return fail(TISPH_ERR_CUDA, "%s failed: %s (%s:%d)", #call,                   \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                      \
    } while (0)

#define CHECK_CTX(ctx)                                                                    \
    do {                                                                                  \
        if (!(ctx)) return fail(TISPH_ERR_INVALID, "null context");                       \
        CU(cudaSetDevice((ctx)->cfg.device));                                             \
    } while (0)

// ------------------------------------------------------------------------------- context
constexpr int MAX_TIMED_STEPS = 64;

struct tisph_ctx {
    tisph_config cfg;
    SimParams sp;
    int n = 0, cap = 0, ncell = 0;
    int color_comp = 3;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    float4 *P[2] = {nullptr, nullptr}, *V[2] = {nullptr, nullptr}, *Q[2] = {nullptr, nullptr};
    int cur = 0;         // which copy holds the authoritative particle records
    int phase = 0;       // 0: between steps, 1: after UPDATE, 2: after DENSITY
    int uphase = 0;      // inside UPDATE: 1 after UPDATE_BIN, 2 after UPDATE_SCAN
    bool skip_sum = false;          // TISPH_P_SKIP_DISCARDED_SUM
    bool walls_pending = false;     // a split force stage ran: TISPH_STAGE_WALLS is due
    int* new_index = nullptr;       // paritcle_index_temp (diagnostics only)
    bool have_sorted = false;
    float4 *D = nullptr, *dvel = nullptr, *a_np = nullptr, *a_p = nullptr;
    float* S = nullptr;
    int *ncount = nullptr, *keys = nullptr, *arrival = nullptr, *ids = nullptr, *keys_sorted = nullptr;
    int *cell_count = nullptr, *cell_end = nullptr, *block_sums = nullptr;
    int* color = nullptr;
    int* err_dev = nullptr;
    // work items of the neighbour walks + neighbour lists handed from walk 1 to walk 2
    int2* items = nullptr;
    int items_cap = 0;
    int pool_rows_cap = 0;
    int occ_dl = 0, occ_fl = 0;        // resident CTAs per SM of the two list kernels
    int arena_rows = 0;                // list-pool rows a CTA of the density walk takes at a time
    int* item_row = nullptr;
    StepCounters* ctr = nullptr;
    int *fb_d = nullptr, *fb_f = nullptr;
    unsigned char* item_flags = nullptr;
    uint32_t* Lg = nullptr;
    int grid_dl = 0, grid_fl = 0, grid_dfb = 0, grid_ffb = 0;   // persistent grids (SMs x resident CTAs)
    void* staging = nullptr;
    size_t staging_bytes = 0;
    // asynchronous host <-> device path (tisph_upload_xv_async / tisph_dump_async): two copy streams and their
    // own staging blocks, so that the copies of step k+1 / step k overlap each other and the kernels
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    cudaEvent_t ev_in_done = nullptr, ev_in_free = nullptr, ev_out_ready = nullptr, ev_out_done = nullptr;
    void *ain = nullptr, *aout = nullptr;
    bool out_pending = false;
    int in_staged = -1;                // particles staged by tisph_upload_xv_stage, -1 = none
    float4 *snapP = nullptr, *snapV = nullptr, *snapQ = nullptr;   // tisph_state_save
    int snap_n = -1;
    int diagnostics = 0;
    int variant = 0;
    bool has_boundary = false;
    int *nbr = nullptr, *nbr_num = nullptr;     // gen-1: particle_neighbors[cap][100], particle_neighbors_num
    float cfl = 0.f;                            // > 0: dt = min(cfg.dt, cfl * h / (c_s + max|v|)) before every step         // any non-fluid particle ever added / announced (TISPH_P_HAS_BOUNDARY)
    // slab sharding (tisph_shard.cuh)
    int *rank_key = nullptr;
    bool sharded = false;
    int row_lo = 0, row_hi = 0, ghost = 1, left_row_lo = -1, right_row_hi = -1;   // cell rows cx * gy + cy; neighbours' far edges, -1 = no neighbour
    int in_off = 0;                    // first record of the input slice inside P/V/Q[cur]
    bool appended = false;             // between tisph_shard_append and the step
    int o_lo = 0, o_hi = 0;            // owned slice of the sorted arrays (host copy)
    bool range_valid = true;
    int* range_dev = nullptr;          // {o_lo, o_hi} on the device
    ShardCounters* shard_ctr = nullptr;
    float4* msg[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
                                       // send_left, send_right, recv_left[0], recv_right[0], recv_left[1], recv_right[1]
    // peer-to-peer halo: the neighbours' receive buffers (two per side, used alternately), mapped
    // into this process with CUDA IPC; the pack kernel then writes the records over NVLink itself
    float4* peer[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [side 0 left / 1 right][parity]
    void* ipc_opened[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int n_ipc_opened = 0;
    unsigned pack_seq = 0;             // packs issued so far; pack k and append k use parity k & 1
    int msg_cap = 0;                   // records per message buffer
    int id_base = 0;                   // original id of the next particle added
    int64_t launches = 0;
    // stage timing
    int timing = 0, timed = 0;
    cudaEvent_t ev[MAX_TIMED_STEPS][4];
    bool ev_made = false;
};

static float d2_cutoff(float h) {
    // smallest f32 t with sqrtf(t) >= h, so that (sqrtf(d2) < h) <=> (d2 < t) in IEEE f32
    float t = h * h;
    while (sqrtf(t) >= h) t = nextafterf(t, 0.0f);
    while (sqrtf(t) < h) t = nextafterf(t, INFINITY);
    return t;
}

static void fill_params(tisph_ctx* c);
// the solver attributes changed (tisph_set_param): refresh the kernels' copy, keep the run-time state
static void fill_physics(tisph_ctx* c) {
    SimParams keep = c->sp;
    fill_params(c);
    c->sp.n = keep.n; c->sp.dt = keep.dt; c->sp.walls = keep.walls; c->sp.lists_only = keep.lists_only;
    c->sp.own_key_lo = keep.own_key_lo; c->sp.own_key_hi = keep.own_key_hi;
    c->sp.walk_key_lo = keep.walk_key_lo; c->sp.walk_key_hi = keep.walk_key_hi;
    c->sp.ghost_walk = keep.ghost_walk;
}

static void fill_params(tisph_ctx* c) {
    const tisph_config& g = c->cfg;
    SimParams& s = c->sp;
    s.n = c->n;
    s.ncell = c->ncell;
    s.gx = g.grid_num[0]; s.gy = g.grid_num[1]; s.gz = g.dim == 3 ? g.grid_num[2] : 1;
    s.dim = g.dim;
    s.h = g.support;
    s.inv_h = 1.0f / g.support;
    s.d2_cut = d2_cutoff(g.support);
    s.k_w = g.k_w; s.k_dw = g.k_dw;
    s.dt = g.dt;
    for (int k = 0; k < 3; ++k) { s.g[k] = g.gravity[k]; s.wall_hi[k] = g.wall_hi[k]; }
    s.pad = g.padding;
    s.rho0 = g.rho0; s.ps_density0 = g.ps_density0;
    s.stiffness = g.stiffness; s.exponent = g.exponent;
    s.visc_fluid_c = g.visc_fluid_c; s.visc_bound_c = g.visc_bound_c; s.eps_h2 = g.eps_h2;
    s.g1_visc_c = g.g1_visc_c; s.g1_mass = g.g1_mass; s.g1_press_c = g.g1_press_c;
    s.m_V0 = g.m_V0;
    s.density_mode = g.density_mode; s.volume_mode = g.volume_mode;
    float e = g.exponent;
    s.int_exponent = (e >= 1.0f && e <= 64.0f && e == floorf(e)) ? (int)e : 0;
    s.one = 1.0f;
    s.walls = 1;
    s.lists_only = 0;
    s.own_key_lo = 0; s.own_key_hi = 0x7fffffff;
    s.walk_key_lo = 0; s.walk_key_hi = 0x7fffffff;
    s.ghost_walk = 1;
}

template <typename T>
static cudaError_t dalloc(T** p, size_t count) {
    return cudaMalloc((void**)p, count * sizeof(T) + 64);
}

static inline int nblocks(int n, int t) { return (n + t - 1) / t; }

// ------------------------------------------------------------------- owned slice (sharding)
// every particle in [0, n) of the current arrays is owned (fresh / restored state)
static int set_owned_all(tisph_ctx* c) {
    c->in_off = 0;
    c->o_lo = 0;
    c->o_hi = c->n;
    c->range_valid = true;
    if (c->sharded) {
        int r[2] = {0, c->n};
        CU(cudaMemcpyAsync(c->range_dev, r, sizeof(r), cudaMemcpyHostToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    return TISPH_OK;
}

// host copy of the owned slice of the sorted arrays; synchronises only when sharded and stale
static int ensure_range(tisph_ctx* c) {
    if (!c->sharded) { c->o_lo = 0; c->o_hi = c->n; return TISPH_OK; }
    if (c->range_valid) return TISPH_OK;
    int r[2];
    CU(cudaMemcpyAsync(r, c->range_dev, sizeof(r), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->o_lo = r[0]; c->o_hi = r[1];
    c->range_valid = true;
    return TISPH_OK;
}

// ------------------------------------------------------------------------------- stages
// gen-1: ps.init() = clear + allocate_particles_to_grid + search_neighbors (partice_system.py:211-215)
static int run_update_gen1(tisph_ctx* c) {
    if (c->n == 0) return fail(TISPH_ERR_INVALID, "no particles");
    cudaStream_t st = c->stream;
    c->sp.n = c->n;
    int a = c->cur;
    int nb_cells = nblocks(c->ncell, SCAN_TILE);
    CU(cudaMemsetAsync(c->cell_count, 0, sizeof(int) * (size_t)c->ncell, st));
    CU(cudaMemsetAsync(c->nbr, 0, sizeof(int) * (size_t)c->n * G1_MAX_NEIGHBORS, st));   // particle_neighbors.fill(0)
    k_bin<<<nblocks(c->n, 256), 256, 0, st>>>(c->sp, c->P[a], c->keys, c->arrival, c->cell_count, c->err_dev);
    k_scan_reduce<<<nb_cells, SCAN_THREADS, 0, st>>>(c->cell_count, c->ncell, c->block_sums);
    k_scan_spine<<<1, 1024, 0, st>>>(c->block_sums, nb_cells);
    k_scan_apply<<<nb_cells, SCAN_THREADS, 0, st>>>(c->cell_count, c->ncell, c->block_sums, c->cell_end);
    k_place<<<nblocks(c->n, 256), 256, 0, st>>>(c->n, c->keys, c->arrival, c->cell_end, c->ids, nullptr, c->rank_key);
    k_g1_order<<<nblocks(c->n, 256), 256, 0, st>>>(c->n, c->keys, c->ids, c->cell_end, c->keys_sorted);
    k_g1_neighbors<<<nblocks(c->n, 128), 128, 0, st>>>(c->sp, c->P[a], c->Q[a], c->cell_end, c->keys_sorted,
                                                       c->nbr, c->nbr_num, c->err_dev);
    c->launches += 7;
    CU(cudaGetLastError());
    c->phase = 1;
    c->have_sorted = true;
    return TISPH_OK;
}

// ps.update() = update_gird_id + PrefixSumExecutor.run + resort (partice_systemv4.py:251-256), as the
// three sub-stages of the ABI or in one go
static int run_update_bin(tisph_ctx* c) {
    if (c->phase != 0 || c->uphase != 0)
        return fail(TISPH_ERR_INVALID, "UPDATE issued out of order (phase %d.%d)", c->phase, c->uphase);
    if (c->walls_pending) return fail(TISPH_ERR_INVALID, "TISPH_STAGE_WALLS is due (the last force stage ran split)");
    cudaStream_t st = c->stream;
    c->sp.n = c->n;
    CU(cudaMemsetAsync(c->cell_count, 0, sizeof(int) * (size_t)c->ncell, st));
    CU(cudaMemsetAsync(c->ctr, 0, sizeof(StepCounters), st));
    k_bin<<<nblocks(c->n, 256), 256, 0, st>>>(c->sp, c->P[c->cur] + c->in_off, c->keys, c->arrival, c->cell_count, c->err_dev);
    c->launches += 1;
    CU(cudaGetLastError());
    c->uphase = 1;
    return TISPH_OK;
}

static int run_update_scan(tisph_ctx* c) {
    if (c->phase != 0 || c->uphase != 1)
        return fail(TISPH_ERR_INVALID, "UPDATE_SCAN issued out of order (phase %d.%d)", c->phase, c->uphase);
    cudaStream_t st = c->stream;
    int nb_cells = nblocks(c->ncell, SCAN_TILE);
    k_scan_reduce<<<nb_cells, SCAN_THREADS, 0, st>>>(c->cell_count, c->ncell, c->block_sums);
    k_scan_spine<<<1, 1024, 0, st>>>(c->block_sums, nb_cells);
    k_scan_apply<<<nb_cells, SCAN_THREADS, 0, st>>>(c->cell_count, c->ncell, c->block_sums, c->cell_end);
    c->launches += 3;
    CU(cudaGetLastError());
    c->uphase = 2;
    return TISPH_OK;
}

static int run_update_sort(tisph_ctx* c) {
    if (c->phase != 0 || c->uphase != 2)
        return fail(TISPH_ERR_INVALID, "UPDATE_SORT issued out of order (phase %d.%d)", c->phase, c->uphase);
    cudaStream_t st = c->stream;
    int a = c->cur, b = c->cur ^ 1;
    const float4 *Pin = c->P[a] + c->in_off, *Vin = c->V[a] + c->in_off, *Qin = c->Q[a] + c->in_off;
    k_place<<<nblocks(c->n, 256), 256, 0, st>>>(c->n, c->keys, c->arrival, c->cell_end, c->ids,
                                                c->sharded ? Qin : nullptr, c->rank_key);
    k_reorder<<<nblocks(c->n, 256), 256, 0, st>>>(c->n, c->keys, c->ids, c->rank_key, c->cell_end, Pin, Vin,
                                                  Qin, c->P[b], c->V[b], c->Q[b], c->keys_sorted,
                                                  c->diagnostics ? c->new_index : nullptr);
    k_items<<<nblocks(c->ncell, 256), 256, 0, st>>>(c->sp, c->sp.walk_key_lo, c->sp.walk_key_hi, c->cell_end,
                                                    c->items, c->ctr);
    c->launches += 3;
    if (c->sharded) {
        k_owned_range<<<1, 1, 0, st>>>(c->cell_end, c->ncell, c->sp.own_key_lo, c->sp.own_key_hi, c->range_dev);
        c->launches += 1;
        c->range_valid = false;
    }
    CU(cudaGetLastError());
    c->in_off = 0;
    c->appended = false;
    c->cur = b;
    c->phase = 1;
    c->uphase = 0;
    c->have_sorted = true;
    return TISPH_OK;
}

static int run_update(tisph_ctx* c) {
    if (c->phase != 0 || c->uphase != 0)
        return fail(TISPH_ERR_INVALID, "UPDATE issued out of order (phase %d.%d)", c->phase, c->uphase);
    if (c->cfg.generation == 1) return run_update_gen1(c);
    if (c->n == 0) {
        if (!c->sharded) return fail(TISPH_ERR_INVALID, "no particles");
        // an empty slab (nothing owned, nothing received) still takes part in every exchange
        cudaStream_t st = c->stream;
        CU(cudaMemsetAsync(c->cell_count, 0, sizeof(int) * (size_t)c->ncell, st));
        CU(cudaMemsetAsync(c->cell_end, 0, sizeof(int) * (size_t)c->ncell, st));
        CU(cudaMemsetAsync(c->ctr, 0, sizeof(StepCounters), st));
        CU(cudaMemsetAsync(c->range_dev, 0, 2 * sizeof(int), st));
        c->range_valid = false;
        c->in_off = 0; c->appended = false;
        c->cur ^= 1; c->phase = 1; c->have_sorted = true;
        return TISPH_OK;
    }
    int rc;
    if ((rc = run_update_bin(c))) return rc;
    if ((rc = run_update_scan(c))) return rc;
    return run_update_sort(c);
}

static int run_walls(tisph_ctx* c) {
    if (c->phase != 0 || !c->walls_pending) return fail(TISPH_ERR_INVALID, "WALLS follows a split FORCE_ADVECT stage");
    int rc = ensure_range(c);
    if (rc) return rc;
    int n = c->o_hi - c->o_lo;
    if (n > 0)
        k_walls<<<nblocks(n, 256), 256, 0, c->stream>>>(c->sp, n, c->P[c->cur] + c->o_lo, c->V[c->cur] + c->o_lo,
                                                        c->Q[c->cur] + c->o_lo);
    c->launches += 1;
    CU(cudaGetLastError());
    c->walls_pending = false;
    return TISPH_OK;
}

static int run_density(tisph_ctx* c) {
    if (c->phase != 1) return fail(TISPH_ERR_INVALID, "DENSITY issued out of order (phase %d)", c->phase);
    int b = c->cur;
    cudaStream_t st = c->stream;
    if (c->cfg.generation == 1) {
        k_g1_density<<<nblocks(c->n, 128), 128, 0, st>>>(c->sp, c->P[b], c->Q[b], c->nbr, c->nbr_num, c->D, c->S,
                                                         c->ncount);
        c->launches += 1;
        CU(cudaGetLastError());
        c->phase = 2;
        return TISPH_OK;
    }
    // TISPH_P_SKIP_DISCARDED_SUM only applies where the sum really is discarded
    c->sp.lists_only = (c->skip_sum && c->sp.density_mode == 0 && c->sp.volume_mode == 0 && !c->diagnostics) ? 1 : 0;
    // ghost cells: their density is needed by the force walk, but in the reference modes it is mass * W(0)
    c->sp.ghost_walk = (c->sp.density_mode == 0 && c->sp.volume_mode == 0) ? 0 : 1;
    auto kd = c->sp.volume_mode == 1 ? k_density_list<true> : k_density_list<false>;
    kd<<<c->grid_dl, NB_THREADS, c->sp.volume_mode == 1 ? DL_SMEM_AKINCI : DL_SMEM, st>>>(
        c->sp, c->cell_end, c->items, c->ctr, c->pool_rows_cap, c->arena_rows, c->variant == 1, c->P[b], c->V[b], c->Q[b],
        c->D, c->S, c->ncount, c->Lg, c->item_row, c->item_flags, c->fb_d, c->fb_f);
    k_density_fb<<<c->grid_dfb, NB_THREADS, DF_SMEM, st>>>(c->sp, c->cell_end, c->items, c->ctr, c->fb_d,
                                                          c->P[b], c->V[b], c->Q[b], c->D, c->S, c->ncount);
    c->launches += 2;
    CU(cudaGetLastError());
    c->phase = 2;
    return TISPH_OK;
}

static int run_force(tisph_ctx* c) {
    if (c->phase != 2) return fail(TISPH_ERR_INVALID, "FORCE_ADVECT issued out of order (phase %d)", c->phase);
    int b = c->cur, a = c->cur ^ 1;
    cudaStream_t st = c->stream;
    float4* dnp = c->diagnostics ? c->a_np : nullptr;
    float4* dp = c->diagnostics ? c->a_p : nullptr;
    if (c->cfg.generation == 1) {
        k_g1_force<<<nblocks(c->n, 128), 128, 0, st>>>(c->sp, c->P[b], c->V[b], c->Q[b], c->D, c->nbr, c->nbr_num,
                                                       c->P[a], c->V[a], c->Q[a], c->dvel, dnp, dp);
        c->launches += 1;
        CU(cudaGetLastError());
        c->cur = a;
        c->phase = 0;
        return TISPH_OK;
    }
    c->walls_pending = c->sp.walls == 0;
    // diagnostics keep the non-pressure and the pressure sums apart (F_A_NONPRESSURE / F_A_PRESSURE); the plain
    // step needs only their total
    auto kf = c->has_boundary ? (c->diagnostics ? k_force_list<true, true> : k_force_list<true, false>)
                              : (c->diagnostics ? k_force_list<false, true> : k_force_list<false, false>);
    kf<<<c->grid_fl, NB_THREADS, FL_SMEM, st>>>(
        c->sp, c->cell_end, c->items, c->ctr, c->P[b], c->V[b], c->Q[b], c->D, c->P[a], c->V[a], c->Q[a],
        c->dvel, dnp, dp, c->Lg, c->item_row, c->item_flags);
    k_force_fb<<<c->grid_ffb, NB_THREADS, FF_SMEM, st>>>(
        c->sp, c->cell_end, c->items, c->ctr, c->fb_f, c->P[b], c->V[b], c->Q[b], c->D, c->P[a], c->V[a],
        c->Q[a], c->dvel, dnp, dp);
    c->launches += 2;
    CU(cudaGetLastError());
    c->cur = a;
    c->phase = 0;
    return TISPH_OK;
}

// ------------------------------------------------------------------------------- C ABI
extern "C" {

const char* tisph_last_error(void) { return g_err; }
int tisph_abi_version(void) { return TISPH_ABI_VERSION; }

int tisph_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int tisph_create(const tisph_config* cfg, tisph_ctx** out) {
    if (!cfg || !out) return fail(TISPH_ERR_INVALID, "null argument");
    if (cfg->struct_size != (int32_t)sizeof(tisph_config))
        return fail(TISPH_ERR_INVALID, "tisph_config size mismatch: got %d, library has %d",
                    cfg->struct_size, (int)sizeof(tisph_config));
    if (!((cfg->generation == 2 && cfg->dim == 3) || (cfg->generation == 1 && cfg->dim == 2)))
        return fail(TISPH_ERR_INVALID, "generation 2 is the 3D path, generation 1 the 2D path (got generation %d, dim %d)",
                    cfg->generation, cfg->dim);
    if (cfg->capacity <= 0 || cfg->support <= 0.f) return fail(TISPH_ERR_INVALID, "bad capacity/support");
    int64_t ncell = (int64_t)cfg->grid_num[0] * cfg->grid_num[1] * (cfg->dim == 3 ? cfg->grid_num[2] : 1);
    if (ncell <= 0 || ncell > 0x7fffffff) return fail(TISPH_ERR_INVALID, "grid too large");
    if (tisph_device_count() <= 0)
        return fail(TISPH_ERR_NO_DEVICE, "no CUDA device visible; libtisph has no CPU fallback");
    CU(cudaSetDevice(cfg->device));
    tisph_ctx* c = new (std::nothrow) tisph_ctx();
    if (!c) return fail(TISPH_ERR_INVALID, "out of host memory");
    c->cfg = *cfg;
    c->cap = cfg->capacity;
    c->ncell = (int)ncell;
    c->color_comp = cfg->generation == 2 ? 3 : 1;
    fill_params(c);
    size_t cap = (size_t)c->cap;
    cudaError_t e = cudaSuccess;
    auto A = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    A(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    for (int k = 0; k < 2; ++k) { A(dalloc(&c->P[k], cap)); A(dalloc(&c->V[k], cap)); A(dalloc(&c->Q[k], cap)); }
    A(dalloc(&c->D, cap)); A(dalloc(&c->dvel, cap));
    A(dalloc(&c->S, cap)); A(dalloc(&c->ncount, cap));
    A(dalloc(&c->keys, cap)); A(dalloc(&c->arrival, cap)); A(dalloc(&c->ids, cap)); A(dalloc(&c->keys_sorted, cap));
    A(dalloc(&c->rank_key, cap)); A(dalloc(&c->range_dev, 2)); A(dalloc(&c->shard_ctr, 1));
    A(dalloc(&c->cell_count, (size_t)c->ncell)); A(dalloc(&c->cell_end, (size_t)c->ncell));
    A(dalloc(&c->block_sums, (size_t)nblocks(c->ncell, SCAN_TILE) + 1024));
    A(dalloc(&c->color, cap * 3));
    A(dalloc(&c->err_dev, 4));
    {
        int64_t occupied_max = ncell < (int64_t)cap ? ncell : (int64_t)cap;
        c->items_cap = (int)(occupied_max + (int64_t)cap / 64 + 2);
        if (cfg->generation == 1) c->items_cap = 1;           // gen-1 walks an explicit table
    }
    if (cfg->generation == 1) {
        A(dalloc(&c->nbr, cap * G1_MAX_NEIGHBORS));
        A(dalloc(&c->nbr_num, cap));
    }
    A(dalloc(&c->items, (size_t)c->items_cap));
    A(dalloc(&c->ctr, 1));
    A(dalloc(&c->fb_d, (size_t)c->items_cap)); A(dalloc(&c->fb_f, (size_t)c->items_cap));
    A(dalloc(&c->item_flags, (size_t)c->items_cap));
    // neighbour-list pool: rows of 32 words (128 B).  A warp of the density walk (4 targets) reserves 1 + the
    // words of its longest lane: ~12 rows at the reference spacing, i.e. ~0.4 KB per particle; sized for
    // 0.75 KB per particle plus slack.  Items that find the pool exhausted take the fallback force kernel.
    {
        int64_t rows = cfg->generation == 1 ? 1 : (int64_t)cap * 6 + 592 * 4 * ARENA_MIN;
        c->pool_rows_cap = (int)(rows < 0x7fffffff / 2 ? rows : 0x7fffffff / 2);
    }
    A(dalloc(&c->Lg, (size_t)c->pool_rows_cap * 32));
    A(dalloc(&c->item_row, 16 * (size_t)c->items_cap));         // one row index per pass and warp
    c->staging_bytes = cap * 16 * 3;
    A(cudaMalloc(&c->staging, c->staging_bytes));
    if (e == cudaSuccess) {
        A(cudaMemsetAsync(c->err_dev, 0, 16, c->stream));
        A(cudaMemsetAsync(c->dvel, 0, cap * 16, c->stream));
        A(cudaMemsetAsync(c->S, 0, cap * 4, c->stream));
        A(cudaMemsetAsync(c->ncount, 0, cap * 4, c->stream));
        A(cudaMemsetAsync(c->D, 0, cap * 16, c->stream));
        A(cudaMemsetAsync(c->keys_sorted, 0, cap * 4, c->stream));
        A(cudaMemsetAsync(c->cell_end, 0, (size_t)c->ncell * 4, c->stream));
        A(cudaMemsetAsync(c->cell_count, 0, (size_t)c->ncell * 4, c->stream));
        if (c->nbr_num) A(cudaMemsetAsync(c->nbr_num, 0, cap * 4, c->stream));
        A(cudaFuncSetAttribute(k_density_list<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DL_SMEM));
        A(cudaFuncSetAttribute(k_density_list<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DL_SMEM_AKINCI));
        A(cudaFuncSetAttribute(k_force_list<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FL_SMEM));
        A(cudaFuncSetAttribute(k_force_list<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FL_SMEM));
        A(cudaFuncSetAttribute(k_force_list<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FL_SMEM));
        A(cudaFuncSetAttribute(k_force_list<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FL_SMEM));
        A(cudaFuncSetAttribute(k_density_fb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DF_SMEM));
        A(cudaFuncSetAttribute(k_force_fb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FF_SMEM));
        int sms = 0, occ = 0;
        A(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device));
        A(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_density_list<false>, NB_THREADS, DL_SMEM));
        c->grid_dl = sms * (occ > 0 ? occ : 1);
        c->occ_dl = occ;
        // every CTA holds at most one partly used chunk: small pools hand out small chunks
        c->arena_rows = c->pool_rows_cap / (4 * c->grid_dl);
        if (c->arena_rows > ARENA_ROWS) c->arena_rows = ARENA_ROWS;
        if (c->arena_rows < 2 * ARENA_MIN) c->arena_rows = 2 * ARENA_MIN;
        A(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_force_list<true, true>, NB_THREADS, FL_SMEM));
        c->grid_fl = sms * (occ > 0 ? occ : 1);
        c->occ_fl = occ;
        A(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_density_fb, NB_THREADS, DF_SMEM));
        c->grid_dfb = sms * (occ > 0 ? occ : 1);
        A(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_force_fb, NB_THREADS, FF_SMEM));
        c->grid_ffb = sms * (occ > 0 ? occ : 1);
        A(cudaStreamSynchronize(c->stream));
    }
    if (e != cudaSuccess) {
        fail(TISPH_ERR_CUDA, "allocation/initialisation failed: %s", cudaGetErrorString(e));
        tisph_destroy(c);
        return TISPH_ERR_CUDA;
    }
    *out = c;
    return TISPH_OK;
}

int tisph_destroy(tisph_ctx* c) {
    if (!c) return TISPH_OK;
    cudaSetDevice(c->cfg.device);
    if (c->own_stream) cudaStreamSynchronize(c->own_stream);
    for (int k = 0; k < 2; ++k) { cudaFree(c->P[k]); cudaFree(c->V[k]); cudaFree(c->Q[k]); }
    cudaFree(c->D); cudaFree(c->dvel); cudaFree(c->a_np); cudaFree(c->a_p); cudaFree(c->new_index); cudaFree(c->S);
    cudaFree(c->ncount); cudaFree(c->keys); cudaFree(c->arrival); cudaFree(c->ids);
    cudaFree(c->keys_sorted); cudaFree(c->cell_count); cudaFree(c->cell_end);
    cudaFree(c->snapP); cudaFree(c->snapV); cudaFree(c->snapQ);
    cudaFree(c->block_sums); cudaFree(c->color); cudaFree(c->err_dev); cudaFree(c->staging);
    if (c->copy_in) {
        cudaStreamSynchronize(c->copy_in); cudaStreamSynchronize(c->copy_out);
        cudaEventDestroy(c->ev_in_done); cudaEventDestroy(c->ev_in_free);
        cudaEventDestroy(c->ev_out_ready); cudaEventDestroy(c->ev_out_done);
        cudaStreamDestroy(c->copy_in); cudaStreamDestroy(c->copy_out);
        cudaFree(c->ain); cudaFree(c->aout);
    }
    cudaFree(c->items); cudaFree(c->ctr); cudaFree(c->fb_d); cudaFree(c->fb_f); cudaFree(c->item_flags);
    cudaFree(c->Lg); cudaFree(c->item_row);
    cudaFree(c->rank_key); cudaFree(c->range_dev); cudaFree(c->shard_ctr);
    cudaFree(c->nbr); cudaFree(c->nbr_num);
    for (int k = 0; k < 6; ++k) cudaFree(c->msg[k]);
    for (int k = 0; k < c->n_ipc_opened; ++k) cudaIpcCloseMemHandle(c->ipc_opened[k]);
    if (c->ev_made)
        for (int s = 0; s < MAX_TIMED_STEPS; ++s)
            for (int k = 0; k < 4; ++k) cudaEventDestroy(c->ev[s][k]);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
    return TISPH_OK;
}

int tisph_add_particles(tisph_ctx* c, int32_t n, const float* pos, const float* vel,
                        const float* density, const float* pressure, const int32_t* material,
                        const int32_t* color) {
    CHECK_CTX(c);
    if (n < 0 || !pos || !vel || !density || !pressure || !material)
        return fail(TISPH_ERR_INVALID, "null/negative argument");
    if (c->phase != 0) return fail(TISPH_ERR_INVALID, "cannot add particles in the middle of a step");
    if (c->appended || (c->sharded && c->have_sorted))
        return fail(TISPH_ERR_INVALID, "a sharded context takes particles only before its first step");
    if ((int64_t)c->n + n > c->cap)
        return fail(TISPH_ERR_CAPACITY, "particle_num %d + %d exceeds particle_max_num %d", c->n, n, c->cap);
    if (n == 0) return TISPH_OK;
    for (int32_t i = 0; i < n && !c->has_boundary; ++i)
        if (material[i] != MAT_FLUID) c->has_boundary = true;
    int dim = c->cfg.dim;
    cudaStream_t st = c->stream;
    // stage in slices so that the staging buffer (48 B/particle) always suffices
    size_t per = (size_t)(2 * dim + 3) * 4;
    int max_chunk = (int)(c->staging_bytes / per);
    for (int done = 0; done < n; done += max_chunk) {
        int m = n - done < max_chunk ? n - done : max_chunk;
        char* s = (char*)c->staging;
        float* d_pos = (float*)s;          s += (size_t)m * dim * 4;
        float* d_vel = (float*)s;          s += (size_t)m * dim * 4;
        float* d_rho = (float*)s;          s += (size_t)m * 4;
        float* d_pr = (float*)s;           s += (size_t)m * 4;
        int* d_mat = (int*)s;
        CU(cudaMemcpyAsync(d_pos, pos + (size_t)done * dim, (size_t)m * dim * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d_vel, vel + (size_t)done * dim, (size_t)m * dim * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d_rho, density + done, (size_t)m * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d_pr, pressure + done, (size_t)m * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d_mat, material + done, (size_t)m * 4, cudaMemcpyHostToDevice, st));
        k_pack_particles<<<nblocks(m, 256), 256, 0, st>>>(m, dim, c->n + done, c->id_base + done, c->cfg.m_V0, d_pos, d_vel,
                                                          d_rho, d_pr, d_mat, c->P[c->cur], c->V[c->cur],
                                                          c->Q[c->cur]);
        c->launches += 1;
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(st));   // staging is reused by the next slice
    }
    size_t cc = (size_t)c->color_comp;
    if (!c->sharded) {   // colour is looked up by original id; a shard sees ids beyond its capacity
        if (color)
            CU(cudaMemcpyAsync(c->color + (size_t)c->n * cc, color, (size_t)n * cc * 4, cudaMemcpyHostToDevice, st));
        else
            CU(cudaMemsetAsync(c->color + (size_t)c->n * cc, 0, (size_t)n * cc * 4, st));
    }
    CU(cudaStreamSynchronize(st));
    c->n += n;
    c->id_base += n;
    c->sp.n = c->n;
    c->have_sorted = false;
    return set_owned_all(c);
}

int tisph_reset(tisph_ctx* c) {
    CHECK_CTX(c);
    CU(cudaStreamSynchronize(c->stream));
    c->n = 0; c->sp.n = 0; c->phase = 0; c->uphase = 0; c->walls_pending = false; c->have_sorted = false; c->id_base = 0;
    CU(cudaMemsetAsync(c->err_dev, 0, 16, c->stream));
    return set_owned_all(c);
}

int tisph_particle_num(tisph_ctx* c, int32_t* n) {
    if (!c || !n) return fail(TISPH_ERR_INVALID, "null argument");
    if (c->sharded) {              // owned particles only (ghosts are not this rank's)
        CU(cudaSetDevice(c->cfg.device));
        int rc = ensure_range(c);
        if (rc) return rc;
        *n = c->o_hi - c->o_lo;
        return TISPH_OK;
    }
    *n = c->n;
    return TISPH_OK;
}

int tisph_state_save(tisph_ctx* c) {
    CHECK_CTX(c);
    if (c->phase != 0) return fail(TISPH_ERR_INVALID, "cannot save in the middle of a step");
    size_t cap = (size_t)c->cap;
    if (!c->snapP) { CU(dalloc(&c->snapP, cap)); CU(dalloc(&c->snapV, cap)); CU(dalloc(&c->snapQ, cap)); }
    if (c->appended) return fail(TISPH_ERR_INVALID, "cannot save between shard_append and the step");
    int rc = ensure_range(c);
    if (rc) return rc;
    int cnt = c->o_hi - c->o_lo;               // owned particles only
    size_t bytes = (size_t)cnt * sizeof(float4);
    CU(cudaMemcpyAsync(c->snapP, c->P[c->cur] + c->o_lo, bytes, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaMemcpyAsync(c->snapV, c->V[c->cur] + c->o_lo, bytes, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaMemcpyAsync(c->snapQ, c->Q[c->cur] + c->o_lo, bytes, cudaMemcpyDeviceToDevice, c->stream));
    c->snap_n = cnt;
    return TISPH_OK;
}

int tisph_state_restore(tisph_ctx* c) {
    CHECK_CTX(c);
    if (c->snap_n < 0) return fail(TISPH_ERR_INVALID, "no saved state");
    if (c->phase != 0) return fail(TISPH_ERR_INVALID, "cannot restore in the middle of a step");
    size_t bytes = (size_t)c->snap_n * sizeof(float4);
    CU(cudaMemcpyAsync(c->P[c->cur], c->snapP, bytes, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaMemcpyAsync(c->V[c->cur], c->snapV, bytes, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaMemcpyAsync(c->Q[c->cur], c->snapQ, bytes, cudaMemcpyDeviceToDevice, c->stream));
    c->n = c->snap_n;
    c->sp.n = c->n;
    c->have_sorted = false;
    return set_owned_all(c);
}

// max |v|^2 over the owned fluid particles (synchronises)
static int max_speed2(tisph_ctx* c, float* v2) {
    int rc = ensure_range(c);
    if (rc) return rc;
    unsigned int bits = 0u;
    unsigned int* d = (unsigned int*)(c->err_dev + 2);
    int n_own = c->o_hi - c->o_lo;
    CU(cudaMemsetAsync(d, 0, 4, c->stream));
    if (n_own > 0)
        k_vmax<<<nblocks(n_own, 256) < 1184 ? nblocks(n_own, 256) : 1184, 256, 0, c->stream>>>(
            n_own, c->V[c->cur] + c->o_lo, c->Q[c->cur] + c->o_lo, d);
    c->launches += 1;
    CU(cudaMemcpyAsync(&bits, d, 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    memcpy(v2, &bits, 4);
    return TISPH_OK;
}

static int make_events(tisph_ctx* c) {
    if (c->ev_made) return TISPH_OK;
    for (int s = 0; s < MAX_TIMED_STEPS; ++s)
        for (int k = 0; k < 4; ++k) CU(cudaEventCreate(&c->ev[s][k]));
    c->ev_made = true;
    return TISPH_OK;
}

int tisph_step(tisph_ctx* c, int32_t nsteps) {
    CHECK_CTX(c);
    if (c->phase != 0 || c->uphase != 0 || c->walls_pending)
        return fail(TISPH_ERR_INVALID, "step issued in the middle of a staged step");
    if (c->sharded && nsteps > 1)
        return fail(TISPH_ERR_INVALID, "a sharded context needs a halo exchange before every step");
    for (int s = 0; s < nsteps; ++s) {
        bool t = c->timing && c->timed < MAX_TIMED_STEPS;
        int rc;
        if (c->cfl > 0.f && !c->appended) {     // optional CFL step (extension): one small reduction + one sync per step
            float v2 = 0.f;
            if ((rc = max_speed2(c, &v2))) return rc;
            float dt = c->cfl * c->cfg.support / (c->cfg.c_s + sqrtf(v2));
            c->sp.dt = dt < c->cfg.dt ? dt : c->cfg.dt;
        }
        if (t) CU(cudaEventRecord(c->ev[c->timed][0], c->stream));
        if ((rc = run_update(c))) return rc;
        if (t) CU(cudaEventRecord(c->ev[c->timed][1], c->stream));
        if ((rc = run_density(c))) return rc;
        if (t) CU(cudaEventRecord(c->ev[c->timed][2], c->stream));
        if ((rc = run_force(c))) return rc;
        if (t) { CU(cudaEventRecord(c->ev[c->timed][3], c->stream)); c->timed++; }
    }
    return TISPH_OK;
}

int tisph_stage_run(tisph_ctx* c, int32_t stage) {
    CHECK_CTX(c);
    switch (stage) {
        case TISPH_STAGE_UPDATE: return run_update(c);
        case TISPH_STAGE_DENSITY: return run_density(c);
        case TISPH_STAGE_FORCE_ADVECT: return run_force(c);
        case TISPH_STAGE_UPDATE_BIN:
        case TISPH_STAGE_UPDATE_SCAN:
        case TISPH_STAGE_UPDATE_SORT:
            if (c->cfg.generation != 2 || c->n == 0)
                return fail(TISPH_ERR_INVALID, "the update sub-stages are those of ParticleSystemV4 with particles");
            return stage == TISPH_STAGE_UPDATE_BIN ? run_update_bin(c)
                   : stage == TISPH_STAGE_UPDATE_SCAN ? run_update_scan(c) : run_update_sort(c);
        case TISPH_STAGE_WALLS: return run_walls(c);
    }
    return fail(TISPH_ERR_INVALID, "unknown stage %d", stage);
}

static int check_device_errors(tisph_ctx* c) {
    int err[4] = {0, 0, 0, 0};
    CU(cudaMemcpyAsync(err, c->err_dev, 16, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (err[0]) {
        CU(cudaMemsetAsync(c->err_dev, 0, 16, c->stream));
        return fail(TISPH_ERR_DOMAIN, "%d particle(s) left the grid (the reference reads out of bounds here)", err[0]);
    }
    if (err[1]) {
        CU(cudaMemsetAsync(c->err_dev, 0, 16, c->stream));
        return fail(TISPH_ERR_CAPACITY, "%d overflow(s) of the gen-1 cell lists (100 per cell) / neighbour lists (100): "
                                        "undefined behaviour in the reference", err[1]);
    }
    return TISPH_OK;
}

int tisph_sync(tisph_ctx* c) {
    CHECK_CTX(c);
    { int rc = tisph_dump_wait(c); if (rc) return rc; }
    CU(cudaStreamSynchronize(c->stream));
    return check_device_errors(c);
}

int tisph_download(tisph_ctx* c, int32_t field, void* dst, size_t bytes) {
    CHECK_CTX(c);
    if (!dst) return fail(TISPH_ERR_INVALID, "null destination");
    cudaStream_t st = c->stream;
    if (c->appended) return fail(TISPH_ERR_INVALID, "cannot download between shard_append and the step");
    { int rc = ensure_range(c); if (rc) return rc; }
    const int off = c->o_lo;                    // owned slice (0 unless sharded)
    int n = c->o_hi - c->o_lo, dim = c->cfg.dim, cur = c->cur;
    const float4* src = nullptr;
    int comp0 = 0, ncomp = 1;
    float floor_x = -INFINITY;
    const void* direct = nullptr;   // already in the reference layout on the device
    size_t count = (size_t)n;
    switch (field) {
        case TISPH_F_X: src = c->P[cur]; comp0 = 0; ncomp = dim; break;
        case TISPH_F_V: src = c->V[cur]; comp0 = 0; ncomp = dim; break;
        case TISPH_F_MASS: src = c->P[cur]; comp0 = 3; break;
        case TISPH_F_VOLUME: src = c->V[cur]; comp0 = 3; break;
        case TISPH_F_DENSITY:          // between the density and the force stage: the clamped density of the D scratch
            if (c->phase == 2 && c->cfg.generation == 1) { src = c->D; comp0 = 2; }
            else if (c->phase == 2) { src = c->D; comp0 = 0; floor_x = c->sp.rho0; }
            else { src = c->Q[cur]; comp0 = 0; }
            break;
        case TISPH_F_PRESSURE: if (c->phase == 2) { src = c->D; comp0 = 3; } else { src = c->Q[cur]; comp0 = 1; } break;
        case TISPH_F_MATERIAL: src = c->Q[cur]; comp0 = 2; break;
        case TISPH_F_ORIG_ID: src = c->Q[cur]; comp0 = 3; break;
        case TISPH_F_D_VELOCITY: src = c->dvel; comp0 = 0; ncomp = dim; break;
        case TISPH_F_DENSITY_RAW: src = c->D; comp0 = 0; break;
        case TISPH_F_A_NONPRESSURE:
        case TISPH_F_A_PRESSURE:
            if (!c->a_np) return fail(TISPH_ERR_INVALID, "enable TISPH_P_DIAGNOSTICS before the step");
            src = field == TISPH_F_A_NONPRESSURE ? c->a_np : c->a_p; comp0 = 0; ncomp = dim; break;
        case TISPH_F_DENSITY_SUM: direct = c->S; break;
        case TISPH_F_NEIGHBOR_COUNT: direct = c->cfg.generation == 1 ? c->nbr_num : c->ncount; break;
        case TISPH_F_NEIGHBORS:
            if (c->cfg.generation != 1) return fail(TISPH_ERR_INVALID, "gen-2 keeps no explicit neighbour table");
            direct = c->nbr; count = (size_t)n * G1_MAX_NEIGHBORS; break;
        case TISPH_F_GRID_IDS: direct = c->uphase ? c->keys : c->keys_sorted; break;   // unsorted between bin and sort
        case TISPH_F_GRID_PARTICLES_NUM:      // gen-2: inclusive scan; gen-1: per-cell counts (partice_system.py:131)
            direct = (c->cfg.generation == 1 || c->uphase == 1) ? c->cell_count : c->cell_end;   // histogram until the scan ran
            count = (size_t)c->ncell; break;
        case TISPH_F_X_IN: src = c->P[c->phase ? cur : cur ^ 1]; comp0 = 0; ncomp = dim; break;
        case TISPH_F_V_IN: src = c->V[c->phase ? cur : cur ^ 1]; comp0 = 0; ncomp = dim; break;
        case TISPH_F_PRESSURE_STORED: src = c->Q[c->phase ? cur : cur ^ 1]; comp0 = 1; break;
        case TISPH_F_PARTICLE_INDEX:
            if (!c->new_index || !c->diagnostics) return fail(TISPH_ERR_INVALID, "enable TISPH_P_DIAGNOSTICS before the step");
            direct = c->new_index; break;
        case TISPH_F_CELL_COUNT: direct = c->cell_count; count = (size_t)c->ncell; break;
        case TISPH_F_COLOR:
            if (c->sharded) return fail(TISPH_ERR_INVALID, "colour is kept by the host side of a sharded run");
            ncomp = c->color_comp; break;
        default: return fail(TISPH_ERR_INVALID, "unknown field %d", field);
    }
    if (src) src += off;
    if (direct && count == (size_t)n) direct = (const char*)direct + (size_t)off * 4;
    size_t need = count * (size_t)ncomp * 4;
    if (bytes != need) return fail(TISPH_ERR_INVALID, "field %d needs %zu bytes, got %zu", field, need, bytes);
    if (need == 0) return TISPH_OK;
    if (direct) {
        CU(cudaMemcpyAsync(dst, direct, need, cudaMemcpyDeviceToHost, st));
    } else {
        if (need > c->staging_bytes) return fail(TISPH_ERR_INVALID, "staging buffer too small");
        if (field == TISPH_F_COLOR)
            k_gather_color<<<nblocks(n, 256), 256, 0, st>>>(n, ncomp, c->Q[cur] + off, c->color, (int*)c->staging);
        else
            k_unpack<<<nblocks(n, 256), 256, 0, st>>>(n, src, comp0, ncomp, (uint32_t*)c->staging, floor_x);
        c->launches += 1;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(dst, c->staging, need, cudaMemcpyDeviceToHost, st));
    }
    CU(cudaStreamSynchronize(st));
    return check_device_errors(c);
}

int tisph_upload_xv(tisph_ctx* c, const float* pos, const float* vel) {
    CHECK_CTX(c);
    if (!pos || !vel) return fail(TISPH_ERR_INVALID, "null argument");
    if (c->phase != 0 || c->appended) return fail(TISPH_ERR_INVALID, "cannot upload in the middle of a step");
    { int rc = ensure_range(c); if (rc) return rc; }
    int n = c->o_hi - c->o_lo, dim = c->cfg.dim;          // the owned particles (all of them unless sharded)
    if (n == 0) return TISPH_OK;
    cudaStream_t st = c->stream;
    float* d_pos = (float*)c->staging;
    float* d_vel = d_pos + (size_t)n * dim;
    CU(cudaMemcpyAsync(d_pos, pos, (size_t)n * dim * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_vel, vel, (size_t)n * dim * 4, cudaMemcpyHostToDevice, st));
    k_upload_xv<<<nblocks(n, 256), 256, 0, st>>>(n, dim, d_pos, d_vel, c->P[c->cur] + c->o_lo, c->V[c->cur] + c->o_lo);
    c->launches += 1;
    CU(cudaGetLastError());
    return TISPH_OK;
}

// ---- asynchronous host <-> device path -----------------------------------------------------------
static int async_setup(tisph_ctx* c) {
    if (c->copy_in) return TISPH_OK;
    CU(cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->copy_out, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->ev_in_done, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_in_free, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_out_ready, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_out_done, cudaEventDisableTiming));
    CU(cudaMalloc(&c->ain, (size_t)c->cap * 24));           // x | v
    CU(cudaMalloc(&c->aout, (size_t)c->cap * 44));          // x | v | material | colour | orig_id
    CU(cudaEventRecord(c->ev_in_free, c->stream));
    return TISPH_OK;
}

// host -> device half of an asynchronous upload: may be issued while a step is still running
int tisph_upload_xv_stage(tisph_ctx* c, const float* pos, const float* vel, int32_t n) {
    CHECK_CTX(c);
    if (!pos || !vel || n < 0 || n > c->cap) return fail(TISPH_ERR_INVALID, "bad argument");
    { int rc = async_setup(c); if (rc) return rc; }
    const int dim = c->cfg.dim;
    float* d_pos = (float*)c->ain;
    float* d_vel = d_pos + (size_t)n * dim;
    CU(cudaStreamWaitEvent(c->copy_in, c->ev_in_free, 0));          // the previous upload's kernel has read the block
    if (n > 0) {
        CU(cudaMemcpyAsync(d_pos, pos, (size_t)n * dim * 4, cudaMemcpyHostToDevice, c->copy_in));
        CU(cudaMemcpyAsync(d_vel, vel, (size_t)n * dim * 4, cudaMemcpyHostToDevice, c->copy_in));
    }
    CU(cudaEventRecord(c->ev_in_done, c->copy_in));
    c->in_staged = n;
    return TISPH_OK;
}

// device half: the staged x, v replace those of the (owned) particles, in their current order
int tisph_upload_xv_commit(tisph_ctx* c) {
    CHECK_CTX(c);
    if (c->in_staged < 0) return fail(TISPH_ERR_INVALID, "nothing staged (tisph_upload_xv_stage)");
    if (c->phase != 0 || c->uphase != 0 || c->appended) return fail(TISPH_ERR_INVALID, "cannot upload in the middle of a step");
    { int rc = ensure_range(c); if (rc) return rc; }
    const int n = c->o_hi - c->o_lo, dim = c->cfg.dim;
    if (n != c->in_staged)
        return fail(TISPH_ERR_INVALID, "%d particles were staged, the context owns %d", c->in_staged, n);
    c->in_staged = -1;
    if (n == 0) return TISPH_OK;
    float* d_pos = (float*)c->ain;
    float* d_vel = d_pos + (size_t)n * dim;
    CU(cudaStreamWaitEvent(c->stream, c->ev_in_done, 0));
    k_upload_xv<<<nblocks(n, 256), 256, 0, c->stream>>>(n, dim, d_pos, d_vel, c->P[c->cur] + c->o_lo, c->V[c->cur] + c->o_lo);
    c->launches += 1;
    CU(cudaGetLastError());
    CU(cudaEventRecord(c->ev_in_free, c->stream));
    return TISPH_OK;
}

int tisph_upload_xv_async(tisph_ctx* c, const float* pos, const float* vel) {
    CHECK_CTX(c);
    if (!pos || !vel) return fail(TISPH_ERR_INVALID, "null argument");
    if (c->phase != 0 || c->uphase != 0 || c->appended) return fail(TISPH_ERR_INVALID, "cannot upload in the middle of a step");
    { int rc = ensure_range(c); if (rc) return rc; }
    int rc = tisph_upload_xv_stage(c, pos, vel, c->o_hi - c->o_lo);
    return rc ? rc : tisph_upload_xv_commit(c);
}

int tisph_dump_wait(tisph_ctx* c) {
    CHECK_CTX(c);
    if (c->out_pending) {
        CU(cudaEventSynchronize(c->ev_out_done));
        c->out_pending = false;
    }
    return TISPH_OK;
}

int tisph_dump_async(tisph_ctx* c, float* pos, float* vel, int32_t* material, int32_t* color, int32_t* orig_id) {
    CHECK_CTX(c);
    if (c->phase != 0 || c->uphase != 0 || c->appended) return fail(TISPH_ERR_INVALID, "cannot dump in the middle of a step");
    if (color && c->sharded) return fail(TISPH_ERR_INVALID, "colour is kept by the host side of a sharded run");
    { int rc = ensure_range(c); if (rc) return rc; }
    { int rc = async_setup(c); if (rc) return rc; }
    { int rc = tisph_dump_wait(c); if (rc) return rc; }               // the staging block and the host arrays are free again
    const int n = c->o_hi - c->o_lo, dim = c->cfg.dim, cc = c->color_comp, cur = c->cur;
    if (n == 0) return TISPH_OK;
    char* s = (char*)c->aout;
    float* o_x = (float*)s;  s += (size_t)n * dim * 4;
    float* o_v = (float*)s;  s += (size_t)n * dim * 4;
    int* o_m = (int*)s;      s += (size_t)n * 4;
    int* o_c = (int*)s;      s += (size_t)n * cc * 4;
    int* o_i = (int*)s;
    k_dump_pack<<<nblocks(n, 256), 256, 0, c->stream>>>(n, dim, cc, c->P[cur] + c->o_lo, c->V[cur] + c->o_lo,
                                                        c->Q[cur] + c->o_lo, c->color, pos ? o_x : nullptr, vel ? o_v : nullptr,
                                                        material ? o_m : nullptr, color ? o_c : nullptr, orig_id ? o_i : nullptr);
    c->launches += 1;
    CU(cudaGetLastError());
    CU(cudaEventRecord(c->ev_out_ready, c->stream));
    CU(cudaStreamWaitEvent(c->copy_out, c->ev_out_ready, 0));
    if (pos) CU(cudaMemcpyAsync(pos, o_x, (size_t)n * dim * 4, cudaMemcpyDeviceToHost, c->copy_out));
    if (vel) CU(cudaMemcpyAsync(vel, o_v, (size_t)n * dim * 4, cudaMemcpyDeviceToHost, c->copy_out));
    if (material) CU(cudaMemcpyAsync(material, o_m, (size_t)n * 4, cudaMemcpyDeviceToHost, c->copy_out));
    if (color) CU(cudaMemcpyAsync(color, o_c, (size_t)n * cc * 4, cudaMemcpyDeviceToHost, c->copy_out));
    if (orig_id) CU(cudaMemcpyAsync(orig_id, o_i, (size_t)n * 4, cudaMemcpyDeviceToHost, c->copy_out));
    CU(cudaEventRecord(c->ev_out_done, c->copy_out));
    // the next kernel that writes the staging block is the next tisph_dump_async's, which waits above
    c->out_pending = true;
    return TISPH_OK;
}

int tisph_device_ptr(tisph_ctx* c, int32_t field, void** ptr, int32_t* stride_bytes) {
    if (!c || !ptr) return fail(TISPH_ERR_INVALID, "null argument");
    CU(cudaSetDevice(c->cfg.device));
    { int rc = ensure_range(c); if (rc) return rc; }
    switch (field) {
        case TISPH_F_X: *ptr = c->P[c->cur] + c->o_lo; break;
        case TISPH_F_V: *ptr = c->V[c->cur] + c->o_lo; break;
        case TISPH_F_D_VELOCITY: *ptr = c->dvel + c->o_lo; break;
        default: return fail(TISPH_ERR_INVALID, "field %d has no zero-copy view", field);
    }
    if (stride_bytes) *stride_bytes = 16;
    return TISPH_OK;
}

int tisph_set_param(tisph_ctx* c, int32_t param, double value) {
    CHECK_CTX(c);
    switch (param) {
        case TISPH_P_DT: c->cfg.dt = (float)value; c->sp.dt = (float)value; return TISPH_OK;
        case TISPH_P_CFL:
            if (c->sharded && value > 0.0)
                return fail(TISPH_ERR_INVALID, "TISPH_P_CFL on a sharded context: the ranks must agree on dt "
                                               "(max-reduce TISPH_P_MAX_SPEED and set TISPH_P_DT)");
            c->cfl = (float)value; if (c->cfl <= 0.f) c->sp.dt = c->cfg.dt; return TISPH_OK;
        case TISPH_P_SPLIT_WALLS:
            if (c->phase == 0 && c->walls_pending) return fail(TISPH_ERR_INVALID, "TISPH_STAGE_WALLS is due");
            c->sp.walls = value != 0.0 ? 0 : 1; return TISPH_OK;
        case TISPH_P_SKIP_DISCARDED_SUM: c->skip_sum = value != 0.0; return TISPH_OK;
        case TISPH_P_STIFFNESS: c->cfg.stiffness = c->sp.stiffness = (float)value; return TISPH_OK;
        case TISPH_P_EXPONENT: c->cfg.exponent = (float)value; fill_physics(c); return TISPH_OK;
        case TISPH_P_VISCOSITY:      // wcsphv2.py:69 (2 nu h c_s) ; sph_base.py:81 (2 (dim+2) nu)
            c->cfg.visc_fluid_c = (float)(2.0 * value * c->cfg.support * c->cfg.c_s);
            c->cfg.g1_visc_c = (float)(2.0 * (c->cfg.dim + 2) * value);
            fill_physics(c); return TISPH_OK;
        case TISPH_P_DENSITY0:       // sph_basev2.py:13 ; gen-1: sph_base.py:16,68
            c->cfg.rho0 = (float)value;
            c->cfg.g1_mass = (float)(c->cfg.m_V0 * value);
            c->cfg.g1_press_c = (float)(-value * c->cfg.m_V0);
            fill_physics(c); return TISPH_OK;
        case TISPH_P_GRAVITY_X: case TISPH_P_GRAVITY_Y: case TISPH_P_GRAVITY_Z:
            c->cfg.gravity[param - TISPH_P_GRAVITY_X] = (float)value; fill_physics(c); return TISPH_OK;
        case TISPH_P_DENSITY_MODE: c->cfg.density_mode = (int)value; c->sp.density_mode = (int)value; return TISPH_OK;
        case TISPH_P_VOLUME_MODE: c->cfg.volume_mode = (int)value; c->sp.volume_mode = (int)value; return TISPH_OK;
        case TISPH_P_DIAGNOSTICS:
            c->diagnostics = value != 0.0;
            if (c->diagnostics && !c->a_np) {
                CU(dalloc(&c->a_np, (size_t)c->cap));
                CU(dalloc(&c->a_p, (size_t)c->cap));
                CU(dalloc(&c->new_index, (size_t)c->cap));
                CU(cudaMemsetAsync(c->new_index, 0, (size_t)c->cap * 4, c->stream));
                CU(cudaMemsetAsync(c->a_np, 0, (size_t)c->cap * 16, c->stream));
                CU(cudaMemsetAsync(c->a_p, 0, (size_t)c->cap * 16, c->stream));
            }
            return TISPH_OK;
        case TISPH_P_KERNEL_VARIANT: c->variant = (int)value; return TISPH_OK;
        case TISPH_P_ID_BASE: c->id_base = (int)value; return TISPH_OK;
        case TISPH_P_HAS_BOUNDARY: c->has_boundary = c->has_boundary || value != 0.0; return TISPH_OK;
    }
    return fail(TISPH_ERR_INVALID, "unknown parameter %d", param);
}

int tisph_get_param(tisph_ctx* c, int32_t param, double* value) {
    if (!c || !value) return fail(TISPH_ERR_INVALID, "null argument");
    switch (param) {
        case TISPH_P_DT: *value = c->sp.dt; return TISPH_OK;      // the dt of the last step when TISPH_P_CFL is on
        case TISPH_P_CFL: *value = c->cfl; return TISPH_OK;
        case TISPH_P_STAT_CHECK_FAILURES: {      // debug builds (-DTISPH_CHECKS): failed device-side bounds checks
            int h[2] = {0, 0};
            CU(cudaSetDevice(c->cfg.device));
            CU(cudaStreamSynchronize(c->stream));
            CU(cudaMemcpyFromSymbol(h, g_tisph_check, sizeof(h)));
#ifdef TISPH_CHECKS
            *value = h[0] ? h[0] + h[1] * 1e-6 : 0.0;     // count + first failing line / 1e6
#else
            *value = -1.0;                                // checks are compiled out
#endif
            return TISPH_OK;
        }
        case TISPH_P_SPLIT_WALLS: *value = c->sp.walls ? 0 : 1; return TISPH_OK;
        case TISPH_P_SKIP_DISCARDED_SUM: *value = c->skip_sum; return TISPH_OK;
        case TISPH_P_PHASE: *value = c->phase + 10 * c->uphase; return TISPH_OK;
        case TISPH_P_STIFFNESS: *value = c->cfg.stiffness; return TISPH_OK;
        case TISPH_P_EXPONENT: *value = c->cfg.exponent; return TISPH_OK;
        case TISPH_P_VISCOSITY: *value = c->cfg.generation == 1 ? c->cfg.g1_visc_c / (2.0 * (c->cfg.dim + 2))
                                                                : c->cfg.visc_fluid_c / (2.0 * c->cfg.support * c->cfg.c_s);
            return TISPH_OK;
        case TISPH_P_DENSITY0: *value = c->cfg.rho0; return TISPH_OK;
        case TISPH_P_GRAVITY_X: case TISPH_P_GRAVITY_Y: case TISPH_P_GRAVITY_Z:
            *value = c->cfg.gravity[param - TISPH_P_GRAVITY_X]; return TISPH_OK;
        case TISPH_P_MAX_SPEED: {
            CU(cudaSetDevice(c->cfg.device));
            if (c->phase != 0 || c->appended) return fail(TISPH_ERR_INVALID, "max |v| is that of a completed step");
            float v2 = 0.f;
            int rc = max_speed2(c, &v2);
            if (rc) return rc;
            *value = sqrtf(v2);
            return TISPH_OK;
        }
        case TISPH_P_DENSITY_MODE: *value = c->cfg.density_mode; return TISPH_OK;
        case TISPH_P_VOLUME_MODE: *value = c->cfg.volume_mode; return TISPH_OK;
        case TISPH_P_DIAGNOSTICS: *value = c->diagnostics; return TISPH_OK;
        case TISPH_P_KERNEL_VARIANT: *value = c->variant; return TISPH_OK;
        case TISPH_P_ID_BASE: *value = c->id_base; return TISPH_OK;
        case TISPH_P_HAS_BOUNDARY: *value = c->has_boundary; return TISPH_OK;
        case TISPH_P_STAT_ITEMS:
        case TISPH_P_STAT_FALLBACK_DENSITY:
        case TISPH_P_STAT_FALLBACK_FORCE: {      // work-item counters of the last step (synchronises)
            StepCounters h;
            CU(cudaSetDevice(c->cfg.device));
            CU(cudaMemcpyAsync(&h, c->ctr, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            *value = param == TISPH_P_STAT_ITEMS ? h.n_items
                     : param == TISPH_P_STAT_FALLBACK_DENSITY ? h.n_fb_d : h.n_fb_f;
            return TISPH_OK;
        }
    }
    return fail(TISPH_ERR_INVALID, "unknown parameter %d", param);
}

int tisph_set_stream(tisph_ctx* c, void* cuda_stream) {
    CHECK_CTX(c);
    CU(cudaStreamSynchronize(c->stream));
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    return TISPH_OK;
}

int tisph_launch_count(tisph_ctx* c, int64_t* launches) {
    if (!c || !launches) return fail(TISPH_ERR_INVALID, "null argument");
    *launches = c->launches;
    return TISPH_OK;
}

int tisph_stage_times(tisph_ctx* c, int32_t enable, float* ms_update, float* ms_density,
                      float* ms_force, int32_t* steps) {
    CHECK_CTX(c);
    CU(cudaStreamSynchronize(c->stream));
    float acc[3] = {0.f, 0.f, 0.f};
    for (int s = 0; s < c->timed; ++s)
        for (int k = 0; k < 3; ++k) {
            float ms = 0.f;
            CU(cudaEventElapsedTime(&ms, c->ev[s][k], c->ev[s][k + 1]));
            acc[k] += ms;
        }
    int cnt = c->timed;
    if (ms_update) *ms_update = cnt ? acc[0] / cnt : 0.f;
    if (ms_density) *ms_density = cnt ? acc[1] / cnt : 0.f;
    if (ms_force) *ms_force = cnt ? acc[2] / cnt : 0.f;
    if (steps) *steps = cnt;
    c->timed = 0;
    c->timing = enable != 0;
    if (c->timing) return make_events(c);
    return TISPH_OK;
}

// ------------------------------------------------------------------ slab sharding (8(e))
int tisph_shard_config(tisph_ctx* c, int32_t plane_lo, int32_t plane_hi, int32_t ghost_planes,
                       int32_t left_lo, int32_t right_hi, int32_t message_capacity) {
    CHECK_CTX(c);
    const int gy = c->sp.gy;
    if (plane_lo < 0 || plane_hi > c->sp.gx || plane_lo >= plane_hi)
        return fail(TISPH_ERR_INVALID, "bad slab [%d,%d) of %d planes", plane_lo, plane_hi, c->sp.gx);
    return tisph_shard_config_rows(c, plane_lo * gy, plane_hi * gy, ghost_planes, left_lo < 0 ? -1 : left_lo * gy,
                                   right_hi < 0 ? -1 : right_hi * gy, message_capacity);
}

int tisph_shard_config_rows(tisph_ctx* c, int32_t row_lo, int32_t row_hi, int32_t ghost_planes,
                            int32_t left_row_lo, int32_t right_row_hi, int32_t message_capacity) {
    CHECK_CTX(c);
    if (c->cfg.generation != 2) return fail(TISPH_ERR_INVALID, "slabs exist in the 3D path only");
    if (c->phase != 0 || c->appended) return fail(TISPH_ERR_INVALID, "slabs can only be (re)configured between steps");
    if (!c->sharded && c->have_sorted) return fail(TISPH_ERR_INVALID, "turn sharding on before the first step");
    const bool reconfig = c->sharded;           // moving the slab faces: the particles that fall outside the new
                                                // rows leave with the next pack, like any other migrant
    const int rows = c->sp.gx * c->sp.gy;
    if (row_lo < 0 || row_hi > rows || row_lo >= row_hi || ghost_planes < 1 || ghost_planes > 2 ||
        message_capacity <= 0 || left_row_lo >= row_lo || (right_row_hi >= 0 && right_row_hi <= row_hi))
        return fail(TISPH_ERR_INVALID, "bad slab rows [%d,%d) / ghost %d / capacity %d", row_lo, row_hi,
                    ghost_planes, message_capacity);
    c->sharded = true;
    c->row_lo = row_lo; c->row_hi = row_hi; c->ghost = ghost_planes;
    c->left_row_lo = left_row_lo < 0 ? -1 : left_row_lo; c->right_row_hi = right_row_hi < 0 ? -1 : right_row_hi;
    const int gz = c->sp.gz, gy = c->sp.gy;
    c->sp.own_key_lo = row_lo * gz;
    c->sp.own_key_hi = row_hi * gz;
    // the density walk covers one cell layer around the owned rows: every cell (cx +- 1, cy +- 1) lies within
    // gy + 1 rows of its centre
    c->sp.walk_key_lo = (row_lo - gy - 1 > 0 ? row_lo - gy - 1 : 0) * gz;
    c->sp.walk_key_hi = (row_hi + gy + 1 < rows ? row_hi + gy + 1 : rows) * gz;
    if (message_capacity != c->msg_cap) {
        if (c->peer[0][0] || c->peer[1][0])
            return fail(TISPH_ERR_INVALID, "the message capacity cannot change once peers have mapped the buffers");
        for (int k = 0; k < 6; ++k) { cudaFree(c->msg[k]); c->msg[k] = nullptr; }
        for (int k = 0; k < 6; ++k) CU(dalloc(&c->msg[k], (size_t)message_capacity * SHARD_REC_F4));
        c->msg_cap = message_capacity;
    }
    if (reconfig && c->have_sorted) return TISPH_OK;     // the owned slice of the sorted arrays stays what it is
    return set_owned_all(c);
}

int tisph_shard_ipc_export(tisph_ctx* c, void* handles, size_t bytes) {
    CHECK_CTX(c);
    if (!c->sharded || !handles || bytes != 4 * sizeof(cudaIpcMemHandle_t))
        return fail(TISPH_ERR_INVALID, "needs a sharded context and room for 4 handles of %zu bytes", sizeof(cudaIpcMemHandle_t));
    cudaIpcMemHandle_t* h = (cudaIpcMemHandle_t*)handles;
    for (int k = 0; k < 4; ++k) CU(cudaIpcGetMemHandle(&h[k], c->msg[2 + k]));   // recv_left[0], recv_right[0], recv_left[1], recv_right[1]
    return TISPH_OK;
}

int tisph_shard_ipc_connect(tisph_ctx* c, int32_t side, const void* handles, size_t bytes) {
    CHECK_CTX(c);
    if (!c->sharded || (side != 0 && side != 1) || !handles || bytes != 4 * sizeof(cudaIpcMemHandle_t))
        return fail(TISPH_ERR_INVALID, "bad argument");
    if (c->n_ipc_opened + 2 > 8) return fail(TISPH_ERR_INVALID, "peers already connected");
    const cudaIpcMemHandle_t* h = (const cudaIpcMemHandle_t*)handles;
    // what I send to my LEFT neighbour it receives from its RIGHT, and vice versa
    for (int par = 0; par < 2; ++par) {
        void* p = nullptr;
        CU(cudaIpcOpenMemHandle(&p, h[2 * par + (side == 0 ? 1 : 0)], cudaIpcMemLazyEnablePeerAccess));
        c->ipc_opened[c->n_ipc_opened++] = p;
        c->peer[side][par] = (float4*)p;
    }
    return TISPH_OK;
}

int tisph_shard_ipc_disconnect(tisph_ctx* c) {
    CHECK_CTX(c);
    CU(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < c->n_ipc_opened; ++k) cudaIpcCloseMemHandle(c->ipc_opened[k]);
    c->n_ipc_opened = 0;
    for (int sd = 0; sd < 2; ++sd) c->peer[sd][0] = c->peer[sd][1] = nullptr;
    return TISPH_OK;
}

int tisph_plane_counts(tisph_ctx* c, int32_t* counts) {
    CHECK_CTX(c);
    if (!counts) return fail(TISPH_ERR_INVALID, "null argument");
    if (c->cfg.generation != 2) return fail(TISPH_ERR_INVALID, "x-planes exist in the 3D path only");
    if (!c->have_sorted || c->phase != 0) return fail(TISPH_ERR_INVALID, "plane counts are those of the last completed step");
    // particles per x-plane at the last sort = differences of the inclusive scan at the plane ends
    const int gx = c->sp.gx, plane = c->sp.gy * c->sp.gz;
    std::vector<int> ends((size_t)gx);
    CU(cudaMemcpy2DAsync(ends.data(), sizeof(int), c->cell_end + (plane - 1), (size_t)plane * sizeof(int), sizeof(int),
                         (size_t)gx, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (int p = 0; p < gx; ++p) counts[p] = ends[p] - (p ? ends[p - 1] : 0);
    return TISPH_OK;
}

int tisph_row_counts(tisph_ctx* c, int32_t* counts) {
    CHECK_CTX(c);
    if (!counts) return fail(TISPH_ERR_INVALID, "null argument");
    if (c->cfg.generation != 2) return fail(TISPH_ERR_INVALID, "cell rows exist in the 3D path only");
    if (!c->have_sorted || c->phase != 0) return fail(TISPH_ERR_INVALID, "row counts are those of the last completed step");
    // particles per cell row (cx, cy) at the last sort = differences of the inclusive scan at the row ends
    const int rows = c->sp.gx * c->sp.gy, gz = c->sp.gz;
    std::vector<int> ends((size_t)rows);
    CU(cudaMemcpy2DAsync(ends.data(), sizeof(int), c->cell_end + (gz - 1), (size_t)gz * sizeof(int), sizeof(int),
                         (size_t)rows, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (int r = 0; r < rows; ++r) counts[r] = ends[r] - (r ? ends[r - 1] : 0);
    return TISPH_OK;
}

int tisph_shard_pack(tisph_ctx* c, int32_t* n_left, int32_t* n_right) {
    CHECK_CTX(c);
    if (!c->sharded) return fail(TISPH_ERR_INVALID, "tisph_shard_config has not been called");
    if (c->phase != 0 || c->appended) return fail(TISPH_ERR_INVALID, "pack issued out of order");
    if (!n_left || !n_right) return fail(TISPH_ERR_INVALID, "null argument");
    cudaStream_t st = c->stream;
    CU(cudaMemsetAsync(c->shard_ctr, 0, sizeof(ShardCounters), st));
    // the owned slice is read from device memory: no host round trip between the step and the pack
    const int n_upper = c->range_valid ? c->o_hi - c->o_lo : c->n;
    const int par = (int)(c->pack_seq & 1u);       // which of the neighbour's two receive buffers this step fills
    c->pack_seq++;
    if (n_upper > 0) {
        k_shard_pack<<<nblocks(n_upper, 256), 256, 0, st>>>(
            c->sp, n_upper, c->range_dev, c->row_lo, c->row_hi, c->ghost, c->left_row_lo, c->right_row_hi,
            c->msg_cap, c->P[c->cur], c->V[c->cur], c->Q[c->cur],
            c->peer[0][par] ? c->peer[0][par] : c->msg[0], c->peer[1][par] ? c->peer[1][par] : c->msg[1], c->shard_ctr);
        c->launches += 1;
        CU(cudaGetLastError());
    }
    struct { ShardCounters k; } h;
    int r[2];
    CU(cudaMemcpyAsync(&h.k, c->shard_ctr, sizeof(ShardCounters), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(r, c->range_dev, sizeof(r), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    c->o_lo = r[0]; c->o_hi = r[1]; c->range_valid = true;
    if (h.k.overflow)
        return fail(TISPH_ERR_CAPACITY, "halo message buffers too small: %d record(s) dropped (capacity %d)",
                    h.k.overflow, c->msg_cap);
    if (h.k.lost)
        return fail(TISPH_ERR_DOMAIN, "%d particle(s) jumped over a whole neighbouring slab in one step", h.k.lost);
    *n_left = h.k.n_left;
    *n_right = h.k.n_right;
    return TISPH_OK;
}

int tisph_shard_buffer(tisph_ctx* c, int32_t which, void** ptr, int32_t* capacity_records) {
    if (!c || !ptr) return fail(TISPH_ERR_INVALID, "null argument");
    if (!c->sharded || which < 0 || which > 3) return fail(TISPH_ERR_INVALID, "no such message buffer");
    // the receive buffers alternate with the step: the one the last pack's step uses
    const int par = c->pack_seq ? (int)((c->pack_seq - 1) & 1u) : 0;
    *ptr = which < 2 ? c->msg[which] : c->msg[which + 2 * par];
    if (capacity_records) *capacity_records = c->msg_cap;
    return TISPH_OK;
}

int tisph_shard_append(tisph_ctx* c, int32_t n_from_left, int32_t n_from_right) {
    CHECK_CTX(c);
    if (!c->sharded) return fail(TISPH_ERR_INVALID, "tisph_shard_config has not been called");
    if (c->phase != 0 || c->appended || !c->range_valid) return fail(TISPH_ERR_INVALID, "append issued out of order");
    if (n_from_left < 0 || n_from_right < 0 || n_from_left > c->msg_cap || n_from_right > c->msg_cap)
        return fail(TISPH_ERR_INVALID, "bad record counts %d / %d", n_from_left, n_from_right);
    int64_t end = (int64_t)c->o_hi + n_from_left + n_from_right;
    if (end > c->cap)
        return fail(TISPH_ERR_CAPACITY, "owned slice [%d,%d) + %d + %d halo records exceed the capacity %d",
                    c->o_lo, c->o_hi, n_from_left, n_from_right, c->cap);
    cudaStream_t st = c->stream;
    int cur = c->cur;
    const int par = c->pack_seq ? (int)((c->pack_seq - 1) & 1u) : 0;
    if (n_from_left > 0)
        k_shard_append<<<nblocks(n_from_left, 256), 256, 0, st>>>(n_from_left, c->msg[2 + 2 * par], c->o_hi, c->P[cur],
                                                                  c->V[cur], c->Q[cur]);
    if (n_from_right > 0)
        k_shard_append<<<nblocks(n_from_right, 256), 256, 0, st>>>(n_from_right, c->msg[3 + 2 * par], c->o_hi + n_from_left,
                                                                   c->P[cur], c->V[cur], c->Q[cur]);
    c->launches += (n_from_left > 0) + (n_from_right > 0);
    CU(cudaGetLastError());
    c->in_off = c->o_lo;
    c->n = (c->o_hi - c->o_lo) + n_from_left + n_from_right;
    c->sp.n = c->n;
    c->appended = true;
    c->have_sorted = false;
    return TISPH_OK;
}

// ------------------------------------------------------------------ mesh sampler (8(f) rank 1)
int tisph_voxelize_mesh(int32_t device, const float* vertices, int32_t nv, const int32_t* faces, int32_t nf,
                        float pitch, int32_t fill, const int32_t* lo, const int32_t* dims,
                        uint8_t* occupancy) {
    if (!vertices || !faces || !lo || !dims || !occupancy || nv <= 0 || nf <= 0 || pitch <= 0.f)
        return fail(TISPH_ERR_INVALID, "bad argument");
    if (tisph_device_count() <= 0)
        return fail(TISPH_ERR_NO_DEVICE, "no CUDA device visible; libtisph has no CPU fallback");
    for (int32_t f = 0; f < 3 * nf; ++f)
        if (faces[f] < 0 || faces[f] >= nv) return fail(TISPH_ERR_INVALID, "face index %d out of range", faces[f]);
    CU(cudaSetDevice(device));
    VoxGrid g;
    size_t total = 1;
    for (int k = 0; k < 3; ++k) {
        if (dims[k] < 3) return fail(TISPH_ERR_INVALID, "grid needs an empty padding layer on every side");
        g.lo[k] = lo[k]; g.dims[k] = dims[k];
        total *= (size_t)dims[k];
    }
    if (total > ((size_t)1 << 33)) return fail(TISPH_ERR_CAPACITY, "voxel grid of %zu cells is too large", total);
    g.pitch = pitch;
    float* d_v = nullptr; int* d_f = nullptr; unsigned char *d_occ = nullptr, *d_out = nullptr; int* d_chg = nullptr;
    cudaError_t e = cudaSuccess;
    auto A = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    A(cudaMalloc(&d_v, (size_t)nv * 12)); A(cudaMalloc(&d_f, (size_t)nf * 12));
    A(cudaMalloc(&d_occ, total)); A(cudaMalloc(&d_out, total)); A(cudaMalloc(&d_chg, 4));
    if (e == cudaSuccess) {
        A(cudaMemcpy(d_v, vertices, (size_t)nv * 12, cudaMemcpyHostToDevice));
        A(cudaMemcpy(d_f, faces, (size_t)nf * 12, cudaMemcpyHostToDevice));
        A(cudaMemset(d_occ, 0, total)); A(cudaMemset(d_out, 0, total));
        k_vox_surface<<<nblocks(nf, 128), 128>>>(g, d_v, d_f, nf, d_occ);
        if (fill) {
            int ncol = dims[0] * dims[1];
            for (int it = 0; it < 4096 && e == cudaSuccess; ++it) {
                int chg = 0;
                A(cudaMemset(d_chg, 0, 4));
                k_vox_flood<<<nblocks(ncol, 128), 128>>>(g, d_occ, d_out, d_chg);
                A(cudaMemcpy(&chg, d_chg, 4, cudaMemcpyDeviceToHost));
                if (!chg) break;
            }
            k_vox_fill<<<(unsigned)((total + 255) / 256), 256>>>(total, d_occ, d_out);
        }
        A(cudaGetLastError());
        A(cudaMemcpy(occupancy, d_occ, total, cudaMemcpyDeviceToHost));
    }
    cudaFree(d_v); cudaFree(d_f); cudaFree(d_occ); cudaFree(d_out); cudaFree(d_chg);
    if (e != cudaSuccess) return fail(TISPH_ERR_CUDA, "voxelisation failed: %s", cudaGetErrorString(e));
    return TISPH_OK;
}

}  // extern "C"
