// tisph_walk.cuh -- the two neighbour walks of a WCSPH step on sm_100a.
//
// Work items.  A walk is cut into ITEMS: one item = up to 64 target particles of one occupied
// grid cell, 32 where the 27-cell neighbourhood is dense (k_items builds the list after the scan).
// Persistent CTAs of 256 threads pull items from an atomic cursor, so the ~97 % empty cells of a
// dam-break grid cost nothing.  Inside an item the CTA's threads are arranged [split][target
// lane]: 64 (32) targets, each walked by 4 (8) "split" threads; split s takes the candidate PAIRS
// s, s + nsplit, ... (pair-wise interleave keeps the splits' survivor counts balanced on
// lattice-like states); the partial sums meet in shared memory.
//
// Candidate tile.  The 27 neighbour cells of a cell are 9 contiguous ranges of the sorted arrays
// (z is the fastest key digit); they are staged into shared memory TCAP candidates at a time (one
// tile at the reference spacing, several where cells are crowded).  Range of cell c is
// [cell_end[max(0,c-1)], cell_end[c])  (partice_systemv4.py:343; cell 0 is therefore invisible as
// a neighbour -- reference quirk, reproduced).  Cells outside the grid are empty (the reference
// reads out of bounds there).  In shared memory the pairs of one split are contiguous (pair_slot),
// so the filter reads consecutive slots and the random gathers of a warp spread over all banks.
//
// Walk 1 (k_density_list): FILTER every candidate with packed f32x2 arithmetic (FADD2 / FMUL2 /
// FFMA2: two candidates per instruction, broadcast loads) against a cutoff widened by 1e-6 -- a
// superset of the neighbours -- pushing the survivors' slots to a per-thread pending list in shared
// memory ([slot][thread]); DRAIN the lists in words of four entries, branch-free: exact IEEE
// test sqrt(d2) < h in the reference's evaluation order (bit-exact neighbour count), kernel sum.
// The drained entries ((tile << 11) | slot, u16, 4 per 8-byte word, [word][thread]) are streamed
// to the item's rows of the global neighbour-list pool.
// Walk 2 (k_force_list): no filter at all -- every thread replays its list and evaluates the
// pair forces branch-free; then advect + walls.  Items that the list path cannot take (more than
// MAX_TOTAL candidates, a list longer than its reservation, pool exhausted) go through the
// self-contained fallback kernels (k_density_fb / k_force_fb); none do in the benchmarks.
#pragma once
#include "tisph_device.cuh"

namespace tisph {

constexpr int LCAP = 64;               // pending-list slots per thread (shared memory)
constexpr int CHUNK = 32;              // candidates filtered between drain checks (fallback kernels)
constexpr float FAR = 1e18f;           // padding candidates / idle targets: never within the cutoff
constexpr int TCAP = 1792;             // candidates of one shared-memory tile (27 cells x 64 at the reference spacing = 1728)
constexpr int KCAP = 96;               // neighbour-list entries per thread and item in global memory
// Neighbour-list pool: an item reserves 1 + need rows of NB_THREADS uint2 words -- row 0 holds the
// per-thread word counts, then `need` <= KCAP/4 rows of list words ([word][thread]).  `need` is
// bounded by the tile size, so sparse cells take a few hundred bytes and dense ones 48 KiB.
constexpr int POOL_ROWS_PER_FULL_ITEM = KCAP / 4 + 1;

struct StepCounters {
    int n_items;        // built by k_items
    int work_d, work_f; // work-stealing cursors of the list kernels
    int n_fb_d, n_fb_f; // items handed to the fallback kernels
    int work_fb_d, work_fb_f;
    int pool_rows;      // bump allocator of the neighbour-list pool (rows of NB_THREADS 8-byte words)
};

// number of candidates in the 27-cell neighbourhood of cell c (9 contiguous ranges, see above)
__device__ __forceinline__ int cell_tile_total(const SimParams& sp, const int* __restrict__ cell_end, int c) {
    const int cz = c % sp.gz, cy = (c / sp.gz) % sp.gy, cx = c / (sp.gz * sp.gy);
    const int zlo = max(cz - 1, 0), zhi = min(cz + 1, sp.gz - 1);
    int total = 0;
    for (int dx = -1; dx <= 1; ++dx)
        for (int dy = -1; dy <= 1; ++dy) {
            const int x = cx + dx, y = cy + dy;
            if (x < 0 || x >= sp.gx || y < 0 || y >= sp.gy) continue;
            const int clo = (x * sp.gy + y) * sp.gz + zlo;
            total += cell_end[clo + (zhi - zlo)] - cell_end[max(clo - 1, 0)];
        }
    return total;
}

// item = {cell, first target relative to the cell's first particle | ITEM_HALF}.  An item holds up
// to 64 targets (4 splits per target) or, where the neighbourhood is dense, 32 (8 splits), which
// keeps the per-thread neighbour lists within KCAP.
constexpr int ITEM_HALF = 1 << 30;
constexpr int DENSE_TOTAL = 1920;

__global__ void __launch_bounds__(256)
k_items(SimParams sp, int key_lo, int key_hi, const int* __restrict__ cell_end, int2* __restrict__ items,
        StepCounters* __restrict__ ctr) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    int cnt = 0;
    if (c < sp.ncell && c >= key_lo && c < key_hi) cnt = cell_end[c] - cell_start(cell_end, c);
    int per = 64;
    if (cnt > 64 || (cnt > 32 && cell_tile_total(sp, cell_end, c) > DENSE_TOTAL)) per = 32;
    int ni = (cnt + per - 1) / per;
    // warp-aggregated reservation keeps the list roughly in cell order (L2 locality of the walks)
    int lane = threadIdx.x & 31;
    int inc = warp_inclusive_scan(ni, lane);
    int tot = __shfl_sync(0xffffffffu, inc, 31);
    int base = 0;
    if (lane == 31 && tot > 0) base = atomicAdd(&ctr->n_items, tot);
    base = __shfl_sync(0xffffffffu, base, 31) + inc - ni;
    for (int k = 0; k < ni; ++k) items[base + k] = make_int2(c, (k * per) | (per == 32 ? ITEM_HALF : 0));
}

struct ItemGeom {
    int c, tb, te, i0, nT, tl, nsplit, total;
};

// ranges + geometry of one item; all threads must call (contains __syncthreads)
__device__ __forceinline__ void item_setup(const SimParams& sp, const int* __restrict__ cell_end,
                                           int2 item, CellRanges& R, ItemGeom& G) {
    G.c = item.x;
    G.tb = cell_start(cell_end, G.c);
    G.te = cell_end[G.c];
    G.i0 = G.tb + (item.y & ~ITEM_HALF);
    G.nT = min(G.te - G.i0, (item.y & ITEM_HALF) ? 32 : 64);
    G.tl = G.nT <= 32 ? 32 : 64;
    G.nsplit = NB_THREADS / G.tl;
    compute_cell_ranges(sp, cell_end, G.c, R);
    G.total = R.off[9];
}

// ---- epilogues shared by the list and fallback kernels ------------------------------------
// density / boundary volume / clamp + Tait EOS   (wcsphv2.py:28-34,45-47 ; sph_basev2.py:195-201)
__device__ __forceinline__ void density_epilogue(const SimParams& sp, int i, float mass_i, int mat_i,
                                                 float wsum, float wbsum, int cnt,
                                                 float4* __restrict__ V, const float4* __restrict__ Q,
                                                 float4* __restrict__ D, float* __restrict__ S,
                                                 int* __restrict__ ncount) {
    float rho_raw, s_i = 0.f;
    if (mat_i == MAT_FLUID) {
        float self = mass_i * sp.k_w;                  // mass_i * W(0)
        s_i = mass_i * (sp.k_w * wsum);                // sum_j mass_i W(r_ij)   (Q2)
        rho_raw = sp.density_mode == 1 ? self + s_i : self;
    } else {
        rho_raw = Q[i].x;                              // boundary keeps its stored density
        float delta = sp.k_w + (sp.volume_mode == 1 ? sp.k_w * wbsum : 0.f);   // sph_basev2.py:195-201
        float4 vi = V[i];
        vi.w = 1.0f / delta;
        V[i] = vi;
    }
    float rho_c = fmaxf(rho_raw, sp.rho0);                                                       // :46
    float pr = sp.stiffness * (eos_pow(rho_c / sp.rho0, sp.exponent, sp.int_exponent) - 1.0f);   // :47
    D[i] = make_float4(rho_raw, pr / (rho_c * rho_c), rho_c, pr);
    S[i] = s_i;
    ncount[i] = cnt;
}

// a = g - non-pressure sums + pressure sums; advert; walls
// (wcsphv2.py:89-93, :53, :95-100 ; sph_basev2.py:151-189)
__device__ __forceinline__ void force_epilogue(const SimParams& sp, int i, bool walker, float4 pi, float4 vi,
                                               float4 di, float4 qi, float anx, float any, float anz,
                                               float apx, float apy, float apz,
                                               float4* __restrict__ Pout, float4* __restrict__ Vout,
                                               float4* __restrict__ Qout, float4* __restrict__ dvel,
                                               float4* __restrict__ a_np_out, float4* __restrict__ a_p_out) {
    float4 pout = pi, vout = vi, acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (walker) {
        float nx = sp.g[0] - anx, ny = sp.g[1] - any, nz = sp.g[2] - anz;
        if (a_np_out) {
            a_np_out[i] = make_float4(nx, ny, nz, 0.f);
            a_p_out[i] = make_float4(apx, apy, apz, 0.f);
        }
        acc.x = nx + apx; acc.y = ny + apy; acc.z = nz + apz;
        vout.x = vi.x + sp.dt * acc.x; vout.y = vi.y + sp.dt * acc.y; vout.z = vi.z + sp.dt * acc.z;
        float px = pi.x + sp.dt * vout.x, py = pi.y + sp.dt * vout.y, pz = pi.z + sp.dt * vout.z;
        float cnx = 0.f, cny = 0.f, cnz = 0.f;       // the wall tests use the pre-clamp position
        pout.x = px; pout.y = py; pout.z = pz;
        if (px > sp.wall_hi[0]) { cnx += 1.f; pout.x = sp.wall_hi[0]; }
        if (px <= sp.pad)       { cnx -= 1.f; pout.x = sp.pad; }
        if (py > sp.wall_hi[1]) { cny += 1.f; pout.y = sp.wall_hi[1]; }
        if (py <= sp.pad)       { cny -= 1.f; pout.y = sp.pad; }
        if (pz > sp.wall_hi[2]) { cnz += 1.f; pout.z = sp.wall_hi[2]; }
        if (pz <= sp.pad)       { cnz -= 1.f; pout.z = sp.pad; }
        float len = sqrtf(cnx * cnx + cny * cny + cnz * cnz);
        if (len > 1e-6f) {
            float ux = cnx / len, uy = cny / len, uz = cnz / len;
            float sdot = 1.5f * (vout.x * ux + vout.y * uy + vout.z * uz);   // :151-156
            vout.x -= sdot * ux; vout.y -= sdot * uy; vout.z -= sdot * uz;
        }
    } else if (a_np_out) {
        a_np_out[i] = acc;
        a_p_out[i] = acc;
    }
    Pout[i] = pout;
    Vout[i] = vout;
    Qout[i] = make_float4(di.z, di.w, qi.z, qi.w);     // clamped rho, p, material, orig id
    dvel[i] = acc;
}

// pair forces of one accepted neighbour (wcsphv2.py:56-80 ; sph_basev2.py:64-78)
//   pj = {x,y,z,psi}: psi = +mass_j (fluid j) / -volume_j (boundary j);  vj = {vx,vy,vz,rho_raw_j}
struct ForceAcc { float anx, any, anz, apx, apy, apz; };

__device__ __forceinline__ void pair_force(const SimParams& sp, float dx, float dy, float dz, float d2,
                                           float4 vi, float psi, float4 vj, float prj, float coh_i,
                                           float rho_i, float pr_i, float nub_i, ForceAcc& A) {
    float rinv = rsqrtf(fmaxf(d2, 1e-30f));
    float r = d2 * rinv;
    float q = r * sp.inv_h;
    // gradW = k_dw dw(q) x_ij / (r h); zero for r <= 1e-5 (sph_basev2.py:53)
    float gfac = r > 1e-5f ? sp.k_dw * spline_dw(q) * rinv * sp.inv_h : 0.f;
    float dot = (vi.x - vj.x) * dx + (vi.y - vj.y) * dy + (vi.z - vj.z) * dz;
    float mn = fminf(dot, 0.f) * fast_rcp(d2 + sp.eps_h2);
    float cn, cp;
    if (psi > 0.f) {                                                   // fluid j
        float w = sp.k_w * spline_w(q);
        float nu = sp.visc_fluid_c * fast_rcp(rho_i + vj.w);           // wcsphv2.py:69
        cn = psi * (coh_i * w - nu * mn * gfac);                       // :64 + :72-73
        cp = -psi * (pr_i + prj) * gfac;                               // sph_basev2.py:71-73
    } else {                                                           // boundary j
        float vol = -psi;
        cn = sp.ps_density0 * vol * (-nub_i * mn) * gfac;              // wcsphv2.py:78-80
        cp = -sp.rho0 * vol * pr_i * gfac;                             // sph_basev2.py:75
    }
    A.anx = fmaf(cn, dx, A.anx); A.any = fmaf(cn, dy, A.any); A.anz = fmaf(cn, dz, A.anz);
    A.apx = fmaf(cp, dx, A.apx); A.apy = fmaf(cp, dy, A.apy); A.apz = fmaf(cp, dz, A.apz);
}

__device__ __forceinline__ int next_item(int* cursor, int* s_slot) {
    __syncthreads();                       // everyone is done with the previous item's shared state
    if (threadIdx.x == 0) *s_slot = atomicAdd(cursor, 1);
    __syncthreads();
    return *s_slot;
}

// =======================================================================================
// Small PTX helpers of the list kernels
// =======================================================================================
__device__ __forceinline__ float rsqrt_approx(float x) {   // one MUFU.RSQ, no denormal fix-up code
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {     // one MUFU.RCP
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// 32-bit shared-window addressing: the generic->shared conversion is done once per kernel instead
// of once per access (ptxas re-derives the window base from SR_CgaCtaId inside hot loops otherwise)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float lds_f32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float4 lds_f32x4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float2 lds_f32x2(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;" :: "r"(a), "r"(v) : "memory");
}

// =======================================================================================
// Walk 1, list path
// =======================================================================================
// Tile, 32 bytes per candidate PAIR p (coordinates NEGATED so that x_i - x_j is one packed add):
//   float index 8p+{0,1} = -x   8p+{2,3} = -y   8p+{4,5} = -z   8p+{6,7} = material (i32 bits)
// A candidate is named by its "pair offset" o = 8p + (0|1) inside the density kernel and by its
// tile index e = 2p + (0|1) in the lists handed to the force kernel.  Pair TCAP/2 is a dummy pair
// that is always FAR away: list padding points at it.
constexpr int DCHUNK = 16;                                    // candidates filtered between drain checks
constexpr int O_DUMMY = 8 * (TCAP / 2);
constexpr int E_DUMMY = TCAP;
// A neighbour-list entry is (tile number << 11) | slot: slots of a tile are 0 .. TCAP-1 (swizzled
// candidate slots, cand_slot) plus E_DUMMY = TCAP, the always-FAR padding candidate of every tile.
constexpr int TSHIFT = 11;
constexpr int MAX_TOTAL = (0x10000 >> TSHIFT) * TCAP;         // 32 tiles: entries must fit 16 bits
constexpr size_t DL_SMEM = (size_t)(TCAP / 2 + 1) * 32 + (size_t)LCAP * NB_THREADS * sizeof(unsigned short);


// Split s of an item walks the candidate pairs s, s + nsplit, s + 2 nsplit, ... of the tile (pair-wise
// interleave keeps the splits' survivor counts balanced on lattice-like states).  In shared memory
// the pairs of one split are stored contiguously -- pair p lives at slot (p % nsplit) * S + p / nsplit,
// S = TCAP/2/nsplit -- so that the filter reads consecutive slots and the later random gathers of a
// warp (all lanes of a warp belong to one split) spread over all banks.
__device__ __forceinline__ int pair_slot(int p, int nsplit) {       // nsplit is 4 or 8
    const int ls = nsplit == 4 ? 2 : 3;
    return (p & (nsplit - 1)) * ((TCAP / 2) >> ls) + (p >> ls);
}
// swizzled slot of tile candidate e (used as the candidate's name in the neighbour lists)
__device__ __forceinline__ int cand_slot(int e, int nsplit) { return 2 * pair_slot(e >> 1, nsplit) + (e & 1); }

// Filter 8 consecutive pair slots and push the survivors' candidate slots (2 * pair slot + 0|1).
__device__ __forceinline__ void filter8(uint32_t a0, uint32_t e0, float2 xi2, float2 yi2, float2 zi2,
                                        float cut_wide, uint32_t sL, int& pend) {
    // loads are issued four pairs ahead of their use: shared-memory stores (the pushes) and loads keep
    // their program order, so an unbatched loop would expose one LDS latency per pair
#pragma unroll
    for (int b = 0; b < 8; b += 4) {
        float4 c[4];
        float2 cz[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            c[k] = lds_f32x4(a0 + 32u * (b + k));
            cz[k] = lds_f32x2(a0 + 32u * (b + k) + 16u);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            TISPH_CHECK(pend + 2 <= LCAP);
            float2 dx = __fadd2_rn(xi2, make_float2(c[k].x, c[k].y));
            float2 dy = __fadd2_rn(yi2, make_float2(c[k].z, c[k].w));
            float2 dz = __fadd2_rn(zi2, cz[k]);
            float2 s = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
            if (s.x < cut_wide) { sts_u16(sL + pend * (2 * NB_THREADS), e0 + 2u * (b + k)); ++pend; }
            if (s.y < cut_wide) { sts_u16(sL + pend * (2 * NB_THREADS), e0 + 2u * (b + k) + 1u); ++pend; }
        }
    }
}

template <bool AKINCI>
__global__ void __launch_bounds__(NB_THREADS, 3)
k_density_list(SimParams sp, const int* __restrict__ cell_end, const int2* __restrict__ items,
               StepCounters* __restrict__ ctr, int pool_rows_cap, int all_to_fallback,
               const float4* __restrict__ P, float4* __restrict__ V, const float4* __restrict__ Q,
               float4* __restrict__ D, float* __restrict__ S, int* __restrict__ ncount,
               uint2* __restrict__ Lg, int* __restrict__ item_row, unsigned char* __restrict__ flags,
               int* __restrict__ fb_d, int* __restrict__ fb_f) {
    extern __shared__ float4 dyn_smem[];
    float4* T = dyn_smem;                                                   // [TCAP/2 + 1][2]
    unsigned short* L = reinterpret_cast<unsigned short*>(dyn_smem + (TCAP / 2 + 1) * 2);
    __shared__ CellRanges R;
    __shared__ float red_w[NB_THREADS];
    __shared__ float red_b[NB_THREADS];
    __shared__ int red_c[NB_THREADS];
    __shared__ int s_slot, s_over, s_row;

    const int tid = threadIdx.x;
    const float cut_wide = sp.d2_cut * 1.000001f;       // superset filter; the drain applies the exact test
    const uint32_t sT = smem_u32(T);
    const uint32_t sL = smem_u32(L) + 2u * tid;         // my pending list: slot k at sL + k * 2 * NB_THREADS
    for (int k = 0; k < LCAP; ++k) L[k * NB_THREADS + tid] = (unsigned short)E_DUMMY;
    if (tid < 2) T[TCAP + tid] = tid == 0 ? make_float4(FAR, FAR, FAR, FAR) : make_float4(FAR, FAR, __int_as_float(MAT_FLUID), __int_as_float(MAT_FLUID));

    for (;;) {
        const int it = next_item(&ctr->work_d, &s_slot);
        if (it >= ctr->n_items) break;
        ItemGeom G;
        item_setup(sp, cell_end, items[it], R, G);
        if (G.total > MAX_TOTAL || all_to_fallback) {   // candidate indices must fit 16 bits
            if (tid == 0) {
                flags[it] = 2;
                fb_d[atomicAdd(&ctr->n_fb_d, 1)] = it;
                fb_f[atomicAdd(&ctr->n_fb_f, 1)] = it;
            }
            continue;
        }
        const bool own = G.c >= sp.own_key_lo && G.c < sp.own_key_hi;   // ghost cells get no force walk
        // rows of list words this item can need: every candidate of my share accepted, + padding per tile
        const int ntile = (G.total + TCAP - 1) / TCAP;
        const int share = (G.total + 2 * G.nsplit - 1) / (2 * G.nsplit) * 2;       // candidates one thread filters
        // ~15 % of them are neighbours (sphere / 27 cells); reserve for 60 % -- a list that still overflows
        // sends the item to the fallback force kernel
        const int need = min(KCAP / 4, (share * 3 / 5 + 3) / 4 + 2 * ntile + 1);
        if (tid == 0) {
            int row = -1;
            if (own) {
                row = atomicAdd(&ctr->pool_rows, need + 1);
                if (row + need + 1 > pool_rows_cap) row = -1;     // pool exhausted: this item takes the fallback force kernel
            }
            s_row = row;
            s_over = row < 0 ? 1 : 0;
            item_row[it] = row;
        }
        const int t_local = tid % G.tl, split = tid / G.tl;
        const int i = G.i0 + t_local;
        const bool active = t_local < G.nT;
        const float4 pi = active ? P[i] : make_float4(-FAR, -FAR, -FAR, 0.f);
        const int self_lo = R.gb[4], self_len = R.off[5] - R.off[4];
        const int self_t = (active && i >= self_lo && i < self_lo + self_len) ? R.off[4] + (i - self_lo) : -1;
        const float2 xi2 = make_float2(pi.x, pi.x), yi2 = make_float2(pi.y, pi.y), zi2 = make_float2(pi.z, pi.z);
        float wsum = 0.f, wbsum = 0.f;
        int cnt = 0, pend = 0, gword = 0;          // gword: 4-entry words already written to the global list
        // the candidates are walked tile by tile (one tile at the reference spacing); a ghost cell whose
        // density does not depend on its neighbours (reference modes) skips the walk altogether
        const int walk_total = (own || sp.ghost_walk) ? G.total : 0;
        for (int tile0 = 0; tile0 < walk_total; tile0 += TCAP) {
            const int tile_n = min(TCAP, walk_total - tile0);
            if (tile0 > 0) __syncthreads();                       // previous tile fully walked
            // ---- stage the tile -------------------------------------------------------------
            const int group = DCHUNK * G.nsplit;                  // candidates per round of all splits
            const int npair = (tile_n + group - 1) / group * (group / 2);      // staged pairs (padded with FAR)
            for (int p = tid; p < npair; p += NB_THREADS) {
                float4 a = make_float4(FAR, FAR, FAR, 0.f), b = a;
                int ma = MAT_FLUID, mb = MAT_FLUID;
                int e = 2 * p;
                if (e < tile_n) {
                    int g = tile_to_global(R, tile0 + e);
                    a = P[g];
                    if (AKINCI) ma = __float_as_int(Q[g].z);
                }
                if (e + 1 < tile_n) {
                    int g = tile_to_global(R, tile0 + e + 1);
                    b = P[g];
                    if (AKINCI) mb = __float_as_int(Q[g].z);
                }
                const int slot = pair_slot(p, G.nsplit);
                TISPH_CHECK(slot >= 0 && slot < TCAP / 2);
                T[2 * slot] = make_float4(-a.x, -b.x, -a.y, -b.y);
                T[2 * slot + 1] = make_float4(-a.z, -b.z, __int_as_float(ma), __int_as_float(mb));
            }
            __syncthreads();
            // ---- walk -----------------------------------------------------------------------
            const bool keep_list = s_row >= 0;                    // (written by thread 0 before the barrier)
            const int self_r = self_t - tile0;                    // my own slot in this tile, if any
            const uint32_t self_e = (self_r >= 0 && self_r < TCAP) ? (uint32_t)cand_slot(self_r, G.nsplit) : 0xffffffffu;
            uint2* gl = Lg + (size_t)(keep_list ? s_row + 1 : 0) * NB_THREADS + tid;
            const int mtot = npair / G.nsplit;                    // my pairs in this tile: split, split + nsplit, ...
            for (int m0 = 0; m0 < mtot; m0 += DCHUNK / 2) {
                const uint32_t p0 = (uint32_t)(split * ((TCAP / 2) >> (G.nsplit == 4 ? 2 : 3)) + m0);   // first pair slot of the chunk
                filter8(sT + 32u * p0, 2u * p0, xi2, yi2, zi2, cut_wide, sL, pend);
                const bool last = m0 + DCHUNK / 2 >= mtot;        // my last chunk of this tile
                if (last || __any_sync(0xffffffffu, pend > LCAP - DCHUNK)) {
                    // ---- drain: whole words of 4 entries; a remainder waits for the next round, the
                    //      last round of a tile pads with the tile's dummy candidate (FAR: q clamps to 1,
                    //      W = 0, not counted).  Branch-free per entry; every lane runs its own word
                    //      count.  The self pair stays in the list (it contributes exact zeros to the
                    //      force walk); its W is masked and its count is taken out at the end.
                    const int nd = last ? (pend + 3) & ~3 : pend & ~3;
                    if (last)
                        for (int k = pend; k < nd; ++k) sts_u16(sL + k * (2 * NB_THREADS), E_DUMMY);
                    uint2* gp = gl + (size_t)gword * NB_THREADS;
                    const int groom = keep_list ? need - gword : 0;       // words that still fit the item's list rows
                    const uint32_t tbase = (uint32_t)(tile0 / TCAP) << TSHIFT;
                    for (int k4 = 0; k4 < nd; k4 += 4, gp += NB_THREADS) {
                        uint32_t ew[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint32_t e = lds_u16(sL + (k4 + j) * (2 * NB_THREADS));
                            TISPH_CHECK(e <= (uint32_t)E_DUMMY && k4 + j < LCAP);
                            const uint32_t a = sT + ((e & ~1u) << 4) + ((e & 1u) << 2);   // pair slot * 32 + lane * 4
                            const float dx = pi.x + lds_f32(a), dy = pi.y + lds_f32(a + 8), dz = pi.z + lds_f32(a + 16);
                            const float d2 = dist2_exact(dx, dy, dz);
                            const float r = d2 * rsqrt_approx(fmaxf(d2, 1e-30f));
                            float w = spline_w(fminf(r * sp.inv_h, 1.0f));
                            w = e == self_e ? 0.f : w;            // p_i != p_j (the count is corrected at the end)
                            wsum += w;
                            if (d2 < sp.d2_cut) ++cnt;
                            if (AKINCI) wbsum += __float_as_int(lds_f32(a + 24)) == MAT_BOUNDARY ? w : 0.f;
                            ew[j] = tbase | e;
                        }
                        if ((k4 >> 2) < groom) {
                            TISPH_CHECK(s_row >= 0 && s_row + 1 + gword + (k4 >> 2) < pool_rows_cap &&
                                        gword + (k4 >> 2) < need);
                            *gp = make_uint2(ew[0] | (ew[1] << 16), ew[2] | (ew[3] << 16));
                        }
                    }
                    gword += nd >> 2;
                    // move the remainder (< 4 entries) to the front
                    const int rem = pend - nd;
                    for (int k = 0; k < rem; ++k) sts_u16(sL + k * (2 * NB_THREADS), lds_u16(sL + (nd + k) * (2 * NB_THREADS)));
                    pend = rem > 0 ? rem : 0;
                }
            }
            // every split drains completely at the end of a tile: pend == 0 at every tile boundary
        }
        if (walk_total <= 0) __syncthreads();              // no tile was staged: publish s_row (uniform branch)
        const bool keep_list = s_row >= 0;
        if (keep_list) {
            Lg[(size_t)s_row * NB_THREADS + tid] = make_uint2((unsigned)min(gword, need), 0u);   // count row, in words
            if (gword > need) s_over = 1;                  // benign race: every writer stores 1
        }
        red_w[tid] = wsum;
        red_b[tid] = wbsum;
        red_c[tid] = cnt;
        __syncthreads();
        if (tid == 0) {
            flags[it] = (unsigned char)s_over;
            if (s_over && own) fb_f[atomicAdd(&ctr->n_fb_f, 1)] = it;
        }
        if (split == 0 && active) {
            for (int s = 1; s < G.nsplit; ++s) {
                wsum += red_w[s * G.tl + t_local];
                wbsum += red_b[s * G.tl + t_local];
                cnt += red_c[s * G.tl + t_local];
            }
            const int mat_i = __float_as_int(Q[i].z);
            if (self_t >= 0 && walk_total > 0) cnt -= 1;   // the self pair was counted (p_i != p_j, partice_systemv4.py:344)
            density_epilogue(sp, i, pi.w, mat_i, wsum, wbsum, cnt, V, Q, D, S, ncount);
        }
    }
}

// =======================================================================================
// Walk 2, list path: forces + advect + walls
//   tile record, 36 B per candidate: tP = {x,y,z,psi}  tV = {vx,vy,vz,rho_raw}  tR = p/rho_c^2
//   slot E_DUMMY is a candidate that is FAR away (list padding points at it)
// =======================================================================================
constexpr int FTILE = TCAP + 4;
constexpr size_t FL_SMEM = (size_t)FTILE * (2 * sizeof(float4) + sizeof(float));

// nsplit > 0: candidate e goes to slot cand_slot(e, nsplit) (the name it has in the neighbour lists);
// nsplit == 0: slot e (fallback kernel)
__device__ __forceinline__ void stage_force_tile(const CellRanges& R, int tile0, int tile_n, int tile_pad, int nsplit,
                                                 const float4* __restrict__ Pin, const float4* __restrict__ Vin,
                                                 const float4* __restrict__ Qin, const float4* __restrict__ D,
                                                 float4* tP, float4* tV, float* tR) {
    for (int e = threadIdx.x; e < tile_pad; e += NB_THREADS) {
        float4 p = make_float4(FAR, FAR, FAR, 0.f), v = make_float4(0.f, 0.f, 0.f, 1.f);
        float pr = 0.f;
        if (e < tile_n) {
            int g = tile_to_global(R, tile0 + e);
            p = Pin[g];
            v = Vin[g];
            float4 d = D[g];
            if (__float_as_int(Qin[g].z) != MAT_FLUID) p.w = -v.w;   // psi = -volume
            v.w = d.x;
            pr = d.y;
        }
        const int slot = nsplit > 0 ? cand_slot(e, nsplit) : e;
        TISPH_CHECK(slot >= 0 && slot < TCAP);
        tP[slot] = p; tV[slot] = v; tR[slot] = pr;
    }
}

// Branch-free pair evaluation of the list kernel.  The self pair and coincident particles give
// exactly zero (x_ij = 0 and gradW = 0 for r <= 1e-5, sph_basev2.py:53), so no index test is needed;
// entries outside the cutoff (filter band, list padding) have q clamped to 1, where W and gradW vanish.
template <bool HAS_BOUNDARY>
__device__ __forceinline__ void pair_force_bf(const SimParams& sp, float kdw_h, float4 pi, float4 vi, float4 pj,
                                              float4 vj, float prj, float coh_i, float rho_i, float pr_i,
                                              float nub_i, ForceAcc& A) {
    const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
    const float d2 = dist2_exact(dx, dy, dz);
    const float rinv = rsqrt_approx(fmaxf(d2, 1e-30f));
    const float r = d2 * rinv;
    const float q = fminf(r * sp.inv_h, 1.0f);      // beyond the support W = gradW = 0: no mask needed
    const float f = 1.0f - q;
    const bool inner = q <= 0.5f;
    const float dw = inner ? q * fmaf(3.0f, q, -2.0f) : -f * f;              // sph_basev2.py:53-60 (/6k)
    const float gfac = r > 1e-5f ? dw * rinv * kdw_h : 0.f;                   // gradW = gfac * x_ij
    const float dot = (vi.x - vj.x) * dx + (vi.y - vj.y) * dy + (vi.z - vj.z) * dz;
    const float mn = fminf(dot, 0.f) * rcp_approx(d2 + sp.eps_h2);
    const float psi = pj.w;
    float cn, cp;
    {                                                                        // fluid j
        const float w = sp.k_w * (inner ? fmaf(6.0f * q * q, q - 1.0f, 1.0f) : 2.0f * f * f * f);
        const float nu = sp.visc_fluid_c * rcp_approx(rho_i + vj.w);          // wcsphv2.py:69
        cn = psi * (coh_i * w - nu * mn * gfac);                              // :64 + :72-73
        cp = -psi * (pr_i + prj) * gfac;                                      // sph_basev2.py:71-73
    }
    if (HAS_BOUNDARY) {
        const float vol = -psi;
        const float cnb = sp.ps_density0 * vol * (-nub_i * mn) * gfac;        // wcsphv2.py:78-80
        const float cpb = -sp.rho0 * vol * pr_i * gfac;                       // sph_basev2.py:75
        cn = psi > 0.f ? cn : cnb;
        cp = psi > 0.f ? cp : cpb;
    }
    A.anx = fmaf(cn, dx, A.anx); A.any = fmaf(cn, dy, A.any); A.anz = fmaf(cn, dz, A.anz);
    A.apx = fmaf(cp, dx, A.apx); A.apy = fmaf(cp, dy, A.apy); A.apz = fmaf(cp, dz, A.apz);
}

template <bool HAS_BOUNDARY>
__global__ void __launch_bounds__(NB_THREADS, 3)
k_force_list(SimParams sp, const int* __restrict__ cell_end, const int2* __restrict__ items,
             StepCounters* __restrict__ ctr, const float4* __restrict__ Pin,
             const float4* __restrict__ Vin, const float4* __restrict__ Qin,
             const float4* __restrict__ D, float4* __restrict__ Pout, float4* __restrict__ Vout,
             float4* __restrict__ Qout, float4* __restrict__ dvel, float4* __restrict__ a_np_out,
             float4* __restrict__ a_p_out, const uint2* __restrict__ Lg,
             const int* __restrict__ item_row, const unsigned char* __restrict__ flags) {
    extern __shared__ float4 dyn_smem[];
    float4* tP = dyn_smem;
    float4* tV = dyn_smem + FTILE;
    float* tR = reinterpret_cast<float*>(dyn_smem + 2 * FTILE);
    __shared__ CellRanges R;
    __shared__ float red[6][NB_THREADS];
    __shared__ int s_slot;
    const int tid = threadIdx.x;
    const uint32_t sP = smem_u32(tP), sR = smem_u32(tR);
    constexpr uint32_t V_OFF = (uint32_t)FTILE * 16u;
    const float kdw_h = sp.k_dw * sp.inv_h;
    if (tid == 0) {
        tP[E_DUMMY] = make_float4(FAR, FAR, FAR, 1.f);
        tV[E_DUMMY] = make_float4(0.f, 0.f, 0.f, 1.f);
        tR[E_DUMMY] = 0.f;
    }

    for (;;) {
        const int it = next_item(&ctr->work_f, &s_slot);
        if (it >= ctr->n_items) break;
        if (flags[it]) continue;                         // handled by k_force_fb
        if (items[it].x < sp.own_key_lo || items[it].x >= sp.own_key_hi) continue;   // ghost cell: not advanced here
        ItemGeom G;
        item_setup(sp, cell_end, items[it], R, G);
        const int t_local = tid % G.tl, split = tid / G.tl;
        const int i = G.i0 + t_local;
        const bool active = t_local < G.nT;
        const float4 pi = active ? Pin[i] : make_float4(-FAR, -FAR, -FAR, 1.f);
        const float4 vi = active ? Vin[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 di = active ? D[i] : make_float4(1.f, 0.f, 1.f, 0.f);
        const float4 qi = active ? Qin[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        const bool walker = active && __float_as_int(qi.z) == MAT_FLUID;
        const float coh_i = 0.01f / pi.w;                         // wcsphv2.py:64
        const float rho_i = di.x, pr_i = di.y;
        const float nub_i = sp.visc_bound_c / (2.0f * rho_i);     // wcsphv2.py:76
        ForceAcc A = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        TISPH_CHECK(item_row[it] >= 0);
        const uint2* gl = Lg + (size_t)item_row[it] * NB_THREADS + tid;
        const int nw = walker ? (int)gl[0].x : 0;                 // words of 4 entries (count row)
        TISPH_CHECK(nw >= 0 && nw <= KCAP / 4);
        gl += NB_THREADS;
        uint2 w = nw > 0 ? gl[0] : make_uint2(0u, 0u);
        if (G.total <= TCAP) {
            // ---- one tile (the common case): replay the whole list
            stage_force_tile(R, 0, G.total, G.total, G.nsplit, Pin, Vin, Qin, D, tP, tV, tR);
            __syncthreads();                                   // tile staged
            for (int k = 0; k < nw; ++k) {
                const uint2 cur = w;
                if (k + 1 < nw) w = gl[(size_t)(k + 1) * NB_THREADS];     // prefetch the next 4 entries
                const uint32_t e4[4] = {cur.x & 0xffffu, cur.x >> 16, cur.y & 0xffffu, cur.y >> 16};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t e = e4[j];                              // slot of tile 0 (or its dummy)
                    TISPH_CHECK(e <= (uint32_t)E_DUMMY);
                    const uint32_t a = sP + 16u * e;
                    const float4 pj = lds_f32x4(a);
                    const float4 vj = lds_f32x4(a + V_OFF);
                    const float prj = lds_f32(sR + 4u * e);
                    pair_force_bf<HAS_BOUNDARY>(sp, kdw_h, pi, vi, pj, vj, prj, coh_i, rho_i, pr_i, nub_i, A);
                }
            }
        } else {
            // ---- several tiles: every list is ascending, so a tile's entries are one run of it;
            //      a word that straddles a tile boundary is replayed in both tiles with the
            //      foreign entries redirected to the dummy candidate
            int k = 0;
            for (int tile0 = 0; tile0 < G.total; tile0 += TCAP) {
                const int tile_n = min(TCAP, G.total - tile0);
                const uint32_t tbase = (uint32_t)(tile0 / TCAP) << TSHIFT;
                if (tile0 > 0) __syncthreads();
                stage_force_tile(R, tile0, tile_n, tile_n, G.nsplit, Pin, Vin, Qin, D, tP, tV, tR);
                __syncthreads();
                while (k < nw) {
                    const uint32_t e4[4] = {w.x & 0xffffu, w.x >> 16, w.y & 0xffffu, w.y >> 16};
                    bool done = true;                          // every real entry of the word lies below the tile's end
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t rel = e4[j] - tbase;
                        const uint32_t e = rel < (1u << TSHIFT) ? rel : (uint32_t)E_DUMMY;
                        done = done && e4[j] < tbase + (1u << TSHIFT);
                        const uint32_t a = sP + 16u * e;
                        const float4 pj = lds_f32x4(a);
                        const float4 vj = lds_f32x4(a + V_OFF);
                        const float prj = lds_f32(sR + 4u * e);
                        pair_force_bf<HAS_BOUNDARY>(sp, kdw_h, pi, vi, pj, vj, prj, coh_i, rho_i, pr_i, nub_i, A);
                    }
                    if (!done) break;
                    ++k;
                    if (k < nw) w = gl[(size_t)k * NB_THREADS];
                }
            }
        }
        red[0][tid] = A.anx; red[1][tid] = A.any; red[2][tid] = A.anz;
        red[3][tid] = A.apx; red[4][tid] = A.apy; red[5][tid] = A.apz;
        __syncthreads();
        if (split == 0 && active) {
            for (int s = 1; s < G.nsplit; ++s) {
                int o = s * G.tl + t_local;
                A.anx += red[0][o]; A.any += red[1][o]; A.anz += red[2][o];
                A.apx += red[3][o]; A.apy += red[4][o]; A.apz += red[5][o];
            }
            force_epilogue(sp, i, walker, pi, vi, di, qi, A.anx, A.any, A.anz, A.apx, A.apy, A.apz,
                           Pout, Vout, Qout, dvel, a_np_out, a_p_out);
        }
    }
}

// =======================================================================================
// Fallback kernels: any tile size (multi-tile loop), scalar exact filter, no global lists.
// They run over the (normally empty) fallback item lists.
// =======================================================================================
constexpr size_t DF_SMEM = (size_t)TCAP * sizeof(float4) + (size_t)LCAP * NB_THREADS * sizeof(unsigned short);

__global__ void __launch_bounds__(NB_THREADS, 3)
k_density_fb(SimParams sp, const int* __restrict__ cell_end, const int2* __restrict__ items,
             StepCounters* __restrict__ ctr, const int* __restrict__ fb_d, const float4* __restrict__ P,
             float4* __restrict__ V, const float4* __restrict__ Q, float4* __restrict__ D,
             float* __restrict__ S, int* __restrict__ ncount) {
    extern __shared__ float4 dyn_smem[];
    float4* tile = dyn_smem;                                           // {x,y,z,material}
    unsigned short* L = reinterpret_cast<unsigned short*>(dyn_smem + TCAP);
    __shared__ CellRanges R;
    __shared__ float red_w[NB_THREADS];
    __shared__ float red_b[NB_THREADS];
    __shared__ int red_c[NB_THREADS];
    __shared__ int s_slot;
    const int tid = threadIdx.x;
    const bool akinci = sp.volume_mode == 1;
    unsigned short* myL = L + tid;

    for (;;) {
        const int w = next_item(&ctr->work_fb_d, &s_slot);
        if (w >= ctr->n_fb_d) break;
        const int it = fb_d[w];
        ItemGeom G;
        item_setup(sp, cell_end, items[it], R, G);
        const int t_local = tid % G.tl, split = tid / G.tl;
        const int i = G.i0 + t_local;
        const bool active = t_local < G.nT;
        const float4 pi = active ? P[i] : make_float4(-FAR, -FAR, -FAR, 0.f);
        const int self_lo = R.gb[4], self_len = R.off[5] - R.off[4];
        const int self_e = (active && i >= self_lo && i < self_lo + self_len) ? R.off[4] + (i - self_lo) : -1;
        float wsum = 0.f, wbsum = 0.f;
        int cnt = 0;
        for (int tile0 = 0; tile0 < G.total; tile0 += TCAP) {
            const int tile_n = min(TCAP, G.total - tile0);
            const int tile_pad = (tile_n + CHUNK - 1) & ~(CHUNK - 1);
            __syncthreads();
            for (int e = tid; e < tile_pad; e += NB_THREADS) {
                float4 p = make_float4(FAR, FAR, FAR, 0.f);
                if (e < tile_n) {
                    int g = tile_to_global(R, tile0 + e);
                    p = P[g];
                    p.w = akinci ? Q[g].z : 0.f;
                }
                tile[e] = p;
            }
            __syncthreads();
            const int self_t = self_e - tile0;
            int pend = 0;
            for (int cb = split * CHUNK; cb < tile_pad; cb += G.nsplit * CHUNK) {
#pragma unroll 8
                for (int k = 0; k < CHUNK; ++k) {
                    float4 cj = tile[cb + k];
                    float d2 = dist2_exact(pi.x - cj.x, pi.y - cj.y, pi.z - cj.z);
                    if (d2 < sp.d2_cut) { myL[pend * NB_THREADS] = (unsigned short)(cb + k); ++pend; }
                }
                const bool last = cb + G.nsplit * CHUNK >= tile_pad;
                if (last || __any_sync(0xffffffffu, pend > LCAP - CHUNK)) {
                    for (int k = 0; k < pend; ++k) {
                        int e = myL[k * NB_THREADS];
                        if (e == self_t) continue;
                        float4 cj = tile[e];
                        float d2 = dist2_exact(pi.x - cj.x, pi.y - cj.y, pi.z - cj.z);
                        float r = d2 * rsqrtf(fmaxf(d2, 1e-30f));
                        float wv = spline_w(r * sp.inv_h);
                        cnt++;
                        wsum += wv;
                        if (__float_as_int(cj.w) == MAT_BOUNDARY) wbsum += wv;
                    }
                    pend = 0;
                }
            }
        }
        red_w[tid] = wsum;
        red_b[tid] = wbsum;
        red_c[tid] = cnt;
        __syncthreads();
        if (split == 0 && active) {
            for (int s = 1; s < G.nsplit; ++s) {
                wsum += red_w[s * G.tl + t_local];
                wbsum += red_b[s * G.tl + t_local];
                cnt += red_c[s * G.tl + t_local];
            }
            density_epilogue(sp, i, pi.w, __float_as_int(Q[i].z), wsum, wbsum, cnt, V, Q, D, S, ncount);
        }
    }
}

constexpr size_t FF_SMEM = FL_SMEM + (size_t)LCAP * NB_THREADS * sizeof(unsigned short);

__global__ void __launch_bounds__(NB_THREADS, 2)
k_force_fb(SimParams sp, const int* __restrict__ cell_end, const int2* __restrict__ items,
           StepCounters* __restrict__ ctr, const int* __restrict__ fb_f, const float4* __restrict__ Pin,
           const float4* __restrict__ Vin, const float4* __restrict__ Qin,
           const float4* __restrict__ D, float4* __restrict__ Pout, float4* __restrict__ Vout,
           float4* __restrict__ Qout, float4* __restrict__ dvel, float4* __restrict__ a_np_out,
           float4* __restrict__ a_p_out) {
    extern __shared__ float4 dyn_smem[];
    float4* tP = dyn_smem;
    float4* tV = dyn_smem + FTILE;
    float* tR = reinterpret_cast<float*>(dyn_smem + 2 * FTILE);
    unsigned short* L = reinterpret_cast<unsigned short*>(tR + FTILE);
    __shared__ CellRanges R;
    __shared__ float red[6][NB_THREADS];
    __shared__ int s_slot;
    const int tid = threadIdx.x;
    unsigned short* myL = L + tid;

    for (;;) {
        const int wk = next_item(&ctr->work_fb_f, &s_slot);
        if (wk >= ctr->n_fb_f) break;
        const int it = fb_f[wk];
        if (items[it].x < sp.own_key_lo || items[it].x >= sp.own_key_hi) continue;   // ghost cell
        ItemGeom G;
        item_setup(sp, cell_end, items[it], R, G);
        const int t_local = tid % G.tl, split = tid / G.tl;
        const int i = G.i0 + t_local;
        const bool active = t_local < G.nT;
        const float4 pi = active ? Pin[i] : make_float4(-FAR, -FAR, -FAR, 1.f);
        const float4 vi = active ? Vin[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 di = active ? D[i] : make_float4(1.f, 0.f, 1.f, 0.f);
        const float4 qi = active ? Qin[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        const bool walker = active && __float_as_int(qi.z) == MAT_FLUID;
        const float xi = walker ? pi.x : -FAR, yi = walker ? pi.y : -FAR, zi = walker ? pi.z : -FAR;
        const int self_lo = R.gb[4], self_len = R.off[5] - R.off[4];
        const int self_e = (active && i >= self_lo && i < self_lo + self_len) ? R.off[4] + (i - self_lo) : -1;
        const float coh_i = 0.01f / pi.w;
        const float rho_i = di.x, pr_i = di.y;
        const float nub_i = sp.visc_bound_c / (2.0f * rho_i);
        ForceAcc A = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int tile0 = 0; tile0 < G.total; tile0 += TCAP) {
            const int tile_n = min(TCAP, G.total - tile0);
            const int tile_pad = (tile_n + CHUNK - 1) & ~(CHUNK - 1);
            __syncthreads();
            stage_force_tile(R, tile0, tile_n, tile_pad, 0, Pin, Vin, Qin, D, tP, tV, tR);
            __syncthreads();
            const int self_t = self_e - tile0;
            int pend = 0;
            for (int cb = split * CHUNK; cb < tile_pad; cb += G.nsplit * CHUNK) {
#pragma unroll 8
                for (int k = 0; k < CHUNK; ++k) {
                    float4 pj = tP[cb + k];
                    float d2 = dist2_exact(xi - pj.x, yi - pj.y, zi - pj.z);
                    if (d2 < sp.d2_cut) { myL[pend * NB_THREADS] = (unsigned short)(cb + k); ++pend; }
                }
                const bool last = cb + G.nsplit * CHUNK >= tile_pad;
                if (last || __any_sync(0xffffffffu, pend > LCAP - CHUNK)) {
                    for (int k = 0; k < pend; ++k) {
                        int e = myL[k * NB_THREADS];
                        if (e == self_t) continue;
                        float4 pj = tP[e];
                        float dx = xi - pj.x, dy = yi - pj.y, dz = zi - pj.z;
                        float d2 = dist2_exact(dx, dy, dz);
                        pair_force(sp, dx, dy, dz, d2, vi, pj.w, tV[e], tR[e], coh_i, rho_i, pr_i, nub_i, A);
                    }
                    pend = 0;
                }
            }
        }
        red[0][tid] = A.anx; red[1][tid] = A.any; red[2][tid] = A.anz;
        red[3][tid] = A.apx; red[4][tid] = A.apy; red[5][tid] = A.apz;
        __syncthreads();
        if (split == 0 && active) {
            for (int s = 1; s < G.nsplit; ++s) {
                int o = s * G.tl + t_local;
                A.anx += red[0][o]; A.any += red[1][o]; A.anz += red[2][o];
                A.apx += red[3][o]; A.apy += red[4][o]; A.apz += red[5][o];
            }
            force_epilogue(sp, i, walker, pi, vi, di, qi, A.anx, A.any, A.anz, A.apx, A.apy, A.apz,
                           Pout, Vout, Qout, dvel, a_np_out, a_p_out);
        }
    }
}

}  // namespace tisph
