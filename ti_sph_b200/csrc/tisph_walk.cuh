// tisph_walk.cuh -- what the neighbour walks of a WCSPH step share, and the fallback kernels.
//
// Work items.  A walk is cut into ITEMS: one item = up to 64 target particles of one occupied grid
// cell (k_items builds the list after the scan).  Persistent CTAs of 256 threads pull items from
// an atomic cursor, so the ~97 % empty cells of a dam-break grid cost nothing.
//
// Candidates.  The 27 neighbour cells of a cell are 9 contiguous ranges of the sorted arrays (z is
// the fastest key digit).  Range of cell c is [cell_end[max(0,c-1)], cell_end[c])
// (partice_systemv4.py:343; cell 0 is therefore invisible as a neighbour -- reference quirk,
// reproduced).  Cells outside the grid are empty (the reference reads out of bounds there).
//
// The list kernels that every benchmark runs are in tisph_lists.cuh.  Items they cannot take (more
// than one tile of candidates, a neighbour list longer than its reservation, list pool exhausted)
// go through the self-contained fallback kernels below (k_density_fb / k_force_fb): threads
// arranged [split][target], 64 targets x 4 splits, candidate tiles of TCAP staged in turn, scalar
// exact filter into per-thread pending lists, no global lists.
#pragma once
#include "tisph_device.cuh"

namespace tisph {

constexpr int LCAP = 64;               // pending-list slots per thread (shared memory)
constexpr int CHUNK = 32;              // candidates filtered between drain checks (fallback kernels)
constexpr float FAR = 1e18f;           // padding candidates / idle targets: never within the cutoff
constexpr int TCAP = 1792;             // candidates of one shared-memory tile of the fallback kernels

struct StepCounters {
    int n_items;        // built by k_items
    int work_d, work_f; // work-stealing cursors of the list kernels
    int n_fb_d, n_fb_f; // items handed to the fallback kernels
    int work_fb_d, work_fb_f;
    int pool_rows;      // bump allocator of the neighbour-list pool (rows of NB_THREADS 32-bit words)
};

// item = {cell, first target relative to the cell's first particle}: up to 64 targets
constexpr int ITEM_HALF = 1 << 30;     // (flag of an item format with 32 targets; not produced any more)

__global__ void __launch_bounds__(256)
k_items(SimParams sp, int key_lo, int key_hi, const int* __restrict__ cell_end, int2* __restrict__ items,
        StepCounters* __restrict__ ctr) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    int cnt = 0;
    if (c < sp.ncell && c >= key_lo && c < key_hi) cnt = cell_end[c] - cell_start(cell_end, c);
    const int per = 64;
    int ni = (cnt + per - 1) / per;
    // warp-aggregated reservation keeps the list roughly in cell order (L2 locality of the walks)
    int lane = threadIdx.x & 31;
    int inc = warp_inclusive_scan(ni, lane);
    int tot = __shfl_sync(0xffffffffu, inc, 31);
    int base = 0;
    if (lane == 31 && tot > 0) base = atomicAdd(&ctr->n_items, tot);
    base = __shfl_sync(0xffffffffu, base, 31) + inc - ni;
    for (int k = 0; k < ni; ++k) items[base + k] = make_int2(c, k * per);
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
    float rho_raw, s_i = 0.f, psi = mass_i;           // psi: what a neighbour's pair terms are weighted with
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
        psi = -vi.w;                                   // boundary neighbours enter with their volume
    }
    float rho_c = fmaxf(rho_raw, sp.rho0);                                                       // :46
    float pr = sp.stiffness * (eos_pow(rho_c / sp.rho0, sp.exponent, sp.int_exponent) - 1.0f);   // :47
    // D = {unclamped density, p / rho_c^2, psi (+mass | -volume), p}; the clamped density is max(D.x, rho0)
    D[i] = make_float4(rho_raw, pr / (rho_c * rho_c), psi, pr);
    S[i] = s_i;
    ncount[i] = cnt;
}

// enforce_boundary_3D_v1 + simulate_collisions_v1 (sph_basev2.py:151-189): clamp to the padded box, reflect
// the velocity about the accumulated wall normal; the wall tests use the pre-clamp position
__device__ __forceinline__ void apply_walls(const SimParams& sp, float4& pout, float4& vout) {
    const float px = pout.x, py = pout.y, pz = pout.z;
    float cnx = 0.f, cny = 0.f, cnz = 0.f;
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
}

// enforce_boundary() as a launch of its own (TISPH_STAGE_WALLS)
__global__ void __launch_bounds__(256)
k_walls(SimParams sp, int n, float4* __restrict__ P, float4* __restrict__ V, const float4* __restrict__ Q) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || __float_as_int(Q[i].z) != MAT_FLUID) return;
    float4 p = P[i], v = V[i];
    apply_walls(sp, p, v);
    P[i] = p; V[i] = v;
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
        pout.x = pi.x + sp.dt * vout.x; pout.y = pi.y + sp.dt * vout.y; pout.z = pi.z + sp.dt * vout.z;
        if (sp.walls) apply_walls(sp, pout, vout);
    } else if (a_np_out) {
        a_np_out[i] = acc;
        a_p_out[i] = acc;
    }
    Pout[i] = pout;
    Vout[i] = vout;
    Qout[i] = make_float4(fmaxf(di.x, sp.rho0), di.w, qi.z, qi.w);     // clamped rho (wcsphv2.py:46), p, material, orig id
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

// ---- item pipeline of the list kernels ----------------------------------------------------------
// Fetching an item is a chain of dependent global accesses (cursor atomic -> item record -> cell_end
// ranges).  Warp 0 therefore fetches item k+1 into registers right after item k has been published, so
// that the chain runs behind item k's walk; two barriers per item publish it to the CTA.
struct ItemFetch {
    int it;         // item index (>= n_items: the list is exhausted)
    int2 item;
    int a, b;       // lane < 9: first sorted index / length of range `lane`;  lane 9: first / end index of the cell;
                    // lane 10: a = the item's flag (force walk)
};
struct ItemMeta { int it, cell, first, tb, te, flag; };

__device__ __forceinline__ ItemFetch fetch_item(const SimParams& sp, const int* __restrict__ cell_end,
                                                const int2* __restrict__ items, int n_items, int* cursor,
                                                const unsigned char* __restrict__ flags = nullptr) {
    const int lane = threadIdx.x & 31;
    ItemFetch f;
    int it = 0;
    if (lane == 0) it = atomicAdd(cursor, 1);
    f.it = __shfl_sync(0xffffffffu, it, 0);
    f.item = make_int2(0, 0);
    f.a = f.b = 0;
    if (f.it < n_items) {
        f.item = items[f.it];
        const int c = f.item.x;
        if (lane < 9) {
            const int cz = c % sp.gz, cy = (c / sp.gz) % sp.gy, cx = c / (sp.gz * sp.gy);
            const int x = cx + lane / 3 - 1, y = cy + lane % 3 - 1;
            if (x >= 0 && x < sp.gx && y >= 0 && y < sp.gy) {
                const int zlo = max(cz - 1, 0), zhi = min(cz + 1, sp.gz - 1);
                const int clo = (x * sp.gy + y) * sp.gz + zlo;
                f.a = cell_end[max(clo - 1, 0)];
                f.b = cell_end[clo + (zhi - zlo)] - f.a;
            }
        } else if (lane == 9) {
            f.a = cell_start(cell_end, c);
            f.b = cell_end[c];
        } else if (lane == 10 && flags) {
            f.a = flags[f.it];
        }
    }
    return f;
}

// warp 0: make the fetched item the CTA's current one (the caller puts a barrier on either side)
__device__ __forceinline__ void publish_item(const ItemFetch& f, CellRanges& R, ItemMeta& M) {
    const int lane = threadIdx.x & 31;
    const int len = lane < 9 ? f.b : 0;
    const int inc = warp_inclusive_scan(len, lane);
    if (lane < 9) { R.gb[lane] = f.a; R.off[lane] = inc - len; }
    if (lane == 8) R.off[9] = inc;
    if (lane == 9) { M.tb = f.a; M.te = f.b; }
    if (lane == 10) M.flag = f.a;
    if (lane == 0) { M.it = f.it; M.cell = f.item.x; M.first = f.item.y; }
}

__device__ __forceinline__ void item_geometry(const ItemMeta& M, const CellRanges& R, ItemGeom& G) {
    G.c = M.cell;
    G.tb = M.tb;
    G.te = M.te;
    G.i0 = G.tb + (M.first & ~ITEM_HALF);
    G.nT = min(G.te - G.i0, (M.first & ITEM_HALF) ? 32 : 64);
    G.tl = G.nT <= 32 ? 32 : 64;
    G.nsplit = NB_THREADS / G.tl;
    G.total = R.off[9];
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
// Fallback kernels: any number of candidates (tile loop), scalar exact filter, no global lists.
// They run over the (normally empty) fallback item lists.
//   force tile record, 36 B per candidate: tP = {x,y,z,psi}  tV = {vx,vy,vz,rho_raw}  tR = p/rho_c^2
// =======================================================================================
constexpr int FTILE = TCAP + 4;
constexpr size_t FB_FL_SMEM = (size_t)FTILE * (2 * sizeof(float4) + sizeof(float));

__device__ __forceinline__ void stage_force_tile(const CellRanges& R, int tile0, int tile_n, int tile_pad,
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
        TISPH_CHECK(e >= 0 && e < TCAP);
        tP[e] = p; tV[e] = v; tR[e] = pr;
    }
}

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

constexpr size_t FF_SMEM = FB_FL_SMEM + (size_t)LCAP * NB_THREADS * sizeof(unsigned short);

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
            stage_force_tile(R, tile0, tile_n, tile_pad, Pin, Vin, Qin, D, tP, tV, tR);
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
